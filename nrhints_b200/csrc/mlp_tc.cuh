// tcgen05 tensor-core MLP engine (NRH_MLP_TCGEN05) -- internal interface.
#pragma once
#include "nrh_common.cuh"

namespace nrh {

// operand scaling shared with the kernels that pre-build operand images (powers of two: exact)
constexpr float TC_W_SCALE = 64.0f;            // weights are stored * 2^6
constexpr float TC_ACT_SCALE = 16.0f;          // forward activations are stored * 2^4
constexpr uint32_t TC_TILE_FEAT_BYTES = 65536; // one tile of features as operand images: 4 chunks x [128 x 64] fp16
constexpr uint32_t TC_TILE_AUX_BYTES = 32768;  // one ray block of non-feature reflectance inputs: 2 chunks

bool tc_available();
size_t tc_packed_bytes(const NrhConfig& cfg);
size_t tc_scratch_bytes(int num_sms);
// builds the tensor-core operand images (fp16 hi/lo, UMMA tile order) from the fp32 section
int tc_pack(const NrhConfig& cfg, const PackedLayout& L, const NrhRawWeights& raw, void* packed, cudaStream_t st);

int sdf_mlp_tc(const void* packed, const PackedLayout& L, Strided3 pts, int64_t N,
               float* sdf, float* gx, float* gy, float* gz, int64_t grad_stride, float* feat, bool feat_as_image,
               float* scratch, size_t scratch_bytes, int num_sms, cudaStream_t st);
// feat_as_image: `feat` receives, per 128-point tile, TC_TILE_FEAT_BYTES of fp16 (x TC_ACT_SCALE) operand images in the
// shared-memory layout of the reflectance kernel's A operand instead of fp32 rows (N must be a multiple of 128).
int color_mlp_tc(const void* packed, const PackedLayout& L, Strided3 pts, Strided3 normals,
                 const float* feat, const float* rayfeat, const void* aux_img, int64_t R, int64_t N,
                 float* cr, float* cg, float* cb, float* scratch, size_t scratch_bytes, int num_sms, cudaStream_t st);
// aux_img != nullptr selects the streamed-input path: `feat` holds operand images (see above) and aux_img holds, per
// block of 128 rays, TC_TILE_AUX_BYTES of operand images of the per-ray inputs (built by k_shade_prep); needs R % 128 == 0.

}  // namespace nrh
