// tcgen05 tensor-core MLP engine (NRH_MLP_TCGEN05) -- internal interface.
#pragma once
#include "nrh_common.cuh"

namespace nrh {

bool tc_available();
size_t tc_packed_bytes(const NrhConfig& cfg);
size_t tc_scratch_bytes(int num_sms);
// builds the tensor-core operand images (fp16 hi/lo, UMMA tile order) from the fp32 section
int tc_pack(const NrhConfig& cfg, const PackedLayout& L, const NrhRawWeights& raw, void* packed, cudaStream_t st);

int sdf_mlp_tc(const void* packed, const PackedLayout& L, Strided3 pts, int64_t N,
               float* sdf, float* gx, float* gy, float* gz, int64_t grad_stride, float* feat,
               float* scratch, size_t scratch_bytes, int num_sms, cudaStream_t st);
int color_mlp_tc(const void* packed, const PackedLayout& L, Strided3 pts, Strided3 normals,
                 const float* feat, const float* rayfeat, int64_t R, int64_t N,
                 float* cr, float* cg, float* cb, float* scratch, size_t scratch_bytes, int num_sms, cudaStream_t st);

}  // namespace nrh
