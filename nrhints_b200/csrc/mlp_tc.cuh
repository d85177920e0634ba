// tcgen05 tensor-core MLP engine (NRH_MLP_TCGEN05) -- internal interface.
#pragma once
#include "nrh_common.cuh"

namespace nrh {

// operand scaling shared with the kernels that pre-build operand images (powers of two: exact)
constexpr float TC_W_SCALE = 64.0f;            // weights are stored * 2^6
constexpr float TC_ACT_SCALE = 16.0f;          // forward activations are stored * 2^4
constexpr uint32_t TC_TILE_FEAT_BYTES = 65536; // one tile of features as operand images: 4 chunks x [128 x 64] fp16
constexpr uint32_t TC_TILE_AUX_BYTES = 32768;  // one ray block of non-feature reflectance inputs: 2 chunks

// host helper (wgrad_tc.cu): CUtensorMap (128 bytes, 64-byte aligned, passed as void*) of a row-major fp16 matrix, SWIZZLE_128B boxes
int encode_tensor_map_f16(void* out_map, const void* ptr, long long ld, long long rows, int box_cols, int box_rows);

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

// ---- training: SDF fine pass with a tape, and its backward (mlp_tc_bwd.inc) -----------------------------------------------
// Buffer layouts (bytes) for N points; P_pad = N rounded up to 128.
//   tape    : [tiles]  per 128-point tile: packed softplus' of the 8 layers, packed reverse adjoints g_1..g_7, g_e
//             [act]    8 x [P_pad][256] fp16, a_1 .. a_8 (x 16), row-major      (layer inputs: dW_l = zb_l^T a_l)
//             [u]      8 x [P_pad][256] fp16, u_0 .. u_7 (x 1024), row-major    (reverse-sweep signals: dW_l += u_l^T gb_l)
//   bwd_out : [gb0]    [P_pad][64] fp16 (39 valid), [gb] 8 x [P_pad][256] fp16 gb_1 .. gb_8, [zb] 8 x [P_pad][256] fp16 zb_0 .. zb_7,
//             all in units of the loss scale S
struct SdfTrainLayout {
    int64_t p_pad;
    size_t tape_tiles_off, tape_act_off, tape_u_off, tape_bytes;
    size_t bwd_gb0_off, bwd_gb_off, bwd_zb_off, bwd_bytes;
    size_t bwd_workspace_bytes;
};
SdfTrainLayout sdf_train_layout(int64_t N, int num_sms);
int sdf_train_forward_tc(const void* packed, const PackedLayout& L, const float* pts, int64_t N, float* sdf, float* grad, float* feat,
                         void* tape, float* scratch, size_t scratch_bytes, int num_sms, cudaStream_t st);
// the same for points / gradients given as strided coordinate arrays (the render pipeline's sample-major SoA buffers)
int sdf_train_forward_tc_strided(const void* packed, const PackedLayout& L, Strided3 pts, int64_t N, float* sdf, float* gx, float* gy,
                                 float* gz, int64_t gstride, float* feat, void* tape, float* scratch, size_t scratch_bytes, int num_sms,
                                 cudaStream_t st, void* feat16 = nullptr, int64_t feat16_ld = 0);
// feat16 != nullptr: the features leave as unscaled fp16 rows [N][feat16_ld] (the feature block of the fused step's reflectance
// operand) and `feat` is not written
// feat16 (fused training step): d_feat as fp16 rows [N][ld] already in units of another power-of-two loss scale, *mul converts to S;
// pts_strided: the points as strided coordinate arrays instead of [N][3]
struct SdfBwdFeat16 { const void* rows; int64_t ld; const float* mul; };
int sdf_train_backward_tc(const void* packed, const PackedLayout& L, const float* pts, int64_t N, const void* tape,
                          const float* d_sdf, const float* d_feat, const float* d_grad, const float* scale, void* bwd_out,
                          float* d_pts, float* scratch, size_t scratch_bytes, int num_sms, cudaStream_t st,
                          const SdfBwdFeat16* feat16 = nullptr, const Strided3* pts_strided = nullptr);

// ---- training: reflectance network forward (with activation dumps) and backward (color_train_tc.inc) ------------------------------
// x16: [P][384] fp16 input in the reference's concatenation order (361 valid columns, rest zero); acts: [4][P][256] fp16 a_1..a_4
// (x TC_ACT_SCALE); y: [P][4] fp32 pre-sigmoid outputs.  Backward: dy [P][3] fp32, *scale = power-of-two loss scale S;
// dz: [4][P][256] fp16 dz_0..dz_3, dy16: [P][8] fp16, dx: [P][384] fp16 -- all in S units.
int color_train_forward_tc(const void* packed, const PackedLayout& L, const void* x16, int64_t P, void* acts, float* y, int num_sms,
                           cudaStream_t st, bool permuted = false);
int color_train_backward_tc(const void* packed, const PackedLayout& L, const float* dy, const float* scale, const void* acts, int64_t P,
                            void* dz, void* dy16, void* dx, int num_sms, cudaStream_t st, bool permuted = false);

}  // namespace nrh
