// Glue kernels of the fused training step (train_step.cu) -- internal interface used by api.cu.
#pragma once
#include <cuda_fp16.h>
#include "nrh_common.cuh"

namespace nrh {

// weight norm of all weight-normed layers in one launch (forward: v, g -> w; backward: v, g, dw -> dg, dv)
struct WnJob {
    const float* v; const float* g; float* w;             // forward
    const float* dw; int dw_ld; float* dg; float* dv;     // backward
    int rows, cols;
    int perm, perm_shadow, perm_spec0;                    // dw columns in the fused step's reflectance operand order (layer 0 only)
};
constexpr int WN_MAX_JOBS = 16;
struct WnTable { WnJob j[WN_MAX_JOBS]; int n; };
int launch_weight_norm(const WnTable& T, bool backward, cudaStream_t st);

struct TrainAssembleArgs {
    int64_t N, R;
    const float* gx; const float* gy; const float* gz;      // grad sdf, SoA [N]
    const float* px; const float* py; const float* pz;      // fine points, SoA [N]
    const float* feat;                                      // [N][256], or nullptr when the feature block of x16 is already in place
    const float* rayfeat;                                   // [99][R] per-ray encodings (k_shade_prep)
    int normalized;
    __half* x16;                                            // [N][384]
    float* grad_aos;                                        // [N][3]
};
struct TrainScatterArgs {
    int64_t R; int S;
    const __half* dx16;                                     // [N][384], S_c units
    const float* scale_c;                                   // [S_c, 1 / S_c, ...]
    const float* gx; const float* gy; const float* gz;
    const float* d_normals;                                 // [R][S][3] adjoint of the analytic_normals output (nullable)
    const float* d_nnormals;                                // [R][S][3] adjoint of the normalized_analytic_normals output (nullable)
    const float* dirs; const float* pl;                     // [R][3]
    int normalized;
    float* d_grad;                                          // [N][3] in: compositor term, out: total adjoint of grad sdf
    float* d_pts;                                           // [N][3] out: position adjoint through the reflectance input
    float* d_dirs_pe;                                       // [R][3]
    float* d_pl;                                            // [R][3] nullable
};

// z: final sample positions, RAY-major [R][S] (NrhOutputs.z_vals) -> dists [R][S] (ray-major), mid_z [S][R] (sample-major)
int launch_train_dists(const float* z, int64_t R, int S, float sample_dist, float* dists, float* mid_z, cudaStream_t st);

// out[i] = src[i] * (*dev_scale or 1) * mul + add[i] * add_mul for up to 24 short vectors in one launch (bias gradients of a step)
struct VecJob { const float* src; const float* dev_scale; float mul; const float* add; float add_mul; float* dst; int n; };
constexpr int VEC_MAX_JOBS = 24;
struct VecTable { VecJob j[VEC_MAX_JOBS]; int n; };
int launch_vec_jobs(const VecTable& T, cudaStream_t st);
// column sums of the window [col0, col0 + 256) of an fp16 matrix [rows][ld] -> out[256] (ACCUMULATED with atomics: zero it first)
int launch_colsum256_f16(const void* a, int64_t rows, int64_t ld, int col0, float* out, cudaStream_t st);
int launch_train_assemble(const TrainAssembleArgs& A, cudaStream_t st);
int launch_color_sigmoid(const float* y, int64_t N, float* color, cudaStream_t st);
int launch_absmax_f32(const float* a, int64_t n, float mul, unsigned int* bits, cudaStream_t st);
int launch_absmax_f16(const void* a, int64_t rows, int ld, int col0, int ncols, const float* dev_mul, unsigned int* bits, cudaStream_t st);
int launch_pow2_scale(const unsigned int* bits, float target, const float* prev, float* scale, cudaStream_t st);
int launch_train_scatter(const TrainScatterArgs& A, cudaStream_t st);
int launch_train_ray_reduce(const float* d_pts_a, const float* d_pts_b, const float* mid_z, int64_t R, int S, const float* d_dirs_c,
                            const float* d_dirs_pe, float* d_o, float* d_d, cudaStream_t st);
int launch_pe_dump(const float* px, const float* py, const float* pz, int64_t N, int64_t P_pad, void* e, cudaStream_t st);
int launch_ds16(const float* d_sdf, int64_t N, int64_t P_pad, const float* scale, void* out, cudaStream_t st);
int launch_sum_f32(const float* a, int64_t n, float mul, float* out, cudaStream_t st);
int launch_colsum3(const float* a, int64_t n, float* out, cudaStream_t st);
int launch_sigmoid_bwd(const float* d_color, const float* color, int64_t n3, float* dy, cudaStream_t st);
int launch_variance_grad(const float* d_inv_s, const float* variance, float* out, cudaStream_t st);

}  // namespace nrh
