// Loss and optimiser behind the ray-march path.
//
//   nrh_train_loss : BaseNRHintPipeline.get_train_loss_dict (/root/reference/pipelines/base_pipeline.py:50-69) and the gradient
//                    autograd would produce for it, in two launches (reduce, then finalise + gradients).  The reference
//                    spends ~20 ATen launches on the forward and as many in the backward; all of it is HBM-bound streaming over
//                    `analytic_normals` [R,S,3] and `relax_inside_sphere` [R,S] (8.4 MB at 4096 x 128).
//   nrh_adam_step  : torch.optim.Adam.step (/root/reference/trainer/trainer.py:99,280) on a flat buffer -- one launch instead of
//                    one multi-tensor pass per operation over 46 tensors.
// Both are bandwidth-trivial next to the MLP kernels; the point is launch count and keeping the step free of host syncs.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "nrh_common.cuh"
#include "composite_train_math.cuh"

namespace nrh {
namespace {

constexpr int TL_THREADS = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// stats[4] += sum m, stats[5] += sum m (|n|-1)^2, stats[6] += sum |rgb - gt|, stats[7] += sum (rgb - gt)^2
__global__ void __launch_bounds__(TL_THREADS)
k_loss_reduce(const float* __restrict__ rgb, const float* __restrict__ gt, const float* __restrict__ nrm,
              const float* __restrict__ mask, int64_t n_rgb, int64_t n_pts, float* __restrict__ stats) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    const int64_t stride = (int64_t)gridDim.x * TL_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * TL_THREADS + threadIdx.x; i < n_pts; i += stride) {
        const float m = mask[i];
        const float x = nrm[3 * i], y = nrm[3 * i + 1], z = nrm[3 * i + 2];
        const float e = sqrtf(x * x + y * y + z * z) - 1.0f;
        a[0] += m; a[1] += m * (e * e);
    }
    for (int64_t i = (int64_t)blockIdx.x * TL_THREADS + threadIdx.x; i < n_rgb; i += stride) {
        const float d = rgb[i] - gt[i];
        a[2] += fabsf(d); a[3] += d * d;
    }
    __shared__ float part[4][TL_THREADS / 32];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float v = warp_sum(a[k]);
        if ((threadIdx.x & 31) == 0) part[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        float v = 0.f;
        for (int w = 0; w < TL_THREADS / 32; ++w) v += part[threadIdx.x][w];
        atomicAdd(stats + 4 + threadIdx.x, v);
    }
}

__global__ void __launch_bounds__(TL_THREADS)
k_loss_finish(const float* __restrict__ rgb, const float* __restrict__ gt, const float* __restrict__ nrm,
              const float* __restrict__ mask, int64_t R, int64_t n_rgb, int64_t n_pts, float igr_weight, float grad_scale,
              float* __restrict__ stats, float* __restrict__ d_rgb, float* __restrict__ d_nrm) {
    const float sum_m = stats[4], sum_e = stats[5], sum_abs = stats[6], sum_sq = stats[7];
    const float inv_r = 1.0f / ((float)R + 1e-5f), inv_m = 1.0f / (sum_m + 1e-5f);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const float rgb_loss = sum_abs * inv_r, eik = sum_e * inv_m;
        stats[0] = rgb_loss + eik * igr_weight;
        stats[1] = rgb_loss;
        stats[2] = eik;
        stats[3] = 10.0f * log10f(1.0f / (sum_sq / (float)n_rgb));          // torchmetrics peak_signal_noise_ratio, data_range 1
    }
    const int64_t stride = (int64_t)gridDim.x * TL_THREADS;
    if (d_nrm) {
        const float c = 2.0f * igr_weight * inv_m * grad_scale;
        for (int64_t i = (int64_t)blockIdx.x * TL_THREADS + threadIdx.x; i < n_pts; i += stride) {
            const float x = nrm[3 * i], y = nrm[3 * i + 1], z = nrm[3 * i + 2];
            const float len = sqrtf(x * x + y * y + z * z);
            const float k = len > 0.0f ? c * mask[i] * (len - 1.0f) / len : 0.0f;      // torch: zero (sub)gradient of the norm at 0
            d_nrm[3 * i] = k * x; d_nrm[3 * i + 1] = k * y; d_nrm[3 * i + 2] = k * z;
        }
    }
    if (d_rgb) {
        const float c = inv_r * grad_scale;
        for (int64_t i = (int64_t)blockIdx.x * TL_THREADS + threadIdx.x; i < n_rgb; i += stride) {
            const float d = rgb[i] - gt[i];
            d_rgb[i] = d > 0.0f ? c : (d < 0.0f ? -c : 0.0f);
        }
    }
}

// torch/optim/adam.py::_single_tensor_adam in its operation order (lerp_, mul_/addcmul_, sqrt/div/add_, addcdiv_)
__global__ void __launch_bounds__(TL_THREADS)
k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
       float w1, float beta2, float w2, float bc2_sqrt, float eps, float neg_step_size, float grad_scale) {
    const int64_t stride = (int64_t)gridDim.x * TL_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * TL_THREADS + threadIdx.x; i < n; i += stride) {
        const float gi = g[i] * grad_scale;
        const float mi = __fmaf_rn(w1, gi - m[i], m[i]);
        const float vi = __fmaf_rn(w2 * gi, gi, v[i] * beta2);
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        m[i] = mi; v[i] = vi;
        p[i] = __fmaf_rn(neg_step_size, mi / denom, p[i]);
    }
}

// The same step with its scalars on the device (CUDA-graph capturable: a replay must see a new step count and learning rate):
// k_adam_prepare advances *step and derives the two step-dependent factors in double, k_adam_dev reads them.
__global__ void k_adam_prepare(long long* __restrict__ step, const float* __restrict__ lr, double beta1, double beta2, float* __restrict__ coef) {
    const long long t = *step + 1;
    *step = t;
    const double bc1 = 1.0 - pow(beta1, (double)t), bc2 = 1.0 - pow(beta2, (double)t);
    coef[0] = (float)(-((double)lr[0] / bc1));      // -step_size
    coef[1] = (float)sqrt(bc2);                      // sqrt(bias_correction2)
}
__global__ void __launch_bounds__(TL_THREADS)
k_adam_dev(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
           float w1, float beta2, float w2, float eps, const float* __restrict__ coef, float grad_scale) {
    const float neg_step_size = __ldg(coef), bc2_sqrt = __ldg(coef + 1);
    const int64_t stride = (int64_t)gridDim.x * TL_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * TL_THREADS + threadIdx.x; i < n; i += stride) {
        const float gi = g[i] * grad_scale;
        const float mi = __fmaf_rn(w1, gi - m[i], m[i]);
        const float vi = __fmaf_rn(w2 * gi, gi, v[i] * beta2);
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        m[i] = mi; v[i] = vi;
        p[i] = __fmaf_rn(neg_step_size, mi / denom, p[i]);
    }
}

// column sums of `n_mats` row-major fp16 matrices [rows][width] -> fp32 [n_mats][width] (bias gradients = point-reductions over
// the fp16 adjoint dumps).  One pass at HBM speed: a thread owns 8 consecutive columns (one 16-byte load per row), a CTA walks a
// contiguous slab of rows, partial sums meet in shared memory and leave as one atomicAdd per column and CTA.
constexpr int CS_THREADS = 256;
__global__ void __launch_bounds__(CS_THREADS)
k_colsum_f16(const __half* __restrict__ mats, int64_t rows, int width, int64_t mat_stride, int slabs_per_mat, float scale,
             float* __restrict__ out) {
    const int mat = blockIdx.x / slabs_per_mat, slab = blockIdx.x % slabs_per_mat;
    const int groups = width >> 3;                          // 8-column groups per row
    const int rows_per_pass = CS_THREADS / groups;          // rows covered by the CTA per iteration
    const int g = threadIdx.x % groups, rslot = threadIdx.x / groups;
    const int64_t r0 = rows * slab / slabs_per_mat, r1 = rows * (slab + 1) / slabs_per_mat;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (rslot < rows_per_pass) {
        const __half* base = mats + (size_t)mat * mat_stride + g * 8;
        for (int64_t r = r0 + rslot; r < r1; r += rows_per_pass) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(base + r * width));
            const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); acc[2 * i] += f.x; acc[2 * i + 1] += f.y; }
        }
    }
    __shared__ float part[CS_THREADS * 8];
#pragma unroll
    for (int i = 0; i < 8; ++i) part[(rslot * groups + g) * 8 + i] = acc[i];
    __syncthreads();
    for (int c = threadIdx.x; c < width; c += CS_THREADS) {
        float v = 0.f;
        for (int rs = 0; rs < rows_per_pass; ++rs) v += part[(rs * groups + (c >> 3)) * 8 + (c & 7)];
        atomicAdd(out + (size_t)mat * width + c, v * scale);
    }
}

// ---- differentiable compositing of the primary ray (composite_train_math.cuh), one thread per ray ---------------------------
struct CtArgs {
    const float* sdf; const float* grad; const float* color;      // per point: [N], [N,3], [N,3]; point (r, j) = r * pr + j * pj
    int64_t pr, pj;
    const float* dists; const float* dirs; const float* inv_s; const float* bg;   // [R,S], [R,3], device scalar, nullable [3]
    float cos_anneal; int64_t R; int S;
};
__device__ __forceinline__ CtRay ct_ray(const CtArgs& A, int64_t r) {
    CtRay Y;
    Y.S = A.S;
    Y.sdf = A.sdf + r * A.pr; Y.sdf_st = A.pj;
    Y.g = A.grad + r * A.pr * 3; Y.g_st = A.pj * 3;
    Y.c = A.color + r * A.pr * 3; Y.c_st = A.pj * 3;
    Y.dist = A.dists + r * A.S; Y.dist_st = 1;
    Y.d[0] = A.dirs[r * 3]; Y.d[1] = A.dirs[r * 3 + 1]; Y.d[2] = A.dirs[r * 3 + 2];
    Y.inv_s = *A.inv_s; Y.cos_anneal = A.cos_anneal;
    Y.has_bg = A.bg != nullptr;
    for (int k = 0; k < 3; ++k) Y.bg[k] = A.bg ? A.bg[k] : 0.f;
    return Y;
}
__global__ void __launch_bounds__(128)
k_composite_train_fwd(CtArgs A, float* __restrict__ w, float* __restrict__ rgb) {
    const int64_t r = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (r >= A.R) return;
    const CtRay Y = ct_ray(A, r);
    float c3[3];
    composite_train_forward(Y, w + r * A.S, 1, c3);
    rgb[r * 3] = c3[0]; rgb[r * 3 + 1] = c3[1]; rgb[r * 3 + 2] = c3[2];
}
__global__ void __launch_bounds__(128)
k_composite_train_bwd(CtArgs A, const float* __restrict__ d_rgb, const float* __restrict__ d_w, float* __restrict__ d_sdf,
                      float* __restrict__ d_grad, float* __restrict__ d_color, float* __restrict__ d_dirs, float* __restrict__ d_inv_s) {
    const int64_t r = (int64_t)blockIdx.x * 128 + threadIdx.x;
    float ds = 0.f;
    if (r < A.R) {
        const CtRay Y = ct_ray(A, r);
        float alpha_s[CT_MAX_S], T_s[CT_MAX_S], dd[3];
        const float g3[3] = {d_rgb[r * 3], d_rgb[r * 3 + 1], d_rgb[r * 3 + 2]};
        ds = composite_train_backward(Y, g3, d_w ? d_w + r * A.S : nullptr, 1, d_sdf + r * A.pr, d_grad + r * A.pr * 3, d_color + r * A.pr * 3,
                                      dd, alpha_s, T_s);
        d_dirs[r * 3] = dd[0]; d_dirs[r * 3 + 1] = dd[1]; d_dirs[r * 3 + 2] = dd[2];
    }
    ds = warp_sum(ds);
    if ((threadIdx.x & 31) == 0 && ds != 0.f) atomicAdd(d_inv_s, ds);
}

int grid_for(int64_t n) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (n + TL_THREADS - 1) / TL_THREADS;
    const int64_t cap = (int64_t)sms * 8;                      // whole waves of 8 resident 256-thread CTAs per SM
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace
}  // namespace nrh

using namespace nrh;

extern "C" {

int nrh_train_loss(const float* rgb, const float* rgb_gt, const float* analytic_normals, const float* relax_inside_sphere,
                   int64_t R, int S, float igr_weight, float grad_scale, float* stats, float* d_rgb, float* d_normals,
                   void* stream) {
    if (!rgb || !rgb_gt || !analytic_normals || !relax_inside_sphere || !stats || R < 0 || S < 1) {
        set_error("nrh_train_loss: null argument or bad size"); return NRH_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    NRH_CUDA_CHECK(cudaMemsetAsync(stats, 0, 8 * sizeof(float), st));
    const int64_t n_pts = R * S, n_rgb = R * 3;
    const int grid = grid_for(n_pts > n_rgb ? n_pts : n_rgb);
    k_loss_reduce<<<grid, TL_THREADS, 0, st>>>(rgb, rgb_gt, analytic_normals, relax_inside_sphere, n_rgb, n_pts, stats);
    NRH_LAUNCH_CHECK();
    k_loss_finish<<<grid, TL_THREADS, 0, st>>>(rgb, rgb_gt, analytic_normals, relax_inside_sphere, R, n_rgb, n_pts, igr_weight,
                                                grad_scale, stats, d_rgb, d_normals);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int nrh_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                  double beta2, double eps, int64_t step, float grad_scale, void* stream) {
    if (n == 0) return NRH_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq || n < 0 || step < 1) { set_error("nrh_adam_step: null argument or step < 1"); return NRH_ERR_INVALID; }
    // scalar bookkeeping in double, as torch does in Python (torch/optim/adam.py)
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    const double step_size = lr / bc1;
    k_adam<<<grid_for(n), TL_THREADS, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, (float)(1.0 - beta1), (float)beta2,
                                                                  (float)(1.0 - beta2), (float)sqrt(bc2), (float)eps, (float)(-step_size),
                                                                  grad_scale);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int nrh_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, const float* lr_dev, double beta1,
                      double beta2, double eps, int64_t* step_dev, float* coef_scratch, float grad_scale, void* stream) {
    if (n == 0) return NRH_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq || !lr_dev || !step_dev || !coef_scratch || n < 0) { set_error("nrh_adam_step_dev: null argument"); return NRH_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    k_adam_prepare<<<1, 1, 0, st>>>(reinterpret_cast<long long*>(step_dev), lr_dev, beta1, beta2, coef_scratch);
    NRH_LAUNCH_CHECK();
    k_adam_dev<<<grid_for(n), TL_THREADS, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
                                                   (float)eps, coef_scratch, grad_scale);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int nrh_colsum_f16(const void* mats, int n_mats, int64_t rows, int width, int64_t mat_stride, float scale, float* out, void* stream) {
    if (n_mats == 0 || rows == 0) return NRH_OK;
    if (!mats || !out || n_mats < 0 || rows < 0 || width < 8 || width > 2048 || (width & 7) || (CS_THREADS % (width >> 3)) != 0 ||
        (mat_stride & 7) || (reinterpret_cast<uintptr_t>(mats) & 15)) {
        set_error("nrh_colsum_f16: need 16-byte aligned fp16 matrices, width a multiple of 8 with 256 %% (width / 8) == 0");
        return NRH_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    NRH_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)n_mats * width * sizeof(float), st));
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int slabs = (sms * 8 + n_mats - 1) / n_mats;                       // ~8 CTAs per SM over all matrices
    const int64_t max_slabs = (rows + 63) / 64;
    if (slabs > max_slabs) slabs = (int)max_slabs;
    if (slabs < 1) slabs = 1;
    k_colsum_f16<<<(unsigned)(n_mats * slabs), CS_THREADS, 0, st>>>(static_cast<const __half*>(mats), rows, width, mat_stride, slabs,
                                                                     scale, out);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

static int ct_fill(CtArgs& A, const float* sdf, const float* grad, const float* color, int64_t pr, int64_t pj, const float* dists,
                   const float* dirs, const float* inv_s, float cos_anneal, const float* bg, int64_t R, int S, const char* who) {
    if (!sdf || !grad || !color || !dists || !dirs || !inv_s || R < 0 || S < 1 || S > CT_MAX_S) {
        set_error("%s: null argument or S outside [1, %d]", who, CT_MAX_S); return NRH_ERR_INVALID;
    }
    A = CtArgs{sdf, grad, color, pr, pj, dists, dirs, inv_s, bg, cos_anneal, R, S};
    return NRH_OK;
}

int nrh_composite_train_forward(const float* sdf, const float* grad, const float* color, int64_t point_stride_ray,
                                int64_t point_stride_sample, const float* dists, const float* dirs, const float* inv_s,
                                float cos_anneal, const float* bg_rgb, int64_t R, int S, float* weights, float* rgb, void* stream) {
    if (R == 0) return NRH_OK;
    CtArgs A; int rc = ct_fill(A, sdf, grad, color, point_stride_ray, point_stride_sample, dists, dirs, inv_s, cos_anneal, bg_rgb, R, S,
                               "nrh_composite_train_forward"); if (rc) return rc;
    if (!weights || !rgb) { set_error("nrh_composite_train_forward: null output"); return NRH_ERR_INVALID; }
    k_composite_train_fwd<<<(unsigned)((R + 127) / 128), 128, 0, (cudaStream_t)stream>>>(A, weights, rgb);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int nrh_composite_train_backward(const float* sdf, const float* grad, const float* color, int64_t point_stride_ray,
                                 int64_t point_stride_sample, const float* dists, const float* dirs, const float* inv_s,
                                 float cos_anneal, const float* bg_rgb, int64_t R, int S, const float* d_rgb, const float* d_weights,
                                 float* d_sdf, float* d_grad, float* d_color, float* d_dirs, float* d_inv_s, void* stream) {
    if (R == 0) return NRH_OK;
    CtArgs A; int rc = ct_fill(A, sdf, grad, color, point_stride_ray, point_stride_sample, dists, dirs, inv_s, cos_anneal, bg_rgb, R, S,
                               "nrh_composite_train_backward"); if (rc) return rc;
    if (!d_rgb || !d_sdf || !d_grad || !d_color || !d_dirs || !d_inv_s) { set_error("nrh_composite_train_backward: null argument"); return NRH_ERR_INVALID; }
    k_composite_train_bwd<<<(unsigned)((R + 127) / 128), 128, 0, (cudaStream_t)stream>>>(A, d_rgb, d_weights, d_sdf, d_grad, d_color, d_dirs, d_inv_s);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

}  // extern "C"
