// Glue kernels of the fused training step (nrh_render_train_forward / nrh_render_backward, orchestrated in api.cu): everything
// between the tensor-core kernels that the reference leaves to ~300 ATen launches and an autograd graph per step
// (models/neus_hint_model.py:504-651 under autograd; pipelines/base_pipeline.py:50-69; trainer/trainer.py:269-283).
//
//   weight norm        W = g v / ||v||_row for all 15 weight-normed layers in ONE launch, and its backward (dW -> dg, dv) in one
//   k_train_dists      section lengths / mid-point depths of the final samples (models/neus_hint_model.py:491-494)
//   k_train_assemble   reflectance input rows x16 = [feature 256 | pts 3 | PE(view) 27 | normal 3 | PE(light) 27 | PE(vis) 9 | PE(spec) 36]
//                      (fields/reflectance_network.py:68-82 with the feature block moved to the front: 16-byte aligned for TMA), the
//                      [N,3] gradient / normal arrays of the compositor
//   k_color_sigmoid    colour = sigmoid(y)
//   k_absmax / k_pow2_scale   power-of-two loss scales chosen on the device (no host sync)
//   k_train_scatter    backward of the assembly: per-ray sums of the view / light encodings' adjoints and their chain rule to the
//                      ray direction / light position, normalisation backward, total adjoint of grad sdf, position adjoint
//   k_train_ray_reduce d origins / d directions from the per-point position adjoints
//   k_pe_dump, k_ds16  small fp16 operands of the weight-gradient reductions
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "nrh_common.cuh"
#include "train_step.cuh"

namespace nrh {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// ---- weight norm ---------------------------------------------------------------------------------------------------------------
// Bit-compatible with torch._weight_norm (dim = 0) on CUDA, which is what the reference's weight-normed layers evaluate
// (fields/sdf_field.py:97-98, torch.nn.utils.weight_norm): one 256-thread block per row, per-thread partial sums with stride 256
// (fused multiply-add), shared-memory tree down to 64 entries, then x[tid] + x[tid + 32] and a shuffle-down tree;
// w = (g * v) * (1 / sqrt(sum)).  Checked bitwise against torch on B200 (tests/test_fused_step.py).
__global__ void __launch_bounds__(256)
k_weight_norm_fwd(const __grid_constant__ WnTable T) {
    int row = blockIdx.x, j = 0;
    while (row >= T.j[j].rows) { row -= T.j[j].rows; ++j; }
    const WnJob& J = T.j[j];
    const float* v = J.v + (size_t)row * J.cols;
    const int tid = threadIdx.x;
    __shared__ float x[256];
    float ss = 0.f;
    for (int c = tid; c < J.cols; c += 256) { const float val = v[c]; ss = fmaf(val, val, ss); }
    x[tid] = ss;
    __syncthreads();
    for (int i = 128; i >= 64; i >>= 1) {
        if (tid < i) x[tid] = x[tid] + x[tid + i];
        __syncthreads();
    }
    if (tid < 32) {
        float fin = x[tid] + x[tid + 32];
#pragma unroll
        for (int i = 16; i >= 1; i >>= 1) fin = fin + __shfl_down_sync(0xffffffffu, fin, i);
        if (tid == 0) x[0] = fin;
    }
    __syncthreads();
    const float norm = sqrtf(x[0]);
    const float g = J.g[row], rnorm = 1.0f / norm;
    float* w = J.w + (size_t)row * J.cols;
    for (int c = tid; c < J.cols; c += 256) w[c] = g * v[c] * rnorm;
}

// natural input column of the reflectance network's first layer -> column of the fused step's operand order
// [feature 256 | pts 3, PE(view) 27, normal 3, PE(light) 27 | PE(vis) 9 at 60 | PE(spec) at 69] (see k_train_assemble)
__device__ __forceinline__ int refl_perm(int c, int shadow, int spec0) {
    if (c < 60) return 256 + c;
    if (c < 316) return c - 60;
    if (shadow && c < 325) return 256 + 60 + (c - 316);
    return 256 + 69 + (c - spec0);
}

// dW (effective-weight gradient, row stride dw_ld, optionally in the permuted column order above) -> dg [rows], dv [rows][cols]
//   dg = sum_c dW v / ||v|| ;  dv = g / ||v|| * (dW - v * (sum_c dW v) / ||v||^2)        (torch's weight_norm backward, dim = 0)
__global__ void __launch_bounds__(128)
k_weight_norm_bwd(const __grid_constant__ WnTable T) {
    int row = blockIdx.x, j = 0;
    while (row >= T.j[j].rows) { row -= T.j[j].rows; ++j; }
    const WnJob& J = T.j[j];
    const float* v = J.v + (size_t)row * J.cols;
    const float* dw = J.dw + (size_t)row * J.dw_ld;
    float ss = 0.f, dot = 0.f;
    for (int c = threadIdx.x; c < J.cols; c += 128) {
        const float vv = v[c], d = dw[J.perm ? refl_perm(c, J.perm_shadow, J.perm_spec0) : c];
        ss += vv * vv; dot += d * vv;
    }
    __shared__ float part[8];
    ss = warp_sum(ss); dot = warp_sum(dot);
    if ((threadIdx.x & 31) == 0) { part[threadIdx.x >> 5] = ss; part[4 + (threadIdx.x >> 5)] = dot; }
    __syncthreads();
    ss = part[0] + part[1] + part[2] + part[3];
    dot = part[4] + part[5] + part[6] + part[7];
    const float norm = sqrtf(ss), gk = J.g[row] / norm, k2 = dot / ss;
    if (threadIdx.x == 0) J.dg[row] = dot / norm;
    float* dv = J.dv + (size_t)row * J.cols;
    for (int c = threadIdx.x; c < J.cols; c += 128) dv[c] = gk * (dw[J.perm ? refl_perm(c, J.perm_shadow, J.perm_spec0) : c] - v[c] * k2);
}

// ---- forward glue --------------------------------------------------------------------------------------------------------------
// z: final sample positions, RAY-major [R][S] -> dists [R][S] (ray-major, the compositor's layout), mid_z [S][R] (sample-major)
// (models/neus_hint_model.py:491-494: dists = z[1:] - z[:-1] with sample_dist appended, mid_z = z + dists / 2)
__global__ void k_train_dists(const float* __restrict__ z, int64_t R, int S, float sample_dist, float* __restrict__ dists, float* __restrict__ mid_z) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * S) return;
    const int64_t r = i / S;
    const int j = (int)(i - r * S);
    const float zc = z[i];
    const float d = j + 1 < S ? z[i + 1] - zc : sample_dist;
    dists[i] = d;
    mid_z[(int64_t)j * R + r] = zc + d * 0.5f;
}

// one warp per point (sample-major p = j R + r): lane c handles columns c, c + 32, ... of the 384-wide row
__global__ void __launch_bounds__(256)
k_train_assemble(TrainAssembleArgs A) {
    const int64_t p = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= A.N) return;
    const int64_t r = p % A.R;
    const float gx = A.gx[p], gy = A.gy[p], gz = A.gz[p];
    const float len = sqrtf(gx * gx + gy * gy + gz * gz);
    const float inv = 1.0f / fmaxf(len, 1e-12f);                       // F.normalize(eps = 1e-12)
    const float nx = gx * inv, ny = gy * inv, nz = gz * inv;
    if (lane == 0) {
        A.grad_aos[p * 3] = gx; A.grad_aos[p * 3 + 1] = gy; A.grad_aos[p * 3 + 2] = gz;
    }
    __half* row = A.x16 + p * 384;
    if (A.feat) {                                         // nullptr: the SDF kernel already wrote the feature block (fp16 rows of x16)
        const float* frow = A.feat + p * 256;
#pragma unroll
        for (int k = 0; k < 8; ++k) row[k * 32 + lane] = __float2half_rn(frow[k * 32 + lane]);
    }
    // auxiliary block, columns 256 .. 383: [pts 3 | PE(view) 27 | normal 3 | PE(light) 27 | PE(vis) 9 | PE(spec) 36 | 0 ...] -- the fixed
    // positions of the inference kernel's layer-0 operand (csrc/mlp_tc.cu::tc_pack); rayfeat rows of a disabled hint are zero
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = k * 32 + lane;
        float v = 0.f;
        if (c < 3) v = (c == 0 ? A.px : (c == 1 ? A.py : A.pz))[p];
        else if (c < 30) v = A.rayfeat[(int64_t)(c - 3) * A.R + r];                              // PE(view)
        else if (c < 33) v = A.normalized ? (c == 30 ? nx : (c == 31 ? ny : nz)) : (c == 30 ? gx : (c == 31 ? gy : gz));
        else if (c < 105) v = A.rayfeat[(int64_t)(27 + c - 33) * A.R + r];                         // PE(light) | PE(vis) | PE(spec)
        row[256 + c] = __float2half_rn(v);
    }
}

__global__ void k_color_sigmoid(const float* __restrict__ y, int64_t N, float* __restrict__ color) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const float4 t = *reinterpret_cast<const float4*>(y + p * 4);
    color[p * 3] = 1.0f / (1.0f + expf(-t.x));
    color[p * 3 + 1] = 1.0f / (1.0f + expf(-t.y));
    color[p * 3 + 2] = 1.0f / (1.0f + expf(-t.z));
}

// ---- loss scales ---------------------------------------------------------------------------------------------------------------
// *out_bits = max(*out_bits, bits(max |a[i] * mul|)) -- non-negative floats order like their bit patterns
__global__ void k_absmax_f32(const float* __restrict__ a, int64_t n, float mul, unsigned int* __restrict__ out_bits) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(a[i]));
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out_bits, __float_as_uint(m * mul));
}
// the same over a column window [col0, col0 + ncols) of an fp16 matrix [rows][ld], values multiplied by *dev_mul; col0, ncols and ld
// are multiples of 8: one 16-byte load per thread and step
__global__ void __launch_bounds__(256)
k_absmax_f16(const __half* __restrict__ a, int64_t rows, int ld, int col0, int ncols, const float* __restrict__ dev_mul,
             unsigned int* __restrict__ out_bits) {
    const int groups = ncols >> 3;
    const int64_t n = rows * groups;
    __half2 m2 = __float2half2_rn(0.f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / groups;
        const int g = (int)(i - r * groups);
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(a + r * ld + col0 + g * 8));
        const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
        for (int k = 0; k < 4; ++k) m2 = __hmax2(m2, __habs2(h[k]));
    }
    float m = fmaxf(__low2float(m2), __high2float(m2));
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out_bits, __float_as_uint(m * __ldg(dev_mul)));
}
// scale[0] = 2^floor(log2(target / amax)) clamped to [2^-40, 2^40], scale[1] = 1 / scale[0], scale[2] = scale[0] / *prev (if prev)
__global__ void k_pow2_scale(const unsigned int* __restrict__ amax_bits, float target, const float* __restrict__ prev, float* __restrict__ scale) {
    const float amax = fmaxf(__uint_as_float(*amax_bits), 1e-30f);
    float s = exp2f(floorf(log2f(target / amax)));
    s = fminf(fmaxf(s, 9.094947e-13f), 1.0995116e12f);
    scale[0] = s; scale[1] = 1.0f / s; scale[2] = prev ? s / prev[0] : 1.0f;
}

// ---- backward glue -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pe_backward(float x, const float* d, int F, int stride_d) {
    // adjoint of x through [x, sin(x f_k), sin(x f_k + pi/2)]: d[0] = d x, d[1 + k] = d sin_k, d[1 + F + k] = d cos_k (caller gathers)
    float acc = d[0], f = 1.0f;
    for (int k = 0; k < F; ++k) {
        const float s = x * f;
        acc += f * (cosf(s) * d[(1 + k) * stride_d] + cosf(s + 1.57079637050628662109375f) * d[(1 + F + k) * stride_d]);
        f *= 2.0f;
    }
    return acc;
}

// One block per ray, one thread per sample (S <= 128).  dx16 rows are in units of the loss scale S_c (inv_sc = 1 / S_c).
__global__ void __launch_bounds__(128)
k_train_scatter(TrainScatterArgs A) {
    const int64_t r = blockIdx.x;
    const int j = threadIdx.x;
    const bool on = j < A.S;
    const int64_t p = (int64_t)j * A.R + r;
    const float inv_sc = A.scale_c[1];
    __shared__ float red[54][4];
    float pe_adj[54];                                  // PE(view) 27 adjoints, PE(light) 27 adjoints of this sample
    float dn[3] = {0.f, 0.f, 0.f}, dp[3] = {0.f, 0.f, 0.f};
    if (on) {
        const __half* row = A.dx16 + p * 384 + 256;
#pragma unroll
        for (int c = 0; c < 3; ++c) dp[c] = __half2float(row[c]) * inv_sc;
#pragma unroll
        for (int c = 0; c < 27; ++c) pe_adj[c] = __half2float(row[3 + c]);
#pragma unroll
        for (int c = 0; c < 3; ++c) dn[c] = __half2float(row[30 + c]) * inv_sc;
#pragma unroll
        for (int c = 0; c < 27; ++c) pe_adj[27 + c] = __half2float(row[33 + c]);
    } else {
#pragma unroll
        for (int c = 0; c < 54; ++c) pe_adj[c] = 0.f;
    }
    // per-ray sums of the encodings' adjoints (the encodings are per-ray quantities broadcast over the samples)
#pragma unroll
    for (int c = 0; c < 54; ++c) {
        const float s = warp_sum(pe_adj[c]);
        if ((j & 31) == 0) red[c][j >> 5] = s;
    }
    if (on) {
        // total adjoint of grad sdf at this point: eikonal term (d analytic_normals) + compositor (true_cos) + reflectance normal input
        const float gx = A.gx[p], gy = A.gy[p], gz = A.gz[p];
        float dg[3] = {A.d_grad[p * 3], A.d_grad[p * 3 + 1], A.d_grad[p * 3 + 2]};
        if (A.d_normals) {
            const float* dnr = A.d_normals + (r * A.S + j) * 3;
            dg[0] += dnr[0]; dg[1] += dnr[1]; dg[2] += dnr[2];
        }
        if (A.normalized) {
            const float len = sqrtf(gx * gx + gy * gy + gz * gz);
            float dnn[3] = {dn[0], dn[1], dn[2]};
            if (A.d_nnormals) {                                           // adjoint of the normalized_analytic_normals OUTPUT, if a loss uses it
                const float* q = A.d_nnormals + (r * A.S + j) * 3;
                dnn[0] += q[0]; dnn[1] += q[1]; dnn[2] += q[2];
            }
            if (len > 1e-12f) {                                           // n = g / |g|: dg += (dn - n (n . dn)) / |g|
                const float inv = 1.0f / len, nx = gx * inv, ny = gy * inv, nz = gz * inv;
                const float nd = nx * dnn[0] + ny * dnn[1] + nz * dnn[2];
                dg[0] += (dnn[0] - nx * nd) * inv; dg[1] += (dnn[1] - ny * nd) * inv; dg[2] += (dnn[2] - nz * nd) * inv;
            } else {                                                      // clamped denominator: n = g / eps
                dg[0] += dnn[0] * 1e12f; dg[1] += dnn[1] * 1e12f; dg[2] += dnn[2] * 1e12f;
            }
        } else {
            dg[0] += dn[0]; dg[1] += dn[1]; dg[2] += dn[2];
        }
        A.d_grad[p * 3] = dg[0]; A.d_grad[p * 3 + 1] = dg[1]; A.d_grad[p * 3 + 2] = dg[2];
        A.d_pts[p * 3] = dp[0]; A.d_pts[p * 3 + 1] = dp[1]; A.d_pts[p * 3 + 2] = dp[2];
    }
    __syncthreads();
    if (j < 6) {
        // chain rule of the two encodings to the ray direction (components 0..2) and the light position (3..5)
        const int which = j / 3, d = j % 3;
        float adj[9];                                                    // [d x, d sin_0..3, d cos_0..3] of this component
        const int base = which * 27;
        auto tot = [&](int c) { return (red[base + c][0] + red[base + c][1] + red[base + c][2] + red[base + c][3]) * inv_sc; };
        adj[0] = tot(d);
#pragma unroll
        for (int k = 0; k < 4; ++k) { adj[1 + k] = tot(3 + d * 4 + k); adj[5 + k] = tot(15 + d * 4 + k); }
        const float x = which == 0 ? A.dirs[r * 3 + d] : A.pl[r * 3 + d];
        const float g = pe_backward(x, adj, 4, 1);
        if (which == 0) A.d_dirs_pe[r * 3 + d] = g;
        else if (A.d_pl) A.d_pl[r * 3 + d] = g;
    }
}

// d origins = sum_j d pts, d directions = sum_j mid_z d pts + compositor term + encoding term; d_pts_a / d_pts_b: the two position
// adjoints (reflectance input, SDF network), [N][3] each, sample-major points
__global__ void __launch_bounds__(128)
k_train_ray_reduce(const float* __restrict__ d_pts_a, const float* __restrict__ d_pts_b, const float* __restrict__ mid_z, int64_t R, int S,
                   const float* __restrict__ d_dirs_c, const float* __restrict__ d_dirs_pe, float* __restrict__ d_o, float* __restrict__ d_d) {
    const int64_t r = blockIdx.x;
    const int j = threadIdx.x;
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (j < S) {
        const int64_t p = (int64_t)j * R + r;
        const float m = mid_z[p];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float d = d_pts_a[p * 3 + c] + d_pts_b[p * 3 + c];
            v[c] = d; v[3 + c] = d * m;
        }
    }
    __shared__ float red[6][4];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const float s = warp_sum(v[c]);
        if ((j & 31) == 0) red[c][j >> 5] = s;
    }
    __syncthreads();
    if (j < 6) {
        const float s = red[j][0] + red[j][1] + red[j][2] + red[j][3];
        if (j < 3) { if (d_o) d_o[r * 3 + j] = s; }
        else if (d_d) d_d[r * 3 + j - 3] = s + d_dirs_c[r * 3 + j - 3] + d_dirs_pe[r * 3 + j - 3];
    }
}

// e [P_pad][64] fp16 = PE(3 x) (39 valid columns) of the fine points: the a_0 operand of dW_0
__global__ void k_pe_dump(const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ pz, int64_t N, int64_t P_pad,
                          __half* __restrict__ e) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P_pad) return;
    __half* row = e + p * 64;
    float v[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) v[c] = 0.f;
    if (p < N) {
        const float x[3] = {px[p] * SDF_SCALE, py[p] * SDF_SCALE, pz[p] * SDF_SCALE};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            v[d] = x[d];
            float f = 1.0f;
#pragma unroll
            for (int k = 0; k < SDF_FREQ; ++k) {
                v[3 + d * SDF_FREQ + k] = sinf(x[d] * f);
                v[3 + 3 * SDF_FREQ + d * SDF_FREQ + k] = sinf(x[d] * f + 1.57079637050628662109375f);
                f *= 2.0f;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 64; c += 8) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { const __half2 h = __floats2half2_rn(v[c + 2 * i], v[c + 2 * i + 1]); w[i] = *reinterpret_cast<const uint32_t*>(&h); }
        *reinterpret_cast<uint4*>(row + c) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ds16 [P_pad][8] fp16: column 0 = d_sdf * S
__global__ void k_ds16(const float* __restrict__ d_sdf, int64_t N, int64_t P_pad, const float* __restrict__ scale, __half* __restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P_pad) return;
    const __half h = __float2half_rn(p < N ? d_sdf[p] * scale[0] : 0.f);
    const uint32_t w0 = (uint32_t)__half_as_ushort(h);
    *reinterpret_cast<uint4*>(out + p * 8) = make_uint4(w0, 0u, 0u, 0u);
}

// short-vector jobs: dst[i] = src[i] * (*dev_scale or 1) * mul + add[i] * add_mul      (one block per job and 256 elements)
__global__ void __launch_bounds__(256)
k_vec_jobs(const __grid_constant__ VecTable T) {
    const VecJob& J = T.j[blockIdx.y];
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= J.n) return;
    float v = J.src[i] * (J.dev_scale ? __ldg(J.dev_scale) : 1.0f) * J.mul;
    if (J.add) v += J.add[i] * J.add_mul;
    J.dst[i] = v;
}

// column sums of a 256-column window of an fp16 matrix: thread = (row slot, 8-column group)
__global__ void __launch_bounds__(256)
k_colsum256_f16(const __half* __restrict__ a, int64_t rows, int64_t ld, int col0, float* __restrict__ out) {
    const int g = threadIdx.x & 31, rslot = threadIdx.x >> 5;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int64_t r = (int64_t)blockIdx.x * 8 + rslot; r < rows; r += (int64_t)gridDim.x * 8) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(a + r * ld + col0 + g * 8));
        const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); acc[2 * i] += f.x; acc[2 * i + 1] += f.y; }
    }
    __shared__ float part[8][256];
#pragma unroll
    for (int i = 0; i < 8; ++i) part[rslot][g * 8 + i] = acc[i];
    __syncthreads();
    float v = 0.f;
#pragma unroll
    for (int rs = 0; rs < 8; ++rs) v += part[rs][threadIdx.x];
    atomicAdd(out + threadIdx.x, v);
}

// column sums of an fp32 matrix [rows][ld] window -> out[c] (+=): small widths (d_feat bias gradient is taken from the fp16 dump instead)
__global__ void k_sum_f32(const float* __restrict__ a, int64_t n, float mul, float* __restrict__ out) {
    float s = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += a[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s * mul);
}
// column sums of dy [N][3]
__global__ void k_colsum3(const float* __restrict__ a, int64_t n, float* __restrict__ out) {
    float s[3] = {0.f, 0.f, 0.f};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        s[0] += a[i * 3]; s[1] += a[i * 3 + 1]; s[2] += a[i * 3 + 2];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float t = warp_sum(s[c]);
        if ((threadIdx.x & 31) == 0) atomicAdd(out + c, t);
    }
}
// dy = d_color * c (1 - c)
__global__ void k_sigmoid_bwd(const float* __restrict__ d_color, const float* __restrict__ color, int64_t n3, float* __restrict__ dy) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3) { const float c = color[i]; dy[i] = d_color[i] * c * (1.0f - c); }
}
// d variance = d inv_s * 10 inv_s when exp(10 v) lies inside the clip range [1e-6, 1e6] (models/neus_hint_model.py:337)
__global__ void k_variance_grad(const float* __restrict__ d_inv_s, const float* __restrict__ variance, float* __restrict__ out) {
    const float s = expf(variance[0] * 10.0f);
    out[0] = (s > 1e-6f && s < 1e6f) ? d_inv_s[0] * 10.0f * s : 0.f;
}

inline int blocks_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

}  // namespace

int launch_weight_norm(const WnTable& T, bool backward, cudaStream_t st) {
    int rows = 0;
    for (int i = 0; i < T.n; ++i) rows += T.j[i].rows;
    if (rows == 0) return NRH_OK;
    if (backward) k_weight_norm_bwd<<<rows, 128, 0, st>>>(T); else k_weight_norm_fwd<<<rows, 256, 0, st>>>(T);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}
int launch_train_dists(const float* z, int64_t R, int S, float sample_dist, float* dists, float* mid_z, cudaStream_t st) {
    k_train_dists<<<blocks_for(R * S, 256), 256, 0, st>>>(z, R, S, sample_dist, dists, mid_z);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}
int launch_train_assemble(const TrainAssembleArgs& A, cudaStream_t st) {
    k_train_assemble<<<blocks_for(A.N * 32, 256), 256, 0, st>>>(A);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}
int launch_color_sigmoid(const float* y, int64_t N, float* color, cudaStream_t st) {
    k_color_sigmoid<<<blocks_for(N, 256), 256, 0, st>>>(y, N, color);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}
int launch_absmax_f32(const float* a, int64_t n, float mul, unsigned int* bits, cudaStream_t st) {
    if (n <= 0) return NRH_OK;
    const int64_t want = (n + 255) / 256;
    k_absmax_f32<<<(int)(want < 1184 ? want : 1184), 256, 0, st>>>(a, n, mul, bits);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}
int launch_absmax_f16(const void* a, int64_t rows, int ld, int col0, int ncols, const float* dev_mul, unsigned int* bits, cudaStream_t st) {
    if (rows <= 0) return NRH_OK;
    k_absmax_f16<<<1184, 256, 0, st>>>(reinterpret_cast<const __half*>(a), rows, ld, col0, ncols, dev_mul, bits);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}
int launch_pow2_scale(const unsigned int* bits, float target, const float* prev, float* scale, cudaStream_t st) {
    k_pow2_scale<<<1, 1, 0, st>>>(bits, target, prev, scale);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}
int launch_train_scatter(const TrainScatterArgs& A, cudaStream_t st) {
    if (A.S > 128) { set_error("fused training step: more than 128 samples per ray"); return NRH_ERR_UNSUPPORTED; }
    k_train_scatter<<<(unsigned)A.R, 128, 0, st>>>(A);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}
int launch_train_ray_reduce(const float* d_pts_a, const float* d_pts_b, const float* mid_z, int64_t R, int S, const float* d_dirs_c,
                            const float* d_dirs_pe, float* d_o, float* d_d, cudaStream_t st) {
    k_train_ray_reduce<<<(unsigned)R, 128, 0, st>>>(d_pts_a, d_pts_b, mid_z, R, S, d_dirs_c, d_dirs_pe, d_o, d_d);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}
int launch_pe_dump(const float* px, const float* py, const float* pz, int64_t N, int64_t P_pad, void* e, cudaStream_t st) {
    k_pe_dump<<<blocks_for(P_pad, 128), 128, 0, st>>>(px, py, pz, N, P_pad, reinterpret_cast<__half*>(e));
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}
int launch_ds16(const float* d_sdf, int64_t N, int64_t P_pad, const float* scale, void* out, cudaStream_t st) {
    k_ds16<<<blocks_for(P_pad, 256), 256, 0, st>>>(d_sdf, N, P_pad, scale, reinterpret_cast<__half*>(out));
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}
int launch_vec_jobs(const VecTable& T, cudaStream_t st) {
    if (T.n <= 0) return NRH_OK;
    int nmax = 1;
    for (int i = 0; i < T.n; ++i) if (T.j[i].n > nmax) nmax = T.j[i].n;
    k_vec_jobs<<<dim3((nmax + 255) / 256, T.n), 256, 0, st>>>(T);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}
int launch_colsum256_f16(const void* a, int64_t rows, int64_t ld, int col0, float* out, cudaStream_t st) {
    if (rows <= 0) return NRH_OK;
    const int64_t want = (rows + 63) / 64;
    k_colsum256_f16<<<(int)(want < 1184 ? want : 1184), 256, 0, st>>>(reinterpret_cast<const __half*>(a), rows, ld, col0, out);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}
int launch_sum_f32(const float* a, int64_t n, float mul, float* out, cudaStream_t st) {
    if (n <= 0) return NRH_OK;
    const int64_t want = (n + 255) / 256;
    k_sum_f32<<<(int)(want < 592 ? want : 592), 256, 0, st>>>(a, n, mul, out);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}
int launch_colsum3(const float* a, int64_t n, float* out, cudaStream_t st) {
    if (n <= 0) return NRH_OK;
    const int64_t want = (n + 255) / 256;
    k_colsum3<<<(int)(want < 592 ? want : 592), 256, 0, st>>>(a, n, out);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}
int launch_sigmoid_bwd(const float* d_color, const float* color, int64_t n3, float* dy, cudaStream_t st) {
    k_sigmoid_bwd<<<blocks_for(n3, 256), 256, 0, st>>>(d_color, color, n3, dy);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}
int launch_variance_grad(const float* d_inv_s, const float* variance, float* out, cudaStream_t st) {
    k_variance_grad<<<1, 1, 0, st>>>(d_inv_s, variance, out);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

}  // namespace nrh
