// fp32 FFMA fused MLP engine (NRH_MLP_FP32_SIMT): the always-correct path that the tcgen05
// engine is validated against.  One CTA owns a tile of 64 points and keeps every activation
// on chip: activations live in shared memory k-major ([feature][point]), weights stream from
// L2 through a cp.async double buffer, each thread owns an 8-point x 8-output register tile.
//
// Reference semantics:
//   SDFNetwork.forward / .gradient   /root/reference/fields/sdf_field.py:106-148
//   ReflectanceNetwork.forward       /root/reference/fields/reflectance_network.py:68-96
//   NeRFEncoding.forward             /root/reference/fields/encodings.py:168-176
// The input gradient is an explicit reverse sweep (what autograd.grad does in the reference);
// softplus' of every layer is parked in an L2-resident per-CTA scratch between the sweeps.
#include <cuda_runtime.h>
#include <math.h>
#include "nrh_common.cuh"
#include "ray_math.cuh"

namespace nrh {
namespace {

constexpr int TM = 64;            // points per tile
constexpr int TMP = 68;           // smem row pitch (floats): 16B-aligned rows, spreads banks
constexpr int NT = 256;           // threads per CTA
constexpr int KC = 8;             // weight rows per cp.async stage
constexpr float INV_SQRT2_DIV = 1.41421354f;   // float(np.sqrt(2)); reference divides by it

struct Acc { float v[8][8]; };    // [point i][output o]

__device__ __forceinline__ int out_index(int lane, int o) { return (o < 4) ? lane * 4 + o : 128 + lane * 4 + (o - 4); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void acc_set_bias(Acc& acc, const float* __restrict__ bias, int lane) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + lane * 4));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + 128 + lane * 4));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc.v[i][0] = b0.x; acc.v[i][1] = b0.y; acc.v[i][2] = b0.z; acc.v[i][3] = b0.w;
        acc.v[i][4] = b1.x; acc.v[i][5] = b1.y; acc.v[i][6] = b1.z; acc.v[i][7] = b1.w;
    }
}
__device__ __forceinline__ void acc_zero(Acc& acc) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int o = 0; o < 8; ++o) acc.v[i][o] = 0.0f;
}

// acc[pt][n] += sum_k A[k][pt] * B[k][n], k < K.  A: smem, pitch TMP.  B: global [Kpad][256],
// Kpad = ceil(K/KC)*KC rows readable (zero padded); A rows up to Kpad must hold finite values.
// Ends with a __syncthreads(): afterwards A may be overwritten.
__device__ __forceinline__ void gemm_accumulate(Acc& acc, const float* __restrict__ A, int K,
                                                const float* __restrict__ B, float* wstage) {
    const int tid = threadIdx.x, lane = tid & 31, tp = tid >> 5;
    const int nchunks = (K + KC - 1) / KC;
    // each stage = KC*256 floats = 512 float4; 2 per thread
    auto prefetch = [&](int c) {
        float* dst = wstage + (c & 1) * (KC * 256);
        const float* src = B + (size_t)c * (KC * 256);
        cp_async16(dst + tid * 4, src + tid * 4);
        cp_async16(dst + 1024 + tid * 4, src + 1024 + tid * 4);
        cp_async_commit();
    };
    prefetch(0);
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) { prefetch(c + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const float* ws = wstage + (c & 1) * (KC * 256);
        const float* a_base = A + (size_t)(c * KC) * TMP + tp * 8;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(a_base + kk * TMP);
            const float4 a1 = *reinterpret_cast<const float4*>(a_base + kk * TMP + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(ws + kk * 256 + lane * 4);
            const float4 b1 = *reinterpret_cast<const float4*>(ws + kk * 256 + 128 + lane * 4);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int o = 0; o < 8; ++o) acc.v[i][o] = fmaf(a[i], b[o], acc.v[i][o]);
        }
        __syncthreads();
    }
}

// softplus(beta=100, threshold=20) and its derivative; e = exp(-100|x|) via MUFU.
// abs error of the value < 3e-9 (everything is divided by beta), of the derivative < 2e-7.
__device__ __forceinline__ float softplus100(float x, float& dsig) {
    const float t = x * 100.0f;
    const float e = __expf(-fabsf(t));
    const float r = __fdividef(1.0f, 1.0f + e);
    dsig = (t > 0.0f) ? r : e * r;
    const float sp = fmaxf(x, 0.0f) + __logf(1.0f + e) * 0.01f;
    if (t > 20.0f) { dsig = 1.0f; return x; }
    return sp;
}

// store an 8-point column segment of one output feature
__device__ __forceinline__ void store_col8(float* dst, const float v[8]) {
    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void load_col8(const float* src, float v[8]) {
    const float4 a = *reinterpret_cast<const float4*>(src);
    const float4 b = *reinterpret_cast<const float4*>(src + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// ===========================================================================================
// SDF network
// ===========================================================================================
struct SdfParams {
    const float* wt[SDF_LAYERS]; const float* b[SDF_LAYERS]; const float* wn[SDF_LAYERS];
    const float* head_w; const float* head_b; const float* feat_wt; const float* feat_b;
};

template <bool WANT_GRAD, bool WANT_FEAT>
__global__ void __launch_bounds__(NT, 2)
sdf_mlp_kernel(SdfParams P, Strided3 pts, int64_t N, float* __restrict__ sdf_out,
               float* __restrict__ gx, float* __restrict__ gy, float* __restrict__ gz, int64_t gstride,
               float* __restrict__ feat_out, float* __restrict__ sig_scratch) {
    extern __shared__ __align__(16) float smem[];
    float* act = smem;                       // [256][TMP]
    float* pe = act + 256 * TMP;             // [40][TMP]
    float* ge = pe + PE_PAD * TMP;           // [40][TMP]   (grad only)
    float* wstage = ge + (WANT_GRAD ? PE_PAD * TMP : 0);   // [2][KC*256]
    const int tid = threadIdx.x, lane = tid & 31, tp = tid >> 5;
    float* sig = sig_scratch + (size_t)blockIdx.x * (SDF_LAYERS * 256 * TM);

    const int64_t ntiles = (N + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t p0 = tile * TM;
        // ---- Fourier encoding of 3*p into pe[0..38], row 39 = 0 ------------------------------
        {
            const int pt = tid & 63, q = tid >> 6;
            const int64_t p = p0 + pt;
            const bool valid = p < N;
            if (q == 0) {
                const float x = valid ? pts.x[p * pts.stride] * SDF_SCALE : 0.f;
                const float y = valid ? pts.y[p * pts.stride] * SDF_SCALE : 0.f;
                const float z = valid ? pts.z[p * pts.stride] * SDF_SCALE : 0.f;
                pe[0 * TMP + pt] = x; pe[1 * TMP + pt] = y; pe[2 * TMP + pt] = z;
                pe[39 * TMP + pt] = 0.f;
            } else {
                const int d = q - 1;
                const float* src = d == 0 ? pts.x : (d == 1 ? pts.y : pts.z);
                const float x = valid ? src[p * pts.stride] * SDF_SCALE : 0.f;
                float f = 1.0f;
#pragma unroll
                for (int k = 0; k < SDF_FREQ; ++k) {
                    const float s = x * f;
                    pe[(3 + d * SDF_FREQ + k) * TMP + pt] = sinf(s);
                    pe[(3 + 3 * SDF_FREQ + d * SDF_FREQ + k) * TMP + pt] = sinf(s + 1.57079637050628662109375f);
                    f *= 2.0f;
                }
            }
        }
        __syncthreads();

        // ---- forward layers ------------------------------------------------------------------
        Acc acc;
#pragma unroll 1
        for (int l = 0; l < SDF_LAYERS; ++l) {
            acc_set_bias(acc, P.b[l], lane);
            if (l == 0) gemm_accumulate(acc, pe, PE_PAD, P.wt[0], wstage);
            else gemm_accumulate(acc, act, 256, P.wt[l], wstage);
            const bool skip_out = (l == SDF_SKIP - 1);       // lin3: 217 outputs, then cat pe, /sqrt2
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                const int n = out_index(lane, o);
                float h[8], s[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    h[i] = softplus100(acc.v[i][o], s[i]);
                    if (skip_out) h[i] = h[i] / INV_SQRT2_DIV;
                }
                if (!skip_out || n < SKIP_H) {
                    store_col8(act + n * TMP + tp * 8, h);
                    if (WANT_GRAD) store_col8(sig + ((size_t)l * 256 + n) * TM + tp * 8, s);
                }
            }
            if (skip_out) {
                for (int idx = tid; idx < PE_DIM * TM; idx += NT) {
                    const int r = idx / TM, pt = idx % TM;
                    act[(SKIP_H + r) * TMP + pt] = pe[r * TMP + pt] / INV_SQRT2_DIV;
                }
            }
            __syncthreads();
        }

        // ---- sdf head: (w . h + b) / 3 ---------------------------------------------------------
        {
            const int pt = tid & 63, part = tid >> 6;
            float s = 0.f;
#pragma unroll 8
            for (int k = part * 64; k < part * 64 + 64; ++k) s = fmaf(act[k * TMP + pt], __ldg(P.head_w + k), s);
            wstage[part * TM + pt] = s;
            __syncthreads();
            if (part == 0) {
                const float tot = ((wstage[pt] + wstage[TM + pt]) + (wstage[2 * TM + pt] + wstage[3 * TM + pt])) + __ldg(P.head_b);
                const int64_t p = p0 + pt;
                if (p < N) sdf_out[p] = tot / SDF_SCALE;
            }
            __syncthreads();
        }

        // ---- feature head ----------------------------------------------------------------------
        if (WANT_FEAT) {
            acc_set_bias(acc, P.feat_b, lane);
            gemm_accumulate(acc, act, 256, P.feat_wt, wstage);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int64_t p = p0 + tp * 8 + i;
                if (p < N) {
                    float* dst = feat_out + p * 256;
                    *reinterpret_cast<float4*>(dst + lane * 4) = make_float4(acc.v[i][0], acc.v[i][1], acc.v[i][2], acc.v[i][3]);
                    *reinterpret_cast<float4*>(dst + 128 + lane * 4) = make_float4(acc.v[i][4], acc.v[i][5], acc.v[i][6], acc.v[i][7]);
                }
            }
        }

        // ---- reverse sweep: d sdf / d p ----------------------------------------------------------
        if (WANT_GRAD) {
            // g_pre7 = (w_s / 3) * softplus'(pre7)
            for (int idx = tid; idx < 256 * (TM / 4); idx += NT) {
                const int n = idx / (TM / 4), c4 = (idx % (TM / 4)) * 4;
                const float w = __ldg(P.head_w + n) / SDF_SCALE;
                float4 s4 = *reinterpret_cast<const float4*>(sig + ((size_t)(SDF_LAYERS - 1) * 256 + n) * TM + c4);
                *reinterpret_cast<float4*>(act + n * TMP + c4) = make_float4(w * s4.x, w * s4.y, w * s4.z, w * s4.w);
            }
            __syncthreads();
#pragma unroll 1
            for (int l = SDF_LAYERS - 1; l >= 1; --l) {
                acc_zero(acc);
                const int K = (l == SDF_SKIP - 1) ? SKIP_H : 256;       // out dim of layer l
                gemm_accumulate(acc, act, K, P.wn[l], wstage);
                const bool skip_in = (l == SDF_SKIP);                   // input was cat([h, pe]) / sqrt2
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    const int n = out_index(lane, o);
                    float g[8];
                    if (skip_in && n >= SKIP_H) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) g[i] = acc.v[i][o] / INV_SQRT2_DIV;
                        store_col8(ge + (n - SKIP_H) * TMP + tp * 8, g);
                    } else {
                        float s[8];
                        load_col8(sig + ((size_t)(l - 1) * 256 + n) * TM + tp * 8, s);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float v = acc.v[i][o];
                            if (skip_in) v = v / INV_SQRT2_DIV;
                            g[i] = v * s[i];
                        }
                        store_col8(act + n * TMP + tp * 8, g);
                    }
                }
                __syncthreads();
            }
            // layer 0: g_e += W0^T g_pre0 (39 inputs), thread = (point, group of 10 inputs)
            {
                const int pt = tid & 63, grp = tid >> 6;
                float s[10];
#pragma unroll
                for (int j = 0; j < 10; ++j) s[j] = 0.f;
                const float* B = P.wn[0] + grp * 10;
                for (int k = 0; k < 256; ++k) {
                    const float a = act[k * TMP + pt];
#pragma unroll
                    for (int j = 0; j < 10; ++j) s[j] = fmaf(a, __ldg(B + k * PE_PAD + j), s[j]);
                }
#pragma unroll
                for (int j = 0; j < 10; ++j) {
                    const int r = grp * 10 + j;
                    if (r < PE_DIM) ge[r * TMP + pt] += s[j];
                }
            }
            __syncthreads();
            // chain through the encoding: d/dx0 [x, sin(x f), sin(x f + pi/2)], then * scale
            if (tid < 3 * TM) {
                const int pt = tid & 63, d = tid >> 6;
                const float x = pe[d * TMP + pt];
                float gsum = ge[d * TMP + pt];
                float f = 1.0f;
#pragma unroll
                for (int k = 0; k < SDF_FREQ; ++k) {
                    const float s = x * f;
                    gsum += ge[(3 + d * SDF_FREQ + k) * TMP + pt] * cosf(s) * f;
                    gsum += ge[(3 + 3 * SDF_FREQ + d * SDF_FREQ + k) * TMP + pt] * cosf(s + 1.57079637050628662109375f) * f;
                    f *= 2.0f;
                }
                const int64_t p = p0 + pt;
                if (p < N) {
                    float* dst = d == 0 ? gx : (d == 1 ? gy : gz);
                    dst[p * gstride] = gsum * SDF_SCALE;
                }
            }
        }
        __syncthreads();
    }
}

constexpr size_t sdf_smem_bytes(bool grad) {
    return (size_t)(256 * TMP + PE_PAD * TMP + (grad ? PE_PAD * TMP : 0) + 2 * KC * 256) * sizeof(float);
}

// ===========================================================================================
// Reflectance network
// ===========================================================================================
struct ColParams {
    const float* wt0a; const float* wt0b; const float* b0;
    const float* wt[3]; const float* b[3];
    const float* w4t; const float* b4;
};

__global__ void __launch_bounds__(NT, 2)
color_mlp_kernel(ColParams P, Strided3 pts, Strided3 nrm, const float* __restrict__ feat,
                 const float* __restrict__ rayfeat, int64_t R, int64_t N,
                 float* __restrict__ cr, float* __restrict__ cg, float* __restrict__ cb) {
    extern __shared__ __align__(16) float smem[];
    float* act = smem;                        // [256][TMP]
    float* wstage = act + 256 * TMP;          // [2][KC*256]
    const int tid = threadIdx.x, lane = tid & 31, tp = tid >> 5;
    const int64_t ntiles = (N + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t p0 = tile * TM;
        // ---- stage the 256 SDF features, transposed to [k][pt] ---------------------------------
        for (int idx = tid; idx < TM * 64; idx += NT) {
            const int pt = idx >> 6, c4 = (idx & 63) * 4;
            const int64_t p = p0 + pt;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < N) v = __ldg(reinterpret_cast<const float4*>(feat + p * 256 + c4));
            act[(c4 + 0) * TMP + pt] = v.x; act[(c4 + 1) * TMP + pt] = v.y;
            act[(c4 + 2) * TMP + pt] = v.z; act[(c4 + 3) * TMP + pt] = v.w;
        }
        __syncthreads();
        Acc acc;
        acc_set_bias(acc, P.b0, lane);
        gemm_accumulate(acc, act, 256, P.wt0a, wstage);
        // ---- stage the 105 (+7 zero) non-feature inputs in act rows 0..111 -----------------------
        for (int idx = tid; idx < AUX_ROWS * TM; idx += NT) {
            const int r = idx / TM, pt = idx % TM;
            const int64_t p = p0 + pt;
            float v = 0.f;
            if (p < N) {
                if (r < 3) v = (r == 0 ? pts.x : (r == 1 ? pts.y : pts.z))[p * pts.stride];
                else if (r < AUX_NORMAL) v = rayfeat[(int64_t)(r - AUX_VIEW) * R + (p % R)];
                else if (r < AUX_LIGHT) { const int d = r - AUX_NORMAL; v = (d == 0 ? nrm.x : (d == 1 ? nrm.y : nrm.z))[p * nrm.stride]; }
                else if (r < AUX_ROWS - 7) v = rayfeat[(int64_t)(r - AUX_LIGHT + COL_PE3) * R + (p % R)];
            }
            act[r * TMP + pt] = v;
        }
        __syncthreads();
        gemm_accumulate(acc, act, AUX_ROWS, P.wt0b, wstage);
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const int n = out_index(lane, o);
            float h[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) h[i] = fmaxf(acc.v[i][o], 0.f);
            store_col8(act + n * TMP + tp * 8, h);
        }
        __syncthreads();
#pragma unroll 1
        for (int l = 0; l < 3; ++l) {
            acc_set_bias(acc, P.b[l], lane);
            gemm_accumulate(acc, act, 256, P.wt[l], wstage);
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                const int n = out_index(lane, o);
                float h[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) h[i] = fmaxf(acc.v[i][o], 0.f);
                store_col8(act + n * TMP + tp * 8, h);
            }
            __syncthreads();
        }
        // ---- output layer (3) + sigmoid ------------------------------------------------------------
        {
            const int pt = tid & 63, ch = tid >> 6;
            if (ch < 3) {
                float s = __ldg(P.b4 + ch);
                for (int k = 0; k < 256; ++k) s = fmaf(act[k * TMP + pt], __ldg(P.w4t + k * 4 + ch), s);
                const float c = 1.0f / (1.0f + expf(-s));
                const int64_t p = p0 + pt;
                if (p < N) (ch == 0 ? cr : (ch == 1 ? cg : cb))[p] = c;
            }
        }
        __syncthreads();
    }
}

constexpr size_t col_smem_bytes() { return (size_t)(256 * TMP + 2 * KC * 256) * sizeof(float); }

// ===========================================================================================
// Outside NeRF (NeRF.forward, /root/reference/fields/nerf_density_field.py:66-89, called from
// render_outside, models/neus_hint_model.py:434-473).  Non-default background model: fp32 FFMA only.
// ===========================================================================================
struct NerfParams {
    const float* wt[NERF_LAYERS]; const float* b[NERF_LAYERS]; const float* wt5e;
    const float* alpha_w; const float* alpha_b; const float* feat_wt; const float* feat_b;
    const float* view_wta; const float* view_wtb; const float* view_b; const float* rgb_wt; const float* rgb_b;
    const float* o[3]; const float* d[3]; const float* pl;       // rays: SoA origins / directions [R], lights [R,3]
};

__device__ __forceinline__ void relu_store(const Acc& acc, float* act, int lane, int tp) {
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        const int n = out_index(lane, o);
        float h[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = fmaxf(acc.v[i][o], 0.f);
        store_col8(act + n * TMP + tp * 8, h);
    }
}

__global__ void __launch_bounds__(NT, 2)
nerf_mlp_kernel(NerfParams P, const float* __restrict__ mid, int64_t R, int64_t N, float* __restrict__ density,
                float* __restrict__ cr, float* __restrict__ cg, float* __restrict__ cb) {
    extern __shared__ __align__(16) float smem[];
    float* act = smem;                        // [256][TMP]
    float* pe = act + 256 * TMP;              // [88][TMP]: PE(10) of the 4-D point; later [56][TMP]: PE(4) of (view, light)
    float* wstage = pe + NERF_PE_PAD * TMP;   // [2][KC*256]
    const int tid = threadIdx.x, lane = tid & 31, tp = tid >> 5;
    const int64_t ntiles = (N + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t p0 = tile * TM;
        // ---- inverted-sphere point and its Fourier encoding: rows [x(4) | sin (d-major, k-minor)(40) | cos(40) | 0(4)] ----
        {
            const int pt = tid & 63, q = tid >> 6;            // q = coordinate of the 4-D point
            const int64_t p = p0 + pt;
            float x = 0.f;
            if (p < N) {
                const int64_t r = p % R;
                const float o[3] = {P.o[0][r], P.o[1][r], P.o[2][r]}, d[3] = {P.d[0][r], P.d[1][r], P.d[2][r]};
                float p4[4];
                outside_point(o, d, mid[p], p4);
                x = p4[q];
            }
            pe[q * TMP + pt] = x;
            pe[(NERF_PE + q) * TMP + pt] = 0.f;
            float f = 1.0f;
#pragma unroll
            for (int k = 0; k < NERF_FREQ; ++k) {
                const float s = x * f;
                pe[(4 + q * NERF_FREQ + k) * TMP + pt] = sinf(s);
                pe[(4 + 4 * NERF_FREQ + q * NERF_FREQ + k) * TMP + pt] = sinf(s + 1.57079637050628662109375f);
                f *= 2.0f;
            }
        }
        __syncthreads();
        // ---- trunk: 8 x 256 ReLU; layer 5 consumes cat([pe, h]) ----
        Acc acc;
#pragma unroll 1
        for (int l = 0; l < NERF_LAYERS; ++l) {
            acc_set_bias(acc, P.b[l], lane);
            if (l == 0) gemm_accumulate(acc, pe, NERF_PE_PAD, P.wt[0], wstage);
            else {
                if (l == NERF_SKIP + 1) gemm_accumulate(acc, pe, NERF_PE_PAD, P.wt5e, wstage);
                gemm_accumulate(acc, act, 256, P.wt[l], wstage);
            }
            relu_store(acc, act, lane, tp);
            __syncthreads();
        }
        // ---- density head (raw; softplus / alpha are applied by the compositor) + (view, light) encoding into `pe` ----
        {
            const int pt = tid & 63, part = tid >> 6;
            float s = 0.f;
#pragma unroll 8
            for (int k = part * 64; k < part * 64 + 64; ++k) s = fmaf(act[k * TMP + pt], __ldg(P.alpha_w + k), s);
            wstage[part * TM + pt] = s;
            const int64_t p = p0 + pt;
            // 6-D input [view(3), light(3)]: thread group `part` encodes coordinates part and part + 4 (< 6)
            for (int c = part; c < 6; c += 4) {
                float x = 0.f;
                if (p < N) { const int64_t r = p % R; x = (c < 3) ? P.d[c][r] : P.pl[r * 3 + (c - 3)]; }
                pe[c * TMP + pt] = x;
                float f = 1.0f;
#pragma unroll
                for (int k = 0; k < NERF_VFREQ; ++k) {
                    const float sv = x * f;
                    pe[(6 + c * NERF_VFREQ + k) * TMP + pt] = sinf(sv);
                    pe[(6 + 6 * NERF_VFREQ + c * NERF_VFREQ + k) * TMP + pt] = sinf(sv + 1.57079637050628662109375f);
                    f *= 2.0f;
                }
            }
            if (part < 2) pe[(NERF_VPE + part) * TMP + pt] = 0.f;
            __syncthreads();
            if (part == 0) {
                const float tot = ((wstage[pt] + wstage[TM + pt]) + (wstage[2 * TM + pt] + wstage[3 * TM + pt])) + __ldg(P.alpha_b);
                if (p < N) density[p] = tot;
            }
            __syncthreads();
        }
        // ---- feature head (linear) -> act ----
        acc_set_bias(acc, P.feat_b, lane);
        gemm_accumulate(acc, act, 256, P.feat_wt, wstage);
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const int n = out_index(lane, o);
            float h[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) h[i] = acc.v[i][o];
            store_col8(act + n * TMP + tp * 8, h);
        }
        __syncthreads();
        // ---- view/light layer: relu(W [feature | PE(view, light)] + b), 128 wide (columns 128..255 are zero weights) ----
        acc_set_bias(acc, P.view_b, lane);
        gemm_accumulate(acc, act, 256, P.view_wta, wstage);
        gemm_accumulate(acc, pe, NERF_VPE_PAD, P.view_wtb, wstage);
        relu_store(acc, act, lane, tp);
        __syncthreads();
        // ---- rgb head + sigmoid (render_outside :460) ----
        {
            const int pt = tid & 63, ch = tid >> 6;
            if (ch < 3) {
                float s = __ldg(P.rgb_b + ch);
                for (int k = 0; k < NERF_VIEW_H; ++k) s = fmaf(act[k * TMP + pt], __ldg(P.rgb_wt + k * 4 + ch), s);
                const float c = 1.0f / (1.0f + expf(-s));
                const int64_t p = p0 + pt;
                if (p < N) (ch == 0 ? cr : (ch == 1 ? cg : cb))[p] = c;
            }
        }
        __syncthreads();
    }
}

constexpr size_t nerf_smem_bytes() { return (size_t)(256 * TMP + NERF_PE_PAD * TMP + 2 * KC * 256) * sizeof(float); }

}  // namespace

size_t sdf_mlp_simt_scratch_bytes(int num_sms) {
    return (size_t)num_sms * 2 * SDF_LAYERS * 256 * TM * sizeof(float);
}

int sdf_mlp_simt(const float* packed, const PackedLayout& L, Strided3 pts, int64_t N,
                 float* sdf, float* gx, float* gy, float* gz, int64_t grad_stride, float* feat,
                 float* scratch, size_t scratch_bytes, int num_sms, cudaStream_t st) {
    if (N <= 0) return NRH_OK;
    SdfParams P;
    for (int l = 0; l < SDF_LAYERS; ++l) { P.wt[l] = packed + L.sdf_wt[l]; P.b[l] = packed + L.sdf_b[l]; P.wn[l] = packed + L.sdf_wn[l]; }
    P.head_w = packed + L.head_w; P.head_b = packed + L.head_b; P.feat_wt = packed + L.feat_wt; P.feat_b = packed + L.feat_b;
    const bool grad = gx != nullptr, wfeat = feat != nullptr;
    int64_t ntiles = (N + TM - 1) / TM;
    int grid = (int)(ntiles < (int64_t)num_sms * 2 ? ntiles : (int64_t)num_sms * 2);
    if (grad && scratch_bytes < (size_t)grid * SDF_LAYERS * 256 * TM * sizeof(float)) {
        set_error("sdf_mlp_simt: scratch too small"); return NRH_ERR_WORKSPACE;
    }
    const size_t smem = sdf_smem_bytes(grad);
#define NRH_LAUNCH_SDF(G, F)                                                                                   \
    do {                                                                                                       \
        NRH_CUDA_CHECK(cudaFuncSetAttribute(sdf_mlp_kernel<G, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        sdf_mlp_kernel<G, F><<<grid, NT, smem, st>>>(P, pts, N, sdf, gx, gy, gz, grad_stride, feat, scratch);  \
    } while (0)
    if (grad && wfeat) NRH_LAUNCH_SDF(true, true);
    else if (grad) NRH_LAUNCH_SDF(true, false);
    else if (wfeat) NRH_LAUNCH_SDF(false, true);
    else NRH_LAUNCH_SDF(false, false);
#undef NRH_LAUNCH_SDF
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int color_mlp_simt(const float* packed, const PackedLayout& L, Strided3 pts, Strided3 normals,
                   const float* feat, const float* rayfeat, int64_t R, int64_t N,
                   float* cr, float* cg, float* cb, int num_sms, cudaStream_t st) {
    if (N <= 0) return NRH_OK;
    ColParams P;
    P.wt0a = packed + L.col_wt0a; P.wt0b = packed + L.col_wt0b; P.b0 = packed + L.col_b0;
    for (int l = 0; l < 3; ++l) { P.wt[l] = packed + L.col_wt[l]; P.b[l] = packed + L.col_b[l]; }
    P.w4t = packed + L.col_w4t; P.b4 = packed + L.col_b4;
    int64_t ntiles = (N + TM - 1) / TM;
    int grid = (int)(ntiles < (int64_t)num_sms * 2 ? ntiles : (int64_t)num_sms * 2);
    const size_t smem = col_smem_bytes();
    NRH_CUDA_CHECK(cudaFuncSetAttribute(color_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    color_mlp_kernel<<<grid, NT, smem, st>>>(P, pts, normals, feat, rayfeat, R, N, cr, cg, cb);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int nerf_mlp_simt(const float* packed, const PackedLayout& L, const float* const o[3], const float* const d[3], const float* pl,
                  const float* mid, int64_t R, int64_t N, float* density, float* cr, float* cg, float* cb, int num_sms,
                  cudaStream_t st) {
    if (N <= 0) return NRH_OK;
    NerfParams P;
    for (int l = 0; l < NERF_LAYERS; ++l) { P.wt[l] = packed + L.nerf_wt[l]; P.b[l] = packed + L.nerf_b[l]; }
    P.wt5e = packed + L.nerf_wt5e;
    P.alpha_w = packed + L.nerf_alpha_w; P.alpha_b = packed + L.nerf_alpha_b;
    P.feat_wt = packed + L.nerf_feat_wt; P.feat_b = packed + L.nerf_feat_b;
    P.view_wta = packed + L.nerf_view_wta; P.view_wtb = packed + L.nerf_view_wtb; P.view_b = packed + L.nerf_view_b;
    P.rgb_wt = packed + L.nerf_rgb_wt; P.rgb_b = packed + L.nerf_rgb_b;
    for (int c = 0; c < 3; ++c) { P.o[c] = o[c]; P.d[c] = d[c]; }
    P.pl = pl;
    int64_t ntiles = (N + TM - 1) / TM;
    int grid = (int)(ntiles < (int64_t)num_sms * 2 ? ntiles : (int64_t)num_sms * 2);
    const size_t smem = nerf_smem_bytes();
    NRH_CUDA_CHECK(cudaFuncSetAttribute(nerf_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nerf_mlp_kernel<<<grid, NT, smem, st>>>(P, mid, R, N, density, cr, cg, cb);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

}  // namespace nrh
