// Multi-resolution hash-grid encoding: HashEncoding.pytorch_fwd
// (/root/reference/fields/encodings.py:306-366; hash_fn :306-322).  The reference never instantiates this
// encoder, so it is exported as a standalone operator (nrh_hash_encode / _backward).
//
// Bound: L2 gathers.  The whole table (16 levels x 2^19 x 2 fp32 = 64 MiB) fits the 126 MB L2, a point needs
// 16 levels x 8 corners x 8 B of table entries (each pulling a 32 B sector) against 12 B in + 128 B out of
// algorithmic HBM traffic.  Layout choices: a warp works on ONE level for 32 consecutive points (neighbouring
// points share grid cells, so corner fetches of a warp coalesce into few sectors on the coarse levels), the
// encoded rows are staged in shared memory and leave the SM as fully coalesced 16 B stores.
//
// Index arithmetic and interpolation follow the reference operation by operation (no FMA contraction), so the
// result is bit-identical to the torch fallback: corners = ceil / floor of x * scaling as int32,
// hash = (x*1) ^ (y*2654435761) ^ (z*805459861) in int64 WITHOUT 32-bit wrap, floor-mod 2^log2_T, + level * T,
// and the weight `offset = scaled - floor` goes to the CEIL corner.
#include <cuda_runtime.h>
#include <stdint.h>
#include "nrh_common.cuh"

namespace nrh {
namespace {

constexpr int HE_PTS = 64;         // points per CTA tile
constexpr int HE_THREADS = 256;
constexpr int HE_MAX_LEVELS = 32;
constexpr int HE_MAX_F = 8;

struct HashArgs {
    float scaling[HE_MAX_LEVELS];
    int n_levels, log2_T, F;
};

__device__ __forceinline__ int64_t hash3(int x, int y, int z, int64_t mask, int64_t level_offset) {
    const int64_t h = ((int64_t)x * 1LL) ^ ((int64_t)y * 2654435761LL) ^ ((int64_t)z * 805459861LL);
    return (h & mask) + level_offset;        // floor-mod by a power of two == two's-complement mask
}

struct Cell {
    int64_t idx[8];       // reference corner order: hashed_0 .. hashed_7
    float ox, oy, oz;
};

__device__ __forceinline__ Cell locate(float x, float y, float z, float scaling, int level, int log2_T) {
    Cell c;
    const float sx = __fmul_rn(x, scaling), sy = __fmul_rn(y, scaling), sz = __fmul_rn(z, scaling);
    const int cx = (int)ceilf(sx), cy = (int)ceilf(sy), cz = (int)ceilf(sz);
    const int fx = (int)floorf(sx), fy = (int)floorf(sy), fz = (int)floorf(sz);
    c.ox = __fsub_rn(sx, (float)fx); c.oy = __fsub_rn(sy, (float)fy); c.oz = __fsub_rn(sz, (float)fz);
    const int64_t T = (int64_t)1 << log2_T, mask = T - 1, off = (int64_t)level * T;
    c.idx[0] = hash3(cx, cy, cz, mask, off);
    c.idx[1] = hash3(cx, fy, cz, mask, off);
    c.idx[2] = hash3(fx, fy, cz, mask, off);
    c.idx[3] = hash3(fx, cy, cz, mask, off);
    c.idx[4] = hash3(cx, cy, fz, mask, off);
    c.idx[5] = hash3(cx, fy, fz, mask, off);
    c.idx[6] = hash3(fx, fy, fz, mask, off);
    c.idx[7] = hash3(fx, cy, fz, mask, off);
    return c;
}

// a*t + b*(1-t) with the reference's separate roundings
__device__ __forceinline__ float mix(float a, float b, float t, float one_minus_t) {
    return __fadd_rn(__fmul_rn(a, t), __fmul_rn(b, one_minus_t));
}

template <int F>
__global__ void __launch_bounds__(HE_THREADS)
k_hash_encode(const float* __restrict__ pts, int64_t N, const float* __restrict__ table, HashArgs A, float* __restrict__ out) {
    extern __shared__ float tile[];                      // [HE_PTS][n_levels*F + 1]
    const int width = A.n_levels * F, pitch = width + 1;
    const int64_t p0 = (int64_t)blockIdx.x * HE_PTS;
    for (int item = threadIdx.x; item < HE_PTS * A.n_levels; item += HE_THREADS) {
        const int level = item / HE_PTS, pt = item % HE_PTS;
        const int64_t p = p0 + pt;
        if (p >= N) continue;
        const Cell c = locate(__ldg(pts + p * 3), __ldg(pts + p * 3 + 1), __ldg(pts + p * 3 + 2), A.scaling[level], level, A.log2_T);
        float f[8][F];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float* row = table + c.idx[k] * F;
            if (F == 2) { const float2 v = __ldg(reinterpret_cast<const float2*>(row)); f[k][0] = v.x; f[k][1] = v.y; }
            else if (F == 4) { const float4 v = __ldg(reinterpret_cast<const float4*>(row)); f[k][0] = v.x; f[k][1] = v.y; f[k][2] = v.z; f[k][3] = v.w; }
            else {
#pragma unroll
                for (int j = 0; j < F; ++j) f[k][j] = __ldg(row + j);
            }
        }
        const float mx = __fsub_rn(1.0f, c.ox), my = __fsub_rn(1.0f, c.oy), mz = __fsub_rn(1.0f, c.oz);
#pragma unroll
        for (int j = 0; j < F; ++j) {
            const float f03 = mix(f[0][j], f[3][j], c.ox, mx);
            const float f12 = mix(f[1][j], f[2][j], c.ox, mx);
            const float f56 = mix(f[5][j], f[6][j], c.ox, mx);
            const float f47 = mix(f[4][j], f[7][j], c.ox, mx);
            const float f0312 = mix(f03, f12, c.oy, my);
            const float f4756 = mix(f47, f56, c.oy, my);
            tile[pt * pitch + level * F + j] = mix(f0312, f4756, c.oz, mz);
        }
    }
    __syncthreads();
    // the tile's rows are consecutive in `out`: one contiguous block of rows*width floats
    const int rows = (int)((N - p0) < HE_PTS ? (N - p0) : HE_PTS);
    float* dst = out + p0 * width;
    for (int i = threadIdx.x; i < rows * width; i += HE_THREADS) dst[i] = tile[(i / width) * pitch + (i % width)];
}

template <int F>
__global__ void __launch_bounds__(HE_THREADS)
k_hash_encode_backward(const float* __restrict__ pts, int64_t N, const float* __restrict__ d_out, HashArgs A,
                       float* __restrict__ d_table) {
    const int width = A.n_levels * F;
    const int64_t p0 = (int64_t)blockIdx.x * HE_PTS;
    for (int item = threadIdx.x; item < HE_PTS * A.n_levels; item += HE_THREADS) {
        const int level = item / HE_PTS, pt = item % HE_PTS;
        const int64_t p = p0 + pt;
        if (p >= N) continue;
        const Cell c = locate(__ldg(pts + p * 3), __ldg(pts + p * 3 + 1), __ldg(pts + p * 3 + 2), A.scaling[level], level, A.log2_T);
        const float mx = 1.0f - c.ox, my = 1.0f - c.oy, mz = 1.0f - c.oz;
        // corner weights implied by the interpolation tree: x: {0,1,5,4 -> ox; 3,2,6,7 -> 1-ox}, y: {03,47 -> oy; 12,56 -> 1-oy},
        // z: {0312 -> oz; 4756 -> 1-oz}
        const float w[8] = {c.ox * c.oy * c.oz, c.ox * my * c.oz, mx * my * c.oz, mx * c.oy * c.oz,
                            c.ox * c.oy * mz,   c.ox * my * mz,   mx * my * mz,   mx * c.oy * mz};
#pragma unroll
        for (int j = 0; j < F; ++j) {
            const float g = __ldg(d_out + p * width + level * F + j);
#pragma unroll
            for (int k = 0; k < 8; ++k) atomicAdd(d_table + c.idx[k] * F + j, w[k] * g);
        }
    }
}

int fill_args(HashArgs& A, const float* host_scalings, int n_levels, int log2_T, int F) {
    if (!host_scalings || n_levels < 1 || n_levels > HE_MAX_LEVELS || F < 1 || F > HE_MAX_F || log2_T < 1 || log2_T > 30) {
        set_error("nrh_hash_encode: need 1 <= n_levels <= %d, 1 <= features_per_level <= %d, 1 <= log2_table_size <= 30",
                  HE_MAX_LEVELS, HE_MAX_F);
        return NRH_ERR_INVALID;
    }
    A.n_levels = n_levels; A.log2_T = log2_T; A.F = F;
    for (int l = 0; l < HE_MAX_LEVELS; ++l) A.scaling[l] = l < n_levels ? host_scalings[l] : 0.f;
    return NRH_OK;
}

}  // namespace
}  // namespace nrh

using namespace nrh;

extern "C" {

int nrh_hash_encode(const float* pts, int64_t N, const float* table, const float* host_scalings, int n_levels,
                    int log2_table_size, int features_per_level, float* out, void* stream) {
    if (N == 0) return NRH_OK;
    if (!pts || !table || !out || N < 0) { set_error("nrh_hash_encode: null argument"); return NRH_ERR_INVALID; }
    HashArgs A; int rc = fill_args(A, host_scalings, n_levels, log2_table_size, features_per_level); if (rc) return rc;
    if (N == 0) return NRH_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((N + HE_PTS - 1) / HE_PTS);
    const size_t smem = (size_t)HE_PTS * (n_levels * features_per_level + 1) * sizeof(float);
#define NRH_HE(FF) case FF:                                                                                            \
        if (smem > 48 * 1024) NRH_CUDA_CHECK(cudaFuncSetAttribute(k_hash_encode<FF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_hash_encode<FF><<<grid, HE_THREADS, smem, st>>>(pts, N, table, A, out); break;
    switch (features_per_level) { NRH_HE(1) NRH_HE(2) NRH_HE(4) NRH_HE(8)
        default: set_error("nrh_hash_encode: features_per_level must be 1, 2, 4 or 8"); return NRH_ERR_UNSUPPORTED; }
#undef NRH_HE
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int nrh_hash_encode_backward(const float* pts, int64_t N, const float* d_out, const float* host_scalings, int n_levels,
                             int log2_table_size, int features_per_level, float* d_table, void* stream) {
    if (N == 0) return NRH_OK;
    if (!pts || !d_out || !d_table || N < 0) { set_error("nrh_hash_encode_backward: null argument"); return NRH_ERR_INVALID; }
    HashArgs A; int rc = fill_args(A, host_scalings, n_levels, log2_table_size, features_per_level); if (rc) return rc;
    if (N == 0) return NRH_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((N + HE_PTS - 1) / HE_PTS);
#define NRH_HE(FF) case FF: k_hash_encode_backward<FF><<<grid, HE_THREADS, 0, st>>>(pts, N, d_out, A, d_table); break;
    switch (features_per_level) { NRH_HE(1) NRH_HE(2) NRH_HE(4) NRH_HE(8)
        default: set_error("nrh_hash_encode_backward: features_per_level must be 1, 2, 4 or 8"); return NRH_ERR_UNSUPPORTED; }
#undef NRH_HE
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

}  // extern "C"
