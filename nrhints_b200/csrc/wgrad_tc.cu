// Weight-gradient reductions of a training step on tcgen05: every dW of the SDF network and of the
// reflectance network is a point-reduction over two fp16 row-major dumps the forward / backward kernels publish anyway,
//
//     out[m, n] += scale * sum_p A[p, a_col0 + m] * B[p, b_col0 + n]          (m < 256, n < N <= 256, p < rows ~ 5e5)
//
// i.e. a [256 x P] x [P x N] GEMM whose reduction dimension is the SLOW axis of both operands.  The reference leaves these to
// autograd (one cuBLAS call per F.linear backward, fields/sdf_field.py:106-123 / fields/reflectance_network.py:84-96); round 1
// called torch.mm.  Here ALL reductions of a step are ONE persistent launch:
//
//   * operands: both matrices are "MN-major" for the MMA (the 64 M/N elements of a row are contiguous, K = the point index is the
//     outer dimension).  Tiles of [64 points x 64 columns] are fetched by the TMA unit through tensor maps
//     (cp.async.bulk.tensor.2d, SWIZZLE_128B, out-of-range rows zero-filled, so ragged P needs no padding) and land exactly in the
//     canonical MN-major SWIZZLE_128B shared-memory layout (8 point-rows x 128 B atoms; column blocks LBO = 8 KB apart, groups
//     of 8 points SBO = 1 KB apart), which tcgen05.mma reads with a_major = b_major = MN in the instruction descriptor;
//   * accumulators: the full 256 x N fp32 result lives in tensor memory (two M = 128 halves x N columns = up to all 512 columns);
//   * split-K without a workspace: the global list of 64-point tiles of all jobs is cut into gridDim.x CONTIGUOUS ranges, so a
//     CTA works on one job (at most two or three) for its whole life, accumulates in TMEM across tiles and flushes once per job
//     with vectorised fp32 reductions (red.global.add.v4.f32) straight into the caller's gradient buffer -- which the caller
//     zeroes once per step (it is the flat buffer of train_ops.FlatAdam's effective-weight gradients);
//   * roofline: the kernel is HBM-bound -- 64 KB of operands per 1 K clk of MMAs per SM is 2.7x what HBM can deliver -- so what
//     matters is that every dump is read exactly once at full bandwidth (3-stage TMA ring, 64 KB per stage).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = flush (one TMEM lane quarter each).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "nrh_common.cuh"
#include "mlp_tc.cuh"
#include "tc_primitives.cuh"

namespace nrh {
namespace {
using namespace tc;

constexpr int WG_MAX_JOBS = 48;
constexpr int WG_MAX_MAPS = 96;
constexpr int WG_THREADS = 192;
constexpr int WG_STAGES = 3;
constexpr uint32_t WG_BLOCK = 8192;                 // one [64 points x 64 columns] fp16 box
constexpr uint32_t WG_A_BYTES = 4 * WG_BLOCK;       // M = 256
constexpr uint32_t WG_B_BYTES = 4 * WG_BLOCK;       // N <= 256
constexpr uint32_t WG_STAGE = WG_A_BYTES + WG_B_BYTES;
constexpr size_t WG_SMEM = WG_STAGES * WG_STAGE + 1024 + 1024;

struct WgJob {
    int map_a, map_b;               // tensor-map indices
    int a_col0, b_col0;             // first column of the operand inside its matrix (multiples of 64 / 8)
    int n;                          // N: 64, 128, 192 or 256
    int rows_valid, cols_valid;     // rows / columns of the result that are written
    long long k_tiles;              // ceil(rows / 64)
    float scale; const float* dev_scale;      // result multiplier: scale * (*dev_scale if set)
    float* out; long long ld_out;
};
struct WgParams {
    CUtensorMap maps[WG_MAX_MAPS];
    WgJob jobs[WG_MAX_JOBS];
    int njobs;
    long long total_tiles;
};
static_assert(sizeof(WgParams) <= 32000, "wgrad parameter block exceeds the kernel parameter space");

// kind::f16, fp16 x fp16 -> fp32, BOTH operands MN-major (bits 15 / 16), dense
__host__ __device__ constexpr uint32_t make_idesc_f16_mn(int M, int N) {
    return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// MN-major SWIZZLE_128B descriptor, low word: start address | LBO (byte stride between 64-column blocks = 8 KB)
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | ((WG_BLOCK >> 4) << 16); }

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ WgParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE);
    uint64_t* full = bars;                     // [WG_STAGES] TMA -> MMA
    uint64_t* empty = bars + WG_STAGES;        // [WG_STAGES] MMA -> TMA
    uint64_t* acc_full = bars + 2 * WG_STAGES;     // MMA -> flush: the accumulator of a job segment is complete
    uint64_t* acc_empty = acc_full + 1;            // flush -> MMA: it has been read out
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;

    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (tid == 32) {
        for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 4);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // this CTA's contiguous range of the global tile list, and the job its first tile belongs to
    const long long t0 = P.total_tiles * blockIdx.x / gridDim.x, t1 = P.total_tiles * (blockIdx.x + 1) / gridDim.x;
    int j0 = 0;
    long long base0 = 0;                                       // global index of job j0's first tile
    while (j0 < P.njobs && base0 + P.jobs[j0].k_tiles <= t0) { base0 += P.jobs[j0].k_tiles; ++j0; }

    if (warp == 0) {
        // ======================= TMA producer =======================
        if (lane == 0) {
            uint32_t it = 0;
            int j = j0;
            long long jb = base0;
            for (long long t = t0; t < t1; ++t, ++it) {
                while (t >= jb + P.jobs[j].k_tiles) { jb += P.jobs[j].k_tiles; ++j; }
                const WgJob& J = P.jobs[j];
                const uint32_t s = it % WG_STAGES, u = it / WG_STAGES;
                mbar_wait(&empty[s], (u & 1) ^ 1);
                const int nb = J.n >> 6;
                mbar_arrive_expect_tx(&full[s], (uint32_t)(4 + nb) * WG_BLOCK);
                uint8_t* a = smem + s * WG_STAGE;
                uint8_t* b = a + WG_A_BYTES;
                const int y = (int)((t - jb) * 64);
                for (int c = 0; c < 4; ++c) tma_load_2d(a + c * WG_BLOCK, &P.maps[J.map_a], J.a_col0 + 64 * c, y, &full[s]);
                for (int c = 0; c < nb; ++c) tma_load_2d(b + c * WG_BLOCK, &P.maps[J.map_b], J.b_col0 + 64 * c, y, &full[s]);
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer (whole warp in lockstep, one elected lane issues) =======================
        uint32_t it = 0, seg = 0;
        int j = j0;
        long long jb = base0;
        const uint32_t a0 = desc_lo_mn(smem_u32(smem)), b0 = desc_lo_mn(smem_u32(smem + WG_A_BYTES));
        bool first = true;
        for (long long t = t0; t < t1; ++t, ++it) {
            if (t >= jb + P.jobs[j].k_tiles) {
                // job boundary inside this CTA's range: hand the finished accumulator to the flush warps, wait until it is read
                while (t >= jb + P.jobs[j].k_tiles) { jb += P.jobs[j].k_tiles; ++j; }
                umma_commit_w(acc_full);
                mbar_wait(acc_empty, seg & 1);
                tc_fence_after();
                ++seg;
                first = true;
            }
            const uint32_t idesc = make_idesc_f16_mn(128, P.jobs[j].n);
            const uint32_t s = it % WG_STAGES, u = it / WG_STAGES;
            mbar_wait(&full[s], u & 1);
            tc_fence_after();
            const uint32_t al = a0 + s * (WG_STAGE >> 4), bl = b0 + s * (WG_STAGE >> 4);
#pragma unroll
            for (uint32_t ks = 0; ks < 4; ++ks) {                 // 16 points per MMA = 2 groups of 8 point-rows = 2 KB
                const uint32_t acc = (first && ks == 0) ? 0u : 1u;
                umma_f16_lo_w(tmem_base, al + ks * 128u, bl + ks * 128u, idesc, acc);                               // rows   0..127
                umma_f16_lo_w(tmem_base + 256u, al + (2 * WG_BLOCK >> 4) + ks * 128u, bl + ks * 128u, idesc, acc);  // rows 128..255
            }
            first = false;
            umma_commit_w(&empty[s]);
        }
        if (t1 > t0) umma_commit_w(acc_full);
    } else {
        // ======================= flush warps: TMEM -> scale -> red.add into the caller's gradient buffer =======================
        const int q = warp & 3;                                    // TMEM lane quarter this warp may read
        uint32_t seg = 0;
        int j = j0;
        long long jb = base0;
        long long t = t0;
        while (t < t1) {
            while (t >= jb + P.jobs[j].k_tiles) { jb += P.jobs[j].k_tiles; ++j; }
            const WgJob& J = P.jobs[j];
            const long long seg_end = (jb + J.k_tiles < t1) ? jb + J.k_tiles : t1;
            mbar_wait(acc_full, seg & 1);
            tc_fence_after();
            const float sc = J.scale * (J.dev_scale ? __ldg(J.dev_scale) : 1.0f);
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int m = h * 128 + q * 32 + lane;
                float* orow = J.out + (long long)m * J.ld_out;
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)h * 256u;
#pragma unroll 1
                for (int c0 = 0; c0 < J.n; c0 += 16) {
                    float v[16];
                    tmem_ld16(ta + c0, v);
                    tmem_wait_ld();
                    if (m < J.rows_valid) {
                        if (c0 + 16 <= J.cols_valid && (J.ld_out & 3) == 0) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4) red_add_v4(orow + c0 + i, v[i] * sc, v[i + 1] * sc, v[i + 2] * sc, v[i + 3] * sc);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (c0 + i < J.cols_valid) atomicAdd(orow + c0 + i, v[i] * sc);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
            ++seg;
            t = seg_end;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---- small-M companion: out[m, n] += scale * sum_p A[p, m] * B[p, n] for m < 8 (the 3-row output layer of the reflectance
//      network, the sdf head's d_sdf row): a bandwidth-trivial streaming reduction on the CUDA cores ----------------------------
template <int MR>
__global__ void __launch_bounds__(256)
k_wgrad_skinny(const __half* __restrict__ A, long long a_ld, const __half* __restrict__ B, long long b_ld, long long rows, int n,
               float scale, const float* __restrict__ dev_scale, float* __restrict__ out, long long ld_out, int rows_valid) {
    static_assert(MR == 8, "one 16-byte load fetches the 8 A values of a row");
    const int col = blockIdx.y * 256 + threadIdx.x;
    float acc[MR];
#pragma unroll
    for (int i = 0; i < MR; ++i) acc[i] = 0.f;
    if (col < n) {
        const long long stride = gridDim.x;
        long long p = blockIdx.x;
        // four rows in flight per trip: the loop is latency-bound otherwise (one dependent 16-byte broadcast load + one 2-byte load per row)
        for (; p + 3 * stride < rows; p += 4 * stride) {
            uint4 a4[4]; float b4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a4[u] = __ldg(reinterpret_cast<const uint4*>(A + (p + u * stride) * a_ld));
                b4[u] = __half2float(B[(p + u * stride) * b_ld + col]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const __half2* h2 = reinterpret_cast<const __half2*>(&a4[u]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(h2[i]);
                    acc[2 * i] = fmaf(f.x, b4[u], acc[2 * i]);
                    acc[2 * i + 1] = fmaf(f.y, b4[u], acc[2 * i + 1]);
                }
            }
        }
        for (; p < rows; p += stride) {
            const uint4 a4 = __ldg(reinterpret_cast<const uint4*>(A + p * a_ld));
            const float b = __half2float(B[p * b_ld + col]);
            const __half2* h2 = reinterpret_cast<const __half2*>(&a4);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(h2[i]);
                acc[2 * i] = fmaf(f.x, b, acc[2 * i]);
                acc[2 * i + 1] = fmaf(f.y, b, acc[2 * i + 1]);
            }
        }
        const float sc = scale * (dev_scale ? __ldg(dev_scale) : 1.0f);
#pragma unroll
        for (int i = 0; i < MR; ++i)
            if (i < rows_valid) atomicAdd(out + (long long)i * ld_out + col, acc[i] * sc);
    }
}

// The same with 16-byte loads of B: a lane owns 8 consecutive columns, a warp one row of a 256-column block, the 8 warps of a CTA
// eight rows per trip and four trips in flight (16 KB of B per CTA on the wire); NV = number of result rows kept (1, 3 or 8).
// Streams B once at HBM speed (the scalar version above is latency-bound: 2 bytes per thread and load).
template <int NV>
__global__ void __launch_bounds__(256)
k_wgrad_skinny_v(const __half* __restrict__ A, long long a_ld, const __half* __restrict__ B, long long b_ld, long long rows, int n,
                 float scale, const float* __restrict__ dev_scale, float* __restrict__ out, long long ld_out, int rows_valid) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = blockIdx.y * 256 + lane * 8;
    float acc[NV][8];
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const bool live = col < n;
    const long long stride = (long long)gridDim.x * 8;
    auto fma_row = [&](const uint4& a4, const uint4& b4) {
        const __half* ah = reinterpret_cast<const __half*>(&a4);
        const __half2* bh = reinterpret_cast<const __half2*>(&b4);
        float bf[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(bh[j]); bf[2 * j] = f.x; bf[2 * j + 1] = f.y; }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float a = __half2float(ah[i]);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a, bf[j], acc[i][j]);
        }
    };
    if (live) {
        long long p = (long long)blockIdx.x * 8 + warp;
        for (; p + 3 * stride < rows; p += 4 * stride) {
            uint4 a4[4], b4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a4[u] = __ldg(reinterpret_cast<const uint4*>(A + (p + u * stride) * a_ld));
                b4[u] = __ldg(reinterpret_cast<const uint4*>(B + (p + u * stride) * b_ld + col));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) fma_row(a4[u], b4[u]);
        }
        for (; p < rows; p += stride)
            fma_row(__ldg(reinterpret_cast<const uint4*>(A + p * a_ld)), __ldg(reinterpret_cast<const uint4*>(B + p * b_ld + col)));
    }
    __shared__ float part[8][NV][256];
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) part[warp][i][lane * 8 + j] = acc[i][j];
    __syncthreads();
    const int c = blockIdx.y * 256 + threadIdx.x;
    if (c < n) {
        const float sc = scale * (dev_scale ? __ldg(dev_scale) : 1.0f);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += part[w][i][threadIdx.x];
            if (i < rows_valid) atomicAdd(out + (long long)i * ld_out + c, v * sc);
        }
    }
}

}  // namespace

// Tensor map of a row-major fp16 matrix [rows][ld] for boxes of [box_rows x box_cols] elements, SWIZZLE_128B (box_cols = 64: the
// 128-byte rows of the canonical K-major / MN-major operand atoms), zero fill outside.  The driver entry point is looked up through the
// runtime (no link dependency on libcuda).
int encode_tensor_map_f16(void* out_map, const void* ptr, long long ld, long long rows, int box_cols, int box_rows) {
    static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return NRH_ERR_CUDA; }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld & 7)) { set_error("tensor-map operands need 16-byte alignment and ld %% 8 == 0"); return NRH_ERR_INVALID; }
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    const CUresult r = enc(reinterpret_cast<CUtensorMap*>(out_map), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for a [%lld x %lld] fp16 matrix", (int)r, rows, ld); return NRH_ERR_CUDA; }
    return NRH_OK;
}

}  // namespace nrh

extern "C" int nrh_wgrad_f16(const NrhWgradJob* jobs, int njobs, void* stream) {
    using namespace nrh;
    if (njobs <= 0) return NRH_OK;
    if (!jobs) { set_error("nrh_wgrad_f16: null job list"); return NRH_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    static thread_local WgParams P;                    // 13 KB: kept off the stack
    P.njobs = 0; P.total_tiles = 0;
    int nmaps = 0;
    struct Key { const void* p; long long ld, rows; } keys[WG_MAX_MAPS];
    auto get_map = [&](const void* ptr, long long ld, long long rows, int* idx) -> int {
        for (int i = 0; i < nmaps; ++i)
            if (keys[i].p == ptr && keys[i].ld == ld && keys[i].rows == rows) { *idx = i; return NRH_OK; }
        if (nmaps >= WG_MAX_MAPS) { set_error("nrh_wgrad_f16: too many distinct operand matrices"); return NRH_ERR_INVALID; }
        const int rc = encode_tensor_map_f16(&P.maps[nmaps], ptr, ld, rows, 64, 64);
        if (rc) return rc;
        keys[nmaps] = Key{ptr, ld, rows};
        *idx = nmaps++;
        return NRH_OK;
    };
    for (int i = 0; i < njobs; ++i) {
        const NrhWgradJob& J = jobs[i];
        if (!J.a || !J.b || !J.out || J.rows <= 0) { set_error("nrh_wgrad_f16: job %d has a null operand or no rows", i); return NRH_ERR_INVALID; }
        if (J.m <= 8) {                                 // skinny reduction on the CUDA cores
            if ((reinterpret_cast<uintptr_t>(J.a) & 15) || (J.a_ld & 7) || (J.a_col0 & 7)) { set_error("nrh_wgrad_f16: job %d: skinny A rows must be 16-byte aligned", i); return NRH_ERR_INVALID; }
            if (J.n <= 0 || J.n > 1024 || J.m < 1) { set_error("nrh_wgrad_f16: job %d: bad shape %d x %d", i, J.m, J.n); return NRH_ERR_INVALID; }
            const __half* A = reinterpret_cast<const __half*>(J.a) + J.a_col0;
            const __half* B = reinterpret_cast<const __half*>(J.b) + J.b_col0;
            dim3 grid(1184, (J.n + 255) / 256);                   // 8 CTAs per SM, ~440 rows each
            const int nv = J.rows_valid > 0 ? (J.rows_valid < J.m ? J.rows_valid : J.m) : J.m;
            const int ncols = J.cols_valid > 0 ? J.cols_valid : J.n;
            const bool vec = !(reinterpret_cast<uintptr_t>(B) & 15) && !(J.b_ld & 7) && !(ncols & 7);
            if (vec && nv == 1)      k_wgrad_skinny_v<1><<<grid, 256, 0, st>>>(A, J.a_ld, B, J.b_ld, J.rows, ncols, J.scale, J.dev_scale, J.out, J.ld_out, nv);
            else if (vec && nv <= 3) k_wgrad_skinny_v<3><<<grid, 256, 0, st>>>(A, J.a_ld, B, J.b_ld, J.rows, ncols, J.scale, J.dev_scale, J.out, J.ld_out, nv);
            else k_wgrad_skinny<8><<<grid, 256, 0, st>>>(A, J.a_ld, B, J.b_ld, J.rows, ncols, J.scale, J.dev_scale, J.out, J.ld_out, nv);
            NRH_LAUNCH_CHECK();
            continue;
        }
        if (J.m != 256 || J.n < 64 || J.n > 256 || (J.n & 63) || (J.a_col0 & 7) || (J.b_col0 & 7)) {
            set_error("nrh_wgrad_f16: job %d: the tensor-core path takes m == 256 (or m <= 8) and n in {64,128,192,256} (got %d x %d)", i, J.m, J.n);
            return NRH_ERR_UNSUPPORTED;
        }
        if (P.njobs >= WG_MAX_JOBS) { set_error("nrh_wgrad_f16: more than %d tensor-core jobs in one call", WG_MAX_JOBS); return NRH_ERR_INVALID; }
        WgJob& W = P.jobs[P.njobs++];
        int rc;
        if ((rc = get_map(J.a, J.a_ld, J.rows, &W.map_a))) return rc;
        if ((rc = get_map(J.b, J.b_ld, J.rows, &W.map_b))) return rc;
        W.a_col0 = J.a_col0; W.b_col0 = J.b_col0; W.n = J.n;
        W.rows_valid = J.rows_valid > 0 ? J.rows_valid : 256;
        W.cols_valid = J.cols_valid > 0 ? J.cols_valid : J.n;
        W.k_tiles = (J.rows + 63) / 64;
        W.scale = J.scale; W.dev_scale = J.dev_scale; W.out = J.out; W.ld_out = J.ld_out;
        P.total_tiles += W.k_tiles;
    }
    if (P.njobs == 0) return NRH_OK;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)(P.total_tiles < sms ? P.total_tiles : sms);
    static bool attr_set = false;
    if (!attr_set) {
        NRH_CUDA_CHECK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM));
        attr_set = true;
    }
    wgrad_tc_kernel<<<grid, WG_THREADS, WG_SMEM, st>>>(P);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}
