// Developer probe (GPU box): tcgen05.mma with the A operand in TENSOR MEMORY (".ts" form), written by tcgen05.st.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tests/_build/tc_probe5 tests/tc_probe5.cu && tests/_build/tc_probe5
// (a) correctness of the operand layout assumed by the MLP engine: row m of A in TMEM lane m, elements k = 2c, 2c+1 packed
//     (low half = even k) in 32-bit column c; the operand is written IN PLACE over the columns of a finished accumulator;
// (b) tensor-pipe cost of an M128 N256 K16 MMA with A from TMEM against A from shared memory;
// (c) cost of publishing 8 activation values per thread: st.shared x2 + fence.proxy.async + arrive  vs
//     tcgen05.st x2 + tcgen05.wait::st + arrive.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../nrhints_b200/csrc/tc_primitives.cuh"

using namespace nrh::tc;

constexpr int M = 128, N = 256, K = 128, KC = 64;

__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__global__ void __launch_bounds__(128, 1) probe_kernel(const float* __restrict__ A, const __half* __restrict__ Bimg, float* __restrict__ D,
                                                        long long* __restrict__ clk) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_tile = smem;                                  // 2 chunks x [128 x 64] fp16
    uint8_t* b_tile = smem + 2 * 16384;                      // 2 chunks x [256 x 64] fp16
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * 16384 + 2 * 32768);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (tid == 32) { for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);

    if (tid == 0) {
        mbar_arrive_expect_tx(&bars[0], 2 * 32768);
        bulk_g2s(b_tile, Bimg, 32768, &bars[0]);
        bulk_g2s(b_tile + 32768, Bimg + 32768 / 2, 32768, &bars[0]);
    }
    for (int c = 0; c < K / KC; ++c)
        for (int k8 = 0; k8 < KC; k8 += 8) {
            __half h[8];
            for (int i = 0; i < 8; ++i) h[i] = __float2half(A[tid * K + c * KC + k8 + i]);
            *reinterpret_cast<uint4*>(a_tile + c * 16384 + sw128_offset(tid, k8)) = *reinterpret_cast<uint4*>(h);
        }
    fence_proxy_async_smem();
    __syncthreads();
    const uint32_t idesc = make_idesc_f16(M, N);
    const uint32_t b_lo0 = desc_lo(smem_u32(b_tile));

    // ---- step 1: accumulator X (columns 0..255) = A B^T with both operands in shared memory
    if (tid == 0) {
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        for (int c = 0; c < K / KC; ++c)
            for (int ks = 0; ks < KC / 16; ++ks)
                umma_f16_lo(tmem_base, desc_lo(smem_u32(a_tile + c * 16384)) + ks * 2, b_lo0 + c * 2048 + ks * 2, idesc, (uint32_t)((c | ks) != 0));
        umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    // ---- step 2: every thread reads its row of X, writes D0, then overwrites columns 0..63 of X with the packed fp16 row of A
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld32(lane_addr + c0, v);
        tmem_wait_ld();
        for (int i = 0; i < 32; ++i) D[tid * N + c0 + i] = v[i];
    }
    for (int c4 = 0; c4 < K / 2; c4 += 4) {
        uint32_t w[4];
        for (int i = 0; i < 4; ++i) {
            const __half2 h = __floats2half2_rn(A[tid * K + 2 * (c4 + i)], A[tid * K + 2 * (c4 + i) + 1]);
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        tmem_st4(lane_addr + c4, w[0], w[1], w[2], w[3]);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    // ---- step 3: accumulator Y (columns 256..511) = A(TMEM, in place in X) B^T
    if (tid == 0) {
        tc_fence_after();
        for (int ks = 0; ks < K / 16; ++ks)
            umma_f16_ts(tmem_base + 256, tmem_base + ks * 8, b_lo0 + (ks >> 2) * 2048 + (ks & 3) * 2, idesc, (uint32_t)(ks != 0));
        umma_commit(&bars[2]);
    }
    mbar_wait(&bars[2], 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld32(lane_addr + 256 + c0, v);
        tmem_wait_ld();
        for (int i = 0; i < 32; ++i) D[(size_t)M * N + tid * N + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    // ---- (b) timing: 480 MMAs, SS then TS, for N = 256, 128, 64 (clk[2 * n + mode])
    if (tid == 0) {
        tc_fence_after();
        uint32_t par = 1;
        for (int n = 0; n < 3; ++n) {
            const uint32_t idn = make_idesc_f16(M, 256 >> n);
            for (int mode = 0; mode < 2; ++mode) {
                const long long t0 = clock64();
                for (int i = 0; i < 480; ++i) {
                    if (mode == 0) umma_f16_lo(tmem_base + 256, desc_lo(smem_u32(a_tile)) + (i & 3) * 2, b_lo0 + (i & 3) * 2, idn, 1u);
                    else umma_f16_ts(tmem_base + 256, tmem_base + (i & 7) * 8, b_lo0 + (i & 3) * 2, idn, 1u);
                }
                umma_commit(&bars[2]);
                mbar_wait(&bars[2], par); par ^= 1;
                clk[2 * n + mode] = clock64() - t0;
            }
        }
    }
    // ---- (b2) the MMA sequence of one MLP layer, 10 layers back to back, accumulator regions alternating:
    //      mode 0: generation 1 (48 x N256, operand from shared memory); 1: the same with the operand in tensor memory;
    //      2: generation 2 (24 x N256 + 2 x 24 x N128, operand in tensor memory); 3: all 96 x N128 from tensor memory
    if (tid == 0) {
        uint32_t par = 1;
        const uint32_t idf = make_idesc_f16(M, 256), idh = make_idesc_f16(M, 128), a_s = desc_lo(smem_u32(a_tile));
        for (int mode = 0; mode < 4; ++mode) {
            const long long t0 = clock64();
            for (int layer = 0; layer < 10; ++layer) {
                const uint32_t D = tmem_base + (layer & 1) * 256, X = tmem_base + ((layer & 1) ^ 1) * 256;
                auto group = [&](int sc, uint32_t d, uint32_t idesc, bool ts) {
                    const uint32_t a0 = X + 32u * sc, s0 = a_s + (sc & 1) * 4, bl = b_lo0 + (sc & 1) * 2048;
                    if (ts) {
                        umma_f16_ts(d, a0, bl, idesc, (uint32_t)(sc != 0)); umma_f16_ts(d, a0 + 16, bl + 2, idesc, 1u);
                        umma_f16_ts(d, a0 + 8, bl, idesc, 1u); umma_f16_ts(d, a0 + 24, bl + 2, idesc, 1u);
                        umma_f16_ts(d, a0, bl + 4, idesc, 1u); umma_f16_ts(d, a0 + 16, bl + 6, idesc, 1u);
                    } else {
                        umma_f16_lo(d, s0, bl, idesc, (uint32_t)(sc != 0)); umma_f16_lo(d, s0 + 2, bl + 2, idesc, 1u);
                        umma_f16_lo(d, s0, bl, idesc, 1u); umma_f16_lo(d, s0 + 2, bl + 2, idesc, 1u);
                        umma_f16_lo(d, s0, bl + 4, idesc, 1u); umma_f16_lo(d, s0 + 2, bl + 6, idesc, 1u);
                    }
                    umma_commit(&bars[3]);                 // stage release, as in the engine (nobody waits on it here)
                };
                if (mode <= 1) { for (int sc = 0; sc < 8; ++sc) group(sc, D, idf, mode == 1); }
                else if (mode == 2) {
                    for (int sc = 0; sc < 4; ++sc) group(sc, D, idf, true);
                    for (int sc = 4; sc < 8; ++sc) group(sc, D, idh, true);
                    for (int sc = 4; sc < 8; ++sc) group(sc, D + 128, idh, true);
                } else {
                    for (int h = 0; h < 2; ++h) for (int sc = 0; sc < 8; ++sc) group(sc, D + 128 * h, idh, true);
                }
            }
            umma_commit(&bars[2]);
            mbar_wait(&bars[2], par); par ^= 1;
            clk[8 + mode] = clock64() - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// (c) publish cost, 512 threads (16 warps), `iters` publishes of 8 values per thread with `work` rounds of 8 independent FMAs
//     of "epilogue math" per publish.  mode 0: math, st.shared x2, fence.proxy.async, arrive (the engine's current order);
//     mode 1: math, tcgen05.st x2, tcgen05.wait::st, arrive; mode 2: DEFERRED hand-off -- st.shared x2 of step i, then the math
//     of step i+1, then fence.proxy.async + arrive for step i (the fence no longer waits for stores issued just before it).
__global__ void __launch_bounds__(512, 1) publish_kernel(int mode, int iters, int work, long long* clk, uint32_t* sink) {
    __shared__ __align__(1024) uint8_t buf[32768];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&slot, 512);
    if (tid == 32) { mbar_init(&bar, 16); fence_mbar_init(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t taddr = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 8;
    const uint32_t sa = smem_u32(buf) + tid * 16;
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = tid + j;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const bool odd = (mode == 3) && ((warp >> 2) & 1);      // mode 3: half of the warps hand off in the MIDDLE of their math
        for (int w = 0; w < work; ++w) {
            if (odd && w == work / 2 && i > 0) {
                fence_proxy_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = a[j] * 1.0001f + 0.5f;
        }
        uint32_t w8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w8[j] = __float_as_uint(a[j]);
        if (mode == 2 && i > 0) {                      // hand off the PREVIOUS step: its stores are long done
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar);
        }
        if (mode == 0 || mode == 2 || mode == 3) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sa), "r"(w8[0]), "r"(w8[1]), "r"(w8[2]), "r"(w8[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sa + 16384), "r"(w8[4]), "r"(w8[5]), "r"(w8[6]), "r"(w8[7]) : "memory");
            if (mode == 0 || (mode == 3 && !odd)) fence_proxy_async_smem();
        } else {
            tmem_st4(taddr + (i & 7) * 32, w8[0], w8[1], w8[2], w8[3]);
            tmem_st4(taddr + (i & 7) * 32 + 4, w8[4], w8[5], w8[6], w8[7]);
            tmem_wait_st();
        }
        if (mode != 2 && !odd) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar);
        }
    }
    const long long t1 = clock64();
    if (tid == 0) clk[0] = t1 - t0;
    sink[tid] = __float_as_uint(a[0] + a[7]);
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 512);
}

int main() {
    std::vector<float> A(M * K), B(N * K), Dref(M * N, 0.f);
    srand(1);
    for (auto& v : A) v = (rand() % 9 - 4) * 0.25f;
    for (auto& v : B) v = (rand() % 9 - 4) * 0.25f;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0.f;
            for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
            Dref[m * N + n] = s;
        }
    std::vector<__half> Bimg((size_t)N * K);
    for (int c = 0; c < K / KC; ++c)
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < KC; ++k)
                Bimg[((size_t)c * 32768 + sw128_offset(n, k)) / 2] = __float2half(B[n * K + c * KC + k]);
    float *dA, *dD; __half* dB; long long* dclk; uint32_t* dsink;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, Bimg.size() * 2); cudaMalloc(&dD, 2 * M * N * 4); cudaMalloc(&dclk, 128); cudaMalloc(&dsink, 4096);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bimg.data(), Bimg.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 2 * M * N * 4);
    const int smem = 2 * 16384 + 2 * 32768 + 64 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, dclk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("PROBE5 CUDA ERROR: %s\n", cudaGetErrorString(e)); return 2; }
    std::vector<float> D(2 * M * N);
    long long clk[16];
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(clk, dclk, 96, cudaMemcpyDeviceToHost);
    int bad1 = 0, bad2 = 0;
    for (int i = 0; i < M * N; ++i) {
        if (D[i] != Dref[i]) { if (bad1 < 5) printf("SS mismatch m=%d n=%d got %f want %f\n", i / N, i % N, D[i], Dref[i]); ++bad1; }
        if (D[M * N + i] != Dref[i]) { if (bad2 < 5) printf("TS mismatch m=%d n=%d got %f want %f\n", i / N, i % N, D[M * N + i], Dref[i]); ++bad2; }
    }
    printf("PROBE5 %s: SS mismatches %d, TS (A in TMEM, in place) mismatches %d of %d\n", (bad1 | bad2) ? "FAIL" : "PASS", bad1, bad2, M * N);
    for (int n = 0; n < 3; ++n)
        printf("PROBE5 480 MMAs M128 N%d K16: SS %lld clk (%.1f / MMA), TS %lld clk (%.1f / MMA)\n", 256 >> n, clk[2 * n], clk[2 * n] / 480.0,
               clk[2 * n + 1], clk[2 * n + 1] / 480.0);
    {
        const char* ln[4] = {"gen 1: 48 x N256, operand in shared memory", "48 x N256, operand in tensor memory",
                             "gen 2: 24 x N256 + 48 x N128, operand in tensor memory", "96 x N128, operand in tensor memory"};
        for (int mode = 0; mode < 4; ++mode) printf("PROBE5 layer MMA sequence (%s): %.0f clk per layer\n", ln[mode], clk[8 + mode] / 10.0);
    }
    const char* names[4] = {"st.shared + fence.proxy.async", "tcgen05.st + wait::st", "st.shared, fence deferred by one step",
                            "st.shared, half of the warps fence in the middle of their math"};
    for (int work = 0; work <= 12; work += 6)
        for (int mode = 0; mode < 4; ++mode) {
            publish_kernel<<<1, 512>>>(mode, 256, work, dclk, dsink);
            e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("PROBE5 publish CUDA ERROR: %s\n", cudaGetErrorString(e)); return 2; }
            cudaMemcpy(clk, dclk, 8, cudaMemcpyDeviceToHost);
            printf("PROBE5 publish, %3d FMAs of math per step, mode %d (%s): %.1f clk per step (16 warps)\n", work * 8, mode, names[mode], clk[0] / 256.0);
        }
    return (bad1 | bad2) ? 1 : 0;
}
