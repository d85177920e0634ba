// Standalone probe of the sm_100a primitives in nrhints_b200/csrc/tc_primitives.cuh (run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tests/_build/tc_probe tests/tc_probe.cu && tests/_build/tc_probe)
// One CTA: B tiles arrive by bulk async copy (pre-swizzled images), A tiles are written by the threads with the
// same swizzled generic stores the MLP epilogue uses, tcgen05.mma accumulates K = 128 into two TMEM accumulators
// (columns 0..255 and 256..511), tcgen05.ld reads them back.  Inputs are small multiples of 1/4, so every product
// and partial sum is exact in fp32 and the comparison against the host result is bitwise.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../nrhints_b200/csrc/tc_primitives.cuh"

using namespace nrh::tc;

constexpr int M = 128, N = 256, K = 128, KC = 64;

__global__ void __launch_bounds__(128, 1) probe_kernel(const float* __restrict__ A, const __half* __restrict__ Bimg, float* __restrict__ D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_tile = smem;                                  // 2 chunks x [128 x 64] fp16 = 2 x 16 KB
    uint8_t* b_tile = smem + 2 * 16384;                      // 2 chunks x [256 x 64] fp16 = 2 x 32 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * 16384 + 2 * 32768);   // [0] b_full, [1] mma_done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (tid == 32) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (tid == 0) {
        mbar_arrive_expect_tx(&bars[0], 2 * 32768);
        bulk_g2s(b_tile, Bimg, 32768, &bars[0]);
        bulk_g2s(b_tile + 32768, Bimg + 32768 / 2, 32768, &bars[0]);
    }
    // A tile: thread = row, generic swizzled 16-byte stores
    for (int c = 0; c < K / KC; ++c)
        for (int k8 = 0; k8 < KC; k8 += 8) {
            __half h[8];
            for (int i = 0; i < 8; ++i) h[i] = __float2half(A[tid * K + c * KC + k8 + i]);
            *reinterpret_cast<uint4*>(a_tile + c * 16384 + sw128_offset(tid, k8)) = *reinterpret_cast<uint4*>(h);
        }
    fence_proxy_async_smem();
    __syncthreads();

    if (tid == 0) {
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_f16(M, N);
        for (int acc = 0; acc < 2; ++acc) {                 // accumulator 1 gets the same product twice (D2 = 2 * D1)
            for (int rep = 0; rep <= acc; ++rep)
                for (int c = 0; c < K / KC; ++c)
                    for (int ks = 0; ks < KC / 16; ++ks) {
                        const uint64_t ad = make_desc_sw128(smem_u32(a_tile + c * 16384) + ks * 32);
                        const uint64_t bd = make_desc_sw128(smem_u32(b_tile + c * 32768) + ks * 32);
                        umma_f16(tmem_base + acc * 256, ad, bd, idesc, (rep | c | ks) != 0);
                    }
        }
        umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    for (int acc = 0; acc < 2; ++acc)
        for (int c0 = 0; c0 < N; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + acc * 256 + c0, v);
            tmem_wait_ld();
            for (int i = 0; i < 32; ++i) D[(size_t)acc * M * N + tid * N + c0 + i] = v[i];
        }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
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
    // pre-swizzled B image: chunk-major, [256 x 64] fp16 SW128 tiles
    std::vector<__half> Bimg((size_t)N * K);
    for (int c = 0; c < K / KC; ++c)
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < KC; ++k)
                Bimg[((size_t)c * 32768 + sw128_offset(n, k)) / 2] = __float2half(B[n * K + c * KC + k]);
    float *dA, *dD; __half* dB;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, Bimg.size() * 2); cudaMalloc(&dD, 2 * M * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bimg.data(), Bimg.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 2 * M * N * 4);
    const int smem = 2 * 16384 + 2 * 32768 + 64 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("PROBE CUDA ERROR: %s\n", cudaGetErrorString(e)); return 2; }
    std::vector<float> D(2 * M * N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad1 = 0, bad2 = 0;
    for (int i = 0; i < M * N; ++i) {
        if (D[i] != Dref[i]) { if (bad1 < 5) printf("acc0 mismatch m=%d n=%d got %f want %f\n", i / N, i % N, D[i], Dref[i]); ++bad1; }
        if (D[M * N + i] != 2.f * Dref[i]) { if (bad2 < 5) printf("acc1 mismatch m=%d n=%d got %f want %f\n", i / N, i % N, D[M * N + i], 2.f * Dref[i]); ++bad2; }
    }
    printf("PROBE %s: acc0 mismatches %d, acc1 mismatches %d of %d\n", (bad1 | bad2) ? "FAIL" : "PASS", bad1, bad2, M * N);
    return (bad1 | bad2) ? 1 : 0;
}
