// Micro-benchmark (developer tool, run on the GPU box): what do the synchronisation operations of the single
// MMA-issuing thread cost the tensor pipe?  Groups of 6 MMAs (128x256x16) separated by different operations.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tests/_build/tc_probe3 tests/tc_probe3.cu
#include <cstdio>
#include <vector>
#include "../nrhints_b200/csrc/tc_primitives.cuh"
using namespace nrh::tc;

__global__ void __launch_bounds__(128, 1) bench(int mode, int groups, int per_group, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_tile = smem;                  // 16 KB
    uint8_t* b_tile = smem + 16384;          // 32 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (warp == 0) tmem_alloc(slot, 512);
    if (tid == 0) slot[1] = 7u;
    if (tid == 32) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); mbar_arrive(&bar[2]); mbar_arrive(&bar[3]); }
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = *slot;
    const uint32_t a_lo = desc_lo(smem_u32(a_tile)), b_lo = desc_lo(smem_u32(b_tile));
    const uint32_t idesc = make_idesc_f16(128, 256);
    if (warp == 1 && lane == 0) {
        long long t0 = clock64();
        uint32_t par4 = 0;
        for (int g = 0; g < groups; ++g) {
            if (mode == 5 && g < 4) {
                for (int i = 0; i < per_group; ++i) { umma_f16_lo(tb, a_lo + (i & 3) * 2, b_lo + (i & 3) * 2, idesc, 1u); out[16 + g * per_group + i] = clock64() - t0; }
                continue;
            }
            for (int i = 0; i < per_group; ++i) umma_f16_lo(tb, a_lo + (i & 3) * 2, b_lo + (i & 3) * 2, idesc, (uint32_t)((g | i) != 0));
            if (mode == 1 || mode == 6) umma_commit(&bar[1]);
            if (mode == 2 || mode == 6) { mbar_wait(&bar[2], 0); }
            if (mode == 6) { mbar_wait(&bar[3], 0); }
            if (mode == 3 || mode == 6) tc_fence_after();
            if (mode == 4) { umma_commit(&bar[4]); mbar_wait(&bar[4], par4); par4 ^= 1; }
            if (mode == 8) { out[8] = clock64(); }
            if (mode == 9) { uint32_t v; do { asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(slot + 1)) : "memory"); } while (v != 7u); }
            if (mode == 10) { uint32_t v; do { asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(slot + 1)) : "memory"); } while (v != 7u); }
            if (mode == 11) { uint32_t ok; do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.relaxed.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar[2])), "r"(0u) : "memory"); } while (!ok); }
            if (mode == 12) { uint32_t ok; do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar[2])), "r"(0u) : "memory"); } while (!ok); }
            if (mode == 13) { uint32_t v; do { asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(slot + 1)) : "memory"); } while (v != 7u); tc_fence_after(); umma_commit(&bar[1]); }
        }
        umma_commit(&bar[0]);
        mbar_wait(&bar[0], 0);
        out[0] = clock64() - t0;
    }
    if (mode == 7 && warp == 2 && lane == 0) {          // a polling neighbour thread (like the weight producer)
        for (int i = 0; i < 200000; ++i) if (mbar_try_wait(&bar[5], 0)) break;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 4096);
    const int smem = 49152 + 128;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* names[] = {"MMAs only", "+ commit (nobody waits)", "+ try_wait on a completed barrier", "+ tcgen05.fence::after_thread_sync",
                           "+ commit and wait for it (drain)", "issue stamps", "+ commit + 2 try_waits + fence", "MMAs only, neighbour thread polls a barrier", "+ clock64 + global store",
                           "+ ld.volatile.shared poll (ready)", "+ ld.acquire.cta.shared poll (ready)", "+ try_wait.relaxed (ready)", "+ test_wait (ready)", "+ ld.volatile poll + fence + commit"};
    for (int pg : {3, 6, 12})
    for (int mode = 0; mode <= 13; ++mode) {
        const int groups = 400;
        cudaMemset(d, 0, 4096);
        bench<<<1, 128, smem>>>(mode, groups, pg, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA ERROR %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(512);
        cudaMemcpy(h.data(), d, 4096, cudaMemcpyDeviceToHost);
        printf("%2d MMAs/group, mode %d (%-45s): %7.1f clk/group, %6.1f clk/MMA\n", pg, mode, names[mode], h[0] / (double)groups, h[0] / (double)(groups * pg));
        if (mode == 5) { printf("   issue stamps of the first %d MMAs:", 4 * pg); for (int i = 0; i < 4 * pg; ++i) printf(" %lld", h[16 + i]); printf("\n"); }
    }
    return 0;
}
