// Micro-benchmark (developer tool, run on the GPU box): do tcgen05.ld reads of one TMEM accumulator overlap with
// tcgen05.mma writes into another one?  Three scenarios per CTA: MMAs only, TMEM loads only, both concurrently.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tests/_build/tc_probe2 tests/tc_probe2.cu
#include <cstdio>
#include <vector>
#include "../nrhints_b200/csrc/tc_primitives.cuh"
using namespace nrh::tc;

__global__ void __launch_bounds__(576, 1) bench(int mode, int n_mma, int n_ld, long long* out, float* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_tile = smem;                  // 16 KB
    uint8_t* b_tile = smem + 16384;          // 32 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (warp == 0) tmem_alloc(slot, 512);
    if (tid == 32) { mbar_init(&bar[0], 1); fence_mbar_init(); }
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = *slot;
    long long t0 = clock64();
    if (warp == 1 && lane == 0 && (mode & 1)) {
        const uint32_t idesc = make_idesc_f16(128, 256);
        for (int i = 0; i < n_mma; ++i)
            umma_f16(tb, make_desc_sw128(smem_u32(a_tile) + (i & 3) * 32), make_desc_sw128(smem_u32(b_tile) + (i & 3) * 32), idesc, i != 0);
        umma_commit(&bar[0]);
        mbar_wait(&bar[0], 0);
        out[blockIdx.x * 4 + 0] = clock64() - t0;
    }
    if (warp >= 2 && (mode & 2)) {
        float acc = 0.f;
        const uint32_t base = tb + ((uint32_t)((warp & 3) * 32) << 16) + 256 + ((warp - 2) >> 2) * 16;
        for (int i = 0; i < n_ld; ++i) {
            float v[16];
            tmem_ld16(base + (i & 3) * 64, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc += v[j];
        }
        if (acc == 123.f) sink[tid] = acc;
        if (warp == 2 && lane == 0) out[blockIdx.x * 4 + 1] = clock64() - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
    long long* d; float* sink; cudaMalloc(&d, 148 * 4 * 8); cudaMalloc(&sink, 4096);
    const int smem = 49152 + 64;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int grid : {1, 148})
        for (int mode = 1; mode <= 3; ++mode) {
            cudaMemset(d, 0, 148 * 4 * 8);
            bench<<<grid, 576, smem>>>(mode, 4800, 6800, d, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA ERROR %s\n", cudaGetErrorString(e)); return 1; }
            std::vector<long long> h(grid * 4);
            cudaMemcpy(h.data(), d, grid * 4 * 8, cudaMemcpyDeviceToHost);
            printf("grid %3d mode %d: 4800 MMAs (128x256x16) take %lld clk (%.1f clk/MMA); 6800 x16-loads per warp (16 warps) take %lld clk (%.1f clk/load-iter)\n",
                   grid, mode, h[0], h[0] / 4800.0, h[1], h[1] / 6800.0);
        }
    return 0;
}
