// Developer probe (GPU box): does st.async (STAS) to the CTA's own shared memory work, with and without a cluster launch,
// and what does it cost next to st.shared + fence.proxy.async + arrive?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tests/_build/tc_probe4 tests/tc_probe4.cu
#include <cstdio>
#include <cstdint>
#include "../nrhints_b200/csrc/tc_primitives.cuh"
using namespace nrh::tc;
// Result (B200): st.async is an illegal instruction unless the kernel is launched with a cluster attribute; its round trip
// (512 x 16 B -> barrier phase complete) takes 666 clk against 347 clk for st.shared + fence.proxy.async + arrive, and used
// for the epilogue's operand stores it made the SDF kernel 1.8x slower (17.1K vs 9.6K clk per layer).  Not used.
__device__ __forceinline__ void st_async128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t bar_saddr) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(bar_saddr) : "memory");
}

__global__ void __launch_bounds__(512, 1) k(int mode, int iters, uint32_t* out, long long* clk) {
    __shared__ __align__(1024) uint8_t buf[16384];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(&bar, mode == 0 ? 1 : 16); fence_mbar_init(); }
    __syncthreads();
    const uint32_t a = smem_u32(buf) + tid * 16, b = smem_u32(&bar);
    long long t0 = clock64();
    uint32_t par = 0;
    for (int i = 0; i < iters; ++i) {
        if (mode == 0) {
            if (tid == 0) mbar_arrive_expect_tx(&bar, 512 * 16);
            st_async128(a, i, tid, 3u, 4u, b);
        } else {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(i), "r"(tid), "r"(3u), "r"(4u) : "memory");
            fence_proxy_async_smem();
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&bar);
        }
        mbar_wait(&bar, par); par ^= 1;
    }
    if (tid == 0) clk[0] = clock64() - t0;
    __syncthreads();
    out[tid] = reinterpret_cast<uint32_t*>(buf)[tid * 4] + reinterpret_cast<uint32_t*>(buf)[tid * 4 + 1];
}

int main() {
    uint32_t* d; long long* c; cudaMalloc(&d, 4096); cudaMalloc(&c, 64);
    for (int cluster = 1; cluster >= 1; --cluster)
        for (int mode = 0; mode < 2; ++mode) {
            cudaLaunchConfig_t lc = {}; lc.gridDim = dim3(1); lc.blockDim = dim3(512);
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            lc.attrs = at; lc.numAttrs = cluster;
            cudaError_t e = cudaLaunchKernelEx(&lc, k, mode, 1000, d, c);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            uint32_t h[4]; long long hc = 0;
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
            printf("cluster-launch %d mode %d (%s): %s  out0=%u out1=%u  %.1f clk/iter\n", cluster, mode, mode == 0 ? "st.async" : "st.shared+fence+arrive",
                   cudaGetErrorString(e), h[0], h[1], hc / 1000.0);
            if (e != cudaSuccess) return 1;
        }
    return 0;
}
