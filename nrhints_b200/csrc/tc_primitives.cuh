// sm_100a primitives used by the tcgen05 MLP engine: mbarrier, bulk async copy (TMA unit, UBLKCP),
// TMEM allocation / load, UMMA shared-memory + instruction descriptors, tcgen05.mma / commit.
// Hand-written inline PTX (no CUTLASS); encodings follow the PTX ISA "tcgen05" chapter
// (cross-checked against cute/arch/mma_sm100_desc.hpp field layouts).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nrh {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded spin: a protocol bug must abort the kernel (trap -> CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 27)) asm volatile("trap;");
    }
}

// ---- bulk async copy global -> shared (TMA unit, no tensor map), completion on an mbarrier ----------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// multicast variant: the bytes land at the same CTA-relative address in every CTA of `cta_mask`, and each of
// those CTAs' mbarrier (same CTA-relative address) receives the complete_tx
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint16_t cta_mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
// all threads of all CTAs of the cluster (a plain launch is a cluster of one CTA)
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-D tiled TMA load through a tensor map (cp.async.bulk.tensor): box at element coordinates (x = column, y = row), completion
// (full box bytes, out-of-range elements zero-filled) on an mbarrier.  `map` points to a CUtensorMap in param / const / global space.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM ------------------------------------------------------------------------------------------------
// one full warp; ncols power of two in [32,512]; the base address lands in *smem_slot
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane (warp%4)*32 + t.
// taddr = (lane_base << 16) | column
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}

// 32 lanes x 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr) : "memory");
}

// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

// 32 lanes x 8 consecutive 32-bit columns, registers -> TMEM (thread t of the warp writes lane (warp%4)*32 + t)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors ------------------------------------------------------------------------------------
// K-major operand tile, 128-byte swizzle: rows of 64 fp16 (128 B), 8-row atoms of 1024 B packed densely
// (SBO = 1024 B), LBO unused (=1) for swizzled K-major, descriptor version 1 (sm_100), layout type 2.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16, A = B = fp16, D = fp32, both K-major, dense; M in {64,128}, N multiple of 16 up to 256
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// bf16 x bf16 -> fp32
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// Same MMA with the descriptors given as 32-bit low words (start address | LBO): the high words of both
// SWIZZLE_128B K-major descriptors are the constant DESC_HI.  Keeps the single issuing thread's loop short.
constexpr uint32_t DESC_HI = 64u | (1u << 14) | (2u << 29);          // SBO = 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ void umma_f16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand is read from TENSOR MEMORY (row m in lane m, elements k = 2c, 2c+1 packed in
// 32-bit column c, low half = even k; verified bitwise in tests/tc_probe5.cu), B as above.
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI) : "memory");
}
// Warp-uniform variants: the WHOLE warp executes the call with identical operands and one elected lane issues.  Called from
// warp-uniform control flow, the operands live in uniform registers and the instruction needs no per-operand
// "elect + R2UR.BROADCAST" waterfall (12 SASS instructions per MMA when a single lane of a diverged warp issues).
__device__ __forceinline__ void umma_f16_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 db;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI) : "memory");
}
__device__ __forceinline__ void umma_f16_lo_w(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI) : "memory");
}
// One operand sub-chunk (2 K-steps) of the fp16 hi/lo split as ONE instruction group behind ONE election: A_hi*W_hi, A_lo*W_hi,
// A_hi*W_lo for both K-steps = 6 MMAs.  Issued one by one through umma_*_w every MMA pays its own ELECT + VOTEU + R2UR.BROADCAST
// preamble (~13 SASS instructions; the issuing warp shares its scheduler with four arithmetic-heavy epilogue warps, so that is
// ~85 clk per MMA -- more than the 64.5 clk an N = 128 MMA occupies the tensor pipe); here the preamble is paid once per group.
// A operand in TENSOR MEMORY: K-step j of the sub-chunk at columns a0 + 16 j (hi) and a0 + 16 j + 8 (lo).
__device__ __forceinline__ void umma_group6_ts_w(uint32_t d_tmem, uint32_t a0, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 b0, b1, b2, b3;\n\t.reg .b32 a1, a2, a3, t;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mov.b64 b0, {%2, %5};\n\t"
        "add.u32 t, %2, 2;\n\tmov.b64 b1, {t, %5};\n\t"
        "add.u32 t, %2, 4;\n\tmov.b64 b2, {t, %5};\n\t"
        "add.u32 t, %2, 6;\n\tmov.b64 b3, {t, %5};\n\t"
        "add.u32 a1, %1, 16;\n\tadd.u32 a2, %1, 8;\n\tadd.u32 a3, %1, 24;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], b0, %3, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %3, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], b0, %3, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], b1, %3, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], b2, %3, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b3, %3, 1;\n\t}"
        ::"r"(d_tmem), "r"(a0), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI) : "memory");
}
// the same with the A operand in SHARED memory: descriptor low words a_hi / a_lo of the sub-chunk's first K-step (second: + 2)
__device__ __forceinline__ void umma_group6_ss_w(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 b0, b1, b2, b3, h0, h1, l0, l1;\n\t.reg .b32 t;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mov.b64 b0, {%3, %6};\n\t"
        "add.u32 t, %3, 2;\n\tmov.b64 b1, {t, %6};\n\t"
        "add.u32 t, %3, 4;\n\tmov.b64 b2, {t, %6};\n\t"
        "add.u32 t, %3, 6;\n\tmov.b64 b3, {t, %6};\n\t"
        "mov.b64 h0, {%1, %6};\n\tadd.u32 t, %1, 2;\n\tmov.b64 h1, {t, %6};\n\t"
        "mov.b64 l0, {%2, %6};\n\tadd.u32 t, %2, 2;\n\tmov.b64 l1, {t, %6};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], h0, b0, %4, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], h1, b1, %4, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], l0, b0, %4, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], l1, b1, %4, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], h0, b2, %4, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], h1, b3, %4, 1;\n\t}"
        ::"r"(d_tmem), "r"(a_hi), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI) : "memory");
}
// two-pass form of the above: A_hi * W_hi + A_hi * W_lo (the A operand is published as a single fp16 tile; reverse sweep)
__device__ __forceinline__ void umma_group4_ss_w(uint32_t d_tmem, uint32_t a_hi, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 b0, b1, b2, b3, h0, h1;\n\t.reg .b32 t;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mov.b64 b0, {%2, %5};\n\t"
        "add.u32 t, %2, 2;\n\tmov.b64 b1, {t, %5};\n\t"
        "add.u32 t, %2, 4;\n\tmov.b64 b2, {t, %5};\n\t"
        "add.u32 t, %2, 6;\n\tmov.b64 b3, {t, %5};\n\t"
        "mov.b64 h0, {%1, %5};\n\tadd.u32 t, %1, 2;\n\tmov.b64 h1, {t, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], h0, b0, %3, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], h1, b1, %3, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], h0, b2, %3, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], h1, b3, %3, 1;\n\t}"
        ::"r"(d_tmem), "r"(a_hi), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI) : "memory");
}
// single-pass form: A_hi * W_hi only (the W_lo half of the image is streamed but not used)
__device__ __forceinline__ void umma_group2_ss_w(uint32_t d_tmem, uint32_t a_hi, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 b0, b1, h0, h1;\n\t.reg .b32 t;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mov.b64 b0, {%2, %5};\n\t"
        "add.u32 t, %2, 2;\n\tmov.b64 b1, {t, %5};\n\t"
        "mov.b64 h0, {%1, %5};\n\tadd.u32 t, %1, 2;\n\tmov.b64 h1, {t, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], h0, b0, %3, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], h1, b1, %3, 1;\n\t}"
        ::"r"(d_tmem), "r"(a_hi), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI) : "memory");
}
__device__ __forceinline__ void umma_commit_w(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on `bar` once every previously issued tcgen05.mma of this thread has completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// the same, arriving on the barrier at this CTA-relative address in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}

// ---- operand images --------------------------------------------------------------------------------------------
// byte offset of element (row, k) inside a [rows x 64] fp16 K-major SWIZZLE_128B tile (rows % 8 == 0)
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t k) {
    return (row >> 3) * 1024u + (row & 7u) * 128u + ((((k >> 3) ^ (row & 7u)) & 7u) << 4) + (k & 7u) * 2u;
}

}  // namespace tc
}  // namespace nrh
