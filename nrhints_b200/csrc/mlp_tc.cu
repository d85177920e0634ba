// tcgen05 tensor-core MLP engine (NRH_MLP_TCGEN05) for sm_100a.
//
// One persistent CTA per SM owns a tile of 128 points (the UMMA M dimension).  Activations never leave the
// SM: they live in shared memory as fp16 K-major SWIZZLE_128B operand tiles, accumulators live in TMEM
// (two 128x256 fp32 accumulators = all 512 columns, ping-ponged across layers so the epilogue of layer l
// overlaps the MMAs of layer l+1 chunk by chunk).  Weights are pre-packed as ready-to-use operand images and
// streamed from L2 by the TMA unit (cp.async.bulk) through a 3-stage mbarrier ring.
//
// Precision: the SDF network needs fp32-grade operands (hard part 1), so every logical product
// is issued as three fp16 MMAs on hi/lo splits (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo, ~22-bit operands, fp32
// accumulate); operands are pre-scaled by powers of two so the lo parts stay normal.  The reflectance network
// enters the pixel through a sigmoid and is safe in a single fp16 pass.
//
// Warp roles (640 threads): warp 0 = weight producer (TMA), warp 1 = MMA issuer (one elected thread), warps 2-3
// idle (keeps the epilogue on whole warpgroups), warps 4..19 = epilogue (TMEM -> registers -> bias/activation/split -> swizzled smem operand for the next layer).
//
// Reference semantics: /root/reference/fields/sdf_field.py:106-148, fields/reflectance_network.py:68-96,
// fields/encodings.py:168-176.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include "mlp_tc.cuh"
#include "tc_primitives.cuh"

namespace nrh {

namespace {
using namespace tc;

constexpr int TM = 128;
constexpr int EPI_WARPS = 16;
constexpr int EPI_THREADS = EPI_WARPS * 32;     // 512
constexpr int NTHREADS = 128 + EPI_THREADS;      // warpgroup 0: warp 0 producer, warp 1 MMA issuer, warps 2-3 idle;
                                                 // warps 4..19 epilogue.
constexpr int EPI_WARP0 = 4;
constexpr int NSTAGES = 3;               // weight ring: 3 x 32 KB images (N < 256 per MMA does not pay: an SS-mode MMA costs
                                         // ~115 clk whatever N is, measured; so a stage is a full [256 x 64] image)
constexpr uint32_t STAGE = 32768;
constexpr uint32_t IMG = 32768;            // one [256 x 64] fp16 weight image
constexpr uint32_t IMG_SMALL = 8192;       // one [64 x 64] fp16 weight image (reverse layer 0)
constexpr uint32_t A_CHUNK = 16384;        // one [128 x 64] fp16 activation chunk
constexpr float W_SCALE = TC_W_SCALE;
constexpr float ACT_SCALE = TC_ACT_SCALE;
constexpr float G_SCALE = 1024.0f;         // reverse-sweep signals are stored * 2^10
constexpr float INV_SQRT2 = 0.70710678f;
constexpr int TC_MAX_JOBS = 208;                     // operand images per weight set (156 today)
constexpr int TC_GENERATION_DEFAULT = 1;
constexpr uint32_t TC_FENCE_MASK_DEFAULT = 0xFFu;     // every sub-chunk handed off on its own
// fp16 MMA passes per logical product outside the forward layers (SdfTcParams::rev_passes / feat_passes).  The forward SDF needs the
// 3-MMA split (its error is multiplied by inv_s ~ 400 in the NeuS alpha); the reverse sweep only feeds the normals and the cosine
// in alpha, whose sensitivity is O(1): see DESIGN.md section 4 for the measured error of each setting.
constexpr int TC_REV_PASSES_DEFAULT = 3;
constexpr int TC_FEAT_PASSES_DEFAULT = 3;
constexpr int TC_BWD_PASSES_DEFAULT = 3;              // sdf_bwd_tc_kernel (training backward): SdfBwdParams::passes
constexpr float OS_F = 1.0f / (W_SCALE * ACT_SCALE);   // accumulator -> forward pre-activation
constexpr float OS_R = 1.0f / W_SCALE;                 // accumulator -> reverse signal (stays in G_SCALE units)

// Developer hooks (kernel timelines, ablations that switch the MMAs / weight loads off, hand-off batching, engine generation).
// They exist ONLY in the developer build of the library (-DNRH_DEV -> libnrhints_b200_dev.so, used by tests/tc_*.py) and are set
// through the explicit nrh_dev_configure() call; the production library compiles them out (the macros below fold to constants)
// and reads nothing from the environment.
struct DevOptions { int gen; int dbg; uint32_t fmask; long long* tlog; };
#ifdef NRH_DEV
#define NRH_DBG(P) ((P).dbg)
#define NRH_TLOG(P) ((P).tlog)
DevOptions g_dev = {TC_GENERATION_DEFAULT, 0, TC_FENCE_MASK_DEFAULT, nullptr};
inline DevOptions dev_options() { return g_dev; }
#else
#define NRH_DBG(P) 0
#define NRH_TLOG(P) (static_cast<long long*>(nullptr))
inline DevOptions dev_options() { return DevOptions{TC_GENERATION_DEFAULT, 0, TC_FENCE_MASK_DEFAULT, nullptr}; }
#endif

// ---- tensor-core section of the packed weight buffer (byte offsets from the section start) ----------------
struct TcLayout {
    uint32_t fwd[SDF_LAYERS];     // l = 0: 2 images (hi, lo); l >= 1: 4 chunks x (hi, lo)
    uint32_t feat;                // 8 images
    uint32_t rev[SDF_LAYERS];     // l >= 1: 8 images of W_l^T; l = 0: 8 small images
    uint32_t col[4];              // reflectance hidden layers: 6 / 4 / 4 / 4 images (single pass)
    uint32_t sdf_bias16;          // [8][256] fp32 biases * ACT_SCALE
    uint32_t col_bias16;          // [4][256] fp32 biases * ACT_SCALE
    uint32_t feat_rev;            // 8 images of W_feat^T (backward of the feature head)
    // reflectance network, training step (color_train_tc.cu): layer 0 in the reference's NATURAL input order
    // [pts | PE(view) | normal | PE(light) | feature | PE(vis) | PE(spec)] (6 chunks, K padded to 384), and the transposed
    // weights of the backward: col_rev[l] (l = 1..3): 4 images [256 in x 64 out]; col_rev0: per 64-wide out-chunk one
    // [256 x 64] image (input rows 0..255) followed by one [128 x 64] image (input rows 256..383)
    uint32_t col_fwd0n;
    uint32_t col_rev[4];          // [0] unused (see col_rev0)
    uint32_t col_rev0;
    uint32_t col_rev0p;           // the same for the fused step's operand order [feature 256 | aux 128] (see train_step.cu::k_train_assemble)
    uint32_t total;
};
__host__ __device__ inline TcLayout tc_layout() {
    TcLayout t{};
    uint32_t off = 0;
    for (int l = 0; l < SDF_LAYERS; ++l) { t.fwd[l] = off; off += (l == 0 ? 2 : 8) * IMG; }
    t.feat = off; off += 8 * IMG;
    for (int l = 0; l < SDF_LAYERS; ++l) { t.rev[l] = off; off += (l == 0 ? 8 * IMG_SMALL : 8 * IMG); }
    for (int l = 0; l < 4; ++l) { t.col[l] = off; off += (l == 0 ? 6 : 4) * IMG; }
    t.sdf_bias16 = off; off += SDF_LAYERS * 256 * 4;
    t.col_bias16 = off; off += 4 * 256 * 4;
    t.feat_rev = off; off += 8 * IMG;
    t.col_fwd0n = off; off += 6 * IMG;
    t.col_rev[0] = 0;
    for (int l = 1; l < 4; ++l) { t.col_rev[l] = off; off += 4 * IMG; }
    t.col_rev0 = off; off += 4 * (IMG + IMG / 2);
    t.col_rev0p = off; off += 4 * (IMG + IMG / 2);
    t.total = off;
    return t;
}

// ---- training tape, per 128-point tile (see SdfTape in mlp_tc.cuh) ----
constexpr uint32_t TAPE_SIG_BYTES = SDF_LAYERS * 128 * TM * 4;        // packed softplus' of the 8 layers  [8][8 sc][4 gq][128 rows] x uint4 (pk_index)
constexpr uint32_t TAPE_G_BYTES = (SDF_LAYERS - 1) * 128 * TM * 4;    // packed g_1 .. g_7 (G_SCALE units) [7][8 sc][4 gq][128 rows] x uint4
constexpr uint32_t TAPE_GE_BYTES = PE_PAD * TM * 4;                   // g_e fp32 (G_SCALE units)          [40][128 rows]
constexpr uint32_t TAPE_TILE_BYTES = TAPE_SIG_BYTES + TAPE_G_BYTES + TAPE_GE_BYTES;

// ---- shared memory map ---------------------------------------------------------------------------------------
constexpr uint32_t SM_A_HI = 0, SM_A_LO = 65536, SM_B = 131072, SM_MISC = SM_B + NSTAGES * STAGE;    // sdf kernel
constexpr uint32_t SMC_A = 0, SMC_B = 6 * A_CHUNK, SMC_MISC = SMC_B + NSTAGES * STAGE;                 // color kernel
constexpr uint32_t SM_MISC_BYTES = 2048, SMC_MISC_BYTES = 12288;
constexpr size_t SDF_SMEM = SM_MISC + SM_MISC_BYTES + 1024;     // + slack for manual 1024-B alignment
constexpr size_t COL_SMEM = SMC_MISC + SMC_MISC_BYTES + 1024;

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
    return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
// (setmaxnreg re-balancing between the control warpgroup and the epilogue was tried: nvcc 12.9 segfaults on this
//  kernel with it, so every thread keeps the 96 registers ptxas grants a 640-thread block.)

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// softplus(beta = 100): max(x,0) + log1p(exp(-100|x|))/100 -- identical to the thresholded torch form in fp32
// (beyond 100x = 20 the log term underflows to 0).  abs error < 3e-9.
__device__ __forceinline__ float softplus100(float x) {
    const float e = ex2_approx(-144.269504f * fabsf(x));
    return fmaf(lg2_approx(1.0f + e), 6.93147181e-3f, fmaxf(x, 0.0f));
}
template <bool WANT_D>
__device__ __forceinline__ float softplus100(float x, float& dsig) {
    const float e = ex2_approx(-144.269504f * fabsf(x));
    if (WANT_D) {
        const float rr = rcp_approx(1.0f + e);
        dsig = x > 0.0f ? rr : e * rr;
    }
    return fmaf(lg2_approx(1.0f + e), 6.93147181e-3f, fmaxf(x, 0.0f));
}

__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// The reverse sweep needs softplus'(x) = sigmoid(100 x) of every forward pre-activation.  The forward epilogue is
// MUFU-bound (ex2 + lg2 per element on 16 SFU lanes/SM), the reverse epilogue has no transcendental at all, so the
// forward parks only e = exp(-100|x|) with the sign of x (one LOP3) and the reverse epilogue finishes the job:
// sigmoid = x > 0 ? 1/(1+e) : e/(1+e)  (one MUFU.RCP there).  Result is pre-divided by W_SCALE.
__device__ __forceinline__ uint32_t pack_sig2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_sig2(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }
__device__ __forceinline__ float dsig_from_packed(float pk) {
    const float e = fabsf(pk);
    const float rr = rcp_approx(1.0f + e);
    return (__float_as_int(pk) >= 0 ? rr : e * rr) * OS_R;      // sign BIT: e underflows to +-0 for |100 x| > 87
}
// Packed per-tile arrays (softplus' of the forward layers, reverse adjoints g_1..g_7): one uint4 = the 8 columns
// [32 sc + 8 gq, +8) of row r as four half2 words, stored [layer][sub-chunk][column group][row] -- exactly the unit an epilogue
// thread produces / consumes per step, so every access is ONE 16-byte LDG / STG and a warp touches 512 contiguous bytes
// (it was four strided 4-byte accesses per step; the fence of the hand-off drains them, so fewer is faster).
__device__ __forceinline__ size_t pk_index(int l, int sc, int gq, int r) { return (((size_t)l * 8 + sc) * 4 + gq) * TM + r; }
__device__ __forceinline__ void pk_load(const uint32_t* base, int l, int sc, int gq, int r, uint32_t (&w)[4]) {
    const uint4 t = reinterpret_cast<const uint4*>(base)[pk_index(l, sc, gq, r)];
    w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
}
// The inference kernels keep softplus' in a per-CTA global scratch that is written once and read once per tile.  After its
// single read the data is dead: `discard.global.L2` drops the (dirty) lines from L2 instead of letting them be written back to
// DRAM on eviction (2.3 GB of write-back per fine pass otherwise).  One lane per 128-byte line.
__device__ __forceinline__ void pk_discard(const uint32_t* base, int l, int sc, int gq, int r) {
    if ((r & 7) == 0) asm volatile("discard.global.L2 [%0], 128;" ::"l"(reinterpret_cast<const uint4*>(base) + pk_index(l, sc, gq, r)) : "memory");
}
__device__ __forceinline__ void pk_store(uint32_t* base, int l, int sc, int gq, int r, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    reinterpret_cast<uint4*>(base)[pk_index(l, sc, gq, r)] = make_uint4(a, b, c, d);
}
// split 8 fp32 values into fp16 hi / lo (value ~= hi + lo) and store both as 16-byte swizzled rows (shared addresses)
__device__ __forceinline__ void store_split8s(uint32_t s_hi, uint32_t s_lo, const float* x) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 h = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
        hw[i] = *reinterpret_cast<const uint32_t*>(&h);
        lw[i] = *reinterpret_cast<const uint32_t*>(&l);
    }
    sts128(s_hi, hw[0], hw[1], hw[2], hw[3]);
    sts128(s_lo, lw[0], lw[1], lw[2], lw[3]);
}
// the same, also returning the four packed hi words (training tape: row-major fp16 dump of the published operand)
__device__ __forceinline__ void store_split8s_hw(uint32_t s_hi, uint32_t s_lo, const float* x, uint32_t (&hw)[4]) {
    uint32_t lw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 h = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
        hw[i] = *reinterpret_cast<const uint32_t*>(&h);
        lw[i] = *reinterpret_cast<const uint32_t*>(&l);
    }
    sts128(s_hi, hw[0], hw[1], hw[2], hw[3]);
    sts128(s_lo, lw[0], lw[1], lw[2], lw[3]);
}
__device__ __forceinline__ void store_half8s_hw(uint32_t s_a, const float* x, uint32_t (&hw)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 h = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
        hw[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    sts128(s_a, hw[0], hw[1], hw[2], hw[3]);
}
__device__ __forceinline__ void store_half8s(uint32_t s_a, const float* x) {
    uint32_t hw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 h = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
        hw[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    sts128(s_a, hw[0], hw[1], hw[2], hw[3]);
}
// (generic-pointer variants, used by the one-off staging code)
__device__ __forceinline__ void store_split8(uint8_t* a_hi, uint8_t* a_lo, uint32_t off, const float* x) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 h = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
        hw[i] = *reinterpret_cast<const uint32_t*>(&h);
        lw[i] = *reinterpret_cast<const uint32_t*>(&l);
    }
    *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}
__device__ __forceinline__ void store_half8(uint8_t* a, uint32_t off, const float* x) {
    uint32_t hw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 h = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
        hw[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(a + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
}
__device__ __forceinline__ void put_split1(uint8_t* a_hi, uint8_t* a_lo, uint32_t row, uint32_t col, float x) {
    const __half h = __float2half_rn(x);
    const __half l = __float2half_rn(x - __half2float(h));
    const uint32_t off = sw128_offset(row, col);
    *reinterpret_cast<__half*>(a_hi + off) = h;
    *reinterpret_cast<__half*>(a_lo + off) = l;
}
__device__ __forceinline__ void ldg16(const float* p, float (&b)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
        b[4 * i] = t.x; b[4 * i + 1] = t.y; b[4 * i + 2] = t.z; b[4 * i + 3] = t.w;
    }
}

// ===============================================================================================================
// SDF network
// ===============================================================================================================
struct SdfTcParams {
    const uint8_t* tc;                 // tensor-core section (operand images)
    const float* bias16;               // [8][256] biases * ACT_SCALE (tensor-core section)
    const float* head_w; const float* head_b; const float* feat_b;
    int feat_image;                    // 1: features leave as fp16 operand images (TC_TILE_FEAT_BYTES per tile)
    long long* tlog;                   // developer timeline (NRH_TC_TLOG): clock64 stamps of block 0, third tile
    uint32_t fmask;                    // hand-off steps of the epilogue (EpiT::fence_mask)
    int dbg;                           // developer ablations (NRH_TC_DEBUG): 1 = epilogue skips the math, 2 = no MMAs, 3 = no weight loads, 4 = neither
    int rev_passes;                    // fp16 MMA passes per product of the REVERSE sweep: 3 = hi/lo split of both operands (as the forward),
                                       // 2 = signal published as one fp16 tile (z_hi W_hi + z_hi W_lo), 1 = z_hi W_hi only
    int feat_passes;                   // the same for the feature head (3 or 1): its result leaves as fp16 anyway
    // training tape (TRAIN instantiation only; see SdfTape in mlp_tc.cuh)
    uint8_t* tape_tiles;               // per tile: TAPE_TILE_BYTES (softplus' of the 8 layers, reverse adjoints g_1..g_7, g_e)
    __half* tape_act;                  // [8][P_pad][256] fp16: a_1 .. a_8 (x ACT_SCALE), row-major
    __half* tape_u;                    // [8][P_pad][256] fp16: u_0 .. u_7 (x G_SCALE), row-major
    int64_t p_pad;
    __half* feat16; int64_t feat16_ld; // fused training step: features leave as unscaled fp16 rows [N][feat16_ld] (the first 256
                                       // columns of the reflectance operand) instead of fp32 rows; nullptr = off
};

// One gemm = `nsub` sub-chunks of 32 K-columns.  Every weight image serves exactly one sub-chunk: its 64 "K" columns are
// [W_hi(k0 .. k0+31) | W_lo(k0 .. k0+31)], so K-steps 0,1 of the image are the hi part and 2,3 the lo part.
struct Gemm { uint32_t b_off; int nsub; int n; uint32_t img_bytes; };

template <bool GRAD, bool FEAT>
__device__ __forceinline__ constexpr int num_gemms() { return SDF_LAYERS + (FEAT ? 1 : 0) + (GRAD ? SDF_LAYERS : 0); }

// gemm sequence of one tile: L0..L7, [feature head], [R7..R1, R0]
template <bool GRAD, bool FEAT>
__device__ __forceinline__ Gemm get_gemm(const TcLayout& T, int idx) {
    Gemm g;
    g.nsub = 8; g.n = 256; g.img_bytes = IMG;
    if (idx < SDF_LAYERS) {
        g.b_off = T.fwd[idx];
        if (idx == 0) g.nsub = 2;
        return g;
    }
    idx -= SDF_LAYERS;
    if (FEAT) { if (idx == 0) { g.b_off = T.feat; return g; } idx -= 1; }
    const int l = SDF_LAYERS - 1 - idx;          // 7 .. 0
    g.b_off = T.rev[l];
    if (l == 0) { g.n = 64; g.img_bytes = IMG_SMALL; }
    return g;
}

// per-thread view of the epilogue: row r of the tile (== TMEM lane), 8-column group gq of every 32-wide sub-chunk.
// The hand-off unit between the epilogue and the MMA issuer is a SUB-CHUNK of 32 activation columns (2 K-steps, 6 MMAs):
// the tensor pipe restarts ~0.5K clk after an accumulator completes and idles only ~0.9K clk behind the last publish.
template <bool TRAIN>
struct EpiT {
    static constexpr bool kTrain = TRAIN;
    bool discard;                          // drop the softplus' scratch lines from L2 after their only read (inference kernels)
    uint8_t* A_hi; uint8_t* A_lo; uint64_t* a_ready;      // a_ready[8]: one per sub-chunk
    // training tape (TRAIN only): `dump` = row of this thread's point in the row-major fp16 [P][256] dump that receives the
    // operand published by the CURRENT epilogue (nullptr: none); `gnx` = packed reverse-sweep adjoints of this tile
    __half* dump; uint32_t* gnx;
    uint32_t* sig;   // [8][128 column pairs][128 rows] packed softplus' of every forward layer: half2 of copysign(exp(-100|x|), x)
    float* pe_s;     // [40][128] fp32 Fourier encoding (final chain)
    float* pk_s;     // [40][128] encoding * ACT_SCALE / sqrt2 (skip concat operand)
    float* ge_s;     // [40][128] skip-path gradient (G_SCALE units)
    int r, gq, lane;
    uint32_t off[2];                       // swizzled byte offset of this thread's 8 columns inside a chunk (even / odd sub-chunk)
    uint32_t s_hi, s_lo;                   // shared-space addresses of A_hi / A_lo
    long long* tl;                         // timeline slot of the current gemm (nullptr = off)
    long long* tw;                         // developer timeline, every epilogue warp: [8] publish stamps of the logged layer (nullptr = off)
    // 8 values -> sub-chunk sc of the A operand (hi/lo split), then signal the MMA issuer (one arrival per warp).
    // NOTE: fence.proxy.async lowers to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC: it drains every outstanding memory
    // operation of the thread, so callers issue their global loads / stores right AFTER publish(), never before.
    // Hand-offs are batched: the proxy fence costs every warp ~20 clk of a CTA-wide serialised resource (16 warps x 8 sub-chunks
    // = 128 fences per layer, tests/tc_probe5.cu), so only the steps in `fence_mask` fence and they signal every sub-chunk stored
    // since the previous hand-off.  Sub-chunks 0 and 7 are always handed off on their own (they bound the pipeline bubble at both
    // ends of a layer).
    uint32_t fence_mask;
    bool rev_lo;                           // reverse-sweep operands are published as hi + lo tiles (SdfTcParams::rev_passes == 3)
    // `lo` = false: the consumer gemm reads only the hi tile of this operand (reduced-pass reverse sweep), so the lo split is skipped
    __device__ __forceinline__ void publish(int sc, const float* o, bool lo = true) const {
        const uint32_t o8 = (uint32_t)(sc >> 1) * A_CHUNK + off[sc & 1];
        uint32_t hw[4];
        if (!lo) store_half8s_hw(s_hi + o8, o, hw);
        else if (TRAIN) store_split8s_hw(s_hi + o8, s_lo + o8, o, hw);
        else store_split8s(s_hi + o8, s_lo + o8, o);
        if ((fence_mask >> sc) & 1u) {
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&a_ready[sc]);
                for (int s = sc - 1; s >= 0 && !((fence_mask >> s) & 1u); --s) mbar_arrive(&a_ready[s]);
            }
        }
        if (TRAIN) { if (dump) *reinterpret_cast<uint4*>(dump + sc * 32 + gq * 8) = make_uint4(hw[0], hw[1], hw[2], hw[3]); }
        if (tw && lane == 0) tw[sc] = clock64();
    }
};
using Epi = EpiT<false>;

__device__ __forceinline__ void ldg8(const float* p, float (&b)[8]) {
    const float4 t0 = __ldg(reinterpret_cast<const float4*>(p)), t1 = __ldg(reinterpret_cast<const float4*>(p) + 1);
    b[0] = t0.x; b[1] = t0.y; b[2] = t0.z; b[3] = t0.w; b[4] = t1.x; b[5] = t1.y; b[6] = t1.z; b[7] = t1.w;
}

// precise trig kept out of line: the range-reduction slow paths are large and would be inlined dozens of times
__device__ __noinline__ float sin_precise(float x) { return sinf(x); }
__device__ __noinline__ float cos_precise(float x) { return cosf(x); }

// forward arithmetic happens in "x16" units (everything pre-multiplied by ACT_SCALE = 16, biases included):
//   x16 = acc / W_SCALE + 16 b ;  softplus16(x16) = max(x16,0) + 16 ln2/100 * lg2(1 + 2^(-100 log2e/16 |x16|))
constexpr float OS_F16 = 1.0f / W_SCALE;
constexpr float SP_K = -144.269504f / ACT_SCALE;
constexpr float SP_L = 6.93147181e-3f * ACT_SCALE;

// All epilogues are software-pipelined over the eight 32-column sub-chunks with two alternating register sets: the
// TMEM load and the L2 loads of sub-chunk sc+1 are in flight while sc is processed.  `wait_acc` blocks until the
// accumulator of this gemm is complete and returns its TMEM address (lane base included); loads that do not
// depend on it (bias, softplus') are issued before it.

// forward layer epilogue.  LT: 0 = plain, 1 = lin3 (skip concat + 1/sqrt2), 2 = lin7 (sdf head dot).
// OUT: 0 = no operand for a next gemm, 1 = activations, 2 = reverse seed (w_s/3 * softplus' * G_SCALE)
template <bool GRAD, int LT, int OUT, class EP, class WaitAcc>
__device__ __forceinline__ void epi_forward(const EP& E, WaitAcc&& wait_acc, int l, const float* __restrict__ bias16,
                                            const float* __restrict__ head_w, float& dot16) {
    float vA[8], bA[8], vB[8], bB[8];
    float wA[LT == 2 ? 8 : 1], wB[LT == 2 ? 8 : 1], pk[LT == 1 ? 16 : 1];
    const int cq = E.gq * 8;
    ldg8(bias16 + cq, bA);
    ldg8(bias16 + cq + 32, bB);
    if (LT == 2) { ldg8(head_w + cq, reinterpret_cast<float (&)[8]>(wA)); ldg8(head_w + cq + 32, reinterpret_cast<float (&)[8]>(wB)); }
    if (LT == 1) {                                     // skip-concat operand of the last two sub-chunks (columns >= 217)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int col = 192 + (i >> 3) * 32 + cq + (i & 7);
            pk[LT == 1 ? i : 0] = col >= SKIP_H ? E.pk_s[(col - SKIP_H) * TM + E.r] : 0.f;
        }
    }
    if (E.tl) E.tl[0] = clock64();
    const uint32_t acc = wait_acc() + cq;
    if (E.tl) E.tl[1] = clock64();
    tmem_ld8(acc, vA);
    auto step = [&](float (&v)[8], float (&b)[8], float (&nv)[8], float* w, const int sc) {
        tmem_wait_ld();
        if (E.tl) E.tl[2 + sc * 3] = clock64();
        if (sc < 7) tmem_ld8(acc + (sc + 1) * 32, nv);
        float s[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float x = fmaf(v[i], OS_F16, b[i]);
            const float e = ex2_approx(SP_K * fabsf(x));
            const float t = 1.0f + e;
            if (GRAD) s[i] = copysignf(e, x);                     // packed softplus' (see dsig_from_packed)
            v[i] = fmaf(lg2_approx(t), SP_L, fmaxf(x, 0.0f));
        }
        if (LT == 1) {
            if (sc * 32 + cq + 8 <= SKIP_H) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] *= INV_SQRT2;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int col = sc * 32 + cq + i;
                    v[i] = col < SKIP_H ? v[i] * INV_SQRT2 : pk[LT == 1 ? ((sc - 6) & 1) * 8 + i : 0];
                }
            }
        }
        if (LT == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dot16 = fmaf(v[i], w[i], dot16);
        }
        if (E.tl) E.tl[3 + sc * 3] = clock64();
        if (OUT == 1) {
            E.publish(sc, v);
        } else if (OUT == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = w[i] * dsig_from_packed(s[i]) * (W_SCALE * G_SCALE / SDF_SCALE);
            E.publish(sc, v, E.rev_lo);
        }
        if (E.tl) E.tl[4 + sc * 3] = clock64();
        // global traffic goes after the fence inside publish(): softplus' stores, bias of the sub-chunk after next
        if (GRAD) pk_store(E.sig, l, sc, E.gq, E.r, pack_sig2(s[0], s[1]), pack_sig2(s[2], s[3]), pack_sig2(s[4], s[5]), pack_sig2(s[6], s[7]));
        if (sc < 6) {
            ldg8(bias16 + cq + (sc + 2) * 32, b);
            if (LT == 2) ldg8(head_w + cq + (sc + 2) * 32, *reinterpret_cast<float (*)[8]>(w));
        }
    };
    step(vA, bA, vB, wA, 0);
    step(vB, bB, vA, wB, 1);
    step(vA, bA, vB, wA, 2);
    step(vB, bB, vA, wB, 3);
    step(vA, bA, vB, wA, 4);
    step(vB, bB, vA, wB, 5);
    step(vA, bA, vB, wA, 6);
    step(vB, bB, vA, wB, 7);
}

// feature head epilogue: write feat (fp32 row-major, or the fp16 operand image of the tile) and, with GRAD, seed the reverse sweep
template <bool GRAD, class EP, class WaitAcc>
__device__ __forceinline__ void epi_feat(const EP& E, WaitAcc&& wait_acc, const float* __restrict__ bias, const float* __restrict__ head_w,
                                         float* __restrict__ feat_row, uint8_t* __restrict__ feat_tile_img, bool valid,
                                         __half* __restrict__ feat16_row = nullptr) {
    float vA[8], bA[8], vB[8], bB[8];
    const int cq = E.gq * 8;
    ldg8(bias + cq, bA);
    const uint32_t acc = wait_acc() + cq;
    tmem_ld8(acc, vA);
    auto step = [&](float (&v)[8], float (&b)[8], float (&nv)[8], float (&nb)[8], const int sc) {
        tmem_wait_ld();
        if (sc < 7) { tmem_ld8(acc + (sc + 1) * 32, nv); ldg8(bias + cq + (sc + 1) * 32, nb); }
        float w[8], s[8];
        if (GRAD) {
            ldg8(head_w + cq + sc * 32, w);
            uint32_t sw[4];
            pk_load(E.sig, SDF_LAYERS - 1, sc, E.gq, E.r, sw);
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float2 t2 = unpack_sig2(sw[i]); s[2 * i] = t2.x; s[2 * i + 1] = t2.y; }
            if (!EP::kTrain && E.discard) pk_discard(E.sig, SDF_LAYERS - 1, sc, E.gq, E.r);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], OS_F, b[i]);
        if (feat_tile_img) {                           // fp16 operand image of this tile (every row is written)
            float t16[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) t16[i] = valid ? v[i] * ACT_SCALE : 0.f;
            store_half8(feat_tile_img + (sc >> 1) * A_CHUNK, E.off[sc & 1], t16);
        } else if (feat16_row) {                       // unscaled fp16 row (same rounding as a later fp32 -> fp16 conversion pass)
            if (valid) store_half8(reinterpret_cast<uint8_t*>(feat16_row + cq + sc * 32), 0u, v);
        } else if (valid) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
                *reinterpret_cast<float4*>(feat_row + cq + sc * 32 + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        if (GRAD) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = w[i] * dsig_from_packed(s[i]) * (W_SCALE * G_SCALE / SDF_SCALE);
            E.publish(sc, v, E.rev_lo);
        }
    };
    step(vA, bA, vB, bB, 0);
    step(vB, bB, vA, bA, 1);
    step(vA, bA, vB, bB, 2);
    step(vB, bB, vA, bA, 3);
    step(vA, bA, vB, bB, 4);
    step(vB, bB, vA, bA, 5);
    step(vA, bA, vB, bB, 6);
    step(vB, bB, vA, bA, 7);
}

// reverse layer epilogue (l = 7..1): g_pre_{l-1} = (W_l^T g_pre_l) * softplus'_{l-1}; SKIP = (l == 4).
// sig holds softplus' / W_SCALE, so acc * sig is already in G_SCALE units.
template <bool SKIP, bool TRAIN, class EP, class WaitAcc>
__device__ __forceinline__ void epi_reverse(const EP& E, WaitAcc&& wait_acc, int l) {
    const int cq = E.gq * 8;
    float vA[8], vB[8];
    uint32_t sA[4], sB[4];
    auto load_sig = [&](uint32_t (&sg)[4], const int sc) {
        pk_load(E.sig, l - 1, sc, E.gq, E.r, sg);
        if (SKIP) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (sc * 32 + cq + 2 * i >= SKIP_H) sg[i] = 0u;
        }
    };
    load_sig(sA, 0);
    load_sig(sB, 1);
    const uint32_t acc = wait_acc() + cq;
    tmem_ld8(acc, vA);
    auto step = [&](float (&v)[8], uint32_t (&sg)[4], float (&nv)[8], const int sc) {
        tmem_wait_ld();
        if (sc < 7) tmem_ld8(acc + (sc + 1) * 32, nv);
        float ge[8];
        uint32_t gpk[TRAIN ? 4 : 1];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 pk = unpack_sig2(sg[i]);
            float graw[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = 2 * i + h, col = sc * 32 + cq + j;
                const float p1 = h ? pk.y : pk.x;
                // adjoint of a_l BEFORE the softplus' gate, G_SCALE units (tape: second-order term of the backward)
                graw[h] = (SKIP && col >= SKIP_H) ? 0.f : v[j] * (SKIP ? OS_R * INV_SQRT2 : OS_R);
                if (!SKIP) v[j] *= dsig_from_packed(p1);
                else if (col >= SKIP_H) { ge[j] = v[j] * (OS_R * INV_SQRT2); v[j] = 0.f; }
                else v[j] = v[j] * INV_SQRT2 * dsig_from_packed(p1);
            }
            if (TRAIN) gpk[TRAIN ? i : 0] = pack_sig2(graw[0], graw[1]);
        }
        E.publish(sc, v, E.rev_lo);
        if (!TRAIN && E.discard) pk_discard(E.sig, l - 1, sc, E.gq, E.r);       // sg of this step has been consumed above
        if (TRAIN) pk_store(E.gnx, l - 1, sc, E.gq, E.r, gpk[0], gpk[TRAIN ? 1 : 0], gpk[TRAIN ? 2 : 0], gpk[TRAIN ? 3 : 0]);
        if (SKIP) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int col = sc * 32 + cq + i;
                if (col >= SKIP_H) E.ge_s[(col - SKIP_H) * TM + E.r] = ge[i];
            }
        }
        if (sc < 6) load_sig(sg, sc + 2);
    };
    step(vA, sA, vB, 0);
    step(vB, sB, vA, 1);
    step(vA, sA, vB, 2);
    step(vB, sB, vA, 3);
    step(vA, sA, vB, 4);
    step(vB, sB, vA, 5);
    step(vA, sA, vB, 6);
    step(vB, sB, vA, 7);
}

template <bool GRAD, bool FEAT, bool TRAIN = false>
__global__ void __launch_bounds__(NTHREADS, 1)
sdf_tc_kernel(SdfTcParams P, Strided3 pts, int64_t N, float* __restrict__ sdf_out, float* __restrict__ gx,
              float* __restrict__ gy, float* __restrict__ gz, int64_t gstride, float* __restrict__ feat_out,
              float* __restrict__ scratch) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    uint8_t* A_hi = smem + SM_A_HI;
    uint8_t* A_lo = smem + SM_A_LO;
    uint8_t* Bst = smem + SM_B;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_MISC);
    uint64_t* b_full = bars;                  // [NSTAGES]
    uint64_t* b_empty = bars + NSTAGES;       // [NSTAGES]
    uint64_t* a_ready = bars + 2 * NSTAGES;   // [8] one per 32-column sub-chunk of the A operand
    uint64_t* acc_full = a_ready + 8;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 2);
    uint32_t* ready = tmem_slot + 1;                             // number of sub-chunks whose operands are in place (scout -> issuer)
    float* part = reinterpret_cast<float*>(bars + 32);          // [3][128]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NG = num_gemms<GRAD, FEAT>();
    const TcLayout T = tc_layout();

    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (tid == 32) {
        *ready = 0;
        for (int i = 0; i < NSTAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 8; ++i) mbar_init(&a_ready[i], EPI_WARPS);
        for (int i = 0; i < 2; ++i) mbar_init(&acc_full[i], 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t ntiles = (N + TM - 1) / TM;
    const int64_t tile0 = blockIdx.x, tstride = gridDim.x, tiles_padded = ntiles;

    if (warp == 0) {
        // ======================= weight producer =======================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = tile0; tile < tiles_padded; tile += tstride)
#pragma unroll 1
                for (int gi = 0; gi < NG; ++gi) {
                    const Gemm G = get_gemm<GRAD, FEAT>(T, gi);
                    for (int img = 0; img < G.nsub; ++img, ++it) {
                        const uint32_t s = it % NSTAGES, u = it / NSTAGES;
                        mbar_wait(&b_empty[s], (u & 1) ^ 1);
                        if (NRH_DBG(P) == 3 || NRH_DBG(P) == 4) { mbar_arrive(&b_full[s]); continue; }      // ablation: no weight traffic (ncta == 1 only)
                        mbar_arrive_expect_tx(&b_full[s], G.img_bytes);
                        bulk_g2s(Bst + s * STAGE, P.tc + G.b_off + (size_t)img * G.img_bytes, G.img_bytes, &b_full[s]);
                    }
                }
        }
    } else if (warp == 2) {
        // ======================= readiness scout =======================
        // Every mbarrier wait executed by the MMA-issuing thread is tensor-pipe idle time (a try_wait that succeeds at once
        // still costs ~280 clk there, a plain shared-memory load ~70: tests/tc_probe3.cu).  This thread does the waiting
        // instead -- operand sub-chunk published by the epilogue, weight image landed -- strictly in issue order, and
        // publishes one monotonic counter that the issuer polls with an ordinary load, usually once per several sub-chunks.
        if (lane == 0) {
            uint32_t it = 0, a_par = 0;
            const uint32_t ready_s = smem_u32(ready);
            for (int64_t tile = tile0; tile < tiles_padded; tile += tstride)
#pragma unroll 1
                for (int gi = 0; gi < NG; ++gi) {
                    const int nsub = get_gemm<GRAD, FEAT>(T, gi).nsub;
                    for (int sc = 0; sc < nsub; ++sc, ++it) {
                        mbar_wait(&a_ready[sc], (a_par >> sc) & 1);
                        a_par ^= (1u << sc);
                        mbar_wait(&b_full[it % NSTAGES], (it / NSTAGES) & 1);
                        asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(ready_s), "r"(it + 1) : "memory");
                    }
                }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        {   // the whole warp runs this loop in lockstep (uniform operands); one elected lane issues inside the *_w primitives:
            // a single lane of a diverged warp pays an elect + R2UR.BROADCAST waterfall loop (12 SASS instructions) per MMA
            uint32_t it = 0, gc = 0, ready_seen = 0;
            const uint32_t a_hi_lo = desc_lo(smem_u32(A_hi)), a_lo_lo = desc_lo(smem_u32(A_lo)), b_lo0 = desc_lo(smem_u32(Bst));
            const uint32_t ready_s = smem_u32(ready);
            const bool mma_on = (NRH_DBG(P) != 2 && NRH_DBG(P) != 4);
            for (int64_t tile = tile0; tile < tiles_padded; tile += tstride)
#pragma unroll 1
                for (int gi = 0; gi < NG; ++gi, ++gc) {
                    const Gemm G = get_gemm<GRAD, FEAT>(T, gi);
                    const uint32_t acc = tmem_base + (gc & 1) * 256;
                    const uint32_t idesc = make_idesc_f16(TM, G.n);
                    // forward layers: always the 3-MMA split; feature head / reverse sweep: SdfTcParams::feat_passes / rev_passes
                    const int passes = gi < SDF_LAYERS ? 3 : ((FEAT && gi == SDF_LAYERS) ? P.feat_passes : P.rev_passes);
                    const int lgi = (NRH_DBG(P) == 9) ? gi - 8 : gi;
                    const bool lg = lane == 0 && NRH_TLOG(P) && blockIdx.x == 0 && tile == tile0 + 2 * tstride && lgi >= 0 && lgi < 8;
                    for (int sc = 0; sc < G.nsub; ++sc, ++it) {
                        if (lg) NRH_TLOG(P)[lgi * 32 + sc * 3 + 0] = clock64();
                        if (ready_seen <= it) {
                            uint32_t spins = 0;
                            while (true) {
                                asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(ready_seen) : "r"(ready_s) : "memory");
                                if (ready_seen > it) break;
                                if (++spins > (1u << 27)) asm volatile("trap;");
                            }
                            __syncwarp();
                            tc_fence_after();
                        }
                        if (lg) NRH_TLOG(P)[lgi * 32 + sc * 3 + 1] = clock64();
                        const uint32_t s = it % NSTAGES;
                        if (mma_on) {
                            // A: K-steps 2 sc, 2 sc + 1 of the 64-wide chunk sc / 2 (32 B per K-step inside the swizzled rows)
                            const uint32_t ko = (uint32_t)(sc >> 1) * (A_CHUNK >> 4) + (uint32_t)(sc & 1) * 4;
                            const uint32_t ah = a_hi_lo + ko, al = a_lo_lo + ko, bl = b_lo0 + s * (STAGE >> 4);
                            if (passes == 3) umma_group6_ss_w(acc, ah, al, bl, idesc, (uint32_t)(sc != 0));  // A_hi * W_hi, A_lo * W_hi, A_hi * W_lo behind one election
                            else if (passes == 2) umma_group4_ss_w(acc, ah, bl, idesc, (uint32_t)(sc != 0));  // A_hi * W_hi, A_hi * W_lo
                            else umma_group2_ss_w(acc, ah, bl, idesc, (uint32_t)(sc != 0));                   // A_hi * W_hi
                        }
                        umma_commit_w(&b_empty[s]);
                        if (lg) NRH_TLOG(P)[lgi * 32 + sc * 3 + 2] = clock64();
                    }
                    umma_commit_w(&acc_full[gc & 1]);
                    if (lg) NRH_TLOG(P)[lgi * 32 + 24] = clock64();
                }
        }
    } else if (warp >= EPI_WARP0) {
        // ======================= epilogue warps =======================
        EpiT<TRAIN> E;
        E.A_hi = A_hi; E.A_lo = A_lo; E.a_ready = a_ready; E.lane = lane; E.tl = nullptr; E.tw = nullptr;
        E.dump = nullptr; E.gnx = nullptr; E.fence_mask = P.fmask; E.discard = NRH_DBG(P) != 7;
        E.rev_lo = P.rev_passes == 3;
        const int q = warp & 3;
        E.gq = (warp - EPI_WARP0) >> 2;
        E.r = q * 32 + lane;                                // row of the tile == TMEM lane
        const int r = E.r, gq = E.gq;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        float* const scr = scratch + (size_t)blockIdx.x * ((SDF_LAYERS * 128 + 3 * PE_PAD) * TM);
        E.sig = reinterpret_cast<uint32_t*>(scr);
        E.pe_s = scr + (size_t)SDF_LAYERS * 128 * TM;
        E.pk_s = E.pe_s + PE_PAD * TM;
        E.ge_s = E.pk_s + PE_PAD * TM;
        E.off[0] = sw128_offset(E.r, E.gq * 8);
        E.off[1] = sw128_offset(E.r, 32 + E.gq * 8);
        E.s_hi = smem_u32(A_hi); E.s_lo = smem_u32(A_lo);
        uint32_t gc = 0;
        auto wait_acc = [&]() -> uint32_t {
            mbar_wait(&acc_full[gc & 1], (gc >> 1) & 1);
            tc_fence_after();
            if (E.tw && lane == 0) E.tw[8] = clock64();
            const uint32_t a = tmem_base + lane_base + (gc & 1) * 256;
            ++gc;
            return a;
        };

        for (int64_t tile = tile0; tile < tiles_padded; tile += tstride) {
            const int64_t p = tile * TM + r;
            const bool valid = p < N;
            if (TRAIN) {                              // softplus' / adjoints of this tile go to the tape instead of the per-CTA scratch
                uint8_t* const tt = P.tape_tiles + (size_t)tile * TAPE_TILE_BYTES;
                E.sig = reinterpret_cast<uint32_t*>(tt);
                E.gnx = reinterpret_cast<uint32_t*>(tt + TAPE_SIG_BYTES);
            }
            // ---------------- Fourier encoding -> A chunk 0 (hi/lo) + fp32 copy ----------------
            {
                float x[3];
                x[0] = valid ? pts.x[p * pts.stride] * SDF_SCALE : 0.f;
                x[1] = valid ? pts.y[p * pts.stride] * SDF_SCALE : 0.f;
                x[2] = valid ? pts.z[p * pts.stride] * SDF_SCALE : 0.f;
                auto put = [&](int col, float v) {
                    put_split1(A_hi, A_lo, r, col, v * ACT_SCALE);
                    E.pe_s[col * TM + r] = v;
                    E.pk_s[col * TM + r] = v * (ACT_SCALE * INV_SQRT2);
                };
                auto put_sin = [&](int d) {
                    float f = 1.0f;
#pragma unroll
                    for (int k = 0; k < SDF_FREQ; ++k) { put(3 + d * SDF_FREQ + k, sin_precise(x[d] * f)); f *= 2.0f; }
                };
                auto put_cos = [&](int d) {
                    float f = 1.0f;
#pragma unroll
                    for (int k = 0; k < SDF_FREQ; ++k) {
                        put(3 + 3 * SDF_FREQ + d * SDF_FREQ + k, sin_precise(x[d] * f + 1.57079637050628662109375f)); f *= 2.0f;
                    }
                };
                if (gq == 0) { put(0, x[0]); put(1, x[1]); put(2, x[2]); put_sin(0); }
                else if (gq == 1) { put_cos(0); put_sin(1); }
                else if (gq == 2) { put_cos(1); put_sin(2); }
                else {
                    put_cos(2);
                    put_split1(A_hi, A_lo, r, 39, 0.f);
                    E.pe_s[39 * TM + r] = 0.f;
                    E.pk_s[39 * TM + r] = 0.f;
                    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int k8 = 40; k8 < 64; k8 += 8) {
                        *reinterpret_cast<uint4*>(A_hi + sw128_offset(r, k8)) = z;
                        *reinterpret_cast<uint4*>(A_lo + sw128_offset(r, k8)) = z;
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&a_ready[0]); mbar_arrive(&a_ready[1]); }
                epi_bar_sync();                       // pe_s visible to every epilogue thread
            }

            // ---------------- forward layers ----------------
            float dot = 0.f;
#pragma unroll 1
            for (int l = 0; l < SDF_LAYERS - 1; ++l) {
                E.tl = (NRH_TLOG(P) && blockIdx.x == 0 && tile == tile0 + 2 * tstride && warp == EPI_WARP0 + 2 && lane == 0) ? NRH_TLOG(P) + 256 + l * 32 : nullptr;
                E.tw = (NRH_TLOG(P) && blockIdx.x == 0 && tile == tile0 + 2 * tstride && l == 2) ? NRH_TLOG(P) + 512 + (warp - EPI_WARP0) * 16 : nullptr;
                if (TRAIN) E.dump = P.tape_act + ((size_t)l * P.p_pad + p) * 256;            // this epilogue publishes a_{l+1}
                if (l == SDF_SKIP - 1) epi_forward<GRAD, 1, 1>(E, wait_acc, l, P.bias16 + l * 256, P.head_w, dot);
                else epi_forward<GRAD, 0, 1>(E, wait_acc, l, P.bias16 + l * 256, P.head_w, dot);
                tc_fence_before();
            }
            {
                constexpr int OUT = FEAT ? 1 : (GRAD ? 2 : 0);
                E.tl = nullptr; E.tw = nullptr;
                if (TRAIN) E.dump = P.tape_act + ((size_t)(SDF_LAYERS - 1) * P.p_pad + p) * 256;   // a_8
                epi_forward<GRAD, 2, OUT>(E, wait_acc, SDF_LAYERS - 1, P.bias16 + (SDF_LAYERS - 1) * 256, P.head_w, dot);
                tc_fence_before();
                // sdf head: combine the four column quarters
                if (gq > 0) part[(gq - 1) * TM + r] = dot;
                epi_bar_sync();
                if (gq == 0 && valid) sdf_out[p] = (((dot + part[r]) + (part[TM + r] + part[2 * TM + r])) * (1.0f / ACT_SCALE) + __ldg(P.head_b)) / SDF_SCALE;
            }
            if (FEAT) {
                if (TRAIN) E.dump = P.tape_u + ((size_t)(SDF_LAYERS - 1) * P.p_pad + p) * 256;    // reverse seed u_7
                epi_feat<GRAD>(E, wait_acc, P.feat_b, P.head_w, feat_out + (valid ? p : 0) * 256,
                               P.feat_image ? reinterpret_cast<uint8_t*>(feat_out) + (size_t)tile * TC_TILE_FEAT_BYTES : nullptr, valid,
                               (TRAIN && P.feat16) ? P.feat16 + (valid ? p : 0) * P.feat16_ld : nullptr);
                tc_fence_before();
            }
            if (GRAD) {
#pragma unroll 1
                for (int l = SDF_LAYERS - 1; l >= 1; --l) {
                    E.tw = (NRH_TLOG(P) && blockIdx.x == 0 && tile == tile0 + 2 * tstride && l == 5) ? NRH_TLOG(P) + 768 + (warp - EPI_WARP0) * 16 : nullptr;
                    if (TRAIN) E.dump = P.tape_u + ((size_t)(l - 1) * P.p_pad + p) * 256;           // publishes u_{l-1}
                    if (l == SDF_SKIP) epi_reverse<true, TRAIN>(E, wait_acc, l);
                    else epi_reverse<false, TRAIN>(E, wait_acc, l);
                    tc_fence_before();
                }
                // ---- reverse layer 0 + chain through the encoding (39 columns; quarter 0 warps) ----
                E.tw = nullptr;
                const uint32_t acc = wait_acc();
                epi_bar_sync();                       // ge_s written by other threads in the R4 epilogue
                if (gq == 0) {
                    float v0[16], v1[16], v2[16];
                    tmem_ld16(acc, v0);
                    tmem_ld16(acc + 16, v1);
                    tmem_ld16(acc + 32, v2);
                    tmem_wait_ld();
                    auto gcol = [&](int j) -> float {
                        const float a = j < 16 ? v0[j] : (j < 32 ? v1[j - 16] : v2[j - 32]);
                        return fmaf(a, OS_R, E.ge_s[j * TM + r]);
                    };
                    if (TRAIN) {                      // adjoint of the encoding (G_SCALE units): second derivative term of the backward
                        float* const ge_t = reinterpret_cast<float*>(P.tape_tiles + (size_t)tile * TAPE_TILE_BYTES + TAPE_SIG_BYTES + TAPE_G_BYTES);
#pragma unroll
                        for (int j = 0; j < PE_DIM; ++j) ge_t[j * TM + r] = gcol(j);
                    }
                    float gsum[3];
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const float x = E.pe_s[d * TM + r];
                        float acc_d = gcol(d);
                        float f = 1.0f;
#pragma unroll
                        for (int k = 0; k < SDF_FREQ; ++k) {
                            const float sarg = x * f;
                            acc_d += gcol(3 + d * SDF_FREQ + k) * cos_precise(sarg) * f;
                            acc_d += gcol(3 + 3 * SDF_FREQ + d * SDF_FREQ + k) * cos_precise(sarg + 1.57079637050628662109375f) * f;
                            f *= 2.0f;
                        }
                        gsum[d] = acc_d;
                    }
                    if (valid) {
                        gx[p * gstride] = gsum[0] * (SDF_SCALE / G_SCALE);
                        gy[p * gstride] = gsum[1] * (SDF_SCALE / G_SCALE);
                        gz[p * gstride] = gsum[2] * (SDF_SCALE / G_SCALE);
                    }
                }
                tc_fence_before();
            }
            epi_bar_sync();                           // scratch / part reuse across tiles
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

#include "mlp_tc_bwd.inc"
#include "mlp_tc2.inc"

// ===============================================================================================================
// Reflectance network (single fp16 pass)
// ===============================================================================================================
struct ColTcParams {
    const uint8_t* tc;
    const float* bias16;               // [4][256] hidden-layer biases * ACT_SCALE
    const float* w4t; const float* b4;
    const uint8_t* feat_img;           // streamed-input path: per-tile feature operand images (else nullptr)
    const uint8_t* aux_img;            //                      per-ray-block operand images of the per-ray inputs
    long long* tlog;                   // developer timeline (NRH_TC_TLOG)
};

__global__ void __launch_bounds__(NTHREADS, 1)
color_tc_kernel(ColTcParams P, Strided3 pts, Strided3 nrm, const float* __restrict__ feat, const float* __restrict__ rayfeat,
                int64_t R, int64_t N, float* __restrict__ cr, float* __restrict__ cg, float* __restrict__ cb) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    uint8_t* A = smem + SMC_A;
    uint8_t* Bst = smem + SMC_B;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMC_MISC);
    uint64_t* b_full = bars;                  // [12]
    uint64_t* b_empty = bars + NSTAGES;       // [12]
    uint64_t* a_ready = bars + 2 * NSTAGES;   // [6]
    uint64_t* acc_full = a_ready + 6;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 2);
    uint64_t* in_full = acc_full + 4;         // streamed inputs of a tile have landed (TMA complete_tx)
    uint64_t* a_free = acc_full + 5;          // the MMAs of the tile's last layer are done: A may be refilled
    float* part = reinterpret_cast<float*>(bars + 40);          // [3 channels][3 quarters][128]
    float* w4s = part + 9 * TM;                                 // [256][4] output layer weights, [4] bias
    const bool streamed = P.aux_img != nullptr;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const TcLayout T = tc_layout();
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (tid == 32) {
        for (int i = 0; i < NSTAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 6; ++i) mbar_init(&a_ready[i], EPI_WARPS);
        for (int i = 0; i < 2; ++i) mbar_init(&acc_full[i], 1);
        mbar_init(in_full, 1); mbar_init(a_free, 1);
        fence_mbar_init();
    }
    for (int i = tid; i < 1024 + 4; i += NTHREADS) w4s[i] = i < 1024 ? __ldg(P.w4t + i) : __ldg(P.b4 + (i - 1024));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t ntiles = (N + TM - 1) / TM;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0, ti = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
                if (streamed) {                    // the tile's A operand for layer 0 comes straight from HBM / L2
                    mbar_wait(a_free, (ti & 1) ^ 1);
                    mbar_arrive_expect_tx(in_full, 6 * A_CHUNK);
                    const uint8_t* fsrc = P.feat_img + (size_t)tile * TC_TILE_FEAT_BYTES;
                    const uint8_t* asrc = P.aux_img + (size_t)(((tile * TM) % R) / TM) * TC_TILE_AUX_BYTES;
                    for (int c = 0; c < 4; ++c) bulk_g2s(A + c * A_CHUNK, fsrc + c * A_CHUNK, A_CHUNK, in_full);
                    for (int c = 0; c < 2; ++c) bulk_g2s(A + (4 + c) * A_CHUNK, asrc + c * A_CHUNK, A_CHUNK, in_full);
                }
                for (int gi = 0; gi < 4; ++gi) {
                    const int nimg = gi == 0 ? 6 : 4;
                    for (int img = 0; img < nimg; ++img, ++it) {
                        const uint32_t s = it % NSTAGES, u = it / NSTAGES;
                        mbar_wait(&b_empty[s], (u & 1) ^ 1);
                        mbar_arrive_expect_tx(&b_full[s], STAGE);
                        bulk_g2s(Bst + s * STAGE, P.tc + T.col[gi] + (size_t)img * IMG, STAGE, &b_full[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, a_par = 0, gc = 0;
            const uint32_t a_lo0 = desc_lo(smem_u32(A)), b_lo0 = desc_lo(smem_u32(Bst));
            const uint32_t idesc = make_idesc_f16(TM, 256);
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int gi = 0; gi < 4; ++gi, ++gc) {
                    const uint32_t acc = tmem_base + (gc & 1) * 256;
                    const int nch = gi == 0 ? 6 : 4;
                    const bool lg = NRH_TLOG(P) && blockIdx.x == 0 && tile == blockIdx.x + 2 * (int64_t)gridDim.x;
                    for (int c = 0; c < nch; ++c) {
                        if (lg && c == 0) NRH_TLOG(P)[gi * 8 + 0] = clock64();
                        mbar_wait(&a_ready[c], (a_par >> c) & 1);
                        if (lg && c == 0) NRH_TLOG(P)[gi * 8 + 1] = clock64();
                        a_par ^= (1u << c);
                        tc_fence_after();
                        const uint32_t al = a_lo0 + c * (A_CHUNK >> 4);
                        {
                            const uint32_t s = it % NSTAGES, u = it / NSTAGES; ++it;
                            mbar_wait(&b_full[s], u & 1);
                            tc_fence_after();
                            const uint32_t bl = b_lo0 + s * (STAGE >> 4), d = acc;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) umma_f16_lo(d, al + ks * 2, bl + ks * 2, idesc, (uint32_t)((c | ks) != 0));
                            umma_commit(&b_empty[s]);
                        }
                    }
                    umma_commit(&acc_full[gc & 1]);
                    if (gi == 3) umma_commit(a_free);
                    if (lg) NRH_TLOG(P)[gi * 8 + 2] = clock64();
                }
        }
    } else if (warp >= EPI_WARP0) {
        const int q = warp & 3, gq = (warp - EPI_WARP0) >> 2;
        const int r = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const uint32_t off0 = sw128_offset(r, gq * 16), off1 = sw128_offset(r, gq * 16 + 8);
        const uint32_t a_s = smem_u32(A);
        uint32_t gc = 0, tcount = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t p = tile * TM + r;
            const bool valid = p < N;
            const bool elg = NRH_TLOG(P) && blockIdx.x == 0 && tile == blockIdx.x + 2 * (int64_t)gridDim.x && warp == EPI_WARP0 + 2 && lane == 0;
            if (elg) NRH_TLOG(P)[64] = clock64();
            if (streamed) {
                // features and per-ray inputs arrive as ready-made operand images; only the six per-point columns
                // (position, normal) of chunk 4 are patched in
                mbar_wait(in_full, tcount & 1);
                ++tcount;
                if (gq < 2) {
                    const Strided3& src = gq == 0 ? pts : nrm;
                    const int col = gq == 0 ? 0 : AUX_NORMAL;
                    const float v3[3] = {src.x[p * src.stride], src.y[p * src.stride], src.z[p * src.stride]};
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                        *reinterpret_cast<__half*>(A + 4 * A_CHUNK + sw128_offset(r, col + i)) = __float2half_rn(v3[i] * ACT_SCALE);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int c = 0; c < 6; ++c) mbar_arrive(&a_ready[c]);
                }
            } else
            // ---- stage the inputs: features -> chunks 0..3, [pts | PE(view) | n | PE(light) | PE(vis) | PE(spec)] -> chunks 4,5 ----
            {
                const float* frow = feat + p * 256 + gq * 64;
#pragma unroll 2
                for (int k8 = 0; k8 < 64; k8 += 8) {
                    float o[8];
                    if (valid) {
                        const float4 a = __ldg(reinterpret_cast<const float4*>(frow + k8));
                        const float4 b = __ldg(reinterpret_cast<const float4*>(frow + k8 + 4));
                        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] = 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] *= ACT_SCALE;
                    store_half8(A + gq * A_CHUNK, sw128_offset(r, k8), o);
                }
                const int64_t ray = valid ? (p % R) : 0;
#pragma unroll 1
                for (int k8 = 0; k8 < 32; k8 += 8) {
                    float o[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int j = gq * 32 + k8 + i;                 // aux row 0..127
                        float v = 0.f;
                        if (valid) {
                            if (j < 3) v = (j == 0 ? pts.x : (j == 1 ? pts.y : pts.z))[p * pts.stride];
                            else if (j < AUX_NORMAL) v = rayfeat[(int64_t)(j - AUX_VIEW) * R + ray];
                            else if (j < AUX_LIGHT) { const int d = j - AUX_NORMAL; v = (d == 0 ? nrm.x : (d == 1 ? nrm.y : nrm.z))[p * nrm.stride]; }
                            else if (j < 105) v = rayfeat[(int64_t)(j - AUX_LIGHT + COL_PE3) * R + ray];
                        }
                        o[i] = v * ACT_SCALE;
                    }
                    store_half8(A + (4 + (gq >> 1)) * A_CHUNK, sw128_offset(r, (gq & 1) * 32 + k8), o);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int c = 0; c < 6; ++c) mbar_arrive(&a_ready[c]);
                }
            }
            if (elg) NRH_TLOG(P)[65] = clock64();
            float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll 1
            for (int gi = 0; gi < 4; ++gi, ++gc) {
                const float* bias16 = P.bias16 + gi * 256 + gq * 16;
                float vA[16], bA[16], vB[16], bB[16];
                ldg16(bias16, bA);
                ldg16(bias16 + 64, bB);
                mbar_wait(&acc_full[gc & 1], (gc >> 1) & 1);
                if (elg) NRH_TLOG(P)[66 + gi * 2] = clock64();
                tc_fence_after();
                const uint32_t acc = tmem_base + lane_base + (gc & 1) * 256 + gq * 16;
                tmem_ld16(acc, vA);
                const bool last = gi == 3;
                auto step = [&](float (&v)[16], float (&b)[16], float (&nv)[16], const int c) {
                    tmem_wait_ld();
                    if (c < 3) tmem_ld16(acc + (c + 1) * 64, nv);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaxf(fmaf(v[i], OS_F16, b[i]), 0.f);      // ReLU, x16 units
                    if (last) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float4 w = *reinterpret_cast<const float4*>(w4s + (c * 64 + gq * 16 + i) * 4);
                            d0 = fmaf(v[i], w.x, d0); d1 = fmaf(v[i], w.y, d1); d2 = fmaf(v[i], w.z, d2);
                        }
                    } else {
                        store_half8s(a_s + c * A_CHUNK + off0, v);
                        store_half8s(a_s + c * A_CHUNK + off1, v + 8);
                        fence_proxy_async_smem();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&a_ready[c]);
                    }
                    if (c < 2) ldg16(bias16 + (c + 2) * 64, b);      // after the fence (it drains outstanding loads)
                };
                step(vA, bA, vB, 0);
                step(vB, bB, vA, 1);
                step(vA, bA, vB, 2);
                step(vB, bB, vA, 3);
                if (elg) NRH_TLOG(P)[67 + gi * 2] = clock64();
                tc_fence_before();
            }
            d0 *= (1.0f / ACT_SCALE); d1 *= (1.0f / ACT_SCALE); d2 *= (1.0f / ACT_SCALE);
            if (gq > 0) { float* pp = part + (gq - 1) * TM + r; pp[0] = d0; pp[3 * TM] = d1; pp[6 * TM] = d2; }
            epi_bar_sync();
            if (gq == 0 && valid) {
                const float s0 = ((d0 + part[r]) + (part[TM + r] + part[2 * TM + r])) + w4s[1024];
                const float s1 = ((d1 + part[3 * TM + r]) + (part[4 * TM + r] + part[5 * TM + r])) + w4s[1025];
                const float s2 = ((d2 + part[6 * TM + r]) + (part[7 * TM + r] + part[8 * TM + r])) + w4s[1026];
                cr[p] = 1.0f / (1.0f + expf(-s0));
                cg[p] = 1.0f / (1.0f + expf(-s1));
                cb[p] = 1.0f / (1.0f + expf(-s2));
            }
            epi_bar_sync();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

#include "color_train_tc.inc"

// ===============================================================================================================
// operand-image builders (weight packing)
// ===============================================================================================================
struct Seg { int dst_k0, src_col0, len, lo; };      // lo: 0 = fp16(x), 1 = the residual x - fp16(x)
struct ImgJob {
    const float* src; int src_ld;        // element (n, kcol) = src[n*src_ld + kcol]   (tr: src[kcol*src_ld + n])
    int tr;
    int row_lo, row_hi;                  // the job writes image rows [row_lo, row_hi) only (source row = n - row_lo); default: all
    int nrows_valid;                     // rows beyond are zero
    int nrows_img;                       // 256 or 64
    Seg seg[4]; int nseg;                // k-range mapping inside this 64-wide chunk
    int lo;                              // 1: every segment stores the residual part
    float scale;
    __half* dst;
};

// every operand image of a weight set in ONE launch (blockIdx.y = job); the job table travels as a kernel parameter (20 KB; sm_70+
// accept 32 KB of parameters since CUDA 12.1), so nothing is staged through device memory and the stream is never synchronised.
// The packing pass of a training step (weights change every step) was 156 launches of the single-image kernel.
struct JobTable { ImgJob j[TC_MAX_JOBS]; };
static_assert(sizeof(JobTable) <= 32000, "job table exceeds the kernel parameter space");
__global__ void k_build_images(const __grid_constant__ JobTable Tb) {
    const ImgJob& J = Tb.j[blockIdx.y];
    const int total = J.nrows_img * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int nrow = i >> 6, k = i & 63;
        if (nrow < J.row_lo || nrow >= J.row_hi) continue;
        const int n = nrow - J.row_lo;
        float v = 0.f;
        int lo = J.lo;
        if (n < J.nrows_valid)
            for (int s = 0; s < J.nseg; ++s)
                if (k >= J.seg[s].dst_k0 && k < J.seg[s].dst_k0 + J.seg[s].len) {
                    const int kc = J.seg[s].src_col0 + (k - J.seg[s].dst_k0);
                    v = (J.tr ? J.src[(size_t)kc * J.src_ld + n] : J.src[(size_t)n * J.src_ld + kc]) * J.scale;
                    lo |= J.seg[s].lo;
                }
        const __half h = __float2half_rn(v);
        const __half out = lo ? __float2half_rn(v - __half2float(h)) : h;
        J.dst[sw128_offset(nrow, k) >> 1] = out;
    }
}

// jobs are collected on the host and flushed by tc_pack in one launch
struct JobList { JobTable t; int n; int tr; int row_lo, row_hi; };   // tr: jobs queued from now on read their source transposed;
                                                                       // row_lo / row_hi: ... and write that row range only (0, 0 = all)
thread_local JobList* g_jobs = nullptr;

int build_image(const float* src, int src_ld, int nrows_valid, int nrows_img, const Seg* segs, int nseg, int lo,
                uint8_t* dst, cudaStream_t st) {
    (void)st;
    if (!g_jobs || g_jobs->n >= TC_MAX_JOBS) { set_error("image job table overflow"); return NRH_ERR_INVALID; }
    ImgJob& J = g_jobs->t.j[g_jobs->n++];
    J.src = src; J.src_ld = src_ld; J.tr = g_jobs->tr; J.nrows_valid = nrows_valid; J.nrows_img = nrows_img; J.nseg = nseg;
    J.row_lo = g_jobs->row_lo; J.row_hi = g_jobs->row_hi > 0 ? g_jobs->row_hi : nrows_img;
    for (int i = 0; i < 4; ++i) J.seg[i] = i < nseg ? segs[i] : Seg{0, 0, 0, 0};
    J.lo = lo; J.scale = W_SCALE; J.dst = reinterpret_cast<__half*>(dst);
    return NRH_OK;
}

int flush_image_jobs(cudaStream_t st) {
    if (!g_jobs || g_jobs->n == 0) return NRH_OK;
    k_build_images<<<dim3(64, (unsigned)g_jobs->n), 256, 0, st>>>(g_jobs->t);
    NRH_LAUNCH_CHECK();
    g_jobs->n = 0;
    return NRH_OK;
}

// a plain K-major matrix [nrows_valid x kcols] -> `nchunks` chunks of 64 K-columns.  split: two images per chunk, one per
// 32-column sub-chunk, each [hi(32) | lo(32)]; else one hi-only image per chunk.
int build_matrix(const float* src, int src_ld, int nrows_valid, int nrows_img, int kcols, int nchunks, bool split,
                 uint8_t* dst, uint32_t img_bytes, cudaStream_t st) {
    int rc;
    for (int c = 0; c < nchunks; ++c) {
        int len = kcols - c * 64; if (len > 64) len = 64; if (len < 0) len = 0;
        Seg s{0, c * 64, len, 0};
        if (split) {
            for (int h = 0; h < 2; ++h) {
                int l32 = kcols - (c * 64 + h * 32); if (l32 > 32) l32 = 32; if (l32 < 0) l32 = 0;
                Seg hl[2] = {{0, c * 64 + h * 32, l32, 0}, {32, c * 64 + h * 32, l32, 1}};
                if ((rc = build_image(src, src_ld, nrows_valid, nrows_img, hl, 2, 0, dst + (size_t)(2 * c + h) * img_bytes, st))) return rc;
            }
        } else {
            if ((rc = build_image(src, src_ld, nrows_valid, nrows_img, &s, 1, 0, dst + (size_t)c * img_bytes, st))) return rc;
        }
    }
    return NRH_OK;
}

}  // namespace

bool tc_available() { return true; }
size_t tc_packed_bytes(const NrhConfig&) { return tc_layout().total; }
size_t tc_scratch_bytes(int num_sms) { return (size_t)num_sms * ((SDF_LAYERS * 128 + 3 * PE_PAD) * TM) * sizeof(float); }

int tc_pack(const NrhConfig& cfg, const PackedLayout& L, const NrhRawWeights& raw, void* packed, cudaStream_t st) {
    uint8_t* tcb = reinterpret_cast<uint8_t*>(packed) + L.tc_offset_bytes;
    const float* Pf = reinterpret_cast<const float*>(packed);
    const TcLayout T = tc_layout();
    int rc;
    static thread_local JobList jobs;              // 20 KB: kept off the stack
    jobs.n = 0; jobs.tr = 0; jobs.row_lo = jobs.row_hi = 0;
    g_jobs = &jobs;
    struct Reset { ~Reset() { g_jobs = nullptr; } } reset_on_exit;
    // forward: B[n = out][k = in] = W native [out][in]
    for (int l = 0; l < SDF_LAYERS; ++l) {
        const int in = (l == 0) ? PE_DIM : 256, out = (l == SDF_SKIP - 1) ? SKIP_H : 256;
        if ((rc = build_matrix(raw.sdf_W[l], in, out, 256, in, l == 0 ? 1 : 4, true, tcb + T.fwd[l], IMG, st))) return rc;
    }
    if ((rc = build_matrix(raw.feat_W, 256, 256, 256, 256, 4, true, tcb + T.feat, IMG, st))) return rc;
    // reverse: B[n = in][k = out] = W^T, taken from the fp32 section's k-major copy Wt[in_pad][256]
    for (int l = 1; l < SDF_LAYERS; ++l) {
        const int out = (l == SDF_SKIP - 1) ? SKIP_H : 256;
        if ((rc = build_matrix(Pf + L.sdf_wt[l], 256, 256, 256, out, 4, true, tcb + T.rev[l], IMG, st))) return rc;
    }
    if ((rc = build_matrix(Pf + L.sdf_wt[0], 256, PE_DIM, 64, 256, 4, true, tcb + T.rev[0], IMG_SMALL, st))) return rc;
    // feature head transposed (backward): B[n = in][k = out] from the fp32 section's k-major copy
    if ((rc = build_matrix(Pf + L.feat_wt, 256, 256, 256, 256, 4, true, tcb + T.feat_rev, IMG, st))) return rc;
    // reflectance layer 0: K order = [feat 256 | pts 3, PE(view) 27, n 3, PE(light) 27, PE(vis) 9, PE(spec) 36 | pad]
    const int cin = 316 + (cfg.shadow_hint ? 9 : 0) + (cfg.specular_hint ? 9 * cfg.n_roughness : 0);
    for (int c = 0; c < 4; ++c) {
        Seg s{0, 60 + c * 64, 64};
        if ((rc = build_image(raw.col_W[0], cin, 256, 256, &s, 1, 0, tcb + T.col[0] + (size_t)c * IMG, st))) return rc;
    }
    {
        // aux rows 0..63 -> chunk 4: rows 0..59 = source columns 0..59, rows 60..63 = PE(vis)[0..3]
        Seg s4[2] = {{0, 0, 60}, {60, 316, cfg.shadow_hint ? 4 : 0}};
        if ((rc = build_image(raw.col_W[0], cin, 256, 256, s4, 2, 0, tcb + T.col[0] + (size_t)4 * IMG, st))) return rc;
        // aux rows 64..127 -> chunk 5: rows 64..68 = PE(vis)[4..8], rows 69..104 = PE(spec)
        const int spec0 = 316 + (cfg.shadow_hint ? 9 : 0);
        Seg s5[2] = {{0, 320, cfg.shadow_hint ? 5 : 0}, {5, spec0, cfg.specular_hint ? 9 * cfg.n_roughness : 0}};
        if ((rc = build_image(raw.col_W[0], cin, 256, 256, s5, 2, 0, tcb + T.col[0] + (size_t)5 * IMG, st))) return rc;
    }
    for (int l = 1; l < 4; ++l)
        if ((rc = build_matrix(raw.col_W[l], 256, 256, 256, 256, 4, false, tcb + T.col[l], IMG, st))) return rc;
    // training images of the reflectance network: layer 0 in natural input order, and W_l^T for the backward chain
    if ((rc = build_matrix(raw.col_W[0], cin, 256, 256, cin, 6, false, tcb + T.col_fwd0n, IMG, st))) return rc;
    jobs.tr = 1;                                          // B[n = in][k = out] = W[k][n]: read the native [out][in] matrix transposed
    for (int l = 1; l < 4; ++l)
        if ((rc = build_matrix(raw.col_W[l], 256, 256, 256, 256, 4, false, tcb + T.col_rev[l], IMG, st))) return rc;
    for (int c = 0; c < 4; ++c) {
        Seg s{0, c * 64, 64};
        uint8_t* dst = tcb + T.col_rev0 + (size_t)c * (IMG + IMG / 2);
        if ((rc = build_image(raw.col_W[0], cin, cin < 256 ? cin : 256, 256, &s, 1, 0, dst, st))) return rc;
        if ((rc = build_image(raw.col_W[0] + 256, cin, cin - 256, 128, &s, 1, 0, dst + IMG, st))) return rc;
    }
    // the same in the fused step's operand order: rows 0..255 = the 256 feature inputs (natural columns 60..315); the [128 x 64] image =
    // aux rows [pts, PE(view), normal, PE(light)] (natural 0..59) | PE(vis) at 60 (natural 316..) | PE(spec) at 69 | zero
    {
        const int spec0 = 316 + (cfg.shadow_hint ? 9 : 0);
        for (int c = 0; c < 4; ++c) {
            Seg s{0, c * 64, 64};
            uint8_t* dst = tcb + T.col_rev0p + (size_t)c * (IMG + IMG / 2);
            if ((rc = build_image(raw.col_W[0] + 60, cin, 256, 256, &s, 1, 0, dst, st))) return rc;
            jobs.row_lo = 0; jobs.row_hi = 60;
            if ((rc = build_image(raw.col_W[0], cin, 60, 128, &s, 1, 0, dst + IMG, st))) return rc;
            jobs.row_lo = 60; jobs.row_hi = 69;
            if ((rc = build_image(raw.col_W[0] + 316, cin, cfg.shadow_hint ? 9 : 0, 128, &s, 1, 0, dst + IMG, st))) return rc;
            jobs.row_lo = 69; jobs.row_hi = 128;
            if ((rc = build_image(raw.col_W[0] + spec0, cin, cfg.specular_hint ? 9 * cfg.n_roughness : 0, 128, &s, 1, 0, dst + IMG, st))) return rc;
            jobs.row_lo = jobs.row_hi = 0;
        }
    }
    jobs.tr = 0;
    if ((rc = flush_image_jobs(st))) return rc;
    // biases pre-multiplied by ACT_SCALE (the forward epilogues work in x16 units)
    for (int l = 0; l < SDF_LAYERS; ++l) {
        const int out = (l == SDF_SKIP - 1) ? SKIP_H : 256;
        if ((rc = queue_scaled_copy(raw.sdf_b[l], reinterpret_cast<float*>(tcb + T.sdf_bias16) + l * 256, out, ACT_SCALE, st))) return rc;
    }
    for (int l = 0; l < 4; ++l) {
        if ((rc = queue_scaled_copy(raw.col_b[l], reinterpret_cast<float*>(tcb + T.col_bias16) + l * 256, 256, ACT_SCALE, st))) return rc;
    }
    return NRH_OK;
}

// epilogue hand-off steps (bit sc = fence after sub-chunk sc; bits 0 and 7 are forced) and engine generation
static uint32_t tc_fence_mask() { return (dev_options().fmask & 0xFFu) | 0x81u; }
static int tc_generation() { return dev_options().gen; }
// developer build: dbg 10 / 11 = reverse sweep with 2 / 1 passes, 12 / 13 = the same with a single-pass feature head, 14 = 3 + 1
static int tc_rev_passes() { const int d = dev_options().dbg; return d == 10 || d == 12 ? 2 : (d == 11 || d == 13 ? 1 : (d == 14 ? 3 : TC_REV_PASSES_DEFAULT)); }
static int tc_bwd_passes() { const int d = dev_options().dbg; return d == 20 ? 2 : (d == 21 ? 1 : (d == 22 ? 3 : TC_BWD_PASSES_DEFAULT)); }   // dev: dbg 20 / 21 / 22
static int tc_feat_passes() { const int d = dev_options().dbg; return d == 12 || d == 13 || d == 14 ? 1 : (d == 10 || d == 11 ? 3 : TC_FEAT_PASSES_DEFAULT); }

int sdf_mlp_tc(const void* packed, const PackedLayout& L, Strided3 pts, int64_t N,
               float* sdf, float* gx, float* gy, float* gz, int64_t grad_stride, float* feat, bool feat_as_image,
               float* scratch, size_t scratch_bytes, int num_sms, cudaStream_t st) {
    if (N <= 0) return NRH_OK;
    if (feat_as_image && (N % TM != 0 || !feat)) { set_error("feature images need N %% 128 == 0"); return NRH_ERR_INVALID; }
    const float* Pf = reinterpret_cast<const float*>(packed);
    SdfTcParams P;
    P.feat_image = feat_as_image ? 1 : 0;
    P.feat16 = nullptr; P.feat16_ld = 0;
    P.tc = reinterpret_cast<const uint8_t*>(packed) + L.tc_offset_bytes;
    P.bias16 = reinterpret_cast<const float*>(P.tc + tc_layout().sdf_bias16);
    P.head_w = Pf + L.head_w; P.head_b = Pf + L.head_b; P.feat_b = Pf + L.feat_b;
    P.tape_tiles = nullptr; P.tape_act = nullptr; P.tape_u = nullptr; P.p_pad = 0;
    P.dbg = dev_options().dbg >= 10 ? 0 : dev_options().dbg; P.tlog = dev_options().tlog;   // codes >= 10 select pass counts (tc_rev_passes)
    P.fmask = tc_fence_mask();
    P.rev_passes = tc_rev_passes(); P.feat_passes = tc_feat_passes();
    const int64_t ntiles = (N + TM - 1) / TM;
    const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    if (scratch_bytes < tc_scratch_bytes(grid)) { set_error("sdf_mlp_tc: scratch too small"); return NRH_ERR_WORKSPACE; }
    const bool grad = gx != nullptr, wfeat = feat != nullptr;
    if (!grad && !wfeat && tc_generation() == 2) {          // second-generation engine (mlp_tc2.inc): sdf-only passes
        NRH_CUDA_CHECK(cudaFuncSetAttribute(sdf_tc2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SDF2_SMEM));
        sdf_tc2_kernel<false, false><<<grid, NTHREADS, SDF2_SMEM, st>>>(P, pts, N, sdf, scratch);
        NRH_LAUNCH_CHECK();
        return NRH_OK;
    }
    if (!grad && !wfeat && tc_generation() == 3) {          // the same with the two-team epilogue
        NRH_CUDA_CHECK(cudaFuncSetAttribute(sdf_tc2_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SDF2_SMEM));
        sdf_tc2_kernel<false, false, true><<<grid, NTHREADS, SDF2_SMEM, st>>>(P, pts, N, sdf, scratch);
        NRH_LAUNCH_CHECK();
        return NRH_OK;
    }
#define NRH_LAUNCH_TC(G, F)                                                                                        \
    do {                                                                                                           \
        NRH_CUDA_CHECK(cudaFuncSetAttribute(sdf_tc_kernel<G, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SDF_SMEM)); \
        sdf_tc_kernel<G, F><<<grid, NTHREADS, SDF_SMEM, st>>>(P, pts, N, sdf, gx, gy, gz, grad_stride, feat, scratch);           \
    } while (0)
    if (grad && wfeat) NRH_LAUNCH_TC(true, true);
    else if (grad) NRH_LAUNCH_TC(true, false);
    else if (wfeat) NRH_LAUNCH_TC(false, true);
    else NRH_LAUNCH_TC(false, false);
#undef NRH_LAUNCH_TC
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

SdfTrainLayout sdf_train_layout(int64_t N, int num_sms) {
    SdfTrainLayout t;
    const int64_t ntiles = (N + TM - 1) / TM;
    t.p_pad = ntiles * TM;
    const size_t dump = (size_t)SDF_LAYERS * t.p_pad * 256 * sizeof(__half);
    t.tape_tiles_off = 0;
    t.tape_act_off = (size_t)ntiles * TAPE_TILE_BYTES;
    t.tape_u_off = t.tape_act_off + dump;
    t.tape_bytes = t.tape_u_off + dump;
    t.bwd_gb0_off = 0;
    t.bwd_gb_off = (size_t)t.p_pad * 64 * sizeof(__half);
    t.bwd_zb_off = t.bwd_gb_off + dump;
    t.bwd_bytes = t.bwd_zb_off + dump;
    const int64_t grid = ntiles < num_sms ? ntiles : num_sms;
    t.bwd_workspace_bytes = (size_t)grid * BWD_SCRATCH_FLOATS * sizeof(float);
    return t;
}

int sdf_train_forward_tc_strided(const void* packed, const PackedLayout& L, Strided3 pts, int64_t N, float* sdf, float* gx, float* gy,
                                 float* gz, int64_t gstride, float* feat, void* tape, float* scratch, size_t scratch_bytes, int num_sms,
                                 cudaStream_t st, void* feat16, int64_t feat16_ld) {
    if (N <= 0) return NRH_OK;
    const float* Pf = reinterpret_cast<const float*>(packed);
    const SdfTrainLayout TL = sdf_train_layout(N, num_sms);
    SdfTcParams P;
    P.feat_image = 0;
    P.tc = reinterpret_cast<const uint8_t*>(packed) + L.tc_offset_bytes;
    P.bias16 = reinterpret_cast<const float*>(P.tc + tc_layout().sdf_bias16);
    P.head_w = Pf + L.head_w; P.head_b = Pf + L.head_b; P.feat_b = Pf + L.feat_b;
    P.dbg = 0; P.tlog = nullptr; P.fmask = tc_fence_mask();
    P.rev_passes = tc_rev_passes(); P.feat_passes = tc_feat_passes();
    uint8_t* tb = reinterpret_cast<uint8_t*>(tape);
    P.tape_tiles = tb + TL.tape_tiles_off;
    P.tape_act = reinterpret_cast<__half*>(tb + TL.tape_act_off);
    P.tape_u = reinterpret_cast<__half*>(tb + TL.tape_u_off);
    P.p_pad = TL.p_pad;
    P.feat16 = reinterpret_cast<__half*>(feat16); P.feat16_ld = feat16_ld;
    if (feat16 && ((reinterpret_cast<uintptr_t>(feat16) & 15) || (feat16_ld & 7))) { set_error("sdf_train_forward_tc: feat16 rows must be 16-byte aligned"); return NRH_ERR_INVALID; }
    const int64_t ntiles = (N + TM - 1) / TM;
    const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    if (scratch_bytes < tc_scratch_bytes(grid)) { set_error("sdf_train_forward_tc: scratch too small"); return NRH_ERR_WORKSPACE; }
    NRH_CUDA_CHECK(cudaFuncSetAttribute(sdf_tc_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SDF_SMEM));
    sdf_tc_kernel<true, true, true><<<grid, NTHREADS, SDF_SMEM, st>>>(P, pts, N, sdf, gx, gy, gz, gstride, feat, scratch);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int sdf_train_forward_tc(const void* packed, const PackedLayout& L, const float* pts, int64_t N, float* sdf, float* grad, float* feat,
                         void* tape, float* scratch, size_t scratch_bytes, int num_sms, cudaStream_t st) {
    return sdf_train_forward_tc_strided(packed, L, Strided3{pts, pts + 1, pts + 2, 3}, N, sdf, grad, grad + 1, grad + 2, 3, feat, tape,
                                        scratch, scratch_bytes, num_sms, st);
}

int sdf_train_backward_tc(const void* packed, const PackedLayout& L, const float* pts, int64_t N, const void* tape,
                          const float* d_sdf, const float* d_feat, const float* d_grad, const float* scale, void* bwd_out,
                          float* d_pts, float* scratch, size_t scratch_bytes, int num_sms, cudaStream_t st,
                          const SdfBwdFeat16* feat16, const Strided3* pts_strided) {
    if (N <= 0) return NRH_OK;
    const float* Pf = reinterpret_cast<const float*>(packed);
    const SdfTrainLayout TL = sdf_train_layout(N, num_sms);
    if (scratch_bytes < TL.bwd_workspace_bytes) { set_error("sdf_train_backward_tc: scratch too small"); return NRH_ERR_WORKSPACE; }
    SdfBwdParams P;
    P.tc = reinterpret_cast<const uint8_t*>(packed) + L.tc_offset_bytes;
    P.head_w = Pf + L.head_w;
    P.scale = scale;
    P.tape_tiles = reinterpret_cast<const uint8_t*>(tape) + TL.tape_tiles_off;
    P.d_sdf = d_sdf; P.d_feat = d_feat; P.d_grad = d_grad;
    P.d_feat16 = nullptr; P.d_feat16_ld = 0; P.d_feat16_mul = nullptr;
    if (feat16) { P.d_feat = nullptr; P.d_feat16 = reinterpret_cast<const __half*>(feat16->rows); P.d_feat16_ld = feat16->ld; P.d_feat16_mul = feat16->mul; }
    uint8_t* ob = reinterpret_cast<uint8_t*>(bwd_out);
    P.gb0 = reinterpret_cast<__half*>(ob + TL.bwd_gb0_off);
    P.gb = reinterpret_cast<__half*>(ob + TL.bwd_gb_off);
    P.zb = reinterpret_cast<__half*>(ob + TL.bwd_zb_off);
    P.d_pts = d_pts;
    P.p_pad = TL.p_pad; P.fmask = tc_fence_mask();
    P.passes = tc_bwd_passes();
    const int64_t ntiles = (N + TM - 1) / TM;
    const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    Strided3 S3{pts, pts + 1, pts + 2, 3};
    if (pts_strided) S3 = *pts_strided;
    NRH_CUDA_CHECK(cudaFuncSetAttribute(sdf_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SDF_SMEM));
    sdf_bwd_tc_kernel<<<grid, NTHREADS, SDF_SMEM, st>>>(P, S3, N, scratch);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int color_train_forward_tc(const void* packed, const PackedLayout& L, const void* x16, int64_t P, void* acts, float* y, int num_sms,
                           cudaStream_t st, bool permuted) {
    if (P <= 0) return NRH_OK;
    const float* Pf = reinterpret_cast<const float*>(packed);
    ColTrainParams A{};
    A.tc = reinterpret_cast<const uint8_t*>(packed) + L.tc_offset_bytes;
    A.bias16 = reinterpret_cast<const float*>(A.tc + tc_layout().col_bias16);
    A.w4t = Pf + L.col_w4t; A.b4 = Pf + L.col_b4;
    A.acts = reinterpret_cast<__half*>(acts); A.y = y; A.P = P; A.permuted = permuted ? 1 : 0;
    alignas(64) CUtensorMap xmap;
    int rc;
    if ((rc = encode_tensor_map_f16(&xmap, x16, 384, P, 64, TM))) return rc;      // boxes of [128 points x 64 columns]: one K-major operand chunk
    const int64_t ntiles = (P + TM - 1) / TM;
    const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    NRH_CUDA_CHECK(cudaFuncSetAttribute(color_train_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COLT_SMEM));
    color_train_fwd_kernel<<<grid, NTHREADS, COLT_SMEM, st>>>(xmap, A);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int color_train_backward_tc(const void* packed, const PackedLayout& L, const float* dy, const float* scale, const void* acts, int64_t P,
                            void* dz, void* dy16, void* dx, int num_sms, cudaStream_t st, bool permuted) {
    if (P <= 0) return NRH_OK;
    const float* Pf = reinterpret_cast<const float*>(packed);
    ColTrainParams A{};
    A.tc = reinterpret_cast<const uint8_t*>(packed) + L.tc_offset_bytes;
    A.w4t = Pf + L.col_w4t;
    A.acts = const_cast<__half*>(reinterpret_cast<const __half*>(acts));
    A.dy = dy; A.scale = scale;
    A.dz = reinterpret_cast<__half*>(dz); A.dy16 = reinterpret_cast<__half*>(dy16); A.dx = reinterpret_cast<__half*>(dx); A.P = P;
    A.permuted = permuted ? 1 : 0;
    const int64_t ntiles = (P + TM - 1) / TM;
    const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    NRH_CUDA_CHECK(cudaFuncSetAttribute(color_train_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COLT_SMEM));
    color_train_bwd_kernel<<<grid, NTHREADS, COLT_SMEM, st>>>(A);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int color_mlp_tc(const void* packed, const PackedLayout& L, Strided3 pts, Strided3 normals,
                 const float* feat, const float* rayfeat, const void* aux_img, int64_t R, int64_t N,
                 float* cr, float* cg, float* cb, float* scratch, size_t scratch_bytes, int num_sms, cudaStream_t st) {
    (void)scratch; (void)scratch_bytes;
    if (N <= 0) return NRH_OK;
    if (aux_img && (R % TM != 0 || N % TM != 0)) { set_error("streamed reflectance inputs need R %% 128 == 0"); return NRH_ERR_INVALID; }
    const float* Pf = reinterpret_cast<const float*>(packed);
    ColTcParams P;
    P.tc = reinterpret_cast<const uint8_t*>(packed) + L.tc_offset_bytes;
    P.bias16 = reinterpret_cast<const float*>(P.tc + tc_layout().col_bias16);
    P.tlog = dev_options().tlog ? dev_options().tlog + 256 : nullptr;
    P.w4t = Pf + L.col_w4t; P.b4 = Pf + L.col_b4;
    P.aux_img = reinterpret_cast<const uint8_t*>(aux_img);
    P.feat_img = aux_img ? reinterpret_cast<const uint8_t*>(feat) : nullptr;
    const int64_t ntiles = (N + TM - 1) / TM;
    const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    NRH_CUDA_CHECK(cudaFuncSetAttribute(color_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COL_SMEM));
    color_tc_kernel<<<grid, NTHREADS, COL_SMEM, st>>>(P, pts, normals, feat, rayfeat, R, N, cr, cg, cb);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

#ifdef NRH_DEV
void dev_configure(int gen, int dbg, unsigned fmask, void* tlog) {
    g_dev = DevOptions{gen, dbg, fmask, reinterpret_cast<long long*>(tlog)};
}
#endif

}  // namespace nrh

#ifdef NRH_DEV
// developer build only: explicit configuration of the hooks above (gen: engine generation, dbg: ablation code, fmask: hand-off
// mask, tlog: DEVICE buffer of >= 512 int64 that receives clock64 stamps, or NULL)
extern "C" int nrh_dev_configure(int gen, int dbg, unsigned fmask, void* tlog) {
    nrh::dev_configure(gen, dbg, fmask, tlog);
    return 0;
}
#endif
