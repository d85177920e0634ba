// C ABI of the nrhints_b200 CUDA library: weight packing, workspace carving, and the kernel
// pipeline that restates NeuSHintRenderer.forward (/root/reference/models/neus_hint_model.py:653-751).
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include "nrh_common.cuh"
#include "sampler_kernels.cuh"
#include "mlp_tc.cuh"
#include "train_step.cuh"

namespace nrh {

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

PackedLayout make_layout(const NrhConfig& cfg) {
    PackedLayout L;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += align_up(n, 4); return o; };
    L.inv_s = take(4);
    for (int l = 0; l < SDF_LAYERS; ++l) {
        L.sdf_wt[l] = take((size_t)(l == 0 ? PE_PAD : 256) * 256);
        L.sdf_b[l] = take(256);
    }
    for (int l = 0; l < SDF_LAYERS; ++l) {
        if (l == 0) L.sdf_wn[l] = take((size_t)256 * PE_PAD);
        else L.sdf_wn[l] = take((size_t)(l == SDF_SKIP - 1 ? 224 : 256) * 256);
    }
    L.head_w = take(256); L.head_b = take(4);
    L.feat_wt = take((size_t)256 * 256); L.feat_b = take(256);
    L.col_wt0a = take((size_t)256 * 256); L.col_wt0b = take((size_t)AUX_ROWS * 256); L.col_b0 = take(256);
    for (int l = 0; l < 3; ++l) { L.col_wt[l] = take((size_t)256 * 256); L.col_b[l] = take(256); }
    L.col_w4t = take((size_t)256 * 4); L.col_b4 = take(4);
    for (int l = 0; l < NERF_LAYERS; ++l) { L.nerf_wt[l] = 0; L.nerf_b[l] = 0; }
    L.nerf_wt5e = L.nerf_alpha_w = L.nerf_alpha_b = L.nerf_feat_wt = L.nerf_feat_b = 0;
    L.nerf_view_wta = L.nerf_view_wtb = L.nerf_view_b = L.nerf_rgb_wt = L.nerf_rgb_b = 0;
    if (cfg.use_outside_nerf) {
        for (int l = 0; l < NERF_LAYERS; ++l) {
            L.nerf_wt[l] = take((size_t)(l == 0 ? NERF_PE_PAD : 256) * 256);
            L.nerf_b[l] = take(256);
        }
        L.nerf_wt5e = take((size_t)NERF_PE_PAD * 256);
        L.nerf_alpha_w = take(256); L.nerf_alpha_b = take(4);
        L.nerf_feat_wt = take((size_t)256 * 256); L.nerf_feat_b = take(256);
        L.nerf_view_wta = take((size_t)256 * 256); L.nerf_view_wtb = take((size_t)NERF_VPE_PAD * 256); L.nerf_view_b = take(256);
        L.nerf_rgb_wt = take((size_t)NERF_VIEW_H * 4); L.nerf_rgb_b = take(4);
    }
    L.total_floats = off;
    L.tc_offset_bytes = align_up(off * sizeof(float), 1024);
    L.total_bytes = L.tc_offset_bytes + tc_packed_bytes(cfg);
    return L;
}

namespace {

// All small fp32 re-layouts of a weight set (transposed slices, padded row copies, pre-scaled biases, scalars) as ONE launch:
//   dst[a * dst_sa + b * dst_sb] = src[a * src_sa + b * src_sb] * scale      for a < d0, b < d1     (job = blockIdx.y)
struct CopyJob { const float* src; float* dst; int d0, d1, src_sa, src_sb, dst_sa, dst_sb; float scale; };
constexpr int COPY_MAX_JOBS = 96;
struct CopyTable { CopyJob j[COPY_MAX_JOBS]; };
__global__ void __launch_bounds__(256)
k_copy_jobs(const __grid_constant__ CopyTable T) {
    const CopyJob& J = T.j[blockIdx.y];
    const int total = J.d0 * J.d1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int a = i / J.d1, b = i - a * J.d1;
        J.dst[(size_t)a * J.dst_sa + (size_t)b * J.dst_sb] = J.src[(size_t)a * J.src_sa + (size_t)b * J.src_sb] * J.scale;
    }
}
struct CopyList { CopyTable t; int n; };
thread_local CopyList g_copy;
__global__ void k_inv_s(const float* __restrict__ variance, float* __restrict__ dst) {
    const float s = fminf(fmaxf(expf(variance[0] * 10.0f), 1e-6f), 1e6f);     // get_alpha :337
    dst[0] = s; dst[1] = 1.0f / s;
}
__global__ void k_copy_scalar(const float* __restrict__ src, float* __restrict__ dst) { dst[0] = src[0]; }

int flush_copy_jobs(cudaStream_t st) {
    if (g_copy.n == 0) return NRH_OK;
    int tmax = 1;
    for (int i = 0; i < g_copy.n; ++i) { const int t = g_copy.t.j[i].d0 * g_copy.t.j[i].d1; if (t > tmax) tmax = t; }
    int bx = (tmax + 255) / 256; if (bx > 64) bx = 64;
    k_copy_jobs<<<dim3(bx, g_copy.n), 256, 0, st>>>(g_copy.t);
    g_copy.n = 0;
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}
int push_copy_job(const CopyJob& J, cudaStream_t st) {
    if (J.d0 * J.d1 == 0) return NRH_OK;
    if (g_copy.n == COPY_MAX_JOBS) { int rc = flush_copy_jobs(st); if (rc) return rc; }
    g_copy.t.j[g_copy.n++] = J;
    return NRH_OK;
}
// dst[(row0 + k)*dst_ld + n] = src[n*src_ld + col0 + k]   (k < ncols, n < nrows)
int transpose_slice(const float* src, int src_ld, int nrows, int col0, int ncols, float* dst, int dst_ld, int row0, cudaStream_t st) {
    return push_copy_job(CopyJob{src + col0, dst + (size_t)row0 * dst_ld, nrows, ncols, src_ld, 1, 1, dst_ld, 1.0f}, st);
}
// dst[r*dst_ld + c] = src[r*cols + c]
int copy_rows(const float* src, int rows, int cols, float* dst, int dst_ld, cudaStream_t st) {
    return push_copy_job(CopyJob{src, dst, rows, cols, cols, 1, dst_ld, 1, 1.0f}, st);
}

int color_in_dim(const NrhConfig& c) { return 316 + (c.shadow_hint ? 9 : 0) + (c.specular_hint ? 9 * c.n_roughness : 0); }

struct Workspace {
    MarchState prim, shad;
    FineBuffers fine;
    RayState rs;
    float* feat;                 // [S*R][256]
    float* ssdf; float* sgx; float* sgy; float* sgz;   // shadow fine pass outputs
    float* rayfeat;              // [RAYFEAT][R]
    float* aux_img;              // [R/128] operand images of the per-ray reflectance inputs (streamed tcgen05 path)
    float* cr; float* cg; float* cb;
    OutsideBuffers ob;           // outside NeRF (null when off)
    float* mlp_scratch; size_t mlp_scratch_bytes;
    size_t total_bytes;
};

// carve (or just measure, base == nullptr) the workspace
Workspace carve(const NrhConfig& cfg, int64_t R, char* base, int num_sms) {
    Workspace w; memset(&w, 0, sizeof(w));
    size_t off = 0;
    auto take = [&](size_t nfloats) -> float* {
        size_t o = off; off += align_up(nfloats * sizeof(float), 256);
        return base ? reinterpret_cast<float*>(base + o) : nullptr;
    };
    const int S = cfg.n_samples + cfg.n_importance;
    const int Ss = cfg.n_shadow_samples + cfg.n_shadow_importance;
    const int Smax = NRH_MAX_SAMPLES;
    const int steps = cfg.up_sample_steps > 0 ? cfg.up_sample_steps : 1;
    int nnew = cfg.n_importance / steps; if (cfg.n_shadow_importance / 4 > nnew) nnew = cfg.n_shadow_importance / 4;
    if (nnew < 1) nnew = 1;
    auto march = [&](MarchState& m) {
        for (int c = 0; c < 3; ++c) { m.o[c] = take(R); m.d[c] = take(R); }
        for (int i = 0; i < 2; ++i) { m.z[i] = take((size_t)Smax * R); m.s[i] = take((size_t)Smax * R); }
        m.znew = take((size_t)nnew * R); m.snew = take((size_t)nnew * R);
        m.wbuf = take((size_t)Smax * R);
        m.px = take((size_t)Smax * R); m.py = take((size_t)Smax * R); m.pz = take((size_t)Smax * R);
    };
    march(w.prim); march(w.shad);
    FineBuffers& f = w.fine;
    f.sdf = take((size_t)S * R); f.gx = take((size_t)S * R); f.gy = take((size_t)S * R); f.gz = take((size_t)S * R);
    const int n_out = cfg.use_outside_nerf ? cfg.n_outside : 0;
    const int St = S + n_out;                                   // weights / colours carry the appended outside samples
    f.w = take((size_t)St * R); f.inside = take((size_t)S * R);
    f.nx = take((size_t)S * R); f.ny = take((size_t)S * R); f.nz = take((size_t)S * R);
    RayState& rs = w.rs;
    rs.depth = take(R); rs.wsum = take(R); rs.vis = take(R); rs.light_dist = take(R);
    for (int c = 0; c < 3; ++c) { rs.hit[c] = take(R); rs.hitn[c] = take(R); }
    for (int i = 0; i < NRH_MAX_ROUGHNESS; ++i) rs.spec[i] = take(R);
    w.feat = take((size_t)S * R * 256);
    w.ssdf = take((size_t)Ss * R); w.sgx = take((size_t)Ss * R); w.sgy = take((size_t)Ss * R); w.sgz = take((size_t)Ss * R);
    w.rayfeat = take((size_t)RAYFEAT * R);
    w.aux_img = take((size_t)((R + 127) / 128) * (TC_TILE_AUX_BYTES / sizeof(float)));
    w.cr = take((size_t)St * R); w.cg = take((size_t)St * R); w.cb = take((size_t)St * R);
    if (n_out > 0) {
        w.ob.mid = take((size_t)St * R); w.ob.dist = take((size_t)St * R); w.ob.density = take((size_t)St * R);
        w.ob.r = take((size_t)St * R); w.ob.g = take((size_t)St * R); w.ob.b = take((size_t)St * R);
    }
    size_t sb = sdf_mlp_simt_scratch_bytes(num_sms);
    size_t tb = tc_scratch_bytes(num_sms);
    w.mlp_scratch_bytes = sb > tb ? sb : tb;
    w.mlp_scratch = take(w.mlp_scratch_bytes / sizeof(float));
    w.total_bytes = off;
    return w;
}

// ---- fused training step: everything that lives from nrh_render_train_forward to nrh_render_backward ----------------------------------
struct TrainWs {
    char* render; size_t render_bytes;          // the render workspace (carve)
    void* tape; size_t tape_bytes;
    float* cap_sdf; float* cap_grad; float* cap_pts;                       // [N], [3][N], [3][N] (the features go straight into x16)
    float* z_rm;                                 // [R][S] final sample positions, ray-major
    float* dists; float* mid_z;                  // [R][S], [N]
    __half* x16; __half* acts; float* y; float* color; float* grad_aos;    // [N][384], [4][N][256], [N][4], [N][3], [N][3]
    // backward
    void* bwd; size_t bwd_bytes; float* sdf_bwd_ws; size_t sdf_bwd_ws_bytes;
    float* d_sdf; float* d_grad; float* d_color; float* dy;                // [N], [N][3], [N][3], [N][3]
    __half* dz; __half* dy16; __half* dx16;                                // [4][N][256], [N][8], [N][384]
    float* d_pts_r; float* d_pts_s; float* d_dirs_c; float* d_dirs_pe;     // [N][3] x2, [R][3] x2
    __half* e16; __half* ds16;                                             // [P_pad][64], [P_pad][8]
    float* dW;                                   // effective-weight gradients (see DW_* offsets), zeroed per step
    float* small;                                // [64]: loss scales, amax bits, d_inv_s, bias sums of the heads
    float* colsum;                               // [13][256]: zb (8), gb_8 (1), dz (4)
    size_t total_bytes;
    SdfTrainLayout lay;
};
// offsets (floats) inside TrainWs::dW
constexpr size_t DW_SDF0 = 0;                                   // [256][64]
constexpr size_t DW_SDF = DW_SDF0 + 256 * 64;                   // l = 1..7: [256][256] each at DW_SDF + (l-1) * 65536
constexpr size_t DW_HEAD = DW_SDF + 7 * 65536;                  // [8][256]: row 0 = d_sdf^T a_8
constexpr size_t DW_FEAT = DW_HEAD + 8 * 256;                   // [256][256]
constexpr size_t DW_FEATB = DW_FEAT + 65536;                    // [8][256]: row 0 = column sums of d_feat
constexpr size_t DW_COL0 = DW_FEATB + 8 * 256;                  // [256][384] (fused operand order)
constexpr size_t DW_COL = DW_COL0 + 256 * 384;                  // l = 1..3: [256][256]
constexpr size_t DW_COL4 = DW_COL + 3 * 65536;                  // [8][256] (3 rows)
constexpr size_t DW_TOTAL = DW_COL4 + 8 * 256;
// TrainWs::small slots
enum { SM_SCALE_C = 0 /* S_c, 1/S_c, . */, SM_SCALE_S = 4 /* S_s, 1/S_s, S_s/S_c */, SM_AMAX_C = 8, SM_AMAX_S = 9, SM_DINVS = 10,
       SM_DBS = 11 /* sum d_sdf */, SM_DB4 = 12 /* 3 */, SM_COUNT = 64 };

TrainWs carve_train(const NrhConfig& cfg, int64_t R, char* base, int num_sms) {
    TrainWs t; memset(&t, 0, sizeof(t));
    const int S = cfg.n_samples + cfg.n_importance;
    const int64_t N = (int64_t)S * R;
    t.lay = sdf_train_layout(N, num_sms);
    const int64_t Pp = t.lay.p_pad;
    size_t off = 0;
    auto take = [&](size_t bytes) -> char* { size_t o = off; off += align_up(bytes, 1024); return base ? base + o : nullptr; };
    t.render_bytes = carve(cfg, R, nullptr, num_sms).total_bytes;
    t.render = take(t.render_bytes);
    t.tape_bytes = t.lay.tape_bytes; t.tape = take(t.tape_bytes);
    t.cap_sdf = (float*)take(4 * N); t.cap_grad = (float*)take(12 * N); t.cap_pts = (float*)take(12 * N);
    t.z_rm = (float*)take(4 * N); t.dists = (float*)take(4 * N); t.mid_z = (float*)take(4 * N);
    t.x16 = (__half*)take(768 * N); t.acts = (__half*)take(2048 * N); t.y = (float*)take(16 * N); t.color = (float*)take(12 * N);
    t.grad_aos = (float*)take(12 * N);
    t.bwd_bytes = t.lay.bwd_bytes; t.bwd = take(t.bwd_bytes);
    t.sdf_bwd_ws_bytes = t.lay.bwd_workspace_bytes; t.sdf_bwd_ws = (float*)take(t.sdf_bwd_ws_bytes);
    t.d_sdf = (float*)take(4 * N); t.d_grad = (float*)take(12 * N); t.d_color = (float*)take(12 * N); t.dy = (float*)take(12 * N);
    t.dz = (__half*)take(2048 * N); t.dy16 = (__half*)take(16 * N); t.dx16 = (__half*)take(768 * N);
    t.d_pts_r = (float*)take(12 * N); t.d_pts_s = (float*)take(12 * N); t.d_dirs_c = (float*)take(12 * R); t.d_dirs_pe = (float*)take(12 * R);
    t.e16 = (__half*)take(128 * Pp); t.ds16 = (__half*)take(16 * Pp);
    t.dW = (float*)take(4 * DW_TOTAL); t.small = (float*)take(4 * SM_COUNT); t.colsum = (float*)take(4 * 13 * 256);
    t.total_bytes = off;
    return t;
}

int device_sms(int* out) {
    int dev = 0;
    NRH_CUDA_CHECK(cudaGetDevice(&dev));
    NRH_CUDA_CHECK(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
    return NRH_OK;
}
int sms_or_default() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 148;                      // B200; sizing queries may run without a device
    }
    return n;
}

int resolve_impl(const NrhConfig& cfg) {
    if (cfg.mlp_impl == NRH_MLP_AUTO) return tc_available() ? NRH_MLP_TCGEN05 : NRH_MLP_FP32_SIMT;
    return cfg.mlp_impl;
}

// engine dispatch ------------------------------------------------------------------------------------
int run_sdf(const NrhConfig& cfg, const void* packed, const PackedLayout& L, Strided3 pts, int64_t N, float* sdf,
            float* gx, float* gy, float* gz, int64_t gstride, float* feat, float* scratch, size_t scratch_bytes,
            int num_sms, cudaStream_t st, bool feat_as_image = false) {
    if (resolve_impl(cfg) == NRH_MLP_TCGEN05)
        return sdf_mlp_tc(packed, L, pts, N, sdf, gx, gy, gz, gstride, feat, feat_as_image, scratch, scratch_bytes, num_sms, st);
    return sdf_mlp_simt(reinterpret_cast<const float*>(packed), L, pts, N, sdf, gx, gy, gz, gstride, feat, scratch,
                        scratch_bytes, num_sms, st);
}
int run_color(const NrhConfig& cfg, const void* packed, const PackedLayout& L, Strided3 pts, Strided3 nrm, const float* feat,
              const float* rayfeat, const void* aux_img, int64_t R, int64_t N, float* cr, float* cg, float* cb, float* scratch,
              size_t scratch_bytes, int num_sms, cudaStream_t st) {
    if (resolve_impl(cfg) == NRH_MLP_TCGEN05)
        return color_mlp_tc(packed, L, pts, nrm, feat, rayfeat, aux_img, R, N, cr, cg, cb, scratch, scratch_bytes, num_sms, st);
    return color_mlp_simt(reinterpret_cast<const float*>(packed), L, pts, nrm, feat, rayfeat, R, N, cr, cg, cb, num_sms, st);
}

// coarse SDF pass + importance steps of one march; leaves the final z in m.z[*cur_out] and the
// section mid-points in m.p{x,y,z}
int run_hierarchical(const NrhConfig& cfg, const void* packed, const PackedLayout& L, int64_t R, const MarchState& m,
                     int n, int n_imp, int steps, float last_dist_const, const float* last_dist_ray,
                     float* scratch, size_t scratch_bytes, int num_sms, cudaStream_t st, int* cur_out) {
    int cur = 0, k = n, rc;
    if (n_imp <= 0 || steps <= 0) {
        if ((rc = launch_sections_only(R, m, cur, n, last_dist_const, last_dist_ray, st))) return rc;
        *cur_out = cur; return NRH_OK;
    }
    const int n_new = n_imp / steps;
    Strided3 P{m.px, m.py, m.pz, 1};
    if ((rc = run_sdf(cfg, packed, L, P, (int64_t)n * R, m.s[0], nullptr, nullptr, nullptr, 1, nullptr, scratch, scratch_bytes, num_sms, st))) return rc;
    for (int i = 0; i < steps; ++i) {
        const bool last = (i + 1 == steps), merge_first = (i > 0);
        const float inv_s = 64.0f * (float)(1 << i);
        if ((rc = launch_importance_step(R, m, cur, k, n_new, merge_first, inv_s, last, last_dist_const, last_dist_ray, st))) return rc;
        if (merge_first) { cur ^= 1; k += n_new; }
        if (!last) {
            if ((rc = run_sdf(cfg, packed, L, P, (int64_t)n_new * R, m.snew, nullptr, nullptr, nullptr, 1, nullptr, scratch, scratch_bytes, num_sms, st))) return rc;
        } else {
            cur ^= 1; k += n_new;
        }
    }
    *cur_out = cur;
    return NRH_OK;
}

}  // namespace

// dst[i] = src[i] * scale, i < n -- joins the copy-job table of the running nrh_pack_weights call (flushed at its end)
int queue_scaled_copy(const float* src, float* dst, int n, float scale, cudaStream_t st) {
    return push_copy_job(CopyJob{src, dst, 1, n, 0, 1, 0, 1, scale}, st);
}
}  // namespace nrh

using namespace nrh;

extern "C" {

int nrh_version(void) { return NRH_ABI_VERSION; }
const char* nrh_last_error(void) { return g_err; }
int nrh_last_launch_count(void) { return g_launches; }

int nrh_check_config(const NrhConfig* c) {
    if (!c) { set_error("config is null"); return NRH_ERR_INVALID; }
    const int S = c->n_samples + c->n_importance, Ss = c->n_shadow_samples + c->n_shadow_importance;
    if (c->n_samples < 2 || S > NRH_MAX_SAMPLES) { set_error("n_samples + n_importance must be in [2,%d]", NRH_MAX_SAMPLES); return NRH_ERR_UNSUPPORTED; }
    if (c->n_importance > 0 && (c->up_sample_steps < 1 || c->n_importance % c->up_sample_steps)) { set_error("n_importance must be a multiple of up_sample_steps"); return NRH_ERR_UNSUPPORTED; }
    if (c->shadow_hint) {
        if (c->n_shadow_samples < 2 || Ss > NRH_MAX_SAMPLES) { set_error("shadow sample counts out of range"); return NRH_ERR_UNSUPPORTED; }
        if (c->n_shadow_importance % 4) { set_error("n_shadow_importance must be a multiple of 4 (get_visibility uses 4 steps)"); return NRH_ERR_UNSUPPORTED; }
    }
    if (c->specular_hint && (c->n_roughness < 1 || c->n_roughness > NRH_MAX_ROUGHNESS)) { set_error("n_roughness must be in [1,%d]", NRH_MAX_ROUGHNESS); return NRH_ERR_UNSUPPORTED; }
    if (c->depth_type < NRH_DEPTH_ALPHA_BLEND || c->depth_type > NRH_DEPTH_SPHERE_TRACE) { set_error("unknown depth_type %d", c->depth_type); return NRH_ERR_INVALID; }
    if (c->use_outside_nerf && (c->n_outside < 1 || c->n_outside > NRH_MAX_OUTSIDE)) { set_error("n_outside must be in [1,%d]", NRH_MAX_OUTSIDE); return NRH_ERR_UNSUPPORTED; }
    if (c->mlp_impl < NRH_MLP_AUTO || c->mlp_impl > NRH_MLP_TCGEN05) { set_error("unknown mlp_impl %d", c->mlp_impl); return NRH_ERR_INVALID; }
    if (c->mlp_impl == NRH_MLP_TCGEN05 && !tc_available()) { set_error("tcgen05 engine not built"); return NRH_ERR_UNSUPPORTED; }
    return NRH_OK;
}

size_t nrh_packed_weights_bytes(const NrhConfig* cfg) {
    if (!cfg) return 0;
    return make_layout(*cfg).total_bytes;
}

int nrh_pack_weights(const NrhConfig* cfg, const NrhRawWeights* raw, void* packed, size_t packed_bytes, void* stream) {
    g_launches = 0;
    int rc = nrh_check_config(cfg); if (rc) return rc;
    if (!raw || !packed) { set_error("null argument"); return NRH_ERR_INVALID; }
    const PackedLayout L = make_layout(*cfg);
    if (packed_bytes < L.total_bytes) { set_error("packed buffer too small: %zu < %zu", packed_bytes, L.total_bytes); return NRH_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    float* P = reinterpret_cast<float*>(packed);
    g_copy.n = 0;
    NRH_CUDA_CHECK(cudaMemsetAsync(packed, 0, L.total_bytes, st));
    k_inv_s<<<1, 1, 0, st>>>(raw->variance, P + L.inv_s); NRH_LAUNCH_CHECK();
    for (int l = 0; l < SDF_LAYERS; ++l) {
        const int in = (l == 0) ? PE_DIM : 256, out = (l == SDF_SKIP - 1) ? SKIP_H : 256;
        if ((rc = transpose_slice(raw->sdf_W[l], in, out, 0, in, P + L.sdf_wt[l], 256, 0, st))) return rc;
        if ((rc = copy_rows(raw->sdf_b[l], 1, out, P + L.sdf_b[l], 256, st))) return rc;
        if ((rc = copy_rows(raw->sdf_W[l], out, in, P + L.sdf_wn[l], l == 0 ? PE_PAD : 256, st))) return rc;
    }
    if ((rc = copy_rows(raw->sdf_out_W, 1, 256, P + L.head_w, 256, st))) return rc;
    if ((rc = copy_rows(raw->sdf_out_b, 1, 1, P + L.head_b, 1, st))) return rc;
    if ((rc = transpose_slice(raw->feat_W, 256, 256, 0, 256, P + L.feat_wt, 256, 0, st))) return rc;
    if ((rc = copy_rows(raw->feat_b, 1, 256, P + L.feat_b, 256, st))) return rc;
    // reflectance layer 0: [pts 3 | PE(view) 27 | normal 3 | PE(light) 27 | feat 256 | PE(vis) 9? | PE(spec) 9*nr?]
    const int cin = color_in_dim(*cfg);
    if ((rc = transpose_slice(raw->col_W[0], cin, 256, 60, 256, P + L.col_wt0a, 256, 0, st))) return rc;
    if ((rc = transpose_slice(raw->col_W[0], cin, 256, 0, 60, P + L.col_wt0b, 256, 0, st))) return rc;
    int col = 316;
    if (cfg->shadow_hint) { if ((rc = transpose_slice(raw->col_W[0], cin, 256, col, 9, P + L.col_wt0b, 256, AUX_VIS, st))) return rc; col += 9; }
    if (cfg->specular_hint) { if ((rc = transpose_slice(raw->col_W[0], cin, 256, col, 9 * cfg->n_roughness, P + L.col_wt0b, 256, AUX_SPEC, st))) return rc; }
    if ((rc = copy_rows(raw->col_b[0], 1, 256, P + L.col_b0, 256, st))) return rc;
    for (int l = 0; l < 3; ++l) {
        if ((rc = transpose_slice(raw->col_W[l + 1], 256, 256, 0, 256, P + L.col_wt[l], 256, 0, st))) return rc;
        if ((rc = copy_rows(raw->col_b[l + 1], 1, 256, P + L.col_b[l], 256, st))) return rc;
    }
    if ((rc = transpose_slice(raw->col_W[4], 256, 3, 0, 256, P + L.col_w4t, 4, 0, st))) return rc;
    if ((rc = copy_rows(raw->col_b[4], 1, 3, P + L.col_b4, 4, st))) return rc;
    if (cfg->use_outside_nerf) {           // fields/nerf_density_field.py:57-64
        for (int l = 0; l < NERF_LAYERS; ++l) {
            if (!raw->nerf_W[l] || !raw->nerf_b[l]) { set_error("use_outside_nerf: null outside-NeRF weight"); return NRH_ERR_INVALID; }
            if (l == 0) { if ((rc = transpose_slice(raw->nerf_W[0], NERF_PE, 256, 0, NERF_PE, P + L.nerf_wt[0], 256, 0, st))) return rc; }
            else if (l == NERF_SKIP + 1) {
                if ((rc = transpose_slice(raw->nerf_W[l], NERF_PE + 256, 256, 0, NERF_PE, P + L.nerf_wt5e, 256, 0, st))) return rc;
                if ((rc = transpose_slice(raw->nerf_W[l], NERF_PE + 256, 256, NERF_PE, 256, P + L.nerf_wt[l], 256, 0, st))) return rc;
            } else { if ((rc = transpose_slice(raw->nerf_W[l], 256, 256, 0, 256, P + L.nerf_wt[l], 256, 0, st))) return rc; }
            if ((rc = copy_rows(raw->nerf_b[l], 1, 256, P + L.nerf_b[l], 256, st))) return rc;
        }
        if (!raw->nerf_alpha_W || !raw->nerf_alpha_b || !raw->nerf_feat_W || !raw->nerf_feat_b || !raw->nerf_view_W ||
            !raw->nerf_view_b || !raw->nerf_rgb_W || !raw->nerf_rgb_b) { set_error("use_outside_nerf: null outside-NeRF head weight"); return NRH_ERR_INVALID; }
        if ((rc = copy_rows(raw->nerf_alpha_W, 1, 256, P + L.nerf_alpha_w, 256, st))) return rc;
        if ((rc = copy_rows(raw->nerf_alpha_b, 1, 1, P + L.nerf_alpha_b, 1, st))) return rc;
        if ((rc = transpose_slice(raw->nerf_feat_W, 256, 256, 0, 256, P + L.nerf_feat_wt, 256, 0, st))) return rc;
        if ((rc = copy_rows(raw->nerf_feat_b, 1, 256, P + L.nerf_feat_b, 256, st))) return rc;
        if ((rc = transpose_slice(raw->nerf_view_W, 256 + NERF_VPE, NERF_VIEW_H, 0, 256, P + L.nerf_view_wta, 256, 0, st))) return rc;
        if ((rc = transpose_slice(raw->nerf_view_W, 256 + NERF_VPE, NERF_VIEW_H, 256, NERF_VPE, P + L.nerf_view_wtb, 256, 0, st))) return rc;
        if ((rc = copy_rows(raw->nerf_view_b, 1, NERF_VIEW_H, P + L.nerf_view_b, 256, st))) return rc;
        if ((rc = transpose_slice(raw->nerf_rgb_W, NERF_VIEW_H, 3, 0, NERF_VIEW_H, P + L.nerf_rgb_wt, 4, 0, st))) return rc;
        if ((rc = copy_rows(raw->nerf_rgb_b, 1, 3, P + L.nerf_rgb_b, 4, st))) return rc;
    }
    if ((rc = flush_copy_jobs(st))) return rc;      // every fp32 re-layout queued above in ONE launch (tc_pack reads some of them)
    if (tc_available()) { if ((rc = tc_pack(*cfg, L, *raw, packed, st))) return rc; }
    return flush_copy_jobs(st);                      // tc_pack's pre-scaled biases
}

size_t nrh_workspace_bytes(const NrhConfig* cfg, int64_t R) {
    if (!cfg || R <= 0) return 0;
    return carve(*cfg, R, nullptr, sms_or_default()).total_bytes;
}

size_t nrh_query_workspace_bytes(const NrhConfig* cfg, int64_t N) {
    (void)N;
    if (!cfg) return 0;
    const int sms = sms_or_default();
    size_t sb = sdf_mlp_simt_scratch_bytes(sms), tb = tc_scratch_bytes(sms);
    return (sb > tb ? sb : tb) + 256;
}

int nrh_sdf_query(const NrhConfig* cfg, const void* packed, const float* pts, int64_t N,
                  float* sdf, float* grad, float* feat, void* workspace, size_t workspace_bytes, void* stream) {
    g_launches = 0;
    int rc = nrh_check_config(cfg); if (rc) return rc;
    if (!packed || !pts || !sdf || N < 0) { set_error("null argument"); return NRH_ERR_INVALID; }
    if (N == 0) return NRH_OK;
    int sms; if ((rc = device_sms(&sms))) return rc;
    if (workspace_bytes < nrh_query_workspace_bytes(cfg, N) - 256 || !workspace) { set_error("workspace too small"); return NRH_ERR_WORKSPACE; }
    const PackedLayout L = make_layout(*cfg);
    Strided3 P{pts, pts + 1, pts + 2, 3};
    return run_sdf(*cfg, packed, L, P, N, sdf, grad, grad ? grad + 1 : nullptr, grad ? grad + 2 : nullptr, 3, feat,
                   reinterpret_cast<float*>(workspace), workspace_bytes, sms, (cudaStream_t)stream);
}

int nrh_sdf_train_layout(const NrhConfig* cfg, int64_t N, NrhTrainLayout* out) {
    if (!cfg || !out || N < 0) { set_error("null argument"); return NRH_ERR_INVALID; }
    const SdfTrainLayout t = sdf_train_layout(N, sms_or_default());
    out->p_pad = t.p_pad;
    out->tape_tiles_off = t.tape_tiles_off; out->tape_act_off = t.tape_act_off; out->tape_u_off = t.tape_u_off; out->tape_bytes = t.tape_bytes;
    out->bwd_gb0_off = t.bwd_gb0_off; out->bwd_gb_off = t.bwd_gb_off; out->bwd_zb_off = t.bwd_zb_off; out->bwd_bytes = t.bwd_bytes;
    out->bwd_workspace_bytes = t.bwd_workspace_bytes;
    return NRH_OK;
}

int nrh_sdf_train_forward(const NrhConfig* cfg, const void* packed, const float* pts, int64_t N, float* sdf, float* grad,
                          float* feat, void* tape, size_t tape_bytes, void* workspace, size_t workspace_bytes, void* stream) {
    g_launches = 0;
    int rc = nrh_check_config(cfg); if (rc) return rc;
    if (resolve_impl(*cfg) != NRH_MLP_TCGEN05) { set_error("nrh_sdf_train_forward needs the tcgen05 engine"); return NRH_ERR_UNSUPPORTED; }
    if (N == 0) return NRH_OK;
    if (!packed || !pts || !sdf || !grad || !feat || !tape || !workspace || N < 0) { set_error("null argument"); return NRH_ERR_INVALID; }
    int sms; if ((rc = device_sms(&sms))) return rc;
    const SdfTrainLayout t = sdf_train_layout(N, sms);
    if (tape_bytes < t.tape_bytes) { set_error("tape too small: %zu < %zu", tape_bytes, t.tape_bytes); return NRH_ERR_WORKSPACE; }
    const PackedLayout L = make_layout(*cfg);
    return sdf_train_forward_tc(packed, L, pts, N, sdf, grad, feat, tape, reinterpret_cast<float*>(workspace), workspace_bytes, sms,
                                (cudaStream_t)stream);
}

int nrh_sdf_train_backward(const NrhConfig* cfg, const void* packed, const float* pts, int64_t N, const void* tape,
                           size_t tape_bytes, const float* d_sdf, const float* d_feat, const float* d_grad,
                           const float* loss_scale, void* bwd_out, size_t bwd_bytes, float* d_pts,
                           void* workspace, size_t workspace_bytes, void* stream) {
    g_launches = 0;
    int rc = nrh_check_config(cfg); if (rc) return rc;
    if (resolve_impl(*cfg) != NRH_MLP_TCGEN05) { set_error("nrh_sdf_train_backward needs the tcgen05 engine"); return NRH_ERR_UNSUPPORTED; }
    if (N == 0) return NRH_OK;
    if (!packed || !pts || !tape || !d_sdf || !d_feat || !d_grad || !loss_scale || !bwd_out || !d_pts || !workspace || N < 0) {
        set_error("null argument"); return NRH_ERR_INVALID;
    }
    int sms; if ((rc = device_sms(&sms))) return rc;
    const SdfTrainLayout t = sdf_train_layout(N, sms);
    if (tape_bytes < t.tape_bytes || bwd_bytes < t.bwd_bytes) { set_error("tape / bwd_out too small"); return NRH_ERR_WORKSPACE; }
    const PackedLayout L = make_layout(*cfg);
    return sdf_train_backward_tc(packed, L, pts, N, tape, d_sdf, d_feat, d_grad, loss_scale, bwd_out, d_pts,
                                 reinterpret_cast<float*>(workspace), workspace_bytes, sms, (cudaStream_t)stream);
}

int nrh_color_train_forward(const NrhConfig* cfg, const void* packed, const void* x16, int64_t P, void* acts, float* y, void* stream) {
    g_launches = 0;
    int rc = nrh_check_config(cfg); if (rc) return rc;
    if (resolve_impl(*cfg) != NRH_MLP_TCGEN05) { set_error("nrh_color_train_forward needs the tcgen05 engine"); return NRH_ERR_UNSUPPORTED; }
    if (P == 0) return NRH_OK;
    if (!packed || !x16 || !acts || !y || P < 0) { set_error("null argument"); return NRH_ERR_INVALID; }
    int sms; if ((rc = device_sms(&sms))) return rc;
    return color_train_forward_tc(packed, make_layout(*cfg), x16, P, acts, y, sms, (cudaStream_t)stream);
}

int nrh_color_train_backward(const NrhConfig* cfg, const void* packed, const float* dy, const float* loss_scale, const void* acts,
                             int64_t P, void* dz, void* dy16, void* dx, void* stream) {
    g_launches = 0;
    int rc = nrh_check_config(cfg); if (rc) return rc;
    if (resolve_impl(*cfg) != NRH_MLP_TCGEN05) { set_error("nrh_color_train_backward needs the tcgen05 engine"); return NRH_ERR_UNSUPPORTED; }
    if (P == 0) return NRH_OK;
    if (!packed || !dy || !loss_scale || !acts || !dz || !dy16 || !dx || P < 0) { set_error("null argument"); return NRH_ERR_INVALID; }
    int sms; if ((rc = device_sms(&sms))) return rc;
    return color_train_backward_tc(packed, make_layout(*cfg), dy, loss_scale, acts, P, dz, dy16, dx, sms, (cudaStream_t)stream);
}

int nrh_sphere_trace(const NrhConfig* cfg, const void* packed, const float* origins, const float* directions, int64_t R,
                     int max_iterations, float threshold, float far_limit, int check_every,
                     float* hit_points, float* hit_depths, void* workspace, size_t workspace_bytes, void* stream) {
    g_launches = 0;
    int rc = nrh_check_config(cfg); if (rc) return rc;
    if (!packed || !origins || !directions || !hit_points || !hit_depths || !workspace || R < 0) { set_error("null argument"); return NRH_ERR_INVALID; }
    if (R == 0) return NRH_OK;
    int sms; if ((rc = device_sms(&sms))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t mlp_bytes = nrh_query_workspace_bytes(cfg, R);
    if (workspace_bytes < mlp_bytes + 32 * (size_t)R) { set_error("workspace too small"); return NRH_ERR_WORKSPACE; }
    char* base = reinterpret_cast<char*>(workspace);
    float* scratch = reinterpret_cast<float*>(base);
    float* sdf = reinterpret_cast<float*>(base + mlp_bytes);                  // [R]
    int* moving = reinterpret_cast<int*>(base + mlp_bytes + 4 * (size_t)R + 16);
    const PackedLayout L = make_layout(*cfg);
    NRH_CUDA_CHECK(cudaMemcpyAsync(hit_points, origins, sizeof(float) * 3 * R, cudaMemcpyDeviceToDevice, st));
    NRH_CUDA_CHECK(cudaMemsetAsync(hit_depths, 0, sizeof(float) * R, st));
    if (check_every < 1) check_every = 1;
    for (int it = 0; it < max_iterations; ++it) {
        const bool check = ((it + 1) % check_every == 0) || (it + 1 == max_iterations);
        if (it % check_every == 0) NRH_CUDA_CHECK(cudaMemsetAsync(moving, 0, sizeof(int), st));
        Strided3 P{hit_points, hit_points + 1, hit_points + 2, 3};
        if ((rc = run_sdf(*cfg, packed, L, P, R, sdf, nullptr, nullptr, nullptr, 1, nullptr, scratch, mlp_bytes, sms, st))) return rc;
        if ((rc = launch_sphere_step(R, directions, sdf, hit_points, hit_depths, threshold, far_limit, moving, st))) return rc;
        if (check) {
            int h = 0;
            NRH_CUDA_CHECK(cudaMemcpyAsync(&h, moving, sizeof(int), cudaMemcpyDeviceToHost, st));
            NRH_CUDA_CHECK(cudaStreamSynchronize(st));
            if (h == 0) break;                     // nothing moved during the last `check_every` updates: every ray converged
        }
    }
    return NRH_OK;
}

int nrh_render_forward(const NrhConfig* cfg, const void* packed, const NrhRays* rays, int64_t R,
                       const float* bg_rgb, const float* jitter_primary, const float* jitter_outside,
                       const float* jitter_shadow, float cos_anneal, int warmup, const NrhOutputs* out,
                       void* workspace, size_t workspace_bytes, void* stream) {
    g_launches = 0;
    int rc = nrh_check_config(cfg); if (rc) return rc;
    if (!packed || !rays || !out || !workspace) { set_error("null argument"); return NRH_ERR_INVALID; }
    if (!rays->origins || !rays->directions || !rays->pl_positions || !rays->nears || !rays->fars) { set_error("null ray field"); return NRH_ERR_INVALID; }
    if (!out->rgb || !out->depth) { set_error("null output field (rgb / depth are mandatory)"); return NRH_ERR_INVALID; }
    if (R < 0) { set_error("negative ray count"); return NRH_ERR_INVALID; }
    if (cfg->depth_type == NRH_DEPTH_SPHERE_TRACE && (!rays->hit_points || !rays->hit_depths)) {
        set_error("depth_type == sphere tracing needs hit_points / hit_depths from nrh_sphere_trace"); return NRH_ERR_INVALID;
    }
    if (R == 0) return NRH_OK;
    int sms; if ((rc = device_sms(&sms))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    Workspace w = carve(*cfg, R, reinterpret_cast<char*>(workspace), sms);
    if (workspace_bytes < w.total_bytes) { set_error("workspace too small: %zu < %zu", workspace_bytes, w.total_bytes); return NRH_ERR_WORKSPACE; }
    const PackedLayout L = make_layout(*cfg);
    const float* Pf = reinterpret_cast<const float*>(packed);
    const float* inv_s = Pf + L.inv_s;
    const int n = cfg->n_samples, S = cfg->n_samples + cfg->n_importance;
    const int ns = cfg->n_shadow_samples, Ss = cfg->n_shadow_samples + cfg->n_shadow_importance;
    const float sample_dist = 2.0f / (float)n;                              // :673

    // ---- primary march ----------------------------------------------------------------------------------
    if ((rc = launch_coarse_primary(*rays, R, n, jitter_primary, w.prim, st))) return rc;
    int cur = 0;
    if ((rc = run_hierarchical(*cfg, packed, L, R, w.prim, n, cfg->n_importance, cfg->up_sample_steps, sample_dist, nullptr,
                               w.mlp_scratch, w.mlp_scratch_bytes, sms, st, &cur))) return rc;
    Strided3 P{w.prim.px, w.prim.py, w.prim.pz, 1};
    // tcgen05 engine with whole 128-ray blocks: features and per-ray inputs travel as fp16 operand images that the
    // reflectance kernel streams straight into its A operand (no conversion pass, half the feature bytes)
    const NrhTrainCapture* cap = out->train_capture;
    if (cap) {
        if (resolve_impl(*cfg) != NRH_MLP_TCGEN05 || cfg->use_outside_nerf) { set_error("train_capture needs the tcgen05 engine without the outside NeRF"); return NRH_ERR_UNSUPPORTED; }
        if (!cap->tape || !cap->sdf || !cap->grad_soa || (!cap->feat && !cap->feat16) || !cap->pts_soa) { set_error("train_capture: null buffer"); return NRH_ERR_INVALID; }
        if (cap->tape_bytes < sdf_train_layout((int64_t)S * R, sms).tape_bytes) { set_error("train_capture: tape too small"); return NRH_ERR_WORKSPACE; }
    }
    const bool streamed = !cap && resolve_impl(*cfg) == NRH_MLP_TCGEN05 && (R % 128 == 0);
    if (out->fine_begin_event) NRH_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(out->fine_begin_event), st));
    if (cap) {
        // the training forward (same arithmetic + tape) straight into the caller's buffers; the compositor reads them there
        const int64_t N = (int64_t)S * R;
        if ((rc = sdf_train_forward_tc_strided(packed, L, P, N, cap->sdf, cap->grad_soa, cap->grad_soa + N, cap->grad_soa + 2 * N, 1,
                                               cap->feat, cap->tape, w.mlp_scratch, w.mlp_scratch_bytes, sms, st, cap->feat16, cap->feat16_ld))) return rc;
        NRH_CUDA_CHECK(cudaMemcpyAsync(w.fine.sdf, cap->sdf, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
        NRH_CUDA_CHECK(cudaMemcpyAsync(w.fine.gx, cap->grad_soa, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
        NRH_CUDA_CHECK(cudaMemcpyAsync(w.fine.gy, cap->grad_soa + N, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
        NRH_CUDA_CHECK(cudaMemcpyAsync(w.fine.gz, cap->grad_soa + 2 * N, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
        NRH_CUDA_CHECK(cudaMemcpyAsync(cap->pts_soa, w.prim.px, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
        NRH_CUDA_CHECK(cudaMemcpyAsync(cap->pts_soa + N, w.prim.py, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
        NRH_CUDA_CHECK(cudaMemcpyAsync(cap->pts_soa + 2 * N, w.prim.pz, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
    } else if ((rc = run_sdf(*cfg, packed, L, P, (int64_t)S * R, w.fine.sdf, w.fine.gx, w.fine.gy, w.fine.gz, 1, w.feat,
                             w.mlp_scratch, w.mlp_scratch_bytes, sms, st, streamed))) return rc;
    if (out->fine_end_event) NRH_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(out->fine_end_event), st));
    // ---- outside NeRF on the merged sample set (render_outside, :716-724): background alpha / colour per section ----
    const int n_out = cfg->use_outside_nerf ? cfg->n_outside : 0;
    const int St = S + n_out;
    if (n_out > 0) {
        if ((rc = launch_outside_setup(R, w.prim, cur, S, n, n_out, rays->fars, jitter_outside, sample_dist, w.ob, st))) return rc;
        const float* oo[3] = {w.prim.o[0], w.prim.o[1], w.prim.o[2]};
        const float* dd[3] = {w.prim.d[0], w.prim.d[1], w.prim.d[2]};
        if ((rc = nerf_mlp_simt(Pf, L, oo, dd, rays->pl_positions, w.ob.mid, R, (int64_t)St * R, w.ob.density, w.ob.r, w.ob.g,
                                w.ob.b, sms, st))) return rc;
    }
    const bool do_shadow = cfg->shadow_hint && !warmup;
    if ((rc = launch_composite_primary(R, w.prim, cur, S, sample_dist, inv_s, cos_anneal, w.fine, w.rs, rays->pl_positions,
                                       do_shadow, w.shad, ns, cfg->shadow_ray_offset, jitter_shadow,
                                       cfg->depth_type, rays->hit_points, rays->hit_depths, n_out, w.ob, st))) return rc;
    if ((rc = launch_specular_cue(R, *cfg, w.rs, rays->pl_positions, rays->directions, warmup, st))) return rc;
    // ---- per-sample geometry block of the RenderOutput, ray-major: final from here on (95 % of the output bytes) ----
    {
        const float* s1[4] = {w.fine.w, nullptr, nullptr, nullptr};
        if (out->weights && (rc = launch_to_ray_major(s1, 1, false, R, St, out->weights, st))) return rc;
        const float* s2[4] = {w.fine.inside, nullptr, nullptr, nullptr};
        if (out->inside_sphere && (rc = launch_to_ray_major(s2, 1, false, R, S, out->inside_sphere, st))) return rc;
        const float* s3[4] = {w.fine.gx, w.fine.gy, w.fine.gz, nullptr};
        if (out->analytic_normals && (rc = launch_to_ray_major(s3, 3, false, R, S, out->analytic_normals, st))) return rc;
        const float* s4[4] = {w.fine.nx, w.fine.ny, w.fine.nz, nullptr};
        if (out->normalized_normals && (rc = launch_to_ray_major(s4, 3, false, R, S, out->normalized_normals, st))) return rc;
        if (cfg->specular_hint && out->specular_cue) {
            const float* s5[4] = {w.rs.spec[0], w.rs.spec[1], w.rs.spec[2], w.rs.spec[3]};
            if ((rc = launch_to_ray_major(s5, cfg->n_roughness, true, R, S, out->specular_cue, st))) return rc;
        }
        if (out->z_vals) {
            const float* s6[4] = {w.prim.z[cur], nullptr, nullptr, nullptr};
            if ((rc = launch_to_ray_major(s6, 1, false, R, S, out->z_vals, st))) return rc;
        }
        if (out->early_event) NRH_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(out->early_event), st));
    }
    // ---- shadow march -------------------------------------------------------------------------------------
    int scur = 0;
    if (do_shadow) {
        if ((rc = run_hierarchical(*cfg, packed, L, R, w.shad, ns, cfg->n_shadow_importance, 4, 0.f, w.rs.light_dist,
                                   w.mlp_scratch, w.mlp_scratch_bytes, sms, st, &scur))) return rc;
        Strided3 Q{w.shad.px, w.shad.py, w.shad.pz, 1};
        if ((rc = run_sdf(*cfg, packed, L, Q, (int64_t)Ss * R, w.ssdf, w.sgx, w.sgy, w.sgz, 1, nullptr,
                          w.mlp_scratch, w.mlp_scratch_bytes, sms, st))) return rc;
    }
    if ((rc = launch_shade_prep(R, *cfg, w.shad, scur, Ss, inv_s, cos_anneal, w.ssdf, w.sgx, w.sgy, w.sgz, w.rs,
                                rays->pl_positions, rays->directions, warmup, do_shadow, w.rayfeat,
                                streamed ? reinterpret_cast<unsigned char*>(w.aux_img) : nullptr, st))) return rc;
    // ---- reflectance + composite ----------------------------------------------------------------------------
    Strided3 Nrm = cfg->normalized_normals ? Strided3{w.fine.nx, w.fine.ny, w.fine.nz, 1} : Strided3{w.fine.gx, w.fine.gy, w.fine.gz, 1};
    if (cap) {            // training capture: reflectance + compositing belong to the caller's differentiable path
        NRH_CUDA_CHECK(cudaMemsetAsync(w.cr, 0, sizeof(float) * (size_t)St * R, st));
        NRH_CUDA_CHECK(cudaMemsetAsync(w.cg, 0, sizeof(float) * (size_t)St * R, st));
        NRH_CUDA_CHECK(cudaMemsetAsync(w.cb, 0, sizeof(float) * (size_t)St * R, st));
    } else if ((rc = run_color(*cfg, packed, L, P, Nrm, w.feat, w.rayfeat, streamed ? w.aux_img : nullptr, R, (int64_t)S * R, w.cr, w.cg, w.cb,
                               w.mlp_scratch, w.mlp_scratch_bytes, sms, st))) return rc;
    if ((rc = launch_final_rgb(R, S, w.fine, w.rs, w.cr, w.cg, w.cb, bg_rgb, out->rgb, out->depth,
                               cfg->shadow_hint ? out->visibilities : nullptr, out->normal_map, out->normalized_normal_map,
                               cfg->specular_hint ? out->specular_cue_ray : nullptr, cfg->n_roughness, n_out, w.ob, st))) return rc;
    // ---- late ray-major RenderOutput fields ---------------------------------------------------------------------
    if (out->z_shadow && do_shadow) {
        const float* s7[4] = {w.shad.z[scur], nullptr, nullptr, nullptr};
        if ((rc = launch_to_ray_major(s7, 1, false, R, Ss, out->z_shadow, st))) return rc;
    }
    if (out->sampled_color) {
        const float* s8[4] = {w.cr, w.cg, w.cb, nullptr};
        if ((rc = launch_to_ray_major(s8, 3, false, R, St, out->sampled_color, st))) return rc;
    }
    if (out->inv_s) { k_copy_scalar<<<1, 1, 0, st>>>(inv_s, out->inv_s); NRH_LAUNCH_CHECK(); }
    return NRH_OK;
}


// ================================================================================================================================
// fused training step (see include/nrhints_b200.h)
// ================================================================================================================================
static int train_check(const NrhConfig* cfg, int64_t R, const char* who) {
    int rc = nrh_check_config(cfg); if (rc) return rc;
    if (resolve_impl(*cfg) != NRH_MLP_TCGEN05 || cfg->use_outside_nerf) { set_error("%s needs the tcgen05 engine without the outside NeRF", who); return NRH_ERR_UNSUPPORTED; }
    if (cfg->n_samples + cfg->n_importance > 128) { set_error("%s: more than 128 samples per ray", who); return NRH_ERR_UNSUPPORTED; }
    if (cfg->n_importance <= 0) { set_error("%s needs importance sampling (the reference then detaches the sample positions)", who); return NRH_ERR_UNSUPPORTED; }
    if (R < 0) { set_error("%s: negative ray count", who); return NRH_ERR_INVALID; }
    return NRH_OK;
}

size_t nrh_train_workspace_bytes(const NrhConfig* cfg, int64_t R) {
    if (!cfg || R <= 0) return 0;
    return carve_train(*cfg, R, nullptr, sms_or_default()).total_bytes;
}

static void fill_wn_table(const NrhConfig& cfg, const NrhTrainParams& p, WnTable& T, float* w_flat, const float* dW, bool backward) {
    // effective weights in w_flat (forward) in the order sdf 0..7, out_sdf, out_feat, col 0..4; dW offsets (backward) as DW_*
    const int cin = color_in_dim(cfg);
    T.n = 0;
    size_t woff = 0;
    auto add = [&](const NrhLayerParams& L, int rows, int cols, const float* dw, int dw_ld, int perm) {
        WnJob& J = T.j[T.n++];
        J.v = L.v; J.g = L.g; J.w = w_flat ? w_flat + woff : nullptr; woff += (size_t)rows * cols;
        J.dw = dw; J.dw_ld = dw_ld; J.dg = L.d_g; J.dv = L.d_v; J.rows = rows; J.cols = cols;
        J.perm = perm; J.perm_shadow = cfg.shadow_hint; J.perm_spec0 = 316 + (cfg.shadow_hint ? 9 : 0);
    };
    for (int l = 0; l < SDF_LAYERS; ++l) {
        const int in = l == 0 ? PE_DIM : 256, out = l == SDF_SKIP - 1 ? SKIP_H : 256;
        add(p.sdf[l], out, in, backward ? (l == 0 ? dW + DW_SDF0 : dW + DW_SDF + (size_t)(l - 1) * 65536) : nullptr, l == 0 ? 64 : 256, 0);
    }
    add(p.sdf_out, 1, 256, backward ? dW + DW_HEAD : nullptr, 256, 0);
    add(p.feat_out, 256, 256, backward ? dW + DW_FEAT : nullptr, 256, 0);
    add(p.col[0], 256, cin, backward ? dW + DW_COL0 : nullptr, 384, 1);
    for (int l = 1; l < 4; ++l) add(p.col[l], 256, 256, backward ? dW + DW_COL + (size_t)(l - 1) * 65536 : nullptr, 256, 0);
    add(p.col[4], 3, 256, backward ? dW + DW_COL4 : nullptr, 256, 0);
}

int nrh_pack_weights_wn(const NrhConfig* cfg, const NrhTrainParams* params, void* wn_scratch, size_t wn_scratch_bytes, void* packed,
                        size_t packed_bytes, void* stream) {
    int rc = nrh_check_config(cfg); if (rc) return rc;
    if (!params || !wn_scratch || !packed) { set_error("null argument"); return NRH_ERR_INVALID; }
    if (cfg->use_outside_nerf) { set_error("nrh_pack_weights_wn does not cover the outside NeRF (plain weights: use nrh_pack_weights)"); return NRH_ERR_UNSUPPORTED; }
    const int cin = color_in_dim(*cfg);
    const size_t need = ((size_t)256 * PE_DIM + 6 * 65536 + (size_t)SKIP_H * 256 + 256 + 65536 + (size_t)256 * cin + 3 * 65536 + 3 * 256) * sizeof(float);
    if (wn_scratch_bytes < need) { set_error("weight-norm scratch too small: %zu < %zu", wn_scratch_bytes, need); return NRH_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    static thread_local WnTable T;
    float* w = reinterpret_cast<float*>(wn_scratch);
    fill_wn_table(*cfg, *params, T, w, nullptr, false);
    for (int i = 0; i < T.n; ++i) if (!T.j[i].v || !T.j[i].g) { set_error("nrh_pack_weights_wn: null parameter"); return NRH_ERR_INVALID; }
    if ((rc = launch_weight_norm(T, false, st))) return rc;
    NrhRawWeights raw; memset(&raw, 0, sizeof(raw));
    for (int l = 0; l < SDF_LAYERS; ++l) { raw.sdf_W[l] = T.j[l].w; raw.sdf_b[l] = params->sdf[l].bias; }
    raw.sdf_out_W = T.j[8].w; raw.sdf_out_b = params->sdf_out.bias;
    raw.feat_W = T.j[9].w; raw.feat_b = params->feat_out.bias;
    for (int l = 0; l < 5; ++l) { raw.col_W[l] = T.j[10 + l].w; raw.col_b[l] = params->col[l].bias; }
    raw.variance = params->variance;
    const int launches = g_launches;
    rc = nrh_pack_weights(cfg, &raw, packed, packed_bytes, stream);
    g_launches += launches;
    return rc;
}

int nrh_render_train_forward(const NrhConfig* cfg, const void* packed, const NrhRays* rays, int64_t R, const float* bg_rgb,
                             const float* jitter_primary, const float* jitter_shadow, float cos_anneal, int warmup,
                             const NrhOutputs* out, void* train_ws, size_t train_ws_bytes, void* stream) {
    g_launches = 0;
    int rc = train_check(cfg, R, "nrh_render_train_forward"); if (rc) return rc;
    if (R == 0) return NRH_OK;
    if (!packed || !rays || !out || !train_ws) { set_error("null argument"); return NRH_ERR_INVALID; }
    if (!out->rgb || !out->weights) { set_error("nrh_render_train_forward: out->rgb and out->weights are required"); return NRH_ERR_INVALID; }
    if (out->sampled_color) { set_error("nrh_render_train_forward: sampled_color is not provided in training mode"); return NRH_ERR_UNSUPPORTED; }
    int sms; if ((rc = device_sms(&sms))) return rc;
    const TrainWs t = carve_train(*cfg, R, reinterpret_cast<char*>(train_ws), sms);
    if (train_ws_bytes < t.total_bytes) { set_error("train workspace too small: %zu < %zu", train_ws_bytes, t.total_bytes); return NRH_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    const int S = cfg->n_samples + cfg->n_importance;
    const int64_t N = (int64_t)S * R;
    // 1. everything the reference computes without gradients + the primary fine pass with its tape (NrhTrainCapture); the final
    //    sample positions come back ray-major in z_rm
    NrhTrainCapture cap{t.tape, t.tape_bytes, t.cap_sdf, t.cap_grad, nullptr, t.cap_pts, t.x16, 384};     // features: fp16 rows of x16
    NrhOutputs o = *out;
    o.train_capture = &cap;
    o.z_vals = t.z_rm;
    if ((rc = nrh_render_forward(cfg, packed, rays, R, bg_rgb, jitter_primary, nullptr, jitter_shadow, cos_anneal, warmup, &o,
                                 t.render, t.render_bytes, stream))) return rc;
    if (out->z_vals) NRH_CUDA_CHECK(cudaMemcpyAsync(out->z_vals, t.z_rm, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
    // 2. the differentiable tail: reflectance network (activations kept), sigmoid, NeuS alpha / weights / compositing
    const Workspace w = carve(*cfg, R, t.render, sms);
    const PackedLayout L = make_layout(*cfg);
    const float* Pf = reinterpret_cast<const float*>(packed);
    if ((rc = launch_train_dists(t.z_rm, R, S, 2.0f / (float)cfg->n_samples, t.dists, t.mid_z, st))) return rc;
    TrainAssembleArgs A{};
    A.N = N; A.R = R; A.gx = t.cap_grad; A.gy = t.cap_grad + N; A.gz = t.cap_grad + 2 * N;
    A.px = t.cap_pts; A.py = t.cap_pts + N; A.pz = t.cap_pts + 2 * N;
    A.feat = nullptr;                              // the feature block of x16 was written by the fine-pass kernel itself
    A.rayfeat = w.rayfeat; A.normalized = cfg->normalized_normals; A.x16 = t.x16; A.grad_aos = t.grad_aos;
    if ((rc = launch_train_assemble(A, st))) return rc;
    if ((rc = color_train_forward_tc(packed, L, t.x16, N, t.acts, t.y, sms, st, true))) return rc;
    if ((rc = launch_color_sigmoid(t.y, N, t.color, st))) return rc;
    if ((rc = nrh_composite_train_forward(t.cap_sdf, t.grad_aos, t.color, 1, R, t.dists, rays->directions, Pf + L.inv_s, cos_anneal,
                                          bg_rgb, R, S, out->weights, out->rgb, stream))) return rc;
    return NRH_OK;
}

int nrh_render_backward(const NrhConfig* cfg, const void* packed, const NrhTrainParams* params, const NrhRays* rays, int64_t R,
                        const float* bg_rgb, float cos_anneal, const NrhTrainAdjoints* adj, void* train_ws, size_t train_ws_bytes,
                        void* stream) {
    g_launches = 0;
    int rc = train_check(cfg, R, "nrh_render_backward"); if (rc) return rc;
    if (R == 0) return NRH_OK;
    if (!packed || !params || !rays || !adj || !adj->d_rgb || !train_ws) { set_error("null argument"); return NRH_ERR_INVALID; }
    static thread_local WnTable WT;
    int sms; if ((rc = device_sms(&sms))) return rc;
    const TrainWs t = carve_train(*cfg, R, reinterpret_cast<char*>(train_ws), sms);
    if (train_ws_bytes < t.total_bytes) { set_error("train workspace too small"); return NRH_ERR_WORKSPACE; }
    fill_wn_table(*cfg, *params, WT, nullptr, t.dW, true);
    for (int i = 0; i < WT.n; ++i)
        if (!WT.j[i].v || !WT.j[i].g || !WT.j[i].dg || !WT.j[i].dv) { set_error("nrh_render_backward: null parameter / gradient pointer"); return NRH_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int S = cfg->n_samples + cfg->n_importance;
    const int64_t N = (int64_t)S * R, Pp = t.lay.p_pad;
    const PackedLayout L = make_layout(*cfg);
    const float* Pf = reinterpret_cast<const float*>(packed);
    float* sm = t.small;
    unsigned int* smu = reinterpret_cast<unsigned int*>(t.small);
    const float* inv_ss = sm + SM_SCALE_S + 1;
    const float* inv_sc = sm + SM_SCALE_C + 1;
    NRH_CUDA_CHECK(cudaMemsetAsync(t.small, 0, sizeof(float) * SM_COUNT, st));
    NRH_CUDA_CHECK(cudaMemsetAsync(t.dW, 0, sizeof(float) * DW_TOTAL, st));
    // 1. compositor backward: d rgb (, d weights) -> d sdf, d grad (true_cos), d colour, d dirs, d inv_s
    if ((rc = nrh_composite_train_backward(t.cap_sdf, t.grad_aos, t.color, 1, R, t.dists, rays->directions, Pf + L.inv_s, cos_anneal, bg_rgb,
                                           R, S, adj->d_rgb, adj->d_weights, t.d_sdf, t.d_grad, t.d_color, t.d_dirs_c, sm + SM_DINVS, stream))) return rc;
    // 2. sigmoid backward, loss scale of the reflectance chain, reflectance backward (tcgen05)
    if ((rc = launch_sigmoid_bwd(t.d_color, t.color, N * 3, t.dy, st))) return rc;
    if ((rc = launch_absmax_f32(t.dy, N * 3, 1.0f, smu + SM_AMAX_C, st))) return rc;
    if ((rc = launch_pow2_scale(smu + SM_AMAX_C, 256.0f, nullptr, sm + SM_SCALE_C, st))) return rc;
    if ((rc = color_train_backward_tc(packed, L, t.dy, sm + SM_SCALE_C, t.acts, N, t.dz, t.dy16, t.dx16, sms, st, true))) return rc;
    // 3. backward of the input assembly: total adjoint of grad sdf, position adjoint, encodings -> d dirs / d light positions
    TrainScatterArgs C{};
    C.R = R; C.S = S; C.dx16 = t.dx16; C.scale_c = sm + SM_SCALE_C;
    C.gx = t.cap_grad; C.gy = t.cap_grad + N; C.gz = t.cap_grad + 2 * N;
    C.d_normals = adj->d_analytic_normals; C.d_nnormals = adj->d_normalized_normals;
    C.dirs = rays->directions; C.pl = rays->pl_positions; C.normalized = cfg->normalized_normals;
    C.d_grad = t.d_grad; C.d_pts = t.d_pts_r; C.d_dirs_pe = t.d_dirs_pe; C.d_pl = adj->d_pl_positions;
    if ((rc = launch_train_scatter(C, st))) return rc;
    // 4. loss scale of the SDF chain (largest adjoint -> ~2^9), SDF second-order backward (tcgen05)
    if ((rc = launch_absmax_f32(t.d_sdf, N, 1.0f, smu + SM_AMAX_S, st))) return rc;
    if ((rc = launch_absmax_f32(t.d_grad, N * 3, SDF_SCALE, smu + SM_AMAX_S, st))) return rc;
    if ((rc = launch_absmax_f16(t.dx16, N, 384, 0, 256, inv_sc, smu + SM_AMAX_S, st))) return rc;
    if ((rc = launch_pow2_scale(smu + SM_AMAX_S, 512.0f, sm + SM_SCALE_C, sm + SM_SCALE_S, st))) return rc;
    {
        const SdfBwdFeat16 f16{t.dx16, 384, sm + SM_SCALE_S + 2};
        const Strided3 P3{t.cap_pts, t.cap_pts + N, t.cap_pts + 2 * N, 1};
        if ((rc = sdf_train_backward_tc(packed, L, nullptr, N, t.tape, t.d_sdf, nullptr, t.d_grad, sm + SM_SCALE_S, t.bwd, t.d_pts_s,
                                        t.sdf_bwd_ws, t.sdf_bwd_ws_bytes, sms, st, &f16, &P3))) return rc;
    }
    // 5. d origins / d directions
    if (adj->d_origins || adj->d_directions)
        if ((rc = launch_train_ray_reduce(t.d_pts_r, t.d_pts_s, t.mid_z, R, S, t.d_dirs_c, t.d_dirs_pe, adj->d_origins, adj->d_directions, st))) return rc;
    // 6. weight gradients: every reduction over the dumps in one tcgen05 launch (+ two skinny ones)
    if ((rc = launch_pe_dump(t.cap_pts, t.cap_pts + N, t.cap_pts + 2 * N, N, Pp, t.e16, st))) return rc;
    if ((rc = launch_ds16(t.d_sdf, N, Pp, sm + SM_SCALE_S, t.ds16, st))) return rc;
    const char* tb = reinterpret_cast<const char*>(t.tape);
    const char* bb = reinterpret_cast<const char*>(t.bwd);
    const __half* act = reinterpret_cast<const __half*>(tb + t.lay.tape_act_off);      // a_1..a_8 (x16)
    const __half* u = reinterpret_cast<const __half*>(tb + t.lay.tape_u_off);          // u_0..u_7 (x1024)
    const __half* gb0 = reinterpret_cast<const __half*>(bb + t.lay.bwd_gb0_off);
    const __half* gb = reinterpret_cast<const __half*>(bb + t.lay.bwd_gb_off);
    const __half* zb = reinterpret_cast<const __half*>(bb + t.lay.bwd_zb_off);
    const size_t M = (size_t)Pp * 256, MC = (size_t)N * 256;
    {
        static thread_local NrhWgradJob J[40];
        int n = 0;
        auto job = [&](const void* a, int64_t a_ld, const void* b, int64_t b_ld, int b_col0, int64_t rows, int m, int nn, float scale,
                       const float* ds, float* outp, int64_t ld_out, int rows_valid) {
            NrhWgradJob& q = J[n++];
            q.a = a; q.a_ld = a_ld; q.a_col0 = 0; q.b = b; q.b_ld = b_ld; q.b_col0 = b_col0; q.rows = rows; q.m = m; q.n = nn;
            q.rows_valid = rows_valid; q.cols_valid = 0; q.scale = scale; q.dev_scale = ds; q.out = outp; q.ld_out = ld_out;
        };
        // SDF network: dW_l = u_l^T gb_{l-1} / (1024 S) + zb_l^T a_{l-1} / (16 S)   (a_0 = the Fourier features, unscaled)
        job(u, 256, gb0, 64, 0, Pp, 256, 64, 1.0f / 1024.0f, inv_ss, t.dW + DW_SDF0, 64, 0);
        job(zb, 256, t.e16, 64, 0, Pp, 256, 64, 1.0f, inv_ss, t.dW + DW_SDF0, 64, 0);
        for (int l = 1; l < SDF_LAYERS; ++l) {
            float* o = t.dW + DW_SDF + (size_t)(l - 1) * 65536;
            const int rv = l == SDF_SKIP - 1 ? SKIP_H : 0;
            job(u + l * M, 256, gb + (l - 1) * M, 256, 0, Pp, 256, 256, 1.0f / 1024.0f, inv_ss, o, 256, rv);
            job(zb + l * M, 256, act + (l - 1) * M, 256, 0, Pp, 256, 256, 1.0f / 16.0f, inv_ss, o, 256, rv);
        }
        // heads: feature head from the fp16 d_feat (S_c units) and a_8; sdf head row from d_sdf * S_s
        job(t.dx16, 384, act + 7 * M, 256, 0, N, 256, 256, 1.0f / 16.0f, inv_sc, t.dW + DW_FEAT, 256, 0);
        job(t.ds16, 8, act + 7 * M, 256, 0, Pp, 1, 256, 1.0f / 16.0f, inv_ss, t.dW + DW_HEAD, 256, 0);
        // reflectance network: dW_0 = dz_0^T x16 / S_c, dW_l = dz_l^T a_l / (16 S_c), dW_out = dy16^T a_4 / (16 S_c)
        job(t.dz, 256, t.x16, 384, 0, N, 256, 256, 1.0f, inv_sc, t.dW + DW_COL0, 384, 0);
        job(t.dz, 256, t.x16, 384, 256, N, 256, 128, 1.0f, inv_sc, t.dW + DW_COL0 + 256, 384, 0);
        for (int l = 1; l < 4; ++l)
            job(t.dz + l * MC, 256, t.acts + (l - 1) * MC, 256, 0, N, 256, 256, 1.0f / 16.0f, inv_sc, t.dW + DW_COL + (size_t)(l - 1) * 65536, 256, 0);
        job(t.dy16, 8, t.acts + 3 * MC, 256, 0, N, 3, 256, 1.0f / 16.0f, inv_sc, t.dW + DW_COL4, 256, 0);
        if ((rc = nrh_wgrad_f16(J, n, stream))) return rc;
    }
    // bias sums of the fp16 dumps: zb_0..zb_7, gb_8, dz_0..dz_3 (nrh_colsum_f16 zeroes its outputs), d_feat, d_sdf, dy
    if ((rc = nrh_colsum_f16(zb, 8, Pp, 256, (int64_t)M, 1.0f, t.colsum, stream))) return rc;
    if ((rc = nrh_colsum_f16(gb + 7 * M, 1, Pp, 256, (int64_t)M, 1.0f, t.colsum + 8 * 256, stream))) return rc;
    if ((rc = nrh_colsum_f16(t.dz, 4, N, 256, (int64_t)MC, 1.0f, t.colsum + 9 * 256, stream))) return rc;
    if ((rc = launch_colsum256_f16(t.dx16, N, 384, 0, t.dW + DW_FEATB, st))) return rc;
    if ((rc = launch_sum_f32(t.d_sdf, N, 1.0f, sm + SM_DBS, st))) return rc;
    if ((rc = launch_colsum3(t.dy, N, sm + SM_DB4, st))) return rc;
    // 7. parameter gradients: biases + the sdf head row d w_s = (sum gb_8 / S + d_sdf^T a_8) / 3 in one launch, weight-norm backward
    //    of all 15 layers in one launch, variance
    {
        static thread_local VecTable V;
        V.n = 0;
        auto vec = [&](const float* src, const float* ds, float mul, const float* add, float add_mul, float* dst, int n) {
            if (!dst) return;
            V.j[V.n++] = VecJob{src, ds, mul, add, add_mul, dst, n};
        };
        for (int l = 0; l < SDF_LAYERS; ++l) vec(t.colsum + l * 256, inv_ss, 1.0f, nullptr, 0.f, params->sdf[l].d_bias, l == SDF_SKIP - 1 ? SKIP_H : 256);
        vec(t.colsum + 8 * 256, inv_ss, 1.0f / SDF_SCALE, t.dW + DW_HEAD, 1.0f / SDF_SCALE, t.dW + DW_HEAD, 256);
        vec(sm + SM_DBS, nullptr, 1.0f / SDF_SCALE, nullptr, 0.f, params->sdf_out.d_bias, 1);
        vec(t.dW + DW_FEATB, inv_sc, 1.0f, nullptr, 0.f, params->feat_out.d_bias, 256);
        for (int l = 0; l < 4; ++l) vec(t.colsum + (9 + l) * 256, inv_sc, 1.0f, nullptr, 0.f, params->col[l].d_bias, 256);
        vec(sm + SM_DB4, nullptr, 1.0f, nullptr, 0.f, params->col[4].d_bias, 3);
        if ((rc = launch_vec_jobs(V, st))) return rc;
    }
    if ((rc = launch_weight_norm(WT, true, st))) return rc;
    if (params->d_variance && (rc = launch_variance_grad(sm + SM_DINVS, params->variance, params->d_variance, st))) return rc;
    return NRH_OK;
}

}  // extern "C"
