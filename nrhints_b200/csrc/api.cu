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

// dst[(row0 + k)*dst_ld + n] = src[n*src_ld + col0 + k]   (k < ncols, n < nrows)
__global__ void k_transpose_slice(const float* __restrict__ src, int src_ld, int nrows, int col0, int ncols,
                                  float* __restrict__ dst, int dst_ld, int row0) {
    const int total = nrows * ncols;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = i / ncols, k = i % ncols;
        dst[(size_t)(row0 + k) * dst_ld + n] = src[(size_t)n * src_ld + col0 + k];
    }
}
// dst[r*dst_ld + c] = src[r*cols + c]
__global__ void k_copy_rows(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst, int dst_ld) {
    const int total = rows * cols;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int r = i / cols, c = i % cols;
        dst[(size_t)r * dst_ld + c] = src[i];
    }
}
__global__ void k_inv_s(const float* __restrict__ variance, float* __restrict__ dst) {
    const float s = fminf(fmaxf(expf(variance[0] * 10.0f), 1e-6f), 1e6f);     // get_alpha :337
    dst[0] = s; dst[1] = 1.0f / s;
}
__global__ void k_copy_scalar(const float* __restrict__ src, float* __restrict__ dst) { dst[0] = src[0]; }

int transpose_slice(const float* src, int src_ld, int nrows, int col0, int ncols, float* dst, int dst_ld, int row0, cudaStream_t st) {
    if (nrows * ncols == 0) return NRH_OK;
    k_transpose_slice<<<(nrows * ncols + 255) / 256, 256, 0, st>>>(src, src_ld, nrows, col0, ncols, dst, dst_ld, row0);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}
int copy_rows(const float* src, int rows, int cols, float* dst, int dst_ld, cudaStream_t st) {
    k_copy_rows<<<(rows * cols + 255) / 256, 256, 0, st>>>(src, rows, cols, dst, dst_ld);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
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
    NRH_CUDA_CHECK(cudaMemsetAsync(packed, 0, L.total_bytes, st));
    k_inv_s<<<1, 1, 0, st>>>(raw->variance, P + L.inv_s); NRH_LAUNCH_CHECK();
    for (int l = 0; l < SDF_LAYERS; ++l) {
        const int in = (l == 0) ? PE_DIM : 256, out = (l == SDF_SKIP - 1) ? SKIP_H : 256;
        if ((rc = transpose_slice(raw->sdf_W[l], in, out, 0, in, P + L.sdf_wt[l], 256, 0, st))) return rc;
        if ((rc = copy_rows(raw->sdf_b[l], 1, out, P + L.sdf_b[l], 256, st))) return rc;
        if ((rc = copy_rows(raw->sdf_W[l], out, in, P + L.sdf_wn[l], l == 0 ? PE_PAD : 256, st))) return rc;
    }
    if ((rc = copy_rows(raw->sdf_out_W, 1, 256, P + L.head_w, 256, st))) return rc;
    k_copy_scalar<<<1, 1, 0, st>>>(raw->sdf_out_b, P + L.head_b); NRH_LAUNCH_CHECK();
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
        k_copy_scalar<<<1, 1, 0, st>>>(raw->nerf_alpha_b, P + L.nerf_alpha_b); NRH_LAUNCH_CHECK();
        if ((rc = transpose_slice(raw->nerf_feat_W, 256, 256, 0, 256, P + L.nerf_feat_wt, 256, 0, st))) return rc;
        if ((rc = copy_rows(raw->nerf_feat_b, 1, 256, P + L.nerf_feat_b, 256, st))) return rc;
        if ((rc = transpose_slice(raw->nerf_view_W, 256 + NERF_VPE, NERF_VIEW_H, 0, 256, P + L.nerf_view_wta, 256, 0, st))) return rc;
        if ((rc = transpose_slice(raw->nerf_view_W, 256 + NERF_VPE, NERF_VIEW_H, 256, NERF_VPE, P + L.nerf_view_wtb, 256, 0, st))) return rc;
        if ((rc = copy_rows(raw->nerf_view_b, 1, NERF_VIEW_H, P + L.nerf_view_b, 256, st))) return rc;
        if ((rc = transpose_slice(raw->nerf_rgb_W, NERF_VIEW_H, 3, 0, NERF_VIEW_H, P + L.nerf_rgb_wt, 4, 0, st))) return rc;
        if ((rc = copy_rows(raw->nerf_rgb_b, 1, 3, P + L.nerf_rgb_b, 4, st))) return rc;
    }
    if (tc_available()) { if ((rc = tc_pack(*cfg, L, *raw, packed, st))) return rc; }
    return NRH_OK;
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
        if (!cap->tape || !cap->sdf || !cap->grad_soa || !cap->feat || !cap->pts_soa) { set_error("train_capture: null buffer"); return NRH_ERR_INVALID; }
        if (cap->tape_bytes < sdf_train_layout((int64_t)S * R, sms).tape_bytes) { set_error("train_capture: tape too small"); return NRH_ERR_WORKSPACE; }
    }
    const bool streamed = !cap && resolve_impl(*cfg) == NRH_MLP_TCGEN05 && (R % 128 == 0);
    if (out->fine_begin_event) NRH_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(out->fine_begin_event), st));
    if (cap) {
        // the training forward (same arithmetic + tape) straight into the caller's buffers; the compositor reads them there
        const int64_t N = (int64_t)S * R;
        if ((rc = sdf_train_forward_tc_strided(packed, L, P, N, cap->sdf, cap->grad_soa, cap->grad_soa + N, cap->grad_soa + 2 * N, 1,
                                               cap->feat, cap->tape, w.mlp_scratch, w.mlp_scratch_bytes, sms, st))) return rc;
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

}  // extern "C"
