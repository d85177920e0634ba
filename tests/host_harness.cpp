// Host build of nrhints_b200/csrc/ray_math.cuh for the CPU test-suite (g++, no CUDA):
// the per-ray sampler / compositor math is header-only host+device code, so the exact
// functions the kernels call are unit-tested against the oracle without a GPU.
// TEST INFRASTRUCTURE ONLY -- never loaded by the product.
#include "../nrhints_b200/csrc/ray_math.cuh"
#include "../nrhints_b200/csrc/raygen_math.cuh"
#include "../nrhints_b200/csrc/composite_train_math.cuh"

using namespace nrh;

extern "C" {

void h_linspace01(int n, float* out) { for (int j = 0; j < n; ++j) out[j] = linspace01(j, n); }

void h_coarse_z(float near, float far, int n, int has_jitter, float jitter, float* z) {
    coarse_z(near, far, n, has_jitter != 0, jitter, SoA{z, 1});
}

void h_fourier_encode(const float* x, int D, int F, float* dst) { fourier_encode(x, D, F, dst, 1); }

void h_upsample(const float* o, const float* d, int k, const float* z, const float* sdf, float inv_s, int n_new,
                float* wbuf, float* z_new) {
    upsample_new_z(o, d, k, CSoA{z, 1}, CSoA{sdf, 1}, inv_s, n_new, SoA{wbuf, 1}, SoA{z_new, 1});
}

void h_merge(int k, const float* z_old, const float* s_old, int n, const float* z_new, const float* s_new,
             float* z_out, float* s_out, int with_sdf) {
    merge_sorted(k, CSoA{z_old, 1}, CSoA{s_old, 1}, n, CSoA{z_new, 1}, CSoA{s_new, 1}, SoA{z_out, 1}, SoA{s_out, 1}, with_sdf != 0);
}

void h_merge_backward(int k, float* z, float* s, int n, const float* z_new, const float* s_new, int with_sdf) {
    merge_sorted_backward(k, SoA{z, 1}, SoA{s, 1}, n, CSoA{z_new, 1}, CSoA{s_new, 1}, with_sdf != 0);
}

// the rank form of the stable merge (what the importance kernel runs on all its threads): every entry computes its own slot
void h_merge_rank(int k, const float* z_old, const float* s_old, int n, const float* z_new, const float* s_new, float* z_out, float* s_out) {
    for (int a = 0; a < k; ++a) { const int pos = a + count_less(CSoA{z_new, 1}, n, z_old[a]); z_out[pos] = z_old[a]; s_out[pos] = s_old[a]; }
    for (int b = 0; b < n; ++b) { const int pos = b + count_less_equal(CSoA{z_old, 1}, k, z_new[b]); z_out[pos] = z_new[b]; s_out[pos] = s_new[b]; }
}

void h_sections(const float* z, int S, float last_dist, float* dist, float* mid) {
    for (int j = 0; j < S; ++j) section(CSoA{z, 1}, j, S, last_dist, dist[j], mid[j]);
}

// returns wsum, depth, nsum[3] in res[5]
void h_composite_primary(const float* o, const float* d, int S, const float* z, float last_dist, const float* sdf,
                         const float* gx, const float* gy, const float* gz, float inv_s, float cos_anneal,
                         float* w, float* inside, float* nx, float* ny, float* nz, float* res) {
    PrimaryComposite pc = composite_primary(o, d, S, CSoA{z, 1}, last_dist, CSoA{sdf, 1}, CSoA{gx, 1}, CSoA{gy, 1}, CSoA{gz, 1},
                                            inv_s, cos_anneal, SoA{w, 1}, SoA{inside, 1}, SoA{nx, 1}, SoA{ny, 1}, SoA{nz, 1});
    res[0] = pc.wsum; res[1] = pc.depth; res[2] = pc.nsum[0]; res[3] = pc.nsum[1]; res[4] = pc.nsum[2];
}

float h_shadow_init(const float* pl, const float* hit, int n, float offset, int has_jitter, const float* jitter,
                    float* dir_out, float* z) {
    return shadow_ray_init(pl, hit, n, offset, has_jitter != 0, CSoA{jitter, 1}, dir_out, SoA{z, 1});
}

float h_shadow_transmittance(const float* d, int S, const float* z, float last_dist, const float* sdf, const float* gx,
                             const float* gy, const float* gz, float inv_s, float cos_anneal) {
    return shadow_transmittance(d, S, CSoA{z, 1}, last_dist, CSoA{sdf, 1}, CSoA{gx, 1}, CSoA{gy, 1}, CSoA{gz, 1}, inv_s, cos_anneal);
}

void h_specular_cue(const float* hit_n, const float* pl, const float* hit, const float* d, int n_rough, const float* rough, float* cue) {
    specular_cue(hit_n, pl, hit, d, n_rough, rough, cue);
}

void h_normalize3(const float* v, float* out) { normalize3(v, out); }

// ---- outside (NeRF++) model ------------------------------------------------------------------------------------
void h_outside_z(float far, int n_samples, int n_out, int has_jitter, const float* jitter, float* zo) {
    outside_z(far, n_samples, n_out, has_jitter != 0, CSoA{jitter, 1}, zo);
}
void h_outside_sections(int S, const float* z, int n_out, const float* zo, float sample_dist, float* dist, float* mid) {
    outside_sections(S, CSoA{z, 1}, n_out, zo, sample_dist, SoA{dist, 1}, SoA{mid, 1});
}
void h_outside_point(const float* o, const float* d, float mid, float* p4) { outside_point(o, d, mid, p4); }
float h_outside_alpha(float density, float dist) { return outside_alpha(density, dist); }

// composite with the background model: w has S + n_out entries
void h_composite_primary_bg(const float* o, const float* d, int S, const float* z, float last_dist, const float* sdf,
                            const float* gx, const float* gy, const float* gz, float inv_s, float cos_anneal,
                            float* w, float* inside, float* nx, float* ny, float* nz, float* res,
                            int n_out, const float* bg_density, const float* bg_dist) {
    PrimaryComposite pc = composite_primary(o, d, S, CSoA{z, 1}, last_dist, CSoA{sdf, 1}, CSoA{gx, 1}, CSoA{gy, 1}, CSoA{gz, 1},
                                            inv_s, cos_anneal, SoA{w, 1}, SoA{inside, 1}, SoA{nx, 1}, SoA{ny, 1}, SoA{nz, 1},
                                            n_out, CSoA{bg_density, 1}, CSoA{bg_dist, 1});
    res[0] = pc.wsum; res[1] = pc.depth; res[2] = pc.nsum[0]; res[3] = pc.nsum[1]; res[4] = pc.nsum[2];
}

// ---- ray generation (raygen_math.cuh): the loops of k_raygen_forward / k_raygen_backward on the host ---------------------------
// cam: fx fy cx cy zn zf.  Tables nullable.  img nullable.
void h_raygen_forward(const float* cam, int mode, int override_nf, long R, const float* w_idx, const float* h_idx, const long* img,
                      const float* poses, const float* pls, const float* pose_noise, const float* pl_noise, const float* adj,
                      const float* pl_adj, float* o, float* d, float* pl, float* near, float* far) {
    const RayGenCamera C{cam[0], cam[1], cam[2], cam[3], cam[4], cam[5]};
    const RayGenTables T{pose_noise, pl_noise, mode != RAYGEN_OFF ? adj : nullptr, pl_adj};
    for (long r = 0; r < R; ++r) {
        float pose[12];
        for (int i = 0; i < 12; ++i) pose[i] = poses[r * 16 + i];
        RayGenState S;
        raygen_forward_one(C, mode, override_nf != 0, w_idx[r], h_idx[r], img ? img[r] : -1, pose, pls + r * 3, T, S, pl + r * 3,
                           near[r], far[r]);
        for (int i = 0; i < 3; ++i) { o[r * 3 + i] = S.o[i]; d[r * 3 + i] = S.d[i]; }
    }
}
void h_raygen_backward(const float* cam, int mode, int override_nf, long R, const float* w_idx, const float* h_idx, const long* img,
                       const float* poses, const float* pose_noise, const float* adj, const float* g_o, const float* g_d,
                       const float* g_pl, const float* g_near, const float* g_far, float* d_adj, float* d_pl_adj) {
    const RayGenCamera C{cam[0], cam[1], cam[2], cam[3], cam[4], cam[5]};
    const RayGenTables T{pose_noise, nullptr, adj, nullptr};
    const float zero3[3] = {0.f, 0.f, 0.f};
    for (long r = 0; r < R; ++r) {
        if (!img || img[r] < 0) continue;
        if (d_adj) {
            float pose[12], pl[3], nr, fr, g[6];
            for (int i = 0; i < 12; ++i) pose[i] = poses[r * 16 + i];
            RayGenState S;
            raygen_forward_one(C, mode, override_nf != 0, w_idx[r], h_idx[r], img[r], pose, zero3, T, S, pl, nr, fr);
            raygen_backward_one(mode, override_nf != 0, S, adj + img[r] * 6, g_o + r * 3, g_d + r * 3, g_near[r], g_far[r], g);
            for (int i = 0; i < 6; ++i) d_adj[img[r] * 6 + i] += g[i];
        }
        if (d_pl_adj) for (int i = 0; i < 3; ++i) d_pl_adj[img[r] * 3 + i] += g_pl[r * 3 + i];
    }
}

// ---- differentiable compositing (composite_train_math.cuh): one ray, contiguous arrays ------------------------------------
static CtRay make_ct(int S, const float* sdf, const float* g, const float* c, const float* dist, const float* d, float inv_s,
                     float cos_anneal, const float* bg) {
    CtRay Y; Y.S = S; Y.sdf = sdf; Y.sdf_st = 1; Y.g = g; Y.g_st = 3; Y.c = c; Y.c_st = 3; Y.dist = dist; Y.dist_st = 1;
    for (int k = 0; k < 3; ++k) { Y.d[k] = d[k]; Y.bg[k] = bg ? bg[k] : 0.f; }
    Y.inv_s = inv_s; Y.cos_anneal = cos_anneal; Y.has_bg = bg != nullptr;
    return Y;
}
void h_composite_train_forward(int S, const float* sdf, const float* g, const float* c, const float* dist, const float* d, float inv_s,
                               float cos_anneal, const float* bg, float* w, float* rgb) {
    composite_train_forward(make_ct(S, sdf, g, c, dist, d, inv_s, cos_anneal, bg), w, 1, rgb);
}
float h_composite_train_backward(int S, const float* sdf, const float* g, const float* c, const float* dist, const float* d, float inv_s,
                                 float cos_anneal, const float* bg, const float* d_rgb, const float* d_w, float* d_sdf, float* d_g,
                                 float* d_c, float* d_dir) {
    float alpha_s[CT_MAX_S], T_s[CT_MAX_S];
    return composite_train_backward(make_ct(S, sdf, g, c, dist, d, inv_s, cos_anneal, bg), d_rgb, d_w, 1, d_sdf, d_g, d_c, d_dir, alpha_s, T_s);
}

}  // extern "C"
