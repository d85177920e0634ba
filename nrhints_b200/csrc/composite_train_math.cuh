// Differentiable compositing of the primary ray for a training step: NeuS alpha (get_alpha,
// /root/reference/models/neus_hint_model.py:339-356), transmittance scan and weights (:521-526), colour compositing with the
// background (:635-637) -- forward AND the hand-derived vector-Jacobian product the reference gets from autograd.
// Header-only host+device code: the CUDA kernels (train_ops.cu) and the CPU test harness (tests/host_harness.cpp) run the same
// functions.  Strided views: element j of this ray lives at p[j * stride].
#pragma once
#include "ray_math.cuh"

namespace nrh {

constexpr int CT_MAX_S = 128;                      // samples per ray handled by one call (= NRH_MAX_SAMPLES)

struct CtRay {                       // one ray's inputs
    int S;
    const float* sdf; int64_t sdf_st;             // [S]
    const float* g; int64_t g_st;                 // [S][3] gradient of the sdf at the section mid-points (xyz contiguous)
    const float* c; int64_t c_st;                 // [S][3] reflectance output (after the sigmoid)
    const float* dist; int64_t dist_st;           // [S] section lengths
    float d[3];                                   // ray direction
    float inv_s, cos_anneal;
    bool has_bg; float bg[3];
};

// forward: weights [S] (stride w_st) and rgb [3]; alpha / T are recomputed by the backward, nothing else is kept
NRH_HD void composite_train_forward(const CtRay& R, float* w, int64_t w_st, float* rgb) {
    float T = 1.0f, wsum = 0.0f, acc[3] = {0.f, 0.f, 0.f};
    for (int j = 0; j < R.S; ++j) {
        const float* g = R.g + j * R.g_st;
        const float a = neus_alpha(R.sdf[j * R.sdf_st], g[0], g[1], g[2], R.d, R.dist[j * R.dist_st], R.inv_s, R.cos_anneal);
        const float wj = a * T;
        T = T * (1.0f - a + 1e-7f);
        w[j * w_st] = wj;
        wsum += wj;
        const float* c = R.c + j * R.c_st;
        acc[0] += c[0] * wj; acc[1] += c[1] * wj; acc[2] += c[2] * wj;
    }
    for (int k = 0; k < 3; ++k) rgb[k] = acc[k] + (R.has_bg ? R.bg[k] * (1.0f - wsum) : 0.0f);
}

// backward: adjoints d_rgb [3] and d_w [S] (nullable) -> d_sdf [S], d_g [S][3], d_c [S][3] (strides as the inputs'), d_dir [3],
// and the ray's contribution to d inv_s (returned).  alpha_s / T_s: caller-provided scratch of S floats each.
NRH_HD float composite_train_backward(const CtRay& R, const float* d_rgb, const float* d_w, int64_t dw_st,
                                      float* d_sdf, float* d_g, float* d_c, float* d_dir, float* alpha_s, float* T_s) {
    float T = 1.0f;
    for (int j = 0; j < R.S; ++j) {               // recompute the scan
        const float* g = R.g + j * R.g_st;
        const float a = neus_alpha(R.sdf[j * R.sdf_st], g[0], g[1], g[2], R.d, R.dist[j * R.dist_st], R.inv_s, R.cos_anneal);
        alpha_s[j] = a; T_s[j] = T;
        T = T * (1.0f - a + 1e-7f);
    }
    float AT = 0.0f, d_s = 0.0f;                  // adjoint of T_{j+1}; accumulated adjoint of inv_s
    d_dir[0] = d_dir[1] = d_dir[2] = 0.0f;
    const float r = R.cos_anneal, s = R.inv_s;
    for (int j = R.S - 1; j >= 0; --j) {
        const float a = alpha_s[j], Tj = T_s[j], q = 1.0f - a + 1e-7f;
        const float* c = R.c + j * R.c_st;
        const float wj = a * Tj;
        float dw = d_w ? d_w[j * dw_st] : 0.0f;
        for (int k = 0; k < 3; ++k) {
            d_c[j * R.c_st + k] = wj * d_rgb[k];
            dw += d_rgb[k] * (c[k] - (R.has_bg ? R.bg[k] : 0.0f));
        }
        const float d_alpha = (dw - AT) * Tj;    // w_j = a T_j ; T_{j+1} = T_j (1 - a + 1e-7)
        AT = dw * a + AT * q;
        // alpha = clip(x, 0, 1), x = (p - n + 1e-5) / (p + 1e-5)
        const float* g = R.g + j * R.g_st;
        const float sdf = R.sdf[j * R.sdf_st], dist = R.dist[j * R.dist_st];
        const float tc = R.d[0] * g[0] + R.d[1] * g[1] + R.d[2] * g[2];
        const float u = -tc * 0.5f + 0.5f;
        const float iter_cos = -(relu_(u) * (1.0f - r) + relu_(-tc) * r);
        const float half = iter_cos * dist * 0.5f;
        const float ap = (sdf - half) * s, an = (sdf + half) * s;
        const float p = sigmoidf_(ap), n = sigmoidf_(an);
        const float den = p + 1e-5f, x = (p - n + 1e-5f) / den;
        const float d_x = (x >= 0.0f && x <= 1.0f) ? d_alpha : 0.0f;
        const float d_p = d_x * n / (den * den), d_n = -d_x / den;
        const float d_ap = d_p * p * (1.0f - p), d_an = d_n * n * (1.0f - n);
        d_sdf[j * R.sdf_st] = (d_ap + d_an) * s;
        d_s += d_ap * (sdf - half) + d_an * (sdf + half);
        const float d_half = (d_an - d_ap) * s;
        const float d_iter = d_half * dist * 0.5f;
        const float d_tc = d_iter * ((u > 0.0f ? 0.5f * (1.0f - r) : 0.0f) + (tc < 0.0f ? r : 0.0f));
        for (int k = 0; k < 3; ++k) {
            d_g[j * R.g_st + k] = d_tc * R.d[k];
            d_dir[k] += d_tc * g[k];
        }
    }
    return d_s;
}

}  // namespace nrh
