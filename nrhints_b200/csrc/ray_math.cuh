// Per-ray sampler / compositor math shared by the CUDA kernels and the host test harness
// (tests/host_harness.cpp compiles this header with g++ so the sampler logic is
// unit-tested against the oracle without a GPU).
//
// Reference semantics restated here (all in /root/reference/models/neus_hint_model.py):
//   sample_pdf(det=True)      :21-65      -> nrh::upsample_new_z (second pass)
//   up_sample                 :269-315    -> nrh::upsample_new_z (first pass)
//   cat_z_vals                :317-331    -> nrh::merge_sorted
//   get_alpha                 :333-357    -> nrh::neus_alpha
//   get_visibility            :373-432    -> nrh::shadow_ray_init / nrh::shadow_transmittance
//   render_core               :475-651    -> nrh::composite_primary / nrh::specular_cue
//   forward (coarse z)        :673-683    -> nrh::coarse_z
//   forward (outside z)       :677-694    -> nrh::outside_z
//   render_outside            :434-473    -> nrh::outside_sections / nrh::outside_point / nrh::outside_alpha
// Encoding: /root/reference/fields/encodings.py:168-176 -> nrh::fourier_encode
//
// All per-ray arrays are "sample-major": element j of ray r lives at base[j*stride + r]
// (stride = number of rays), so that one-thread-per-ray kernels are fully coalesced.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define NRH_HD __host__ __device__ __forceinline__
#else
#define NRH_HD inline
#endif

namespace nrh {

struct SoA {                      // strided view of one ray's samples
    float* p; int64_t stride;
    NRH_HD float& operator[](int j) const { return p[(int64_t)j * stride]; }
};
struct CSoA {
    const float* p; int64_t stride;
    NRH_HD float operator[](int j) const { return p[(int64_t)j * stride]; }
};

// torch.linspace(0, 1, n)[j] in fp32 (ATen RangeFactories: symmetric evaluation about the middle;
// the upper half is a fused multiply-add in ATen's vectorised kernel -- verified bitwise in tests).
NRH_HD float linspace01(int j, int n) {
    if (n <= 1) return 0.0f;
    const float step = 1.0f / (float)(n - 1);
    return (j < n / 2) ? step * (float)j : fmaf(-step, (float)(n - 1 - j), 1.0f);
}

NRH_HD float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
NRH_HD float relu_(float x) { return x > 0.0f ? x : 0.0f; }
NRH_HD float norm3(float x, float y, float z) { return sqrtf(x * x + y * y + z * z); }

// [x, sin(x 2^k), sin(x 2^k + pi/2)] for ONE scalar input, written with the reference's
// interleaving handled by the caller: out_sin[k], out_cos[k], k < F.
NRH_HD void fourier_scalar(float x, int F, float* out_sin, float* out_cos) {
    float f = 1.0f;
    for (int k = 0; k < F; ++k) {
        const float s = x * f;
        out_sin[k] = sinf(s);
        out_cos[k] = sinf(s + 1.57079637050628662109375f);
        f *= 2.0f;
    }
}

// NeRFEncoding over a D-vector: layout [x(D), sin(d-major,k-minor)(D*F), cos(D*F)]; dst strided.
NRH_HD void fourier_encode(const float* x, int D, int F, float* dst, int64_t dst_stride) {
    for (int d = 0; d < D; ++d) dst[(int64_t)d * dst_stride] = x[d];
    for (int d = 0; d < D; ++d) {
        float f = 1.0f;
        for (int k = 0; k < F; ++k) {
            const float s = x[d] * f;
            dst[(int64_t)(D + d * F + k) * dst_stride] = sinf(s);
            dst[(int64_t)(D + D * F + d * F + k) * dst_stride] = sinf(s + 1.57079637050628662109375f);
            f *= 2.0f;
        }
    }
}

// ------------------------------------------------------------------------------------------
// coarse samples: z_j = near + (far-near) * linspace01(j) (+ (jit-0.5)*2/n when training)
// ------------------------------------------------------------------------------------------
NRH_HD void coarse_z(float near, float far, int n, bool has_jitter, float jitter, SoA z) {
    const float span = far - near;
    const float off = has_jitter ? (jitter - 0.5f) * 2.0f / (float)n : 0.0f;
    for (int j = 0; j < n; ++j) {
        float v = near + span * linspace01(j, n);
        if (has_jitter) v = v + off;
        z[j] = v;
    }
}

// ------------------------------------------------------------------------------------------
// one importance step: (z sorted[k], sdf[k]) -> n_new new z (non-decreasing), in two parts so that the expensive,
// independent per-interval part can be spread over many threads:
//   interval_alpha(j)     : NeuS alpha of interval [z_j, z_{j+1}] with the fixed sharpness inv_s (up_sample :276-310)
//   sample_from_alphas()  : weights = alpha * exclusive cumprod, pdf/cdf, inverse-CDF at linspace(0,1,n) (:311-314, :21-65)
// upsample_new_z() = both, sequentially (host harness / reference order of operations).
// ------------------------------------------------------------------------------------------
NRH_HD float interval_alpha(const float o[3], const float d[3], int j, CSoA z, CSoA sdf, float inv_s) {
    const float zj = z[j], zn = z[j + 1], sj = sdf[j], sn = sdf[j + 1];
    const float rj = norm3(o[0] + d[0] * zj, o[1] + d[1] * zj, o[2] + d[2] * zj);
    const float rn = norm3(o[0] + d[0] * zn, o[1] + d[1] * zn, o[2] + d[2] * zn);
    const float inside = (rj < 1.0f || rn < 1.0f) ? 1.0f : 0.0f;
    const float mid_sdf = (sj + sn) * 0.5f;
    const float dist = zn - zj;
    const float cos_raw = (sn - sj) / (dist + 1e-5f);
    float prev_cos = 0.0f;
    if (j > 0) { const float zp = z[j - 1]; prev_cos = (sj - sdf[j - 1]) / (zj - zp + 1e-5f); }
    float c = fminf(prev_cos, cos_raw);
    c = fminf(fmaxf(c, -1e3f), 0.0f) * inside;
    const float half = c * dist * 0.5f;
    const float prev_cdf = sigmoidf_((mid_sdf - half) * inv_s);
    const float next_cdf = sigmoidf_((mid_sdf + half) * inv_s);
    return (prev_cdf - next_cdf + 1e-5f) / (prev_cdf + 1e-5f);
}

// The sampler of one importance step in three parts, so that the kernel can run the middle one on all its threads:
//   weights_from_alphas : wbuf[j] (k-1 alphas) -> w_j = alpha_j * T_j + 1e-5 (exclusive transmittance product), returns sum w
//   normalize_weights   : wbuf[j] -> w_j / wsum   (the pdf; one IEEE division per interval, independent of each other)
//   inverse_cdf_samples : running cdf over the pdf + inverse CDF at u = linspace(0,1,n_new)  (:21-65)
// `c_below + w / wsum` of the one-piece formulation is a division followed by an addition (nothing to contract), so hoisting
// the divisions out of the walk leaves every bit unchanged.
NRH_HD float weights_from_alphas(int k, SoA wbuf) {
    float T = 1.0f, wsum = 0.0f;
    for (int j = 0; j + 1 < k; ++j) {
        const float alpha = wbuf[j];
        const float w = alpha * T + 1e-5f;
        T = T * (1.0f - alpha + 1e-7f);
        wbuf[j] = w;
        wsum += w;
    }
    return wsum;
}
NRH_HD void normalize_weights(int k, SoA wbuf, float wsum) {
    for (int j = 0; j + 1 < k; ++j) wbuf[j] = wbuf[j] / wsum;
}
NRH_HD void inverse_cdf_samples(int k, CSoA z, int n_new, CSoA pdf, SoA z_new) {
    // cdf has k entries, cdf[0]=0.
    int m = 0;                 // running searchsorted(right=True) result
    float c_m = 0.0f;          // cdf[m]
    float c_below = 0.0f;      // cdf[m-1]
    for (int t = 0; t < n_new; ++t) {
        const float u = linspace01(t, n_new);
        while (m < k && c_m <= u) {
            c_below = c_m;
            ++m;
            if (m < k) c_m = c_below + pdf[m - 1];
        }
        const int below = m - 1 > 0 ? m - 1 : 0;
        const int above = m < k - 1 ? m : k - 1;
        const float c_above = (m < k) ? c_m : c_below;
        float denom = c_above - c_below;
        if (denom < 1e-5f) denom = 1.0f;
        const float tt = (u - c_below) / denom;
        const float zb = z[below], za = z[above];
        z_new[t] = zb + tt * (za - zb);
    }
}
// wbuf holds the k-1 alphas on entry and is overwritten with the pdf.
NRH_HD void sample_from_alphas(int k, CSoA z, int n_new, SoA wbuf, SoA z_new) {
    const float wsum = weights_from_alphas(k, wbuf);
    normalize_weights(k, wbuf, wsum);
    inverse_cdf_samples(k, z, n_new, CSoA{wbuf.p, wbuf.stride}, z_new);
}

NRH_HD void upsample_new_z(const float o[3], const float d[3], int k, CSoA z, CSoA sdf,
                           float inv_s, int n_new, SoA wbuf, SoA z_new) {
    for (int j = 0; j + 1 < k; ++j) wbuf[j] = interval_alpha(o, d, j, z, sdf, inv_s);
    sample_from_alphas(k, z, n_new, wbuf, z_new);
}

// ------------------------------------------------------------------------------------------
// merge two sorted runs (ties: old first). sdf arrays optional (last step merges z only).
// ------------------------------------------------------------------------------------------
NRH_HD void merge_sorted(int k, CSoA z_old, CSoA s_old, int n, CSoA z_new, CSoA s_new,
                         SoA z_out, SoA s_out, bool with_sdf) {
    int a = 0, b = 0;
    float za = z_old[0], zb = z_new[0];
    for (int i = 0; i < k + n; ++i) {
        const bool take_old = (b >= n) || (a < k && za <= zb);
        if (take_old) {
            z_out[i] = za;
            if (with_sdf) s_out[i] = s_old[a];
            ++a; if (a < k) za = z_old[a];
        } else {
            z_out[i] = zb;
            if (with_sdf) s_out[i] = s_new[b];
            ++b; if (b < n) zb = z_new[b];
        }
    }
}

// Parallel form of the same stable merge (ties: old first): the final slot of old entry a is a + #{new < z_old[a]}, that of new
// entry b is k + ... = b + #{old <= z_new[b]}; every (ray, entry) pair can compute its slot on its own.
NRH_HD int count_less(CSoA v, int n, float x) {           // #{i < n : v[i] < x}, v ascending
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (v[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}
NRH_HD int count_less_equal(CSoA v, int n, float x) {     // #{i < n : v[i] <= x}
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (v[mid] <= x) lo = mid + 1; else hi = mid; }
    return lo;
}

// The same merge, in place and from the back: z/s hold the k old entries in slots [0,k) and have room for
// k+n; the n new entries come from z_new/s_new.  Ties keep old before new, exactly like merge_sorted.
NRH_HD void merge_sorted_backward(int k, SoA z, SoA s, int n, CSoA z_new, CSoA s_new, bool with_sdf) {
    int a = k - 1, b = n - 1;
    for (int i = k + n - 1; i >= 0 && b >= 0; --i) {
        const float zb = z_new[b];
        const bool take_new = (a < 0) || (zb >= z[a]);
        if (take_new) {
            z[i] = zb;
            if (with_sdf) s[i] = s_new[b];
            --b;
        } else {
            z[i] = z[a];
            if (with_sdf) s[i] = s[a];
            --a;
        }
    }
}

// section length / mid-point of sample j (render_core :491-493, get_visibility :416-418)
NRH_HD void section(CSoA z, int j, int S, float last_dist, float& dist, float& mid) {
    const float zj = z[j];
    dist = (j + 1 < S) ? (z[j + 1] - zj) : last_dist;
    mid = zj + dist * 0.5f;
}

// get_alpha :339-356
NRH_HD float neus_alpha(float sdf, float gx, float gy, float gz, const float d[3], float dist,
                        float inv_s, float cos_anneal) {
    const float true_cos = d[0] * gx + d[1] * gy + d[2] * gz;
    const float iter_cos = -(relu_(-true_cos * 0.5f + 0.5f) * (1.0f - cos_anneal) + relu_(-true_cos) * cos_anneal);
    const float half = iter_cos * dist * 0.5f;
    const float prev_cdf = sigmoidf_((sdf - half) * inv_s);
    const float next_cdf = sigmoidf_((sdf + half) * inv_s);
    const float a = (prev_cdf - next_cdf + 1e-5f) / (prev_cdf + 1e-5f);
    return fminf(fmaxf(a, 0.0f), 1.0f);
}

// ------------------------------------------------------------------------------------------
// outside (NeRF++) model: sample positions beyond `far`, inverse-depth spaced (forward :677-694)
//   zo = linspace(1e-3, 1 - 1/(n_out+1), n_out); training: stratified jitter inside [lower, upper] bins;
//   z_out = far / flip(zo) + 1/n_samples.  `zo` (n_out entries) receives z_out in ascending order.
// ------------------------------------------------------------------------------------------
NRH_HD float linspace_general(float start, float end, int j, int n) {       // torch.linspace, scalar formula
    if (n <= 1) return start;
    const float step = (end - start) / (float)(n - 1);
    return (j < n / 2) ? start + step * (float)j : end - step * (float)(n - 1 - j);
}

NRH_HD void outside_z(float far, int n_samples, int n_out, bool has_jitter, CSoA jitter, float* zo) {
    const float start = 1e-3f, end = (float)(1.0 - 1.0 / ((double)n_out + 1.0));
    for (int j = 0; j < n_out; ++j) zo[j] = linspace_general(start, end, j, n_out);
    if (has_jitter) {
        float prev = zo[0];
        for (int j = 0; j < n_out; ++j) {
            const float cur = zo[j];
            const float next = (j + 1 < n_out) ? zo[j + 1] : cur;
            const float lower = (j == 0) ? cur : 0.5f * (cur + prev);
            const float upper = (j + 1 < n_out) ? 0.5f * (next + cur) : cur;
            prev = cur;
            zo[j] = lower + (upper - lower) * jitter[j];
        }
    }
    const float add = (float)(1.0 / (double)n_samples);
    // flip, then far / . + 1/n
    for (int j = 0; j < n_out / 2; ++j) { const float t = zo[j]; zo[j] = zo[n_out - 1 - j]; zo[n_out - 1 - j] = t; }
    for (int j = 0; j < n_out; ++j) zo[j] = far / zo[j] + add;
    // torch.sort of the concatenation follows; the run is ascending whenever far > 0 -- make it so in any case
    for (int i = 1; i < n_out; ++i) {
        const float v = zo[i]; int k = i - 1;
        while (k >= 0 && zo[k] > v) { zo[k + 1] = zo[k]; --k; }
        zo[k + 1] = v;
    }
}

// z_feed = sort(cat(z[S], z_out[n_out])); section lengths (last = sample_dist) and mid-points (render_outside :441-444)
NRH_HD void outside_sections(int S, CSoA z, int n_out, const float* zo, float sample_dist, SoA dist_out, SoA mid_out) {
    int a = 0, b = 0;
    float prev = 0.f;
    const int St = S + n_out;
    for (int i = 0; i < St; ++i) {
        float cur;
        if (b >= n_out || (a < S && z[a] <= zo[b])) { cur = z[a]; ++a; } else { cur = zo[b]; ++b; }
        if (i > 0) { const float dd = cur - prev; dist_out[i - 1] = dd; mid_out[i - 1] = prev + dd * 0.5f; }
        prev = cur;
    }
    dist_out[St - 1] = sample_dist; mid_out[St - 1] = prev + sample_dist * 0.5f;
}

// inverted-sphere parametrisation of a section mid-point (render_outside :447-450)
NRH_HD void outside_point(const float o[3], const float d[3], float mid, float p4[4]) {
    const float x = o[0] + d[0] * mid, y = o[1] + d[1] * mid, z = o[2] + d[2] * mid;
    const float dis = fminf(fmaxf(norm3(x, y, z), 1.0f), 1e10f);
    p4[0] = x / dis; p4[1] = y / dis; p4[2] = z / dis; p4[3] = 1.0f / dis;
}

// alpha = 1 - exp(-softplus(density) * dist)   (render_outside :461; F.softplus beta 1, threshold 20)
NRH_HD float outside_alpha(float density, float dist) {
    const float sp = density > 20.0f ? density : log1pf(expf(density));
    return 1.0f - expf(-sp * dist);
}

struct PrimaryComposite {
    float wsum, depth, nsum[3];
    float max_w, max_mid;          // largest weight and its section mid-point (first maximum, like torch.argmax)
};

// weights, inside mask, normals, depth (render_core :508-533, :583-587).
// sdf/grad are the fine-pass MLP outputs at the section mid-points.
// With the outside model (n_out > 0; bg_density / bg_dist hold S + n_out entries from render_outside) the NeuS alpha is
// replaced by the background alpha outside the unit sphere and the n_out far samples are appended (:517-519);
// w_out then has S + n_out entries, depth / normals still use the first S (`neus_weights`, :525).
// composite_primary in two parts, so that the kernel can evaluate the per-sample terms (two sigmoids, four divisions, two square roots
// per sample, independent of each other) on all its threads and keep only the transmittance scan on one thread per ray:
//   composite_sample     : alpha of sample j (blended with the background alpha outside the unit sphere), inside flag, unit normal
//   composite_accumulate : w = alpha * T, T update, running sums / arg-max -- in sample order
struct SampleTerms { float a, inside, n0, n1, n2, mid; };
NRH_HD SampleTerms composite_sample(const float o[3], const float d[3], int S, CSoA z, int j, float last_dist, float sdf_j,
                                    float g0, float g1, float g2, float inv_s, float cos_anneal, int n_out, float bg_density_j, float bg_dist_j) {
    SampleTerms t;
    float dist; section(z, j, S, last_dist, dist, t.mid);
    const float px = o[0] + d[0] * t.mid, py = o[1] + d[1] * t.mid, pz = o[2] + d[2] * t.mid;
    float a = neus_alpha(sdf_j, g0, g1, g2, d, dist, inv_s, cos_anneal);
    t.inside = norm3(px, py, pz) < 1.0f ? 1.0f : 0.0f;
    if (n_out > 0) a = a * t.inside + outside_alpha(bg_density_j, bg_dist_j) * (1.0f - t.inside);
    t.a = a;
    const float gn = fmaxf(norm3(g0, g1, g2), 1e-12f);       // F.normalize eps
    t.n0 = g0 / gn; t.n1 = g1 / gn; t.n2 = g2 / gn;
    return t;
}
NRH_HD float composite_accumulate(PrimaryComposite& r, float& T, float a, float mid, float n0, float n1, float n2) {
    const float w = a * T;
    T = T * (1.0f - a + 1e-7f);
    if (w > r.max_w) { r.max_w = w; r.max_mid = mid; }
    r.wsum += w; r.depth += mid * w;
    r.nsum[0] += n0 * w; r.nsum[1] += n1 * w; r.nsum[2] += n2 * w;
    return w;
}
NRH_HD void composite_init(PrimaryComposite& r) {
    r.wsum = 0.f; r.depth = 0.f; r.nsum[0] = r.nsum[1] = r.nsum[2] = 0.f;
    r.max_w = -1.f; r.max_mid = 0.f;
}
// the n_out samples of the outside model appended behind the S primary ones (:517-519)
NRH_HD void composite_append_outside(PrimaryComposite& r, float& T, int S, int n_out, CSoA bg_density, CSoA bg_dist, SoA w_out) {
    for (int j = S; j < S + n_out; ++j) {
        const float a = outside_alpha(bg_density[j], bg_dist[j]);
        const float w = a * T;
        T = T * (1.0f - a + 1e-7f);
        w_out[j] = w;
        r.wsum += w;
    }
}
NRH_HD PrimaryComposite composite_primary(const float o[3], const float d[3], int S, CSoA z, float last_dist,
                                          CSoA sdf, CSoA gx, CSoA gy, CSoA gz, float inv_s, float cos_anneal,
                                          SoA w_out, SoA inside_out, SoA nx, SoA ny, SoA nz,
                                          int n_out = 0, CSoA bg_density = CSoA{nullptr, 0}, CSoA bg_dist = CSoA{nullptr, 0}) {
    PrimaryComposite r; composite_init(r);
    float T = 1.0f;
    for (int j = 0; j < S; ++j) {
        const SampleTerms t = composite_sample(o, d, S, z, j, last_dist, sdf[j], gx[j], gy[j], gz[j], inv_s, cos_anneal, n_out,
                                               n_out > 0 ? bg_density[j] : 0.f, n_out > 0 ? bg_dist[j] : 0.f);
        w_out[j] = composite_accumulate(r, T, t.a, t.mid, t.n0, t.n1, t.n2);
        inside_out[j] = t.inside;
        nx[j] = t.n0; ny[j] = t.n1; nz[j] = t.n2;
    }
    composite_append_outside(r, T, S, n_out, bg_density, bg_dist, w_out);
    return r;
}

// shadow ray from the light to the hit point (get_visibility :380-395). Returns L (light distance).
NRH_HD float shadow_ray_init(const float pl[3], const float hit[3], int n, float offset, bool has_jitter,
                             CSoA jitter, float dir_out[3], SoA z) {
    const float vx = hit[0] - pl[0], vy = hit[1] - pl[1], vz = hit[2] - pl[2];
    const float L = norm3(vx, vy, vz);
    dir_out[0] = vx / L; dir_out[1] = vy / L; dir_out[2] = vz / L;
    const float keep = 1.0f - offset;
    if (!has_jitter) {
        for (int j = 0; j < n; ++j) z[j] = linspace01(j, n) * L * keep;
    } else {
        float zprev = 0.f, zcur = linspace01(0, n) * L * keep;
        for (int j = 0; j < n; ++j) {
            const float znext = (j + 1 < n) ? linspace01(j + 1, n) * L * keep : zcur;
            const float lower = (j == 0) ? zcur : 0.5f * (zcur + zprev);
            const float upper = (j + 1 < n) ? 0.5f * (znext + zcur) : zcur;
            z[j] = lower + (upper - lower) * jitter[j];
            zprev = zcur; zcur = znext;
        }
    }
    return L;
}

// taus[:, -1]: transmittance in front of the last sample (get_visibility :427-432); shadow_alpha = its per-sample term
NRH_HD float shadow_alpha(const float d[3], int S, CSoA z, int j, float last_dist, float sdf_j, float g0, float g1, float g2,
                          float inv_s, float cos_anneal) {
    float dist, mid; section(z, j, S, last_dist, dist, mid);
    return neus_alpha(sdf_j, g0, g1, g2, d, dist, inv_s, cos_anneal);
}
NRH_HD float shadow_transmittance(const float d[3], int S, CSoA z, float last_dist, CSoA sdf,
                                  CSoA gx, CSoA gy, CSoA gz, float inv_s, float cos_anneal) {
    float T = 1.0f;
    for (int j = 0; j + 1 < S; ++j) {
        const float a = shadow_alpha(d, S, z, j, last_dist, sdf[j], gx[j], gy[j], gz[j], inv_s, cos_anneal);
        T = T * (1.0f - a + 1e-7f);
    }
    return T;
}

NRH_HD void normalize3(const float v[3], float out[3]) {
    const float n = fmaxf(norm3(v[0], v[1], v[2]), 1e-12f);
    out[0] = v[0] / n; out[1] = v[1] / n; out[2] = v[2] / n;
}
NRH_HD float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

// Cook-Torrance cue for each roughness (render_core :588-616)
NRH_HD void specular_cue(const float hit_n[3], const float pl[3], const float hit[3], const float d[3],
                         int n_rough, const float* rough, float* cue) {
    float lv[3] = {pl[0] - hit[0], pl[1] - hit[1], pl[2] - hit[2]}, l[3], v[3], h[3];
    normalize3(lv, l);
    float nd[3] = {-d[0], -d[1], -d[2]};
    normalize3(nd, v);
    float hv[3] = {l[0] + v[0], l[1] + v[1], l[2] + v[2]};
    normalize3(hv, h);
    const float n_l = clamp01(hit_n[0] * l[0] + hit_n[1] * l[1] + hit_n[2] * l[2]);
    const float n_v = clamp01(hit_n[0] * v[0] + hit_n[1] * v[1] + hit_n[2] * v[2]);
    const float n_h = clamp01(hit_n[0] * h[0] + hit_n[1] * h[1] + hit_n[2] * h[2]);
    const float h_v = clamp01(h[0] * v[0] + h[1] * v[1] + h[2] * v[2]);
    const float n_h2 = n_h * n_h;
    const float om = 1.0f - h_v;
    const float om5 = (om * om) * (om * om) * om;
    const float fres = 0.04f + 0.96f * om5;
    for (int i = 0; i < n_rough; ++i) {
        const float r = rough[i];
        const float k = (r + 1.0f) * (r + 1.0f) / 8.0f;
        const float g = (n_v / (n_v * (1.0f - k) + k)) * (n_l / (n_l * (1.0f - k) + k));
        const float a2 = r * r;
        const float den = n_h2 * (a2 - 1.0f) + 1.0f;
        const float ndf = a2 / (3.14159274101257324f * (den * den));
        cue[i] = ndf * g * fres / (4.0f * n_v + 1e-3f);
    }
}

}  // namespace nrh
