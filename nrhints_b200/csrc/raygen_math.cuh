// Ray generation math shared by the CUDA kernels (raygen.cu) and the host test harness (tests/host_harness.cpp).
//
// Reference semantics restated here:
//   RayGenerator.forward        /root/reference/camera/ray_generator.py:75-150
//   exp_map_SO3xR3              /root/reference/camera/lie_groups.py:26-61
//   exp_map_SE3                 /root/reference/camera/lie_groups.py:65-116
// plus the hand-derived vector-Jacobian products of all three (the reference gets them from autograd): gradients of a
// ray's (origin, direction, light position, near, far) w.r.t. the per-image `cam_pose_adjustment` [6] and `pl_adjustment` [3].
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define NRH_HD __host__ __device__ __forceinline__
#else
#define NRH_HD inline
#endif

namespace nrh {

struct RayGenCamera { float fx, fy, cx, cy, zn, zf; };      // CameraModel (camera/camera_model.py:5-24)

constexpr int RAYGEN_OFF = 0, RAYGEN_SO3XR3 = 1, RAYGEN_SE3 = 2;

// ---- exp maps: tangent [6] = (translation-like 3, rotation 3) -> M [3][4] row-major ------------------------------------
// lie_groups.py:38-60
NRH_HD void exp_map_so3xr3(const float* tv, float* M) {
    const float w0 = tv[3], w1 = tv[4], w2 = tv[5];
    const float nrms = w0 * w0 + w1 * w1 + w2 * w2;
    const float th = sqrtf(nrms < 1e-4f ? 1e-4f : nrms);
    const float inv = 1.0f / th;
    const float fac1 = inv * sinf(th);
    const float fac2 = inv * inv * (1.0f - cosf(th));
    const float K[9] = {0.f, -w2, w1, w2, 0.f, -w0, -w1, w0, 0.f};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            const float k2 = K[i * 3] * K[j] + K[i * 3 + 1] * K[3 + j] + K[i * 3 + 2] * K[6 + j];
            M[i * 4 + j] = fac1 * K[i * 3 + j] + fac2 * k2 + (i == j ? 1.0f : 0.0f);
        }
    M[3] = tv[0]; M[7] = tv[1]; M[11] = tv[2];
}

// coefficients of exp_map_SE3 and their derivatives w.r.t. theta (the branch torch.where selects, lie_groups.py:80-109)
struct Se3Coef {
    float c, s1r, c2r, s1t, c2t, c3t;           // cosine, sine_by_theta, one_minus_cosine_by_theta2 (rotation / translation), theta_minus_sine_by_theta3
    float dc, ds1r, dc2r, ds1t, dc2t, dc3t;
};
NRH_HD Se3Coef se3_coef(float th) {
    Se3Coef q;
    const float th2 = th * th;
    if (th < 1e-2f) {
        q.c = 8.0f / (4.0f + th2) - 1.0f;
        q.s1r = 0.5f * q.c + 0.5f;
        q.c2r = 0.5f * q.s1r;
        q.s1t = 1.0f - th2 / 6.0f;
        q.c2t = 0.5f - th2 / 24.0f;
        q.c3t = 1.0f / 6.0f - th2 / 120.0f;
        q.dc = -16.0f * th / ((4.0f + th2) * (4.0f + th2));
        q.ds1r = 0.5f * q.dc; q.dc2r = 0.25f * q.dc;
        q.ds1t = -th / 3.0f; q.dc2t = -th / 12.0f; q.dc3t = -th / 60.0f;
    } else {
        const float s = sinf(th), c = cosf(th), th3 = th2 * th;
        q.c = c;
        q.s1r = s / th;
        q.c2r = (1.0f - c) / th2;
        q.s1t = q.s1r; q.c2t = q.c2r;
        q.c3t = (th - s) / th3;
        q.dc = -s;
        q.ds1r = (th * c - s) / th2;
        q.dc2r = (th * s - 2.0f * (1.0f - c)) / th3;
        q.ds1t = q.ds1r; q.dc2t = q.dc2r;
        q.dc3t = ((1.0f - c) * th - 3.0f * (th - s)) / (th2 * th2);
    }
    return q;
}
// lie_groups.py:76-115
NRH_HD void exp_map_se3(const float* tv, float* M) {
    const float v0 = tv[0], v1 = tv[1], v2 = tv[2], w0 = tv[3], w1 = tv[4], w2 = tv[5];
    const float th = sqrtf(w0 * w0 + w1 * w1 + w2 * w2);
    const Se3Coef q = se3_coef(th);
    const float w[3] = {w0, w1, w2};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[i * 4 + j] = q.c2r * w[i] * w[j] + (i == j ? q.c : 0.0f);
    const float t0 = q.s1r * w0, t1 = q.s1r * w1, t2 = q.s1r * w2;
    M[1] -= t2; M[4] += t2; M[2] += t1; M[8] -= t1; M[6] -= t0; M[9] += t0;
    const float cx = w1 * v2 - w2 * v1, cy = w2 * v0 - w0 * v2, cz = w0 * v1 - w1 * v0;     // w x v
    const float wv = w0 * v0 + w1 * v1 + w2 * v2;
    M[3] = q.s1t * v0 + q.c2t * cx + q.c3t * (w0 * wv);
    M[7] = q.s1t * v1 + q.c2t * cy + q.c3t * (w1 * wv);
    M[11] = q.s1t * v2 + q.c2t * cz + q.c3t * (w2 * wv);
}

// ---- vector-Jacobian products of the exp maps: gM [3][4] -> g_tv [6] -----------------------------------------------------
NRH_HD void exp_map_so3xr3_vjp(const float* tv, const float* gM, float* g) {
    const float w0 = tv[3], w1 = tv[4], w2 = tv[5];
    const float nrms = w0 * w0 + w1 * w1 + w2 * w2;
    const bool clamped = nrms < 1e-4f;                 // torch.clamp passes the gradient where nrms >= min
    const float th = sqrtf(clamped ? 1e-4f : nrms);
    const float inv = 1.0f / th, s = sinf(th), c = cosf(th);
    const float fac1 = inv * s, fac2 = inv * inv * (1.0f - c);
    const float K[9] = {0.f, -w2, w1, w2, 0.f, -w0, -w1, w0, 0.f};
    float G[9], K2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            G[i * 3 + j] = gM[i * 4 + j];
            K2[i * 3 + j] = K[i * 3] * K[j] + K[i * 3 + 1] * K[3 + j] + K[i * 3 + 2] * K[6 + j];
        }
    float g_fac1 = 0.f, g_fac2 = 0.f;
    for (int i = 0; i < 9; ++i) { g_fac1 += G[i] * K[i]; g_fac2 += G[i] * K2[i]; }
    // gK = fac1 G + fac2 (G K^T + K^T G)
    float gK[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float a = 0.f;
            for (int k = 0; k < 3; ++k) a += G[i * 3 + k] * K[j * 3 + k] + K[k * 3 + i] * G[k * 3 + j];
            gK[i * 3 + j] = fac1 * G[i * 3 + j] + fac2 * a;
        }
    float gw0 = gK[7] - gK[5], gw1 = gK[2] - gK[6], gw2 = gK[3] - gK[1];
    if (!clamped) {
        const float dfac1 = (th * c - s) * inv * inv;
        const float dfac2 = (th * s - 2.0f * (1.0f - c)) * inv * inv * inv;
        const float g_n = (g_fac1 * dfac1 + g_fac2 * dfac2) * 0.5f * inv;      // d theta / d nrms = 1 / (2 theta)
        gw0 += 2.0f * w0 * g_n; gw1 += 2.0f * w1 * g_n; gw2 += 2.0f * w2 * g_n;
    }
    g[0] = gM[3]; g[1] = gM[7]; g[2] = gM[11];
    g[3] = gw0; g[4] = gw1; g[5] = gw2;
}

NRH_HD void exp_map_se3_vjp(const float* tv, const float* gM, float* g) {
    const float v[3] = {tv[0], tv[1], tv[2]}, w[3] = {tv[3], tv[4], tv[5]};
    const float th = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const Se3Coef q = se3_coef(th);
    float G[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) G[i * 3 + j] = gM[i * 4 + j];
    const float gt[3] = {gM[3], gM[7], gM[11]};
    const float wxv[3] = {w[1] * v[2] - w[2] * v[1], w[2] * v[0] - w[0] * v[2], w[0] * v[1] - w[1] * v[0]};
    const float wv = w[0] * v[0] + w[1] * v[1] + w[2] * v[2];
    const float wg = w[0] * gt[0] + w[1] * gt[1] + w[2] * gt[2];
    // coefficient adjoints
    float g_c2r = 0.f;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) g_c2r += G[i * 3 + j] * w[i] * w[j];
    const float g_c = G[0] + G[4] + G[8];
    const float vee[3] = {G[7] - G[5], G[2] - G[6], G[3] - G[1]};            // adjoint of the skew part per unit of sine_by_theta
    const float g_s1r = vee[0] * w[0] + vee[1] * w[1] + vee[2] * w[2];
    const float g_s1t = gt[0] * v[0] + gt[1] * v[1] + gt[2] * v[2];
    const float g_c2t = gt[0] * wxv[0] + gt[1] * wxv[1] + gt[2] * wxv[2];
    const float g_c3t = wg * wv;
    const float g_th = g_c * q.dc + g_s1r * q.ds1r + g_c2r * q.dc2r + g_s1t * q.ds1t + g_c2t * q.dc2t + g_c3t * q.dc3t;
    // direct terms
    const float vxg[3] = {v[1] * gt[2] - v[2] * gt[1], v[2] * gt[0] - v[0] * gt[2], v[0] * gt[1] - v[1] * gt[0]};   // v x gt
    const float gxw[3] = {gt[1] * w[2] - gt[2] * w[1], gt[2] * w[0] - gt[0] * w[2], gt[0] * w[1] - gt[1] * w[0]};   // gt x w
    for (int i = 0; i < 3; ++i) {
        float sym = 0.f;
        for (int j = 0; j < 3; ++j) sym += (G[i * 3 + j] + G[j * 3 + i]) * w[j];
        float gw = q.c2r * sym + q.s1r * vee[i] + q.c2t * vxg[i] + q.c3t * (wv * gt[i] + wg * v[i]);
        if (th > 0.0f) gw += g_th * w[i] / th;            // torch.linalg.norm has a zero (sub)gradient at the origin
        g[3 + i] = gw;
        g[i] = q.s1t * gt[i] + q.c2t * gxw[i] + q.c3t * w[i] * wg;
    }
}

// ---- one ray ------------------------------------------------------------------------------------------------------
struct RayGenTables {               // per-image tables, all nullable (ray_generator.py:49-73)
    const float* cam_pose_noise;    // [Ncam,3,4]  buffer `cam_pose_noise`
    const float* pl_noise;          // [Ncam,3]    buffer `pl_noise`
    const float* cam_pose_adjustment;   // [Ncam,6] parameter (cam_opt_mode != off)
    const float* pl_adjustment;     // [Ncam,3]    parameter (pl_opt)
};

NRH_HD void apply_delta(const float* D, float* R, float* t) {     // R <- dR R ; t <- dt + dR t   (ray_generator.py:95-98,112-116)
    float Rn[9], tn[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Rn[i * 3 + j] = D[i * 4] * R[j] + D[i * 4 + 1] * R[3 + j] + D[i * 4 + 2] * R[6 + j];
        tn[i] = D[i * 4 + 3] + (D[i * 4] * t[0] + D[i * 4 + 1] * t[1] + D[i * 4 + 2] * t[2]);
    }
    for (int i = 0; i < 9; ++i) R[i] = Rn[i];
    for (int i = 0; i < 3; ++i) t[i] = tn[i];
}

struct RayGenState {                // forward intermediates the backward needs
    float dirs[3];                  // camera-space direction
    float R1[9], t1[3];             // pose after the noise stage, before the learned delta
    float v[3], vnorm;              // un-normalised world direction
    float o[3], d[3];
};

// pose: [4][4] row-major camera-to-world of this ray.  img < 0: no image index (video views: ray_generator.py:103-105).
NRH_HD void raygen_forward_one(const RayGenCamera& cam, int cam_opt_mode, bool override_near_far, float w_idx, float h_idx,
                               int64_t img, const float* pose, const float* pl_in, const RayGenTables& T,
                               RayGenState& S, float* pl, float& near, float& far) {
    const float x = w_idx + 0.5f, y = h_idx + 0.5f;
    S.dirs[0] = (x - cam.cx) / cam.fx;
    S.dirs[1] = -(y - cam.cy) / cam.fy;
    S.dirs[2] = -1.0f;
    float R[9], t[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = pose[i * 4 + j];
        t[i] = pose[i * 4 + 3];
    }
    if (T.cam_pose_noise && img >= 0) apply_delta(T.cam_pose_noise + img * 12, R, t);
    for (int i = 0; i < 9; ++i) S.R1[i] = R[i];
    for (int i = 0; i < 3; ++i) S.t1[i] = t[i];
    if (cam_opt_mode != RAYGEN_OFF && T.cam_pose_adjustment && img >= 0) {
        float M[12];
        if (cam_opt_mode == RAYGEN_SO3XR3) exp_map_so3xr3(T.cam_pose_adjustment + img * 6, M);
        else exp_map_se3(T.cam_pose_adjustment + img * 6, M);
        apply_delta(M, R, t);
    }
    for (int i = 0; i < 3; ++i) pl[i] = pl_in[i];
    if (T.pl_noise && img >= 0) for (int i = 0; i < 3; ++i) pl[i] = pl[i] + T.pl_noise[img * 3 + i];
    if (T.pl_adjustment && img >= 0) for (int i = 0; i < 3; ++i) pl[i] = pl[i] + T.pl_adjustment[img * 3 + i];
    for (int i = 0; i < 3; ++i) S.v[i] = S.dirs[0] * R[i * 3] + S.dirs[1] * R[i * 3 + 1] + S.dirs[2] * R[i * 3 + 2];
    S.vnorm = sqrtf(S.v[0] * S.v[0] + S.v[1] * S.v[1] + S.v[2] * S.v[2]);
    const float den = S.vnorm > 1e-12f ? S.vnorm : 1e-12f;              // F.normalize eps
    for (int i = 0; i < 3; ++i) { S.d[i] = S.v[i] / den; S.o[i] = t[i]; }
    if (override_near_far) {
        const float a = S.d[0] * S.d[0] + S.d[1] * S.d[1] + S.d[2] * S.d[2];
        const float b = 2.0f * (S.o[0] * S.d[0] + S.o[1] * S.d[1] + S.o[2] * S.d[2]);
        const float mid = 0.5f * (-b) / a;
        near = mid - 1.0f; far = mid + 1.0f;
    } else {
        near = cam.zn; far = cam.zf;
    }
}

// adjoints (g_o, g_d [3], g_near, g_far) -> adjoint of the learned delta tangent g_adj[6] (g_pl passes straight to pl_adjustment)
NRH_HD void raygen_backward_one(int cam_opt_mode, bool override_near_far, const RayGenState& S, const float* adj,
                                const float* g_o_in, const float* g_d_in, float g_near, float g_far, float* g_adj) {
    float go[3] = {g_o_in[0], g_o_in[1], g_o_in[2]}, gd[3] = {g_d_in[0], g_d_in[1], g_d_in[2]};
    if (override_near_far) {
        const float a = S.d[0] * S.d[0] + S.d[1] * S.d[1] + S.d[2] * S.d[2];
        const float b = 2.0f * (S.o[0] * S.d[0] + S.o[1] * S.d[1] + S.o[2] * S.d[2]);
        const float g_mid = g_near + g_far;
        const float g_b = -0.5f / a * g_mid, g_a = 0.5f * b / (a * a) * g_mid;
        for (int i = 0; i < 3; ++i) {
            go[i] += 2.0f * S.d[i] * g_b;
            gd[i] += 2.0f * S.o[i] * g_b + 2.0f * S.d[i] * g_a;
        }
    }
    // d = v / |v|
    float gv[3] = {0.f, 0.f, 0.f};
    if (S.vnorm > 1e-12f) {
        const float dg = S.d[0] * gd[0] + S.d[1] * gd[1] + S.d[2] * gd[2];
        for (int i = 0; i < 3; ++i) gv[i] = (gd[i] - S.d[i] * dg) / S.vnorm;
    } else {
        for (int i = 0; i < 3; ++i) gv[i] = gd[i] / 1e-12f;
    }
    // R = dR R1, t = dt + dR t1 ; v_i = sum_j R[i][j] dirs[j]  =>  g_dR[i][k] = gv[i] (R1 dirs)[k] + go[i] t1[k]
    float r1d[3];
    for (int k = 0; k < 3; ++k) r1d[k] = S.R1[k * 3] * S.dirs[0] + S.R1[k * 3 + 1] * S.dirs[1] + S.R1[k * 3 + 2] * S.dirs[2];
    float gM[12];
    for (int i = 0; i < 3; ++i) {
        for (int k = 0; k < 3; ++k) gM[i * 4 + k] = gv[i] * r1d[k] + go[i] * S.t1[k];
        gM[i * 4 + 3] = go[i];
    }
    if (cam_opt_mode == RAYGEN_SO3XR3) exp_map_so3xr3_vjp(adj, gM, g_adj);
    else exp_map_se3_vjp(adj, gM, g_adj);
}

}  // namespace nrh
