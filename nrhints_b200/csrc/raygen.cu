// Ray generation in front of the ray-march path: RayGenerator.forward
// (/root/reference/camera/ray_generator.py:75-150) with the Lie-group exp maps of camera/lie_groups.py:26-116, and its
// backward (gradients of the rays w.r.t. the per-image `cam_pose_adjustment` / `pl_adjustment` parameters).
//
// The reference issues ~25 small ATen kernels per batch (stack, gather, bmm, normalize, sums) plus the autograd graph of the
// exp map; here it is ONE launch forward and ONE launch backward, one thread per ray.  Both are trivially HBM/launch bound:
// 104 B in + 44 B out per ray forward.  The backward re-derives the forward intermediates from the same inputs (cheaper than a
// tape) and reduces the per-image gradients inside the warp before touching global memory when all lanes of a warp share an image
// index (SAME_IMAGE pixel sampling, data/data_loader.py:66-68; the reference trainer's ALL_IMAGES sampling, trainer/trainer.py:118-125,
// mixes images inside a warp and takes the per-lane atomic path).
#include <cuda_runtime.h>
#include <stdint.h>
#include "nrh_common.cuh"
#include "raygen_math.cuh"

namespace nrh {
namespace {

constexpr int RG_THREADS = 128;

struct RayGenArgs {
    RayGenCamera cam;
    int cam_opt_mode, override_near_far;
    const float* w_idx; const float* h_idx; const int64_t* img_idx;
    const float* poses; const float* pls;
    RayGenTables T;
    int64_t n_cameras;
};

__device__ __forceinline__ int64_t image_of(const RayGenArgs& A, int64_t r) {
    if (!A.img_idx) return -1;
    int64_t i = A.img_idx[r];
    if (i < 0) i += A.n_cameras;                        // negative indices wrap, as torch indexing does in the reference (:92-98,:121-126)
    // indices outside [-n, n) raise an IndexError in the reference; RayGenerator.forward asserts the range on the device
    // (validate_indices), and the kernel never dereferences such an index: the ray then gets no noise / adjustment / gradient
    return (i >= 0 && i < A.n_cameras) ? i : -1;
}

__global__ void __launch_bounds__(RG_THREADS)
k_raygen_forward(RayGenArgs A, int64_t R, float* __restrict__ o, float* __restrict__ d, float* __restrict__ pl,
                 float* __restrict__ near, float* __restrict__ far) {
    const int64_t r = (int64_t)blockIdx.x * RG_THREADS + threadIdx.x;
    if (r >= R) return;
    float pose[12];
    const float4* p4 = reinterpret_cast<const float4*>(A.poses + r * 16);
#pragma unroll
    for (int i = 0; i < 3; ++i) { const float4 q = __ldg(p4 + i); pose[i * 4] = q.x; pose[i * 4 + 1] = q.y; pose[i * 4 + 2] = q.z; pose[i * 4 + 3] = q.w; }
    const float pl_in[3] = {A.pls[r * 3], A.pls[r * 3 + 1], A.pls[r * 3 + 2]};
    RayGenState S;
    float plo[3], nr, fr;
    raygen_forward_one(A.cam, A.cam_opt_mode, A.override_near_far != 0, A.w_idx[r], A.h_idx[r], image_of(A, r), pose, pl_in, A.T,
                       S, plo, nr, fr);
#pragma unroll
    for (int i = 0; i < 3; ++i) { o[r * 3 + i] = S.o[i]; d[r * 3 + i] = S.d[i]; pl[r * 3 + i] = plo[i]; }
    near[r] = nr; far[r] = fr;
}

__global__ void __launch_bounds__(RG_THREADS)
k_raygen_backward(RayGenArgs A, int64_t R, const float* __restrict__ g_o, const float* __restrict__ g_d,
                  const float* __restrict__ g_pl, const float* __restrict__ g_near, const float* __restrict__ g_far,
                  float* __restrict__ d_cam, float* __restrict__ d_pl) {
    const int64_t r = (int64_t)blockIdx.x * RG_THREADS + threadIdx.x;
    const bool live = r < R;
    int64_t img = -1;
    float g[9];                                       // [0..5] cam_pose_adjustment row, [6..8] pl_adjustment row
#pragma unroll
    for (int i = 0; i < 9; ++i) g[i] = 0.f;
    if (live) {
        img = image_of(A, r);
        if (img >= 0) {
            if (d_cam) {
                float pose[12];
                const float4* p4 = reinterpret_cast<const float4*>(A.poses + r * 16);
#pragma unroll
                for (int i = 0; i < 3; ++i) { const float4 q = __ldg(p4 + i); pose[i * 4] = q.x; pose[i * 4 + 1] = q.y; pose[i * 4 + 2] = q.z; pose[i * 4 + 3] = q.w; }
                const float pl_in[3] = {0.f, 0.f, 0.f};
                RayGenState S;
                float plo[3], nr, fr;
                raygen_forward_one(A.cam, A.cam_opt_mode, A.override_near_far != 0, A.w_idx[r], A.h_idx[r], img, pose, pl_in, A.T,
                                   S, plo, nr, fr);
                const float go[3] = {g_o ? g_o[r * 3] : 0.f, g_o ? g_o[r * 3 + 1] : 0.f, g_o ? g_o[r * 3 + 2] : 0.f};
                const float gd[3] = {g_d ? g_d[r * 3] : 0.f, g_d ? g_d[r * 3 + 1] : 0.f, g_d ? g_d[r * 3 + 2] : 0.f};
                raygen_backward_one(A.cam_opt_mode, A.override_near_far != 0, S, A.T.cam_pose_adjustment + img * 6, go, gd,
                                    g_near ? g_near[r] : 0.f, g_far ? g_far[r] : 0.f, g);
            }
            if (d_pl && g_pl) { g[6] = g_pl[r * 3]; g[7] = g_pl[r * 3 + 1]; g[8] = g_pl[r * 3 + 2]; }
        }
    }
    // warp-level reduction when the whole warp writes the same image row (the common case), atomics otherwise
    const unsigned full = 0xffffffffu;
    const int64_t img0 = __shfl_sync(full, img, 0);
    const bool uniform = __all_sync(full, img == img0);
    if (uniform) {
        if (img0 < 0) return;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            float v = g[i];
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(full, v, s);
            g[i] = v;
        }
        if ((threadIdx.x & 31) == 0) {
            if (d_cam) for (int i = 0; i < 6; ++i) atomicAdd(d_cam + img0 * 6 + i, g[i]);
            if (d_pl) for (int i = 0; i < 3; ++i) atomicAdd(d_pl + img0 * 3 + i, g[6 + i]);
        }
    } else if (img >= 0) {
        if (d_cam) for (int i = 0; i < 6; ++i) atomicAdd(d_cam + img * 6 + i, g[i]);
        if (d_pl) for (int i = 0; i < 3; ++i) atomicAdd(d_pl + img * 3 + i, g[6 + i]);
    }
}

int fill_args(RayGenArgs& A, const NrhCamera* cam, int cam_opt_mode, int override_near_far, const NrhRayGenInputs* in,
              const char* who) {
    if (!cam || !in || !in->w_indices || !in->h_indices || !in->poses || !in->pls) {
        set_error("%s: null argument", who); return NRH_ERR_INVALID;
    }
    if (cam_opt_mode < NRH_CAM_OPT_OFF || cam_opt_mode > NRH_CAM_OPT_SE3) { set_error("%s: unknown cam_opt_mode %d", who, cam_opt_mode); return NRH_ERR_INVALID; }
    if (cam_opt_mode != NRH_CAM_OPT_OFF && in->img_indices && !in->cam_pose_adjustment) {
        set_error("%s: cam_opt_mode needs cam_pose_adjustment", who); return NRH_ERR_INVALID;
    }
    if ((reinterpret_cast<uintptr_t>(in->poses) & 15) != 0) { set_error("%s: poses must be 16-byte aligned", who); return NRH_ERR_INVALID; }
    A.cam = RayGenCamera{cam->fx, cam->fy, cam->cx, cam->cy, cam->zn, cam->zf};
    A.cam_opt_mode = cam_opt_mode; A.override_near_far = override_near_far;
    A.w_idx = in->w_indices; A.h_idx = in->h_indices; A.img_idx = in->img_indices;
    A.poses = in->poses; A.pls = in->pls;
    A.T = RayGenTables{in->cam_pose_noise, in->pl_noise, cam_opt_mode != NRH_CAM_OPT_OFF ? in->cam_pose_adjustment : nullptr, in->pl_adjustment};
    A.n_cameras = in->n_cameras;
    return NRH_OK;
}

}  // namespace
}  // namespace nrh

using namespace nrh;

extern "C" {

int nrh_raygen_forward(const NrhCamera* cam, int cam_opt_mode, int override_near_far, const NrhRayGenInputs* in, int64_t R,
                       float* origins, float* directions, float* pl_positions, float* nears, float* fars, void* stream) {
    if (R == 0) return NRH_OK;
    RayGenArgs A; int rc = fill_args(A, cam, cam_opt_mode, override_near_far, in, "nrh_raygen_forward"); if (rc) return rc;
    if (R < 0 || !origins || !directions || !pl_positions || !nears || !fars) { set_error("nrh_raygen_forward: null output"); return NRH_ERR_INVALID; }
    k_raygen_forward<<<(unsigned)((R + RG_THREADS - 1) / RG_THREADS), RG_THREADS, 0, (cudaStream_t)stream>>>(
        A, R, origins, directions, pl_positions, nears, fars);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int nrh_raygen_backward(const NrhCamera* cam, int cam_opt_mode, int override_near_far, const NrhRayGenInputs* in, int64_t R,
                        const float* d_origins, const float* d_directions, const float* d_pl_positions, const float* d_nears,
                        const float* d_fars, float* d_cam_pose_adjustment, float* d_pl_adjustment, void* stream) {
    if (R == 0) return NRH_OK;
    RayGenArgs A; int rc = fill_args(A, cam, cam_opt_mode, override_near_far, in, "nrh_raygen_backward"); if (rc) return rc;
    if (R < 0) { set_error("nrh_raygen_backward: R < 0"); return NRH_ERR_INVALID; }
    if (!in->img_indices || (!d_cam_pose_adjustment && !d_pl_adjustment)) return NRH_OK;        // nothing is differentiable
    if (d_cam_pose_adjustment && cam_opt_mode == NRH_CAM_OPT_OFF) { set_error("nrh_raygen_backward: d_cam_pose_adjustment with cam_opt_mode off"); return NRH_ERR_INVALID; }
    k_raygen_backward<<<(unsigned)((R + RG_THREADS - 1) / RG_THREADS), RG_THREADS, 0, (cudaStream_t)stream>>>(
        A, R, d_origins, d_directions, d_pl_positions, d_nears, d_fars, d_cam_pose_adjustment, d_pl_adjustment);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

}  // extern "C"
