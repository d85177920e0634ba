/* nrhints_b200 -- C ABI of the B200-native NRHints ray-march hot path.
 *
 * The reference (iamNCJ/NRHints, commit 291800d) has no FFI layer: its boundary for this path
 * is the Python nn.Module `NeuSHintRenderer` (models/neus_hint_model.py:236-267, :653-758).
 * This header is the C boundary a maintainer binds (ctypes stub in INTEGRATION.md) to replace
 * the bodies of:
 *   NeuSHintRenderer.forward            models/neus_hint_model.py:653-751   -> nrh_render_forward
 *   SDFNetwork.forward/.sdf/.gradient   fields/sdf_field.py:106-148         -> nrh_sdf_query
 *   extract_fields (mesh grid queries)  models/neus_hint_model.py:68-83     -> nrh_sdf_query
 *   weight_norm materialisation         fields/sdf_field.py:81-82,100-101   -> done by the caller
 *                                       (torch._weight_norm), then nrh_pack_weights
 *   render_outside + NeRF.forward       models/neus_hint_model.py:434-473,
 *                                       fields/nerf_density_field.py:66-89  -> inside nrh_render_forward
 *                                       when NrhConfig.use_outside_nerf (off by default in the reference)
 *   HashEncoding.pytorch_fwd            fields/encodings.py:306-366         -> nrh_hash_encode
 *   RayGenerator.forward (+ autograd)   camera/ray_generator.py:75-150,
 *                                       camera/lie_groups.py:26-116         -> nrh_raygen_forward / _backward
 *   get_train_loss_dict (+ autograd)    pipelines/base_pipeline.py:50-69    -> nrh_train_loss
 *   torch.optim.Adam.step               trainer/trainer.py:99,278-281       -> nrh_adam_step (flat buffer)
 *   get_alpha / weights / compositing   models/neus_hint_model.py:339-356,
 *   (+ autograd), training step         :521-526,:635-637                   -> nrh_composite_train_forward / _backward
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer unless named host_*; the caller owns every buffer.
 *  - The library allocates nothing persistent, never synchronises the stream, keeps no
 *    global state besides a thread-local error string.
 *  - All tensors are contiguous fp32, row-major, shapes as in the reference (torch) API.
 *  - Return value: 0 on success, negative NRH_ERR_* otherwise; nrh_last_error() describes it.
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *  - Network architecture is the reference default and fixed at compile time:
 *    SDF  : Fourier(6) -> 8 x 256 softplus(beta=100), skip-concat at layer 4, heads 1 + 256
 *    Color: 361(316/325/352)-in -> 4 x 256 ReLU -> 3 sigmoid, Fourier(4) on view/light/hints
 *    NeRF : Fourier(10) of the 4-D inverted-sphere point -> 8 x 256 ReLU, input re-concatenated after layer 4,
 *           density head, feature head, 128-wide view/light layer (Fourier(4) of 6-D), rgb head
 *    nrh_check_config() reports anything else as NRH_ERR_UNSUPPORTED.
 */
#ifndef NRHINTS_B200_H
#define NRHINTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRH_ABI_VERSION 7

#define NRH_OK 0
#define NRH_ERR_INVALID (-1)     /* bad argument (null pointer, size, alignment)         */
#define NRH_ERR_UNSUPPORTED (-2) /* config outside what the kernels implement             */
#define NRH_ERR_WORKSPACE (-3)   /* workspace too small                                    */
#define NRH_ERR_CUDA (-4)        /* a CUDA runtime call / kernel launch failed             */

#define NRH_MAX_ROUGHNESS 4
#define NRH_MAX_SAMPLES 128      /* n_samples + n_importance (and the shadow pair) <= 128   */
#define NRH_MAX_OUTSIDE 64       /* n_outside_samples of the outside NeRF <= 64             */

/* MLP engine selection */
#define NRH_MLP_AUTO 0
#define NRH_MLP_FP32_SIMT 1      /* fp32 FFMA register-tiled fused MLP (always-correct path)   */
#define NRH_MLP_TCGEN05 2        /* tcgen05 tensor-core fused MLP, fp16 hi/lo split operands   */

/* depth / hit-point estimators (models/neus_hint_model.py:528-538) */
#define NRH_DEPTH_ALPHA_BLEND 0  /* sum_j mid_z_j w_j (default)                                 */
#define NRH_DEPTH_MAX_WEIGHT 1   /* mid_z of the sample with the largest weight                 */
#define NRH_DEPTH_SPHERE_TRACE 2 /* hit point supplied by the caller from nrh_sphere_trace      */

/* Knobs of NeuSRendererConfig (models/neus_hint_model.py:133-174) the path honours. */
typedef struct NrhConfig {
    int32_t n_samples;            /* coarse samples per primary ray (64)                      */
    int32_t n_importance;         /* importance samples per primary ray (64)                  */
    int32_t up_sample_steps;      /* importance steps (4); n_importance % steps == 0          */
    int32_t n_shadow_samples;     /* coarse samples per shadow ray (64)                       */
    int32_t n_shadow_importance;  /* importance samples per shadow ray (64), 4 steps          */
    int32_t shadow_hint;          /* reflectance net consumes the shadow hint                 */
    int32_t specular_hint;        /* reflectance net consumes the specular hint               */
    int32_t n_roughness;          /* <= NRH_MAX_ROUGHNESS                                     */
    float roughness[NRH_MAX_ROUGHNESS];
    float shadow_ray_offset;      /* 1e-2                                                     */
    int32_t normalized_normals;   /* 1: NormalizedAnalytic feeds the reflectance net, 0: Analytic */
    int32_t mlp_impl;             /* NRH_MLP_*                                                */
    int32_t depth_type;           /* NRH_DEPTH_* (DepthComputationType, models/neus_hint_model.py:113-121) */
    int32_t use_outside_nerf;     /* NeRF++ background model (models/neus_hint_model.py:137, default 0)   */
    int32_t n_outside;            /* its samples per ray (32), <= NRH_MAX_OUTSIDE                          */
} NrhConfig;

/* Effective (weight-norm already applied) weights, torch layout W[out][in], b[out]. */
typedef struct NrhRawWeights {
    const float* sdf_W[8];        /* [256,39] [256,256]x2 [217,256] [256,256]x4               */
    const float* sdf_b[8];
    const float* sdf_out_W;       /* [1,256]                                                  */
    const float* sdf_out_b;       /* [1]                                                      */
    const float* feat_W;          /* [256,256]                                                */
    const float* feat_b;          /* [256]                                                    */
    const float* col_W[5];        /* [256,Cin] [256,256]x3 [3,256]; Cin = 316 + 9*shadow + 9*n_rough*specular */
    const float* col_b[5];
    const float* variance;        /* device scalar: deviation_network.variance                */
    /* outside NeRF (fields/nerf_density_field.py:57-64); read only when use_outside_nerf            */
    const float* nerf_W[8];       /* pts_linears: [256,84] [256,256]x4 [256,340] [256,256]x2         */
    const float* nerf_b[8];
    const float* nerf_alpha_W;    /* alpha_linear   [1,256]                                          */
    const float* nerf_alpha_b;    /* [1]                                                             */
    const float* nerf_feat_W;     /* feature_linear [256,256]                                        */
    const float* nerf_feat_b;
    const float* nerf_view_W;     /* views_linears[0] [128,310] = [feature 256 | PE(view,light) 54]  */
    const float* nerf_view_b;
    const float* nerf_rgb_W;      /* rgb_linear [3,128]                                              */
    const float* nerf_rgb_b;
} NrhRawWeights;

typedef struct NrhRays {          /* RayBundle fields (camera/ray_utils.py:214-235)           */
    const float* origins;         /* [R,3] */
    const float* directions;      /* [R,3] unit */
    const float* pl_positions;    /* [R,3] */
    const float* nears;           /* [R,1] */
    const float* fars;            /* [R,1] */
    const float* hit_points;      /* [R,3] nullable: required when depth_type == NRH_DEPTH_SPHERE_TRACE */
    const float* hit_depths;      /* [R,1] nullable: required when depth_type == NRH_DEPTH_SPHERE_TRACE */
} NrhRays;

/* Training capture (tcgen05 engine, no outside NeRF): when NrhOutputs.train_capture is set, the primary fine pass of
 * nrh_render_forward IS the training forward of nrh_sdf_train_forward -- it writes the tape the hand-written backward needs
 * and hands its results to the caller, so a training step evaluates the 128 fine samples of every ray once instead of twice.
 * Point order is the pipeline's: SAMPLE-major, point p = j * R + r (sample j of ray r), N = S * R points.
 * In this mode the reflectance network and the final compositing are left to the caller's differentiable path:
 * NrhOutputs.rgb / sampled_color / normal maps are NOT produced (depth, visibilities, the per-sample geometry block,
 * specular cue and z_vals are). */
typedef struct NrhTrainCapture {
    void* tape; size_t tape_bytes;   /* nrh_sdf_train_layout(cfg, S * R).tape_bytes                         */
    float* sdf;                      /* [N]                                                                 */
    float* grad_soa;                 /* [3][N]  d sdf / d x, one plane per coordinate                       */
    float* feat;                     /* [N,256] fp32 feature head output (nullable when feat16 is given)    */
    float* pts_soa;                  /* [3][N]  the section mid-points the kernels evaluated                */
    void* feat16; int64_t feat16_ld; /* nullable: the features as unscaled fp16 rows [N][feat16_ld] instead  */
                                     /* (16-byte aligned, ld a multiple of 8): the feature block of the      */
                                     /* reflectance operand, written in place by the fine-pass kernel        */
} NrhTrainCapture;

typedef struct NrhOutputs {       /* RenderOutput fields (models/neus_hint_model.py:216-233); S = n_samples+n_importance;
                                   * with the outside NeRF `weights` and `sampled_color` have S + n_outside entries per ray
                                   * (the reference returns the concatenated weights, :521-524,:640) */
    float* rgb;                   /* [R,3]   */
    float* depth;                 /* [R,1]   */
    float* weights;               /* [R,S]   nullable (see the per-ray maps below)                          */
    float* inside_sphere;         /* [R,S]   nullable; relax_inside_sphere aliases it (reference quirk :746)  */
    float* analytic_normals;      /* [R,S,3] nullable */
    float* normalized_normals;    /* [R,S,3] nullable */
    float* visibilities;          /* [R,1]   nullable                                          */
    float* specular_cue;          /* [R,S,n_roughness] nullable                                */
    float* inv_s;                 /* [1]     exp(10*variance) clipped to [1e-6,1e6]; s_val = 1/inv_s broadcast by the caller */
    float* z_vals;                /* [R,S]   nullable; final primary sample positions (debug / backward) */
    float* z_shadow;              /* [R,Ss]  nullable; final shadow-ray sample positions                   */
    float* sampled_color;         /* [R,S,3] nullable; per-sample reflectance output                       */
    /* per-ray maps for full-image evaluation (pipelines/base_pipeline.py:126-148 computes them on the host from the
     * per-sample tensors): with these, the per-sample fields above (weights ... specular_cue) may all be NULL and only
     * 60 B/ray instead of 7 KB/ray leave the device. */
    float* normal_map;            /* [R,3] nullable: sum_j analytic_normal_j * w_j * inside_j               */
    float* normalized_normal_map; /* [R,3] nullable: sum_j normalized_normal_j * w_j * inside_j             */
    float* specular_cue_ray;      /* [R,n_roughness] nullable: the per-ray cue (before broadcast)           */
    /* nullable cudaEvent_t: recorded on `stream` as soon as the per-sample geometry block (weights, inside_sphere,
     * analytic_normals, normalized_normals, specular_cue, z_vals -- 95 % of the RenderOutput bytes) is final, i.e. before
     * the shadow march and the reflectance network run, so that a caller that moves the result to the host
     * (pipelines/base_pipeline.py:120) can overlap that copy with the rest of the render on a second stream. */
    void* early_event;
    const NrhTrainCapture* train_capture;   /* nullable: see NrhTrainCapture */
    /* nullable cudaEvent_t pair recorded on `stream` right before / after the primary fine-pass SDF kernel (forward + feature head
     * + reverse sweep over R*S points, the dominant kernel): lets a caller time that launch INSIDE a step (bench.py roofline). */
    void* fine_begin_event;
    void* fine_end_event;
} NrhOutputs;

int nrh_version(void);
const char* nrh_last_error(void);

/* Validate a config against what this build implements. */
int nrh_check_config(const NrhConfig* cfg);

/* Packed (kernel-layout) weights: size in bytes, and the packing pass (device-side transposes
 * / splits, asynchronous on `stream`). Re-run whenever the parameters change. */
size_t nrh_packed_weights_bytes(const NrhConfig* cfg);
int nrh_pack_weights(const NrhConfig* cfg, const NrhRawWeights* raw, void* packed, size_t packed_bytes, void* stream);

/* Scratch requirement of one nrh_render_forward call over R rays / one nrh_sdf_query over N points. */
size_t nrh_workspace_bytes(const NrhConfig* cfg, int64_t R);
size_t nrh_query_workspace_bytes(const NrhConfig* cfg, int64_t N);

/* NeuSHintRenderer.forward.  jitter_primary [R], jitter_outside [R,n_outside] (outside NeRF only) and
 * jitter_shadow [R,n_shadow_samples] are the torch.rand draws of training mode (:682, :689, :394), made by
 * the caller in that order; NULL = inference (no perturbation).
 * bg_rgb: device [3] or NULL.  cos_anneal = min(1, step/anneal_end) in training, 1 otherwise.
 * warmup != 0 zeroes both hints (geometry warm-up, :577-579,:617-619). */
int nrh_render_forward(const NrhConfig* cfg, const void* packed, const NrhRays* rays, int64_t R,
                       const float* bg_rgb, const float* jitter_primary, const float* jitter_outside,
                       const float* jitter_shadow, float cos_anneal, int warmup, const NrhOutputs* out,
                       void* workspace, size_t workspace_bytes, void* stream);

/* SDFNetwork.sdf / .gradient / .forward on arbitrary points: pts [N,3] ->
 * sdf [N] (required), grad [N,3] (nullable), feat [N,256] (nullable). */
int nrh_sdf_query(const NrhConfig* cfg, const void* packed, const float* pts, int64_t N,
                  float* sdf, float* grad, float* feat, void* workspace, size_t workspace_bytes, void* stream);

/* sphere_trace (models/neus_hint_model.py:359-371): p <- p + sdf(p) d from the ray origin until |sdf| < threshold or
 * depth > far_limit, at most max_iterations times.  Converged points are fixed points of the update, so the result does
 * not depend on when the loop stops; like the reference (`converged.all()` is a host sync there) this call checks for
 * completion on the host every `check_every` iterations and therefore SYNCHRONISES the stream (the only entry point
 * that does).  Outputs: hit_points [R,3], hit_depths [R,1].  Workspace: nrh_query_workspace_bytes(cfg, R) + 32*R bytes. */
int nrh_sphere_trace(const NrhConfig* cfg, const void* packed, const float* origins, const float* directions, int64_t R,
                     int max_iterations, float threshold, float far_limit, int check_every,
                     float* hit_points, float* hit_depths, void* workspace, size_t workspace_bytes, void* stream);

/* ---- training: SDF fine pass with a tape + its hand-written backward (tcgen05 engine only) ------------------------------
 * Replaces, for a training step, the reference's `sdf_network(pts)` + `sdf_network.gradient(pts)` pair of render_core /
 * get_alpha (models/neus_hint_model.py:504-508,:335-336; fields/sdf_field.py:106-148 with create_graph=True) and the
 * autograd double-backward through it:  (sdf, feat, grad) = f(pts; W)  and its vector-Jacobian product.
 *   nrh_sdf_train_forward : pts [N,3] -> sdf [N], grad [N,3], feat [N,256] (fp32) + `tape`
 *   nrh_sdf_train_backward: adjoints d_sdf [N], d_feat [N,256], d_grad [N,3] (fp32) -> d_pts [N,3] and the fp16 operand
 *       dumps gb_l / zb_l in `bwd_out`, all in units of the power-of-two loss scale *loss_scale (device scalar chosen by the
 *       caller so that |adjoint| * scale stays within fp16 range, e.g. 2^floor(log2(512 / max|adjoint|))).
 * The weight gradients are point-reductions over the dumps (plain GEMMs, left to the caller / cuBLAS):
 *       dW_l = (u_l^T gb_l) / (1024 S) + (zb_l^T a_l) / (16 S),   db_l = sum_p zb_l / S        (l = 0: a_0 = PE(3 pts), no 1/16)
 *       dw_sdf = (sum_p gb_8 / S + d_sdf^T a_8 / 16) / 3,  dW_feat = d_feat^T a_8 / 16
 * with a_l, u_l from the tape (fp16, x16 / x1024) and gb_l, zb_l from bwd_out; offsets from nrh_sdf_train_layout.
 * Math: oracle/nrh_oracle.py::sdf_mlp_backward (phase A / phase B).  Workspace: forward nrh_query_workspace_bytes,
 * backward NrhTrainLayout.bwd_workspace_bytes. */
typedef struct NrhTrainLayout {
    int64_t p_pad;                 /* N rounded up to a multiple of 128: row count of every dump                      */
    uint64_t tape_tiles_off;       /* per-tile packed softplus' / reverse adjoints (opaque to the caller)             */
    uint64_t tape_act_off;         /* 8 x [p_pad][256] fp16: a_1 .. a_8, x16                                          */
    uint64_t tape_u_off;           /* 8 x [p_pad][256] fp16: u_0 .. u_7, x1024                                        */
    uint64_t tape_bytes;
    uint64_t bwd_gb0_off;          /* [p_pad][64] fp16: gb_0 (39 valid columns)                                       */
    uint64_t bwd_gb_off;           /* 8 x [p_pad][256] fp16: gb_1 .. gb_8                                             */
    uint64_t bwd_zb_off;           /* 8 x [p_pad][256] fp16: zb_0 .. zb_7                                             */
    uint64_t bwd_bytes;
    uint64_t bwd_workspace_bytes;
} NrhTrainLayout;
int nrh_sdf_train_layout(const NrhConfig* cfg, int64_t N, NrhTrainLayout* out);
int nrh_sdf_train_forward(const NrhConfig* cfg, const void* packed, const float* pts, int64_t N, float* sdf, float* grad,
                          float* feat, void* tape, size_t tape_bytes, void* workspace, size_t workspace_bytes, void* stream);
int nrh_sdf_train_backward(const NrhConfig* cfg, const void* packed, const float* pts, int64_t N, const void* tape,
                           size_t tape_bytes, const float* d_sdf, const float* d_feat, const float* d_grad,
                           const float* loss_scale, void* bwd_out, size_t bwd_bytes, float* d_pts,
                           void* workspace, size_t workspace_bytes, void* stream);

/* HashEncoding.pytorch_fwd (fields/encodings.py:306-366; unreachable from the reference's shipped presets, kept as a
 * standalone operator): multi-resolution hash-grid lookup with trilinear interpolation.
 *   pts [N,3] in [0,1]; table [n_levels * 2^log2_T, F] fp32; scalings [n_levels] fp32 (host array: floor(min_res * g^l));
 *   out [N, n_levels * F].  Index arithmetic is the reference's: corners ceil/floor of pts*scaling as int32, hash
 *   (x*1) ^ (y*2654435761) ^ (z*805459861) in int64 (no 32-bit wrap), mod 2^log2_T, + level * 2^log2_T; the
 *   interpolation weight `offset = scaled - floor` goes to the CEIL corner.  F <= 8, n_levels <= 32. */
int nrh_hash_encode(const float* pts, int64_t N, const float* table, const float* host_scalings, int n_levels,
                    int log2_table_size, int features_per_level, float* out, void* stream);

/* Gradient of nrh_hash_encode w.r.t. the table: d_table[idx] += w * d_out (fp32 atomics; d_table must be zeroed
 * by the caller). */
int nrh_hash_encode_backward(const float* pts, int64_t N, const float* d_out, const float* host_scalings, int n_levels,
                             int log2_table_size, int features_per_level, float* d_table, void* stream);

/* ---- ray generation in front of the path ---------------------------------------------------------------
 * RayGenerator.forward (camera/ray_generator.py:75-150): pixel indices + per-ray camera-to-world pose -> RayBundle fields,
 * with the initial pose / light noise buffers (:62-73,:92-98,:121-122) and the learned per-image deltas
 * `cam_pose_adjustment` (exp_map_SO3xR3 / exp_map_SE3, camera/lie_groups.py:26-116) and `pl_adjustment` (:124-126) applied
 * when `img_indices` is given (video views pass NULL and get neither, :103-105).  near/far from the unit sphere when
 * override_near_far (:133-139), else the camera's zn/zf. */
#define NRH_CAM_OPT_OFF 0
#define NRH_CAM_OPT_SO3XR3 1
#define NRH_CAM_OPT_SE3 2
typedef struct NrhCamera { float fx, fy, cx, cy, zn, zf; } NrhCamera;      /* CameraModel (camera/camera_model.py:5-24)  */
typedef struct NrhRayGenInputs {  /* RawPixelBundle (data/data_loader.py:80-89) + the generator's tables                     */
    const float* w_indices;       /* [R]   pixel column (the reference's [R,1] tensor, as fp32)                              */
    const float* h_indices;       /* [R]   pixel row                                                                          */
    const int64_t* img_indices;   /* [R]   nullable: image index of every ray, in [0, n_cameras)                              */
    const float* poses;           /* [R,4,4] camera-to-world, 16-byte aligned                                                 */
    const float* pls;             /* [R,3] point-light positions                                                              */
    const float* cam_pose_noise;  /* [n_cameras,3,4] nullable: buffer `cam_pose_noise`                                        */
    const float* pl_noise;        /* [n_cameras,3]   nullable: buffer `pl_noise`                                              */
    const float* cam_pose_adjustment; /* [n_cameras,6] parameter; required when cam_opt_mode != off and img_indices given    */
    const float* pl_adjustment;   /* [n_cameras,3]   nullable: parameter (pl_opt)                                             */
    int64_t n_cameras;
} NrhRayGenInputs;
int nrh_raygen_forward(const NrhCamera* cam, int cam_opt_mode, int override_near_far, const NrhRayGenInputs* in, int64_t R,
                       float* origins, float* directions, float* pl_positions, float* nears, float* fars, void* stream);
/* Vector-Jacobian product of the above w.r.t. the two parameters: d_cam_pose_adjustment [n_cameras,6] and d_pl_adjustment
 * [n_cameras,3] are ACCUMULATED into (fp32 atomics after an in-warp reduction; zero them first); either may be NULL, and so
 * may any incoming adjoint (treated as zero).  Index rows that no ray touches are left alone. */
int nrh_raygen_backward(const NrhCamera* cam, int cam_opt_mode, int override_near_far, const NrhRayGenInputs* in, int64_t R,
                        const float* d_origins, const float* d_directions, const float* d_pl_positions, const float* d_nears,
                        const float* d_fars, float* d_cam_pose_adjustment, float* d_pl_adjustment, void* stream);

/* ---- loss and optimiser behind the path ----------------------------------------------------------------
 * get_train_loss_dict (pipelines/base_pipeline.py:50-69) and its gradient in two launches, no host sync:
 *   rgb_loss = sum |rgb - rgb_gt| / (R + 1e-5);  eikonal = sum(m (|n| - 1)^2) / (sum m + 1e-5), m = relax_inside_sphere;
 *   loss = rgb_loss + igr_weight * eikonal;  psnr = 10 log10(1 / mean((rgb - rgb_gt)^2))
 * stats (device, 8 floats): [loss, rgb_loss, eikonal_loss, psnr, sum m, sum m (|n|-1)^2, sum |d rgb|, sum (d rgb)^2].
 * d_rgb [R,3] and d_normals [R,S,3] (both nullable) receive d loss / d rgb and d loss / d analytic_normals times
 * `grad_scale` (a host scalar, e.g. a loss scale; 1 for plain backward). */
int nrh_train_loss(const float* rgb, const float* rgb_gt, const float* analytic_normals, const float* relax_inside_sphere,
                   int64_t R, int S, float igr_weight, float grad_scale, float* stats, float* d_rgb, float* d_normals,
                   void* stream);

/* One torch.optim.Adam step (amsgrad off, weight_decay 0; trainer/trainer.py:99,280) over a FLAT parameter buffer of n floats:
 *   m <- m + (g - m)(1 - beta1);  v <- beta2 v + (1 - beta2) g g;  p <- p - (lr / (1 - beta1^step)) m / (sqrt(v) / sqrt(1 - beta2^step) + eps)
 * with g = grad * grad_scale (undo a loss scale or average over ranks), in torch's operation order.  `step` is the 1-based
 * step count after the increment; the scalar bookkeeping (bias corrections, step size) is done in double like torch's.  lr comes from the caller's scheduler (trainer/trainer.py:101-113). */
int nrh_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                  double beta2, double eps, int64_t step, float grad_scale, void* stream);

/* The same step with the step count and the learning rate on the DEVICE, for callers that capture the training step in a CUDA graph
 * (a replay must see a new step count / scheduler value): *step_dev (int64) is incremented by the call, *lr_dev (fp32) is read at
 * execution time, coef_scratch = 2 floats of device scratch owned by the caller.  Same arithmetic (bias corrections in double). */
int nrh_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, const float* lr_dev, double beta1,
                      double beta2, double eps, int64_t* step_dev, float* coef_scratch, float grad_scale, void* stream);

/* Column sums of n_mats row-major fp16 matrices [rows][width] (consecutive matrices `mat_stride` elements apart) -> fp32
 * out [n_mats][width] = scale * sum over rows: the bias gradients of a training step are such point-reductions over the fp16
 * adjoint dumps (db_l = sum_p zb_l / S above).  width: multiple of 8 with 256 % (width / 8) == 0; matrices 16-byte aligned. */
int nrh_colsum_f16(const void* mats, int n_mats, int64_t rows, int width, int64_t mat_stride, float scale, float* out, void* stream);

/* Reflectance network of a training step (fields/reflectance_network.py:68-96 under autograd in the reference), tcgen05 engine.
 * Forward: x16 = the concatenated network input [P][384] fp16, row-major, in the reference's order [points 3 | PE(view) 27 | normal 3
 * | PE(light) 27 | feature 256 | PE(visibility) 9 | PE(specular cue) 36] (columns beyond the configuration's input width must be zero)
 * -> acts [4][P][256] fp16 (post-ReLU hidden activations x16, kept for the backward and as weight-gradient operands) and y [P][4]
 * fp32 (columns 0..2 = the pre-sigmoid output).  Backward: dy [P][3] fp32 and the power-of-two loss scale *loss_scale (device) ->
 * dz [4][P][256] fp16 (adjoints of the hidden pre-activations, layer 0 first), dy16 [P][8] fp16 (dy * S in columns 0..2) and
 * dx [P][384] fp16 (adjoint of x16), all in units of S.  The weight gradients follow with nrh_wgrad_f16:
 *   dW_l = dz_l^T a_l / (16 S) (a_0 = x16: / S), dW_out = dy16^T a_4 / (16 S); biases: column sums of dz_l / S. */
int nrh_color_train_forward(const NrhConfig* cfg, const void* packed, const void* x16, int64_t P, void* acts, float* y, void* stream);
int nrh_color_train_backward(const NrhConfig* cfg, const void* packed, const float* dy, const float* loss_scale, const void* acts,
                             int64_t P, void* dz, void* dy16, void* dx, void* stream);

/* ---- the fused training step: forward, and ONE backward entry point that writes parameter gradients ----------
 * What the reference does per step under autograd (pipelines/base_pipeline.py:39-48 -> models/neus_hint_model.py:653-751, then
 * loss.backward(), trainer/trainer.py:269-283) as two calls on caller-owned memory, tcgen05 engine, no outside NeRF:
 *
 *   nrh_render_train_forward  = nrh_render_forward in training mode (samplers on both rays, shadow visibility, depth, specular cue:
 *       everything the reference keeps under no_grad) with the primary fine pass writing its tape, followed by the differentiable
 *       tail: reflectance network (tensor cores, activations kept), sigmoid, NeuS alpha / weights / compositing.  `out` receives the
 *       RenderOutput fields as for nrh_render_forward (rgb, weights from the differentiable tail).
 *   nrh_render_backward       : d rgb [R,3] (+ optional adjoints of the analytic_normals / normalized_analytic_normals / weights
 *       outputs) -> gradients of ALL 46 parameter tensors, written straight to the pointers of NrhTrainParams (e.g. views of an
 *       optimizer's flat gradient buffer: the all-reduce operand), and d origins / d directions / d light positions [R,3] (nullable).
 *       Chain: compositor backward -> sigmoid -> reflectance backward (tcgen05) -> input scatter -> SDF second-order backward
 *       (tcgen05) -> all weight-gradient reductions (tcgen05, one launch) -> bias column sums -> weight-norm backward (one launch).
 *       Loss scales are chosen on the device; the call never synchronises.
 *
 * `train_ws` (nrh_train_workspace_bytes) carries the render workspace, the tape and every intermediate from the forward call to the
 * backward call of the same step; the caller must not touch it in between.  Parameters are given as the module stores them: weight-
 * normed layers as (v [out,in], g [out,1], bias [out]) -- the library applies W = g v / ||v||_row itself (nrh_pack_weights_wn). */
typedef struct NrhLayerParams {
    const float* v; const float* g; const float* bias;     /* parameters (device, fp32, contiguous) */
    float* d_v; float* d_g; float* d_bias;                 /* gradients, OVERWRITTEN by nrh_render_backward (same shapes) */
} NrhLayerParams;
typedef struct NrhTrainParams {
    NrhLayerParams sdf[8];        /* sdf_network.lin0..7 */
    NrhLayerParams sdf_out;       /* sdf_network.out_sdf  ([1,256]) */
    NrhLayerParams feat_out;      /* sdf_network.out_feat ([256,256]) */
    NrhLayerParams col[5];        /* color_network.lin0..4 */
    const float* variance; float* d_variance;              /* deviation_network.variance */
} NrhTrainParams;
typedef struct NrhTrainAdjoints {
    const float* d_rgb;                /* [R,3] */
    const float* d_analytic_normals;   /* [R,S,3] nullable (eikonal loss) */
    const float* d_normalized_normals; /* [R,S,3] nullable */
    const float* d_weights;            /* [R,S] nullable */
    float* d_origins; float* d_directions; float* d_pl_positions;   /* [R,3] each, nullable */
} NrhTrainAdjoints;
size_t nrh_train_workspace_bytes(const NrhConfig* cfg, int64_t R);
/* effective weights from (v, g) for all weight-normed layers in one launch, then nrh_pack_weights; `wn_scratch`: >= 3.4 MB */
int nrh_pack_weights_wn(const NrhConfig* cfg, const NrhTrainParams* params, void* wn_scratch, size_t wn_scratch_bytes, void* packed,
                        size_t packed_bytes, void* stream);
int nrh_render_train_forward(const NrhConfig* cfg, const void* packed, const NrhRays* rays, int64_t R, const float* bg_rgb,
                             const float* jitter_primary, const float* jitter_shadow, float cos_anneal, int warmup,
                             const NrhOutputs* out, void* train_ws, size_t train_ws_bytes, void* stream);
int nrh_render_backward(const NrhConfig* cfg, const void* packed, const NrhTrainParams* params, const NrhRays* rays, int64_t R,
                        const float* bg_rgb, float cos_anneal, const NrhTrainAdjoints* adj, void* train_ws, size_t train_ws_bytes,
                        void* stream);

/* Weight-gradient reductions of a training step (the dW = delta^T h products the reference leaves to autograd behind every
 * F.linear: fields/sdf_field.py:106-123, fields/reflectance_network.py:84-96; trainer/trainer.py:279), ALL of them in one call:
 *     out[m, n] += scale * (*dev_scale if non-NULL) * sum_{p < rows} A[p, a_col0 + m] * B[p, b_col0 + n]
 * A, B: row-major fp16 device matrices [rows][a_ld] / [rows][b_ld] (the operand dumps the forward / backward kernels publish),
 * out: fp32 [m][ld_out], ACCUMULATED (the caller zeroes its gradient buffer once per step).  m == 256 with n in {64,128,192,256}
 * runs on tcgen05 (one persistent launch for all such jobs: TMA tensor-map loads of MN-major operand tiles, 256 x n fp32
 * accumulators in tensor memory, split-K over contiguous tile ranges, one vectorised red.add flush per CTA and job);
 * m <= 8 (A needs 8 readable columns) is a streaming reduction on the CUDA cores.  rows_valid / cols_valid (0 = all) limit what is
 * written.  Operands: 16-byte aligned, leading dimensions multiples of 8, a_col0 / b_col0 multiples of 8. */
typedef struct NrhWgradJob {
    const void* a; int64_t a_ld; int32_t a_col0;
    const void* b; int64_t b_ld; int32_t b_col0;
    int64_t rows;
    int32_t m, n, rows_valid, cols_valid;
    float scale; const float* dev_scale;
    float* out; int64_t ld_out;
} NrhWgradJob;
int nrh_wgrad_f16(const NrhWgradJob* jobs, int n_jobs, void* stream);

/* Differentiable compositing of the primary ray for a training step: get_alpha (models/neus_hint_model.py:339-356), the
 * transmittance scan / weights (:521-526) and rgb = sum_j w_j c_j + bg (1 - sum_j w_j) (:635-637), and the vector-Jacobian
 * product torch autograd derives for them.  Per-point inputs sdf [N], grad [N,3], color [N,3] (N = R*S) are addressed as
 * point(r, j) = r * point_stride_ray + j * point_stride_sample, so both the ray-major (S, 1) and the pipeline's sample-major
 * (1, R) orders are accepted; dists [R,S] and weights [R,S] are ray-major, dirs [R,3], inv_s a device scalar, bg_rgb [3] or NULL.
 * Backward: d_rgb [R,3], d_weights [R,S] (nullable) -> d_sdf / d_grad / d_color in the inputs' point order, d_dirs [R,3], and the
 * scalar *d_inv_s (ACCUMULATED: zero it first).  One thread per ray, no tape: the scan is recomputed. */
int nrh_composite_train_forward(const float* sdf, const float* grad, const float* color, int64_t point_stride_ray,
                                int64_t point_stride_sample, const float* dists, const float* dirs, const float* inv_s,
                                float cos_anneal, const float* bg_rgb, int64_t R, int S, float* weights, float* rgb, void* stream);
int nrh_composite_train_backward(const float* sdf, const float* grad, const float* color, int64_t point_stride_ray,
                                 int64_t point_stride_sample, const float* dists, const float* dirs, const float* inv_s,
                                 float cos_anneal, const float* bg_rgb, int64_t R, int S, const float* d_rgb, const float* d_weights,
                                 float* d_sdf, float* d_grad, float* d_color, float* d_dirs, float* d_inv_s, void* stream);

/* Kernel launches issued by the last nrh_render_forward / nrh_sdf_query call on this thread. */
int nrh_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NRHINTS_B200_H */
