// Workspace views and launchers of the per-ray sampler / compositor kernels (internal).
#pragma once
#include "nrh_common.cuh"

namespace nrh {

// One ray march (primary rays, or shadow rays from the light): sample-major arrays over R rays.
struct MarchState {
    float* o[3]; float* d[3];         // [R] ray origin / unit direction (SoA)
    float* z[2]; float* s[2];         // ping-pong sorted sample positions / SDF at them, [Smax][R]
    float* znew; float* snew;         // [n_new_max][R] freshly drawn samples / their SDF
    float* wbuf;                      // [Smax][R] interval-weight scratch
    float* px; float* py; float* pz;  // [Smax*R] points handed to the MLP engine (p = j*R + r)
};

struct FineBuffers {                  // fine pass of the primary march, [S*R] each (p = j*R + r)
    float* sdf; float* gx; float* gy; float* gz;   // MLP outputs at the section mid-points
    float* w; float* inside; float* nx; float* ny; float* nz;   // compositor outputs
};

struct OutsideBuffers {               // outside NeRF, [(S + n_out) * R] each (p = j*R + r); all null when the model is off
    float* mid; float* dist;          // section mid-points / lengths of the merged sample set (render_outside :441-444)
    float* density; float* r; float* g; float* b;   // NeRF outputs (raw density, sigmoid colour)
};

struct RayState {                     // per-ray scalars, [R] each
    float* depth; float* wsum; float* hit[3]; float* hitn[3]; float* vis; float* light_dist;
    float* spec[NRH_MAX_ROUGHNESS];
};

int launch_coarse_primary(const NrhRays& rays, int64_t R, int n, const float* jitter, const MarchState& m, cudaStream_t st);
int launch_importance_step(int64_t R, const MarchState& m, int cur, int k_old, int n_new, bool merge_first, float inv_s,
                           bool last, float last_dist_const, const float* last_dist_ray, cudaStream_t st);
int launch_sections_only(int64_t R, const MarchState& m, int cur, int S, float last_dist_const, const float* last_dist_ray, cudaStream_t st);
int launch_composite_primary(int64_t R, const MarchState& m, int cur, int S, float last_dist, const float* inv_s,
                             float cos_anneal, const FineBuffers& f, const RayState& rs, const float* pl, bool do_shadow,
                             const MarchState& sh, int n_shadow, float shadow_offset, const float* jitter_shadow,
                             int depth_type, const float* hit_pts, const float* hit_depth, int n_out, const OutsideBuffers& ob,
                             cudaStream_t st);
int launch_outside_setup(int64_t R, const MarchState& m, int cur, int S, int n_samples, int n_out, const float* fars,
                         const float* jitter_outside, float sample_dist, const OutsideBuffers& ob, cudaStream_t st);
int launch_sphere_step(int64_t R, const float* dirs, const float* sdf, float* pts, float* depth, float threshold, float far_limit,
                       int* moving, cudaStream_t st);
int launch_specular_cue(int64_t R, const NrhConfig& cfg, const RayState& rs, const float* pl, const float* dirs, int warmup, cudaStream_t st);
int launch_shade_prep(int64_t R, const NrhConfig& cfg, const MarchState& sh, int cur, int S_shadow, const float* inv_s,
                      float cos_anneal, const float* ssdf, const float* sgx, const float* sgy, const float* sgz,
                      const RayState& rs, const float* pl, const float* dirs, int warmup, bool shadow_marched,
                      float* rayfeat, unsigned char* aux_img, cudaStream_t st);
int launch_final_rgb(int64_t R, int S, const FineBuffers& f, const RayState& rs, float* cr, float* cg,
                     float* cb, const float* bg, float* rgb, float* depth, float* vis_out, float* nmap, float* nnmap,
                     float* spec_ray, int n_rough, int n_out, const OutsideBuffers& ob, cudaStream_t st);
int launch_to_ray_major(const float* const* src, int C, bool broadcast, int64_t R, int S, float* dst, cudaStream_t st);

}  // namespace nrh
