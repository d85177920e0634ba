// Per-ray sampler / compositor kernels: one thread per ray over sample-major workspace arrays
// (element j of ray r at [j*R + r]) so every access is coalesced.  All arithmetic lives in
// ray_math.cuh (shared with the host test harness); this file only moves data.
//
// Reference call sites: /root/reference/models/neus_hint_model.py:673-713 (coarse + importance
// loop), :475-651 (render_core), :373-432 (get_visibility).
#include "nrh_common.cuh"
#include "ray_math.cuh"
#include "sampler_kernels.cuh"
#include "mlp_tc.cuh"
#include "tc_primitives.cuh"
#include <cuda_fp16.h>

namespace nrh {
namespace {

constexpr int TPB = 128;

__device__ __forceinline__ void write_points(const float o[3], const float d[3], float z, int64_t idx,
                                             float* px, float* py, float* pz) {
    px[idx] = o[0] + d[0] * z; py[idx] = o[1] + d[1] * z; pz[idx] = o[2] + d[2] * z;
}

// ---- primary rays: load the bundle, coarse z, coarse points --------------------------------------
__global__ void k_coarse_primary(NrhRays rays, int64_t R, int n, const float* __restrict__ jitter,
                                 MarchState m) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= R) return;
    float o[3], d[3];
    for (int c = 0; c < 3; ++c) {
        o[c] = rays.origins[r * 3 + c]; d[c] = rays.directions[r * 3 + c];
        m.o[c][r] = o[c]; m.d[c][r] = d[c];
    }
    SoA z{m.z[0] + r, R};
    coarse_z(rays.nears[r], rays.fars[r], n, jitter != nullptr, jitter ? jitter[r] : 0.f, z);
    for (int j = 0; j < n; ++j) write_points(o, d, z[j], (int64_t)j * R + r, m.px, m.py, m.pz);
}

// ---- one importance step (merge previous new samples, draw new ones, emit their points; on the
//      last step also merge them and emit the section mid-points of the final sample set) ---------
// The per-ray algorithm is a chain of short dependent steps over <= 128 samples, so each thread first stages its
// ray's arrays in shared memory (coalesced, fully overlapped global loads), runs ray_math.cuh on them, and writes
// the results back: the dependent chain then pays shared-memory latency instead of L2 latency per step.
constexpr int IS_RAYS = 32;                                 // rays per CTA (128 CTAs for 4096 rays: the parallel phases use most SMs)
constexpr int IS_THREADS = 512;                             // 16 threads per ray for the parallel phases
constexpr int IS_NEW = 32;                                   // max samples drawn per step
constexpr int IS_ROWS = 3 * NRH_MAX_SAMPLES + 2 * IS_NEW;   // z[128] s[128] w[128] z_new[32] s_new[32]
constexpr size_t IS_SMEM = (size_t)IS_ROWS * IS_RAYS * sizeof(float);

// A CTA owns IS_RAYS rays.  Data movement, the per-interval alphas (two precise sigmoids + three divisions each, the bulk
// of the arithmetic), the merges (rank form) and the pdf divisions are spread over all 512 threads; only the two genuinely
// sequential parts (transmittance product + weight sum, cdf walk) run one thread per ray on the shared-memory copies.
// Arithmetic and its order are exactly those of ray_math.cuh.
// Stable merge of the n_new sorted new entries (szn / ssn) into the k_old sorted old ones (sz / ss), in place, on ALL threads: every
// (ray, entry) pair finds its own slot by binary search in the other run (ray_math.cuh: count_less / count_less_equal), all slots
// are computed before any is written.  Same result as merge_sorted / merge_sorted_backward (tests/test_ray_math_host.py).
constexpr int IS_ITEMS = NRH_MAX_SAMPLES * IS_RAYS / IS_THREADS;           // (ray, entry) pairs per thread
__device__ __forceinline__ void merge_parallel(int tid, int nr, int k_old, int n_new, float* sz, float* ss, const float* szn,
                                               const float* ssn, bool with_sdf) {
    float vz[IS_ITEMS], vs[IS_ITEMS];
    int vp[IS_ITEMS];
    const int total = (k_old + n_new) * IS_RAYS;
#pragma unroll
    for (int i = 0; i < IS_ITEMS; ++i) {
        const int idx = tid + i * IS_THREADS;
        vp[i] = -1; vz[i] = 0.f; vs[i] = 0.f;
        if (idx < total) {
            const int j = idx / IS_RAYS, t = idx % IS_RAYS;
            if (t < nr) {
                if (j < k_old) {
                    vz[i] = sz[idx];
                    if (with_sdf) vs[i] = ss[idx];
                    vp[i] = (j + count_less(CSoA{szn + t, IS_RAYS}, n_new, vz[i])) * IS_RAYS + t;
                } else {
                    const int b = j - k_old;
                    vz[i] = szn[b * IS_RAYS + t];
                    if (with_sdf) vs[i] = ssn[b * IS_RAYS + t];
                    vp[i] = (b + count_less_equal(CSoA{sz + t, IS_RAYS}, k_old, vz[i])) * IS_RAYS + t;
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < IS_ITEMS; ++i)
        if (vp[i] >= 0) { sz[vp[i]] = vz[i]; if (with_sdf) ss[vp[i]] = vs[i]; }
    __syncthreads();
}

__global__ void __launch_bounds__(IS_THREADS)
k_importance_step(int64_t R, MarchState m, int cur, int k_old, int n_new, bool merge_first,
                  float inv_s, bool last, float last_dist_const, const float* last_dist_ray) {
    extern __shared__ float sm[];
    const int tid = threadIdx.x;
    const int64_t r0 = blockIdx.x * (int64_t)IS_RAYS;
    const int nr = (int)((R - r0) < IS_RAYS ? (R - r0) : IS_RAYS);          // rays of this CTA
    float* sz = sm;                                                       // element j of local ray t at [j * IS_RAYS + t]
    float* ss = sz + NRH_MAX_SAMPLES * IS_RAYS;
    float* sw = ss + NRH_MAX_SAMPLES * IS_RAYS;
    float* szn = sw + NRH_MAX_SAMPLES * IS_RAYS;
    float* ssn = szn + IS_NEW * IS_RAYS;
    // ---- phase 0: stage (coalesced: consecutive threads = consecutive rays) ----
    for (int idx = tid; idx < k_old * IS_RAYS; idx += IS_THREADS) {
        const int j = idx / IS_RAYS, t = idx % IS_RAYS;
        if (t < nr) { sz[idx] = m.z[cur][(int64_t)j * R + r0 + t]; ss[idx] = m.s[cur][(int64_t)j * R + r0 + t]; }
    }
    if (merge_first)
        for (int idx = tid; idx < n_new * IS_RAYS; idx += IS_THREADS) {
            const int j = idx / IS_RAYS, t = idx % IS_RAYS;
            if (t < nr) { szn[idx] = m.znew[(int64_t)j * R + r0 + t]; ssn[idx] = m.snew[(int64_t)j * R + r0 + t]; }
        }
    __syncthreads();
    int k = k_old;
    // ---- phase 1: merge the samples drawn by the previous step (all threads) ----
    if (merge_first) {
        merge_parallel(tid, nr, k_old, n_new, sz, ss, szn, ssn, true);
        k = k_old + n_new;
        cur ^= 1;
        if (!last)
            for (int idx = tid; idx < k * IS_RAYS; idx += IS_THREADS) {
                const int j = idx / IS_RAYS, t = idx % IS_RAYS;
                if (t < nr) { m.z[cur][(int64_t)j * R + r0 + t] = sz[idx]; m.s[cur][(int64_t)j * R + r0 + t] = ss[idx]; }
            }
    }
    // ---- phase 2: per-interval alphas, all threads ----
    for (int idx = tid; idx < (k - 1) * IS_RAYS; idx += IS_THREADS) {
        const int j = idx / IS_RAYS, t = idx % IS_RAYS;
        if (t < nr) {
            const int64_t r = r0 + t;
            const float o[3] = {m.o[0][r], m.o[1][r], m.o[2][r]}, d[3] = {m.d[0][r], m.d[1][r], m.d[2][r]};
            sw[idx] = interval_alpha(o, d, j, CSoA{sz + t, IS_RAYS}, CSoA{ss + t, IS_RAYS}, inv_s);
        }
    }
    __syncthreads();
    // ---- phase 3: sample_from_alphas in its three parts: transmittance product + weights (one thread per ray: a 2-flop chain),
    //      the k - 1 divisions by the weight sum (all threads), the cdf walk (one thread per ray: adds and compares only);
    //      last step: final merge of the positions (all threads) ----
    if (tid < nr) ssn[tid] = weights_from_alphas(k, SoA{sw + tid, IS_RAYS});      // ssn is free after phase 1
    __syncthreads();
    for (int idx = tid; idx < (k - 1) * IS_RAYS; idx += IS_THREADS) {
        const int t = idx % IS_RAYS;
        if (t < nr) sw[idx] = sw[idx] / ssn[t];
    }
    __syncthreads();
    if (tid < nr) inverse_cdf_samples(k, CSoA{sz + tid, IS_RAYS}, n_new, CSoA{sw + tid, IS_RAYS}, SoA{szn + tid, IS_RAYS});
    __syncthreads();
    if (last) merge_parallel(tid, nr, k, n_new, sz, ss, szn, ssn, false);
    // ---- phase 4: write back (all threads) ----
    if (!last) {
        for (int idx = tid; idx < n_new * IS_RAYS; idx += IS_THREADS) {
            const int j = idx / IS_RAYS, t = idx % IS_RAYS;
            if (t < nr) {
                const int64_t r = r0 + t;
                const float o[3] = {m.o[0][r], m.o[1][r], m.o[2][r]}, d[3] = {m.d[0][r], m.d[1][r], m.d[2][r]};
                const float zv = szn[idx];
                m.znew[(int64_t)j * R + r] = zv;
                write_points(o, d, zv, (int64_t)j * R + r, m.px, m.py, m.pz);
            }
        }
    } else {
        cur ^= 1;
        const int S = k + n_new;
        for (int idx = tid; idx < S * IS_RAYS; idx += IS_THREADS) {
            const int j = idx / IS_RAYS, t = idx % IS_RAYS;
            if (t < nr) {
                const int64_t r = r0 + t;
                const float o[3] = {m.o[0][r], m.o[1][r], m.o[2][r]}, d[3] = {m.d[0][r], m.d[1][r], m.d[2][r]};
                const float last_dist = last_dist_ray ? last_dist_ray[r] : last_dist_const;
                float dist, mid; section(CSoA{sz + t, IS_RAYS}, j, S, last_dist, dist, mid);
                m.z[cur][(int64_t)j * R + r] = sz[idx];
                write_points(o, d, mid, (int64_t)j * R + r, m.px, m.py, m.pz);
            }
        }
    }
}

// no importance samples at all: just emit the section mid-points of the coarse set
__global__ void k_sections_only(int64_t R, MarchState m, int cur, int S, float last_dist_const, const float* last_dist_ray) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= R) return;
    float o[3], d[3];
    for (int c = 0; c < 3; ++c) { o[c] = m.o[c][r]; d[c] = m.d[c][r]; }
    const float last_dist = last_dist_ray ? last_dist_ray[r] : last_dist_const;
    CSoA z{m.z[cur] + r, R};
    for (int j = 0; j < S; ++j) {
        float dist, mid; section(z, j, S, last_dist, dist, mid);
        write_points(o, d, mid, (int64_t)j * R + r, m.px, m.py, m.pz);
    }
}

// ---- staging for the one-thread-per-ray compositors ----------------------------------------------------------------------
// composite_primary / shadow_transmittance walk 128 samples per ray with a short dependent chain (the transmittance) and five
// loads per step; with one thread per ray straight on the global arrays every step paid an exposed L2 round trip (131 us / 109 us
// for 4096 rays, ncu r2q).  A CTA of ST_THREADS threads therefore first copies the sample-major input arrays of its ST_RAYS rays
// into shared memory (coalesced 128-byte rows, all loads in flight at once), then one warp runs the UNCHANGED ray_math.cuh function
// on the shared-memory views: same arithmetic, same order, bitwise the same results.
constexpr int ST_RAYS = 32;
constexpr int ST_THREADS = 256;
__device__ __forceinline__ void stage_rows(float* __restrict__ dst, const float* __restrict__ src, int S, int64_t R, int64_t r0, int nr) {
    for (int idx = threadIdx.x; idx < S * ST_RAYS; idx += ST_THREADS) {
        const int j = idx / ST_RAYS, t = idx % ST_RAYS;
        if (t < nr) dst[idx] = src[(int64_t)j * R + r0 + t];
    }
}
inline size_t staged_smem(int S) { return (size_t)5 * S * ST_RAYS * sizeof(float); }
inline int staged_blocks(int64_t R) { return (int)((R + ST_RAYS - 1) / ST_RAYS); }

// ---- primary composite + shadow-ray setup ----------------------------------------------------------
__global__ void __launch_bounds__(ST_THREADS)
k_composite_primary(int64_t R, MarchState m, int cur, int S, float last_dist,
                                    const float* __restrict__ inv_s_ptr, float cos_anneal, FineBuffers f,
                                    RayState rs, const float* __restrict__ pl, bool do_shadow, MarchState sh,
                                    int n_shadow, float shadow_offset, const float* __restrict__ jitter_shadow,
                                    int depth_type, const float* __restrict__ hit_pts, const float* __restrict__ hit_depth,
                                    int n_out, OutsideBuffers ob) {
    extern __shared__ float st_sm[];
    const int64_t r0 = blockIdx.x * (int64_t)ST_RAYS;
    const int nr = (int)((R - r0) < ST_RAYS ? (R - r0) : ST_RAYS);
    const int rows = S * ST_RAYS;
    stage_rows(st_sm, m.z[cur], S, R, r0, nr);
    stage_rows(st_sm + rows, f.sdf, S, R, r0, nr);
    stage_rows(st_sm + 2 * rows, f.gx, S, R, r0, nr);
    stage_rows(st_sm + 3 * rows, f.gy, S, R, r0, nr);
    stage_rows(st_sm + 4 * rows, f.gz, S, R, r0, nr);
    __syncthreads();
    const float inv_s = inv_s_ptr[0];
    // per-sample terms on all threads (composite_sample): alpha replaces sdf, the unit normal replaces the gradient in shared memory
    // (every (sample, ray) slot is read and written by the same thread; z is read-only: section() looks at the next sample)
    for (int idx = threadIdx.x; idx < rows; idx += ST_THREADS) {
        const int j = idx / ST_RAYS, tt = idx % ST_RAYS;
        if (tt < nr) {
            const int64_t rr = r0 + tt;
            const float oo[3] = {m.o[0][rr], m.o[1][rr], m.o[2][rr]}, dd[3] = {m.d[0][rr], m.d[1][rr], m.d[2][rr]};
            const SampleTerms ts = composite_sample(oo, dd, S, CSoA{st_sm + tt, ST_RAYS}, j, last_dist, st_sm[rows + idx], st_sm[2 * rows + idx],
                                                    st_sm[3 * rows + idx], st_sm[4 * rows + idx], inv_s, cos_anneal, n_out,
                                                    n_out > 0 ? ob.density[(int64_t)j * R + rr] : 0.f, n_out > 0 ? ob.dist[(int64_t)j * R + rr] : 0.f);
            st_sm[rows + idx] = ts.a; st_sm[2 * rows + idx] = ts.n0; st_sm[3 * rows + idx] = ts.n1; st_sm[4 * rows + idx] = ts.n2;
            const int64_t g = (int64_t)j * R + rr;
            f.inside[g] = ts.inside; f.nx[g] = ts.n0; f.ny[g] = ts.n1; f.nz[g] = ts.n2;
        }
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= nr) return;
    const int64_t r = r0 + t;
    float o[3], d[3];
    for (int c = 0; c < 3; ++c) { o[c] = m.o[c][r]; d[c] = m.d[c][r]; }
    // the scan, one thread per ray, in sample order (composite_accumulate)
    PrimaryComposite pc; composite_init(pc);
    {
        float T = 1.0f;
        const CSoA z{st_sm + t, ST_RAYS};
        for (int j = 0; j < S; ++j) {
            float dist, mid; section(z, j, S, last_dist, dist, mid);
            const int idx = j * ST_RAYS + t;
            f.w[(int64_t)j * R + r] = composite_accumulate(pc, T, st_sm[rows + idx], mid, st_sm[2 * rows + idx], st_sm[3 * rows + idx], st_sm[4 * rows + idx]);
        }
        composite_append_outside(pc, T, S, n_out, CSoA{n_out > 0 ? ob.density + r : nullptr, R}, CSoA{n_out > 0 ? ob.dist + r : nullptr, R},
                                 SoA{f.w + r, R});
    }
    float hit[3], hn[3];
    float depth = pc.depth;                                              // AlphaBlend (:530-533)
    if (depth_type == NRH_DEPTH_MAX_WEIGHT) depth = pc.max_mid;           // MaximalWeightPoint (:534-538)
    if (depth_type == NRH_DEPTH_SPHERE_TRACE) {                           // SphereTracing (:528-529), traced by the caller
        depth = hit_depth[r];
        for (int c = 0; c < 3; ++c) hit[c] = hit_pts[r * 3 + c];
    } else {
        for (int c = 0; c < 3; ++c) hit[c] = o[c] + d[c] * depth;
    }
    rs.depth[r] = depth; rs.wsum[r] = pc.wsum;
    normalize3(pc.nsum, hn);
    for (int c = 0; c < 3; ++c) { rs.hit[c][r] = hit[c]; rs.hitn[c][r] = hn[c]; }
    if (do_shadow) {
        float l[3] = {pl[r * 3 + 0], pl[r * 3 + 1], pl[r * 3 + 2]}, sd[3];
        SoA z{sh.z[0] + r, R};
        const float L = shadow_ray_init(l, hit, n_shadow, shadow_offset, jitter_shadow != nullptr,
                                        CSoA{jitter_shadow ? jitter_shadow + r * (int64_t)n_shadow : nullptr, 1}, sd, z);
        rs.light_dist[r] = L / (float)n_shadow;        // sample_dist of the shadow ray (:383)
        for (int c = 0; c < 3; ++c) { sh.o[c][r] = l[c]; sh.d[c][r] = sd[c]; }
        for (int j = 0; j < n_shadow; ++j) write_points(l, sd, z[j], (int64_t)j * R + r, sh.px, sh.py, sh.pz);
    }
}

// ---- outside NeRF: sample positions beyond `far`, merged with the primary samples; section lengths / mid-points ----
__global__ void k_outside_setup(int64_t R, MarchState m, int cur, int S, int n_samples, int n_out, const float* __restrict__ fars,
                                const float* __restrict__ jitter_outside, float sample_dist, OutsideBuffers ob) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= R) return;
    float zo[NRH_MAX_OUTSIDE];
    outside_z(fars[r], n_samples, n_out, jitter_outside != nullptr,
              CSoA{jitter_outside ? jitter_outside + r * (int64_t)n_out : nullptr, 1}, zo);
    outside_sections(S, CSoA{m.z[cur] + r, R}, n_out, zo, sample_dist, SoA{ob.dist + r, R}, SoA{ob.mid + r, R});
}

// ---- specular cue of the hit point (needs only the primary compositor's hit point / normal) --------------------
__global__ void k_specular_cue(int64_t R, NrhConfig cfg, RayState rs, const float* __restrict__ pl, const float* __restrict__ dirs,
                               int warmup) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= R) return;
    float cue[NRH_MAX_ROUGHNESS] = {0.f, 0.f, 0.f, 0.f};
    if (cfg.specular_hint && !warmup) {
        float d[3] = {dirs[r * 3 + 0], dirs[r * 3 + 1], dirs[r * 3 + 2]};
        float l[3] = {pl[r * 3 + 0], pl[r * 3 + 1], pl[r * 3 + 2]};
        float hit[3] = {rs.hit[0][r], rs.hit[1][r], rs.hit[2][r]};
        float hn[3] = {rs.hitn[0][r], rs.hitn[1][r], rs.hitn[2][r]};
        specular_cue(hn, l, hit, d, cfg.n_roughness, cfg.roughness, cue);
    }
    for (int i = 0; i < NRH_MAX_ROUGHNESS; ++i) rs.spec[i][r] = cue[i];
}

// ---- shadow transmittance, per-ray reflectance inputs ---------------------------------------------------------
__global__ void __launch_bounds__(ST_THREADS)
k_shade_prep(int64_t R, NrhConfig cfg, MarchState sh, int cur, int S_shadow,
                             const float* __restrict__ inv_s_ptr, float cos_anneal, const float* __restrict__ ssdf,
                             const float* __restrict__ sgx, const float* __restrict__ sgy, const float* __restrict__ sgz,
                             RayState rs, const float* __restrict__ pl, const float* __restrict__ dirs, int warmup,
                             bool shadow_marched, float* __restrict__ rayfeat, unsigned char* __restrict__ aux_img) {
    extern __shared__ float st_sm[];
    const int64_t r0 = blockIdx.x * (int64_t)ST_RAYS;
    const int nr = (int)((R - r0) < ST_RAYS ? (R - r0) : ST_RAYS);
    const int rows = S_shadow * ST_RAYS;
    if (shadow_marched) {                                 // kernel argument: uniform over the grid
        stage_rows(st_sm, sh.z[cur], S_shadow, R, r0, nr);
        stage_rows(st_sm + rows, ssdf, S_shadow, R, r0, nr);
        stage_rows(st_sm + 2 * rows, sgx, S_shadow, R, r0, nr);
        stage_rows(st_sm + 3 * rows, sgy, S_shadow, R, r0, nr);
        stage_rows(st_sm + 4 * rows, sgz, S_shadow, R, r0, nr);
        __syncthreads();
    }
    if (shadow_marched) {
        // per-sample alphas on all threads (shadow_alpha), written over the sdf slots; the last sample does not enter taus[:, -1]
        const float inv_s = inv_s_ptr[0];
        for (int idx = threadIdx.x; idx < rows - ST_RAYS; idx += ST_THREADS) {
            const int j = idx / ST_RAYS, tt = idx % ST_RAYS;
            if (tt < nr) {
                const int64_t rr = r0 + tt;
                const float sd[3] = {sh.d[0][rr], sh.d[1][rr], sh.d[2][rr]};
                st_sm[rows + idx] = shadow_alpha(sd, S_shadow, CSoA{st_sm + tt, ST_RAYS}, j, rs.light_dist[rr], st_sm[rows + idx],
                                                 st_sm[2 * rows + idx], st_sm[3 * rows + idx], st_sm[4 * rows + idx], inv_s, cos_anneal);
            }
        }
        __syncthreads();
    }
    const int t = threadIdx.x;
    if (t >= nr) return;
    const int64_t r = r0 + t;
    float vis = 0.f;
    if (shadow_marched) {                                  // the transmittance product, in sample order (shadow_transmittance)
        float T = 1.0f;
        for (int j = 0; j + 1 < S_shadow; ++j) T = T * (1.0f - st_sm[rows + j * ST_RAYS + t] + 1e-7f);
        vis = T;
    }
    rs.vis[r] = vis;
    float d[3] = {dirs[r * 3 + 0], dirs[r * 3 + 1], dirs[r * 3 + 2]};
    float l[3] = {pl[r * 3 + 0], pl[r * 3 + 1], pl[r * 3 + 2]};
    float cue[NRH_MAX_ROUGHNESS];                        // written by k_specular_cue right after the primary compositor
    for (int i = 0; i < NRH_MAX_ROUGHNESS; ++i) cue[i] = rs.spec[i][r];
    // per-ray encoded reflectance inputs: PE(view) 27 | PE(light) 27 | PE(vis) 9 | PE(spec) 36
    fourier_encode(d, 3, COL_FREQ, rayfeat + r, R);
    fourier_encode(l, 3, COL_FREQ, rayfeat + (int64_t)COL_PE3 * R + r, R);
    float* vis_dst = rayfeat + (int64_t)(2 * COL_PE3) * R + r;
    float* spec_dst = rayfeat + (int64_t)(2 * COL_PE3 + 9) * R + r;
    if (cfg.shadow_hint) fourier_encode(&vis, 1, COL_FREQ, vis_dst, R);
    else for (int i = 0; i < 9; ++i) vis_dst[(int64_t)i * R] = 0.f;
    for (int i = 0; i < 36; ++i) spec_dst[(int64_t)i * R] = 0.f;
    if (cfg.specular_hint) fourier_encode(cue, cfg.n_roughness, COL_FREQ, spec_dst, R);
    if (aux_img) {
        // operand image row of this ray for the streamed reflectance kernel: 128 aux columns, fp16 x TC_ACT_SCALE, in the
        // K-major SWIZZLE_128B layout of two [128 x 64] chunks; the per-point columns (position 0..2, normal 30..32)
        // are left zero and patched in by that kernel
        unsigned char* img = aux_img + (size_t)(r / 128) * TC_TILE_AUX_BYTES;
        const unsigned row = (unsigned)(r % 128);
        for (int k8 = 0; k8 < 128; k8 += 8) {
            __align__(16) __half h[8];
            for (int i = 0; i < 8; ++i) {
                const int j = k8 + i;
                float v = 0.f;
                if (j >= AUX_VIEW && j < AUX_NORMAL) v = rayfeat[(int64_t)(j - AUX_VIEW) * R + r];
                else if (j >= AUX_LIGHT && j < 105) v = rayfeat[(int64_t)(j - AUX_LIGHT + COL_PE3) * R + r];
                h[i] = __float2half_rn(v * TC_ACT_SCALE);
            }
            *reinterpret_cast<uint4*>(img + (k8 >> 6) * 16384 + tc::sw128_offset(row, k8 & 63)) = *reinterpret_cast<const uint4*>(h);
        }
    }
}

// ---- rgb = sum_j w_j c_j + bg (1 - sum w) ---------------------------------------------------------------
// With the outside NeRF (n_out > 0) the sampled colour is blended with the background colour outside the unit sphere and the
// n_out far samples are appended (render_core :630-633); the blended colours are written back to cr/cg/cb ([(S+n_out)*R]).
__global__ void k_final_rgb(int64_t R, int S, FineBuffers f, RayState rs, float* __restrict__ cr,
                            float* __restrict__ cg, float* __restrict__ cb, const float* __restrict__ bg,
                            float* __restrict__ rgb, float* __restrict__ depth, float* __restrict__ vis_out,
                            float* __restrict__ nmap, float* __restrict__ nnmap, float* __restrict__ spec_ray, int n_rough,
                            int n_out, OutsideBuffers ob) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= R) return;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    float m[3] = {0.f, 0.f, 0.f}, mn[3] = {0.f, 0.f, 0.f};
    const bool maps = nmap != nullptr || nnmap != nullptr;
    for (int j = 0; j < S; ++j) {
        const int64_t i = (int64_t)j * R + r;
        const float w = f.w[i];
        float c0 = cr[i], c1 = cg[i], c2 = cb[i];
        if (n_out > 0) {
            const float ins = f.inside[i], ou = 1.0f - ins;
            c0 = c0 * ins + ob.r[i] * ou; c1 = c1 * ins + ob.g[i] * ou; c2 = c2 * ins + ob.b[i] * ou;
            cr[i] = c0; cg[i] = c1; cb[i] = c2;
        }
        a0 += c0 * w; a1 += c1 * w; a2 += c2 * w;
        if (maps) {                                   // einsum('...ij,...i,...i->...j', normals, weights, inside_sphere)
            const float wi = w * f.inside[i];
            m[0] += f.gx[i] * wi; m[1] += f.gy[i] * wi; m[2] += f.gz[i] * wi;
            mn[0] += f.nx[i] * wi; mn[1] += f.ny[i] * wi; mn[2] += f.nz[i] * wi;
        }
    }
    for (int j = S; j < S + n_out; ++j) {
        const int64_t i = (int64_t)j * R + r;
        const float w = f.w[i];
        const float c0 = ob.r[i], c1 = ob.g[i], c2 = ob.b[i];
        cr[i] = c0; cg[i] = c1; cb[i] = c2;
        a0 += c0 * w; a1 += c1 * w; a2 += c2 * w;
    }
    if (nmap) { nmap[r * 3 + 0] = m[0]; nmap[r * 3 + 1] = m[1]; nmap[r * 3 + 2] = m[2]; }
    if (nnmap) { nnmap[r * 3 + 0] = mn[0]; nnmap[r * 3 + 1] = mn[1]; nnmap[r * 3 + 2] = mn[2]; }
    if (spec_ray) for (int c = 0; c < n_rough; ++c) spec_ray[r * n_rough + c] = rs.spec[c][r];
    if (bg) {
        const float rest = 1.0f - rs.wsum[r];
        a0 += bg[0] * rest; a1 += bg[1] * rest; a2 += bg[2] * rest;
    }
    rgb[r * 3 + 0] = a0; rgb[r * 3 + 1] = a1; rgb[r * 3 + 2] = a2;
    depth[r] = rs.depth[r];
    if (vis_out) vis_out[r] = rs.vis[r];
}

// ---- sample-major [S][R] (x C channels) -> ray-major [R][S][C] -------------------------------------------
struct TransposeArgs { const float* src[4]; int C; int broadcast; };   // broadcast: src[c] is per-ray [R]

__global__ void k_to_ray_major(TransposeArgs a, int64_t R, int S, float* __restrict__ dst) {
    __shared__ float tile[4][32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int j0 = blockIdx.y * 32;
    for (int c = 0; c < a.C; ++c)
        for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
            const int j = j0 + jj; const int64_t r = r0 + threadIdx.x;
            float v = 0.f;
            if (j < S && r < R) v = a.broadcast ? a.src[c][r] : a.src[c][(int64_t)j * R + r];
            tile[c][jj][threadIdx.x] = v;
        }
    __syncthreads();
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        const int64_t r = r0 + rr; const int j = j0 + threadIdx.x;
        if (r < R && j < S)
            for (int c = 0; c < a.C; ++c) dst[(r * S + j) * a.C + c] = tile[c][threadIdx.x][rr];
    }
}

// one sphere-tracing update (models/neus_hint_model.py:365-368); counts the rays that are still moving
__global__ void k_sphere_step(int64_t R, const float* __restrict__ dirs, const float* __restrict__ sdf, float* __restrict__ pts,
                              float* __restrict__ depth, float threshold, float far_limit, int* __restrict__ moving) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float s = sdf[r], dep = depth[r];
    const bool converged = (fabsf(s) < threshold) || (dep > far_limit);
    if (!converged) {
        for (int c = 0; c < 3; ++c) pts[r * 3 + c] = pts[r * 3 + c] + s * dirs[r * 3 + c];
        depth[r] = dep + s;
        atomicAdd(moving, 1);
    }
}

inline int blocks_for(int64_t R) { return (int)((R + TPB - 1) / TPB); }

}  // namespace

int launch_coarse_primary(const NrhRays& rays, int64_t R, int n, const float* jitter, const MarchState& m, cudaStream_t st) {
    k_coarse_primary<<<blocks_for(R), TPB, 0, st>>>(rays, R, n, jitter, m);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int launch_importance_step(int64_t R, const MarchState& m, int cur, int k_old, int n_new, bool merge_first, float inv_s,
                           bool last, float last_dist_const, const float* last_dist_ray, cudaStream_t st) {
    if (n_new > IS_NEW) { set_error("importance step draws at most %d samples per step", IS_NEW); return NRH_ERR_UNSUPPORTED; }
    const int k_after = k_old + (merge_first ? n_new : 0);            // merge_parallel handles at most NRH_MAX_SAMPLES entries per ray
    if (k_after + (last ? n_new : 0) > NRH_MAX_SAMPLES) {
        set_error("importance step: more than %d samples per ray", NRH_MAX_SAMPLES); return NRH_ERR_UNSUPPORTED;
    }
    NRH_CUDA_CHECK(cudaFuncSetAttribute(k_importance_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IS_SMEM));
    k_importance_step<<<(unsigned)((R + IS_RAYS - 1) / IS_RAYS), IS_THREADS, IS_SMEM, st>>>(R, m, cur, k_old, n_new, merge_first, inv_s, last, last_dist_const, last_dist_ray);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int launch_sections_only(int64_t R, const MarchState& m, int cur, int S, float last_dist_const, const float* last_dist_ray, cudaStream_t st) {
    k_sections_only<<<blocks_for(R), TPB, 0, st>>>(R, m, cur, S, last_dist_const, last_dist_ray);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int launch_composite_primary(int64_t R, const MarchState& m, int cur, int S, float last_dist, const float* inv_s,
                             float cos_anneal, const FineBuffers& f, const RayState& rs, const float* pl, bool do_shadow,
                             const MarchState& sh, int n_shadow, float shadow_offset, const float* jitter_shadow,
                             int depth_type, const float* hit_pts, const float* hit_depth, int n_out, const OutsideBuffers& ob,
                             cudaStream_t st) {
    if (S > NRH_MAX_SAMPLES) { set_error("composite: more than %d samples per ray", NRH_MAX_SAMPLES); return NRH_ERR_UNSUPPORTED; }
    NRH_CUDA_CHECK(cudaFuncSetAttribute(k_composite_primary, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_smem(NRH_MAX_SAMPLES)));
    k_composite_primary<<<staged_blocks(R), ST_THREADS, staged_smem(S), st>>>(R, m, cur, S, last_dist, inv_s, cos_anneal, f, rs, pl, do_shadow,
                                                                              sh, n_shadow, shadow_offset, jitter_shadow, depth_type, hit_pts,
                                                                              hit_depth, n_out, ob);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int launch_outside_setup(int64_t R, const MarchState& m, int cur, int S, int n_samples, int n_out, const float* fars,
                         const float* jitter_outside, float sample_dist, const OutsideBuffers& ob, cudaStream_t st) {
    k_outside_setup<<<blocks_for(R), TPB, 0, st>>>(R, m, cur, S, n_samples, n_out, fars, jitter_outside, sample_dist, ob);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int launch_sphere_step(int64_t R, const float* dirs, const float* sdf, float* pts, float* depth, float threshold, float far_limit,
                       int* moving, cudaStream_t st) {
    k_sphere_step<<<blocks_for(R), TPB, 0, st>>>(R, dirs, sdf, pts, depth, threshold, far_limit, moving);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int launch_specular_cue(int64_t R, const NrhConfig& cfg, const RayState& rs, const float* pl, const float* dirs, int warmup, cudaStream_t st) {
    k_specular_cue<<<blocks_for(R), TPB, 0, st>>>(R, cfg, rs, pl, dirs, warmup);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int launch_shade_prep(int64_t R, const NrhConfig& cfg, const MarchState& sh, int cur, int S_shadow, const float* inv_s,
                      float cos_anneal, const float* ssdf, const float* sgx, const float* sgy, const float* sgz,
                      const RayState& rs, const float* pl, const float* dirs, int warmup, bool shadow_marched,
                      float* rayfeat, unsigned char* aux_img, cudaStream_t st) {
    if (S_shadow > NRH_MAX_SAMPLES) { set_error("shade prep: more than %d shadow samples per ray", NRH_MAX_SAMPLES); return NRH_ERR_UNSUPPORTED; }
    NRH_CUDA_CHECK(cudaFuncSetAttribute(k_shade_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_smem(NRH_MAX_SAMPLES)));
    k_shade_prep<<<staged_blocks(R), ST_THREADS, shadow_marched ? staged_smem(S_shadow) : 0, st>>>(
        R, cfg, sh, cur, S_shadow, inv_s, cos_anneal, ssdf, sgx, sgy, sgz, rs, pl, dirs, warmup, shadow_marched, rayfeat, aux_img);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int launch_final_rgb(int64_t R, int S, const FineBuffers& f, const RayState& rs, float* cr, float* cg,
                     float* cb, const float* bg, float* rgb, float* depth, float* vis_out, float* nmap, float* nnmap,
                     float* spec_ray, int n_rough, int n_out, const OutsideBuffers& ob, cudaStream_t st) {
    k_final_rgb<<<blocks_for(R), TPB, 0, st>>>(R, S, f, rs, cr, cg, cb, bg, rgb, depth, vis_out, nmap, nnmap, spec_ray, n_rough,
                                               n_out, ob);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

int launch_to_ray_major(const float* const* src, int C, bool broadcast, int64_t R, int S, float* dst, cudaStream_t st) {
    TransposeArgs a; a.C = C; a.broadcast = broadcast ? 1 : 0;
    for (int c = 0; c < 4; ++c) a.src[c] = c < C ? src[c] : nullptr;
    dim3 grid((unsigned)((R + 31) / 32), (unsigned)((S + 31) / 32)), block(32, 8);
    k_to_ray_major<<<grid, block, 0, st>>>(a, R, S, dst);
    NRH_LAUNCH_CHECK();
    return NRH_OK;
}

}  // namespace nrh
