// Shared declarations of the nrhints_b200 CUDA library (internal).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "../../include/nrhints_b200.h"

namespace nrh {

// ---- fixed network architecture (reference defaults; fields/sdf_field.py:11-36,
//      fields/reflectance_network.py:9-22) -------------------------------------------------
constexpr int HID = 256;            // hidden width of both MLPs
constexpr int SDF_LAYERS = 8;
constexpr int SDF_SKIP = 4;         // layer whose input is cat([h, pe]) / sqrt(2)
constexpr int SDF_FREQ = 6;
constexpr int PE_DIM = 39;          // 3 * (2*6 + 1)
constexpr int PE_PAD = 40;
constexpr int SKIP_H = HID - PE_DIM;  // 217 outputs of lin3
constexpr float SDF_SCALE = 3.0f;
constexpr int COL_FREQ = 4;
constexpr int COL_PE3 = 27;         // 3 * (2*4 + 1)
constexpr int AUX_ROWS = 112;       // non-feature reflectance inputs: 3+27+3+27+9+36 = 105, padded
constexpr int AUX_PTS = 0, AUX_VIEW = 3, AUX_NORMAL = 30, AUX_LIGHT = 33, AUX_VIS = 60, AUX_SPEC = 69;
constexpr int RAYFEAT = 99;         // per-ray encoded inputs: PE(view) 27, PE(light) 27, PE(vis) 9, PE(spec) 36
// outside NeRF (fields/nerf_density_field.py:12-64, reference defaults)
constexpr int NERF_LAYERS = 8;
constexpr int NERF_SKIP = 4;        // the encoded point is re-concatenated (in front) after layer 4 -> layer 5 has 340 inputs
constexpr int NERF_FREQ = 10;
constexpr int NERF_PE = 84;         // 4 * (2*10 + 1)
constexpr int NERF_PE_PAD = 88;
constexpr int NERF_VFREQ = 4;
constexpr int NERF_VPE = 54;        // 6 * (2*4 + 1)
constexpr int NERF_VPE_PAD = 56;
constexpr int NERF_VIEW_H = 128;    // width of the view/light layer

// ---- packed weight buffer (fp32 section; offsets in floats) --------------------------------
struct PackedLayout {
    // scalars
    size_t inv_s;                    // [0] inv_s  [1] 1/inv_s
    // SDF forward: Wt[l] is [Kpad][256] (k-major, zero padded), bias [256]
    size_t sdf_wt[SDF_LAYERS];
    size_t sdf_b[SDF_LAYERS];
    // SDF reverse: native W[l] [rows_pad][256] for l = 1..7 (rows = out dim, padded to x8), l = 0: [256][40]
    size_t sdf_wn[SDF_LAYERS];
    size_t head_w;                   // out_sdf weight [256]
    size_t head_b;                   // [1] (+pad)
    size_t feat_wt;                  // [256][256]
    size_t feat_b;                   // [256]
    // reflectance: layer 0 split into feature part [256][256] and aux part [112][256]
    size_t col_wt0a, col_wt0b, col_b0;
    size_t col_wt[3], col_b[3];      // hidden layers 1..3
    size_t col_w4t;                  // [256][4]
    size_t col_b4;                   // [4]
    // outside NeRF (present only when cfg.use_outside_nerf): trunk Wt[l] [Kpad][256] (layer 0: 88 rows; layer 5: the 256
    // hidden inputs, its 84 encoded-point inputs live in nerf_wt5e [88][256]), heads, view layer split [feature | PE]
    size_t nerf_wt[NERF_LAYERS], nerf_b[NERF_LAYERS], nerf_wt5e;
    size_t nerf_alpha_w, nerf_alpha_b;       // [256], [1]
    size_t nerf_feat_wt, nerf_feat_b;        // [256][256], [256]
    size_t nerf_view_wta, nerf_view_wtb, nerf_view_b;   // [256][256] (128 real columns), [56][256], [256]
    size_t nerf_rgb_wt, nerf_rgb_b;          // [128][4], [4]
    size_t total_floats;
    // tcgen05 section (bytes from the start of the packed buffer); 0 if absent
    size_t tc_offset_bytes;
    size_t total_bytes;
};

PackedLayout make_layout(const NrhConfig& cfg);

struct Strided3 {                    // three coordinate arrays with a common element stride
    const float* x; const float* y; const float* z; int64_t stride;
};

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// api.cu: small fp32 copy queued into the one re-layout launch of the running nrh_pack_weights call
int queue_scaled_copy(const float* src, float* dst, int n, float scale, cudaStream_t st);

#define NRH_CUDA_CHECK(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            nrh::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return NRH_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

#define NRH_LAUNCH_CHECK()                                                                \
    do {                                                                                  \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            nrh::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return NRH_ERR_CUDA;                                                          \
        }                                                                                 \
        nrh::count_launch();                                                              \
    } while (0)

// ---- MLP engines -----------------------------------------------------------------------------
// SDF MLP over N points given as strided coordinate arrays. Outputs are SoA (sdf[N], gx/gy/gz[N])
// and feat row-major [N][256]; any of grad / feat may be null. sdf is dense [N]; gx/gy/gz are
// written at [p*grad_stride].
int sdf_mlp_simt(const float* packed, const PackedLayout& L, Strided3 pts, int64_t N,
                 float* sdf, float* gx, float* gy, float* gz, int64_t grad_stride, float* feat,
                 float* scratch, size_t scratch_bytes, int num_sms, cudaStream_t st);
size_t sdf_mlp_simt_scratch_bytes(int num_sms);

// Reflectance MLP over N = S*R sample-major points (p = j*R + r).
// normals n{x,y,z}[N], feat [N][256], rayfeat [RAYFEAT][R]; outputs c{r,g,b}[N].
int color_mlp_simt(const float* packed, const PackedLayout& L, Strided3 pts, Strided3 normals,
                   const float* feat, const float* rayfeat, int64_t R, int64_t N,
                   float* cr, float* cg, float* cb, int num_sms, cudaStream_t st);

// Outside NeRF over N = St*R sample-major section mid-points (p = j*R + r) of the merged sample set: the kernel forms the
// inverted-sphere point from (o, d, mid) itself.  o/d: SoA [R] each; pl: [R,3]; outputs density[N] (raw), c{r,g,b}[N] (sigmoid).
int nerf_mlp_simt(const float* packed, const PackedLayout& L, const float* const o[3], const float* const d[3], const float* pl,
                  const float* mid, int64_t R, int64_t N, float* density, float* cr, float* cg, float* cb, int num_sms,
                  cudaStream_t st);

}  // namespace nrh
