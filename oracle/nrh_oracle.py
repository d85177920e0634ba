"""CPU oracle for the NRHints ray-march hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a functional (no nn.Module, explicit weight dict) CPU restatement of
the reference algorithm, written against /root/reference (commit 291800d):

    models/neus_hint_model.py:21-65     sample_pdf            -> sample_pdf_det
    models/neus_hint_model.py:269-315   up_sample             -> up_sample
    models/neus_hint_model.py:317-331   cat_z_vals            -> merge_samples
    models/neus_hint_model.py:333-357   get_alpha             -> neus_alpha
    models/neus_hint_model.py:373-432   get_visibility        -> shadow_visibility
    models/neus_hint_model.py:475-651   render_core           -> render_core
    models/neus_hint_model.py:653-751   forward               -> render_forward
    fields/sdf_field.py:106-148         SDFNetwork.forward/.gradient -> sdf_mlp (manual reverse pass)
    fields/encodings.py:155-176         NeRFEncoding.forward  -> fourier_encode
    fields/reflectance_network.py:68-96 ReflectanceNetwork.forward -> reflectance_mlp
    fields/nerf_density_field.py:66-89  NeRF.forward (outside model)  -> nerf_mlp
    models/neus_hint_model.py:434-473   render_outside        -> render_outside
    models/neus_hint_model.py:677-694   outside sample positions -> outside_z

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it.  The product (nrhints_b200/) never does: the product path is the
CUDA library and it fails loudly when that library is missing.

PARITY PINNING: the reference ships no golden vectors or tests, so this
oracle is pinned against the reference module itself, imported unmodified in the
build container: tests/golden/make_golden.py runs reference and oracle on the same
seeded rays/weights, asserts agreement, and commits the reference outputs as
fixtures under tests/golden/*.npz.  tests/test_oracle_golden.py re-checks the oracle
against those fixtures wherever the tests run (the GPU box has no /root/reference).

dtype: float32 by default; pass dtype=torch.float64 for a "truth" evaluation.  Every tensor is created on the
device of its inputs, so the same restatement also runs with torch CUDA ops (tools_bench_extra.py times it on the B200
as the "unfused PyTorch on the same GPU" baseline of BASELINE.md section 3.6); the CPU is the default and what the tests use.
The SDF input gradient is computed by an explicit reverse pass (not autograd) so
that it doubles as the executable spec of the CUDA kernel's reverse sweep; all ops are
differentiable torch ops, so autograd through it yields the second-order terms
needed by training-mode goldens.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class OracleConfig:
    """Knobs of the hot path (names follow NeuSRendererConfig / SDFNetConfig /
    ReflectanceNetConfig, models/neus_hint_model.py:133-174, fields/sdf_field.py:11-36)."""
    n_samples: int = 64
    n_importance_samples: int = 64
    up_sample_steps: int = 4
    n_shadow_samples: int = 64
    n_shadow_importance_samples: int = 64
    shadow_hint: bool = True
    specular_hint: bool = True
    shadow_ray_offset: float = 1e-2
    specular_roughness: List[float] = field(default_factory=lambda: [0.02, 0.05, 0.13, 0.34])
    normalized_normals: bool = True      # NormalComputationType.NormalizedAnalytic
    depth_type: str = "alpha_blending"   # DepthComputationType value: alpha_blending | maximum_point | sphere_tracing
    # SDF network
    sdf_n_layers: int = 8
    sdf_hidden: int = 256
    sdf_skip_in: tuple = (4,)
    sdf_multires: int = 6
    sdf_scale: float = 3.0
    sdf_feat: int = 256
    # reflectance network
    refl_n_layers: int = 4
    refl_multires: int = 4
    refl_squeeze_out: bool = True
    # outside NeRF (NeRF++ background, fields/nerf_density_field.py:12-26; off by default)
    use_outside_nerf: bool = False
    n_outside_samples: int = 32
    nerf_n_layers: int = 8
    nerf_multires: int = 10
    nerf_multires_view: int = 4
    nerf_skips: tuple = (4,)

    @staticmethod
    def from_model_config(cfg) -> "OracleConfig":
        r, s, c = cfg.renderer, cfg.sdf_network, cfg.reflectance_network
        nerf = getattr(cfg, "outside_nerf", None)
        extra = {}
        if nerf is not None:
            extra = dict(nerf_n_layers=nerf.n_layers, nerf_multires=nerf.multi_res, nerf_multires_view=nerf.multi_res_view,
                         nerf_skips=tuple(nerf.skips))
        return OracleConfig(
            use_outside_nerf=bool(r.use_outside_nerf), n_outside_samples=r.n_outside_samples, **extra,
            n_samples=r.n_samples, n_importance_samples=r.n_importance_samples,
            up_sample_steps=r.up_sample_steps, n_shadow_samples=r.n_shadow_samples,
            n_shadow_importance_samples=r.n_shadow_importance_samples,
            shadow_hint=r.shadow_hint, specular_hint=r.specular_hint,
            shadow_ray_offset=r.shadow_ray_offset, specular_roughness=list(r.specular_roughness),
            normalized_normals=(getattr(r.normal_type, "value", r.normal_type) == "normalized_analytic"),
            depth_type=getattr(r.depth_type, "value", r.depth_type),
            sdf_n_layers=s.n_layers, sdf_hidden=s.d_hidden, sdf_skip_in=tuple(s.skip_in),
            sdf_multires=s.multi_res, sdf_scale=s.scale, sdf_feat=s.d_out_feat,
            refl_n_layers=c.n_layers, refl_multires=c.multi_res, refl_squeeze_out=c.squeeze_out)


# --------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------
def effective_weights(state: Dict[str, Tensor], dtype=torch.float32) -> Dict[str, Tensor]:
    """state_dict (weight_g / weight_v / bias per layer) -> plain W, b.

    nn.utils.weight_norm(dim=0): W = g * v / ||v||_2 over every dim but 0
    (fields/sdf_field.py:81-82,100-101; fields/reflectance_network.py:58-59)."""
    out: Dict[str, Tensor] = {}
    prefixes = sorted({k.rsplit(".", 1)[0] for k in state if k.endswith("weight_v") or k.endswith(".weight")})
    for p in prefixes:
        if p + ".weight_v" in state:
            v = state[p + ".weight_v"].to(dtype)
            g = state[p + ".weight_g"].to(dtype)
            w = v * (g / torch.linalg.vector_norm(v, ord=2, dim=1, keepdim=True))
        else:
            w = state[p + ".weight"].to(dtype)
        out[p + ".W"] = w
        out[p + ".b"] = state[p + ".bias"].to(dtype)
    out["variance"] = state["deviation_network.variance"].to(dtype)
    return out


# --------------------------------------------------------------------------------------
# encodings and MLPs
# --------------------------------------------------------------------------------------
def fourier_encode(x: Tensor, n_freq: int) -> Tensor:
    """[N,D] -> [N, D*(2F+1)] = [x, sin(x_d 2^k) (d-major,k-minor), sin(x_d 2^k + pi/2)]
    (fields/encodings.py:168-176; the cosine is sin(.+pi/2) evaluated in the working dtype)."""
    freqs = 2 ** torch.linspace(0.0, n_freq - 1, n_freq, dtype=x.dtype, device=x.device)
    s = (x[..., None] * freqs).reshape(*x.shape[:-1], -1)
    enc = torch.sin(torch.cat([s, s + torch.pi / 2.0], dim=-1))
    return torch.cat([x, enc], dim=-1)


def fourier_encode_jvp_T(x: Tensor, n_freq: int, g_enc: Tensor) -> Tensor:
    """Transposed Jacobian of fourier_encode applied to g_enc: [N, D*(2F+1)] -> [N, D]."""
    D = x.shape[-1]
    freqs = 2 ** torch.linspace(0.0, n_freq - 1, n_freq, dtype=x.dtype, device=x.device)
    s = x[..., None] * freqs                                            # [N,D,F]
    g_id = g_enc[..., :D]
    g_sin = g_enc[..., D:D + D * n_freq].reshape(*x.shape, n_freq)
    g_cos = g_enc[..., D + D * n_freq:].reshape(*x.shape, n_freq)
    return g_id + (g_sin * torch.cos(s) * freqs).sum(-1) + (g_cos * torch.cos(s + torch.pi / 2.0) * freqs).sum(-1)


def _softplus100(x: Tensor) -> Tensor:
    return F.softplus(x, beta=100.0, threshold=20.0)


def _softplus100_grad(x: Tensor) -> Tensor:
    """d softplus(beta=100)/dx = z/(z+1), z = exp(100 x), i.e. sigmoid(100 x); 1 beyond the linear threshold.
    Written with sigmoid so that autograd THROUGH this expression (second-order terms of the training loss) stays
    finite: exp(100 x) overflows for x > 0.887 and would poison the where() with inf/inf."""
    return torch.where(x * 100.0 > 20.0, torch.ones_like(x), torch.sigmoid(x * 100.0))


def sdf_mlp(W: Dict[str, Tensor], pts: Tensor, cfg: OracleConfig, want_feat=False, want_grad=False):
    """SDFNetwork.forward (+ .gradient) restated (fields/sdf_field.py:106-148).

    returns dict(sdf [N,1], feat [N,256]?, grad [N,3]?)"""
    x0 = pts * cfg.sdf_scale
    e = fourier_encode(x0, cfg.sdf_multires)
    inv_sqrt2 = 1.0 / math.sqrt(2.0)
    h = e
    sig = []                                            # softplus' at every layer, for the reverse sweep
    for l in range(cfg.sdf_n_layers):
        if l in cfg.sdf_skip_in:
            h = torch.cat([h, e], dim=1) / math.sqrt(2.0)
        pre = F.linear(h, W[f"sdf_network.lin{l}.W"], W[f"sdf_network.lin{l}.b"])
        h = _softplus100(pre)
        if want_grad:
            sig.append(_softplus100_grad(pre))
    out = {"sdf": F.linear(h, W["sdf_network.out_sdf.W"], W["sdf_network.out_sdf.b"]) / cfg.sdf_scale}
    if want_feat:
        out["feat"] = F.linear(h, W["sdf_network.out_feat.W"], W["sdf_network.out_feat.b"])
    if want_grad:
        n_e = e.shape[1]
        g = (W["sdf_network.out_sdf.W"] / cfg.sdf_scale).expand(pts.shape[0], -1)     # d sdf / d h_last
        g_e = torch.zeros_like(e)
        for l in reversed(range(cfg.sdf_n_layers)):
            g = (g * sig[l]) @ W[f"sdf_network.lin{l}.W"]                           # d / d (layer input)
            if l in cfg.sdf_skip_in:
                g_e = g_e + g[:, -n_e:] * inv_sqrt2
                g = g[:, :-n_e] * inv_sqrt2
        g_e = g_e + g
        out["grad"] = fourier_encode_jvp_T(x0, cfg.sdf_multires, g_e) * cfg.sdf_scale
    return out


def reflectance_mlp(W, pts, normals, view_dirs, feat, lights, vis, spec, cfg: OracleConfig) -> Tensor:
    """ReflectanceNetwork.forward (fields/reflectance_network.py:68-96); concat order
    [pts, PE(view), normal, PE(light), feat, PE(vis)?, PE(spec)?]."""
    F_ = cfg.refl_multires
    parts = [pts, fourier_encode(view_dirs, F_), normals, fourier_encode(lights, F_), feat]
    if cfg.shadow_hint:
        parts.append(fourier_encode(vis, F_))
    if cfg.specular_hint:
        parts.append(fourier_encode(spec, F_))
    x = torch.cat(parts, dim=-1)
    n = cfg.refl_n_layers + 1
    for l in range(n):
        x = F.linear(x, W[f"color_network.lin{l}.W"], W[f"color_network.lin{l}.b"])
        if l < n - 1:
            x = torch.relu(x)
    return torch.sigmoid(x) if cfg.refl_squeeze_out else x


def nerf_mlp(W, pts4: Tensor, views: Tensor, pls: Tensor, cfg: OracleConfig):
    """NeRF.forward of the outside model (fields/nerf_density_field.py:66-89): PE(10) of the 4-D inverted-sphere point,
    8 x 256 ReLU with the encoded point re-concatenated IN FRONT after layer 4, density head, feature head,
    one 128-wide view/light layer on [feature, PE(4)(view, light)], rgb head.  Returns (density [N,1], rgb_raw [N,3])."""
    e = fourier_encode(pts4, cfg.nerf_multires)
    ev = fourier_encode(torch.cat([views, pls], dim=-1), cfg.nerf_multires_view)
    h = e
    for i in range(cfg.nerf_n_layers):
        h = torch.relu(F.linear(h, W[f"outside_nerf.pts_linears.{i}.W"], W[f"outside_nerf.pts_linears.{i}.b"]))
        if i in cfg.nerf_skips:
            h = torch.cat([e, h], dim=-1)
    density = F.linear(h, W["outside_nerf.alpha_linear.W"], W["outside_nerf.alpha_linear.b"])
    feature = F.linear(h, W["outside_nerf.feature_linear.W"], W["outside_nerf.feature_linear.b"])
    h = torch.cat([feature, ev], dim=-1)
    h = torch.relu(F.linear(h, W["outside_nerf.views_linears.0.W"], W["outside_nerf.views_linears.0.b"]))
    rgb = F.linear(h, W["outside_nerf.rgb_linear.W"], W["outside_nerf.rgb_linear.b"])
    return density, rgb


def outside_z(cfg: OracleConfig, far: Tensor, jitter_outside: Optional[Tensor], dtype) -> Tensor:
    """Sample positions of the outside model (models/neus_hint_model.py:677-694): inverse-depth spacing beyond `far`."""
    n_out = cfg.n_outside_samples
    zo = torch.linspace(1e-3, 1.0 - 1.0 / (n_out + 1.0), n_out, dtype=dtype, device=far.device)
    if jitter_outside is not None:
        mids = 0.5 * (zo[..., 1:] + zo[..., :-1])
        upper = torch.cat([mids, zo[..., -1:]], -1)
        lower = torch.cat([zo[..., :1], mids], -1)
        zo = lower[None, :] + (upper - lower)[None, :] * jitter_outside.to(dtype)
    return far / torch.flip(zo, dims=[-1]) + 1.0 / cfg.n_samples


def render_outside(W, cfg: OracleConfig, o, d, pl, z_feed, sample_dist):
    """render_outside (models/neus_hint_model.py:434-473) -> (sigmoid colour [R,St,3], alpha [R,St])."""
    R, St = z_feed.shape
    dists = torch.cat([z_feed[..., 1:] - z_feed[..., :-1], torch.full((R, 1), sample_dist, dtype=z_feed.dtype, device=z_feed.device)], -1)
    mid_z = z_feed + dists * 0.5
    pts = o[:, None, :] + d[:, None, :] * mid_z[..., :, None]
    dis = torch.linalg.norm(pts, ord=2, dim=-1, keepdim=True).clip(1.0, 1e10)
    pts4 = torch.cat([pts / dis, 1.0 / dis], dim=-1).reshape(-1, 4)
    dirs = d[:, None, :].expand(R, St, 3).reshape(-1, 3)
    pls = pl[:, None, :].expand(R, St, 3).reshape(-1, 3)
    density, rgb = nerf_mlp(W, pts4, dirs, pls, cfg)
    color = torch.sigmoid(rgb).reshape(R, St, 3)
    alpha = 1.0 - torch.exp(-F.softplus(density.reshape(R, St)) * dists)
    return color, alpha


# --------------------------------------------------------------------------------------
# sampler
# --------------------------------------------------------------------------------------
def sample_pdf_det(bins: Tensor, weights: Tensor, n: int) -> Tensor:
    """Deterministic inverse-CDF sampling (models/neus_hint_model.py:21-65 with det=True)."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = torch.linspace(0.0, 1.0, steps=n, dtype=bins.dtype, device=bins.device).expand(list(cdf.shape[:-1]) + [n]).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bin_b, bin_a = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    return bin_b + t * (bin_a - bin_b)


def _excl_cumprod(one_minus_alpha: Tensor) -> Tensor:
    ones = torch.ones_like(one_minus_alpha[:, :1])
    return torch.cumprod(torch.cat([ones, one_minus_alpha], -1), -1)[:, :-1]


def up_sample(o: Tensor, d: Tensor, z: Tensor, sdf: Tensor, n_new: int, inv_s: float) -> Tensor:
    """NeuS importance step with a fixed sharpness (models/neus_hint_model.py:269-315)."""
    pts = o[:, None, :] + d[:, None, :] * z[..., :, None]
    radius = torch.linalg.norm(pts, ord=2, dim=-1)
    inside = (radius[:, :-1] < 1.0) | (radius[:, 1:] < 1.0)
    prev_sdf, next_sdf = sdf[:, :-1], sdf[:, 1:]
    prev_z, next_z = z[:, :-1], z[:, 1:]
    mid_sdf = (prev_sdf + next_sdf) * 0.5
    cos = (next_sdf - prev_sdf) / (next_z - prev_z + 1e-5)
    prev_cos = torch.cat([torch.zeros_like(cos[:, :1]), cos[:, :-1]], dim=-1)
    cos = torch.minimum(prev_cos, cos).clip(-1e3, 0.0) * inside
    dist = next_z - prev_z
    prev_cdf = torch.sigmoid((mid_sdf - cos * dist * 0.5) * inv_s)
    next_cdf = torch.sigmoid((mid_sdf + cos * dist * 0.5) * inv_s)
    alpha = (prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)
    w = alpha * _excl_cumprod(1.0 - alpha + 1e-7)
    return sample_pdf_det(z, w, n_new).detach()


def merge_samples(W, cfg, o, d, z, z_new, sdf, last: bool):
    """cat_z_vals (models/neus_hint_model.py:317-331): sort, evaluate the SDF at the new
    z themselves (not mid-points) unless this is the last step."""
    R = z.shape[0]
    z_cat, index = torch.sort(torch.cat([z, z_new], dim=-1), dim=-1)
    if not last:
        pts = (o[:, None, :] + d[:, None, :] * z_new[..., :, None]).reshape(-1, 3)
        new_sdf = sdf_mlp(W, pts, cfg)["sdf"].reshape(R, -1)
        sdf = torch.gather(torch.cat([sdf, new_sdf], dim=-1), 1, index)
    return z_cat, sdf


def hierarchical_z(W, cfg, o, d, z, n_importance: int, steps: int):
    """coarse SDF pass + `steps` importance steps with sharpness 64*2^i
    (models/neus_hint_model.py:696-713 and :397-412)."""
    R, n = z.shape
    if n_importance <= 0:
        return z
    pts = (o[:, None, :] + d[:, None, :] * z[..., :, None]).reshape(-1, 3)
    sdf = sdf_mlp(W, pts, cfg)["sdf"].reshape(R, n)
    for i in range(steps):
        z_new = up_sample(o, d, z, sdf, n_importance // steps, 64.0 * 2 ** i)
        z, sdf = merge_samples(W, cfg, o, d, z, z_new, sdf, last=(i + 1 == steps))
    return z


def neus_alpha(W, cfg, pts, dists, dirs, cos_anneal: float):
    """get_alpha (models/neus_hint_model.py:333-357). Returns alpha [N,1], sdf, grad [N,3], inv_s."""
    r = sdf_mlp(W, pts, cfg, want_grad=True)
    sdf, grad = r["sdf"], r["grad"]
    inv_s = torch.exp(W["variance"] * 10.0).clip(1e-6, 1e6)
    true_cos = (dirs * grad).sum(-1, keepdim=True)
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal) + F.relu(-true_cos) * cos_anneal)
    est_next = sdf + iter_cos * dists.reshape(-1, 1) * 0.5
    est_prev = sdf - iter_cos * dists.reshape(-1, 1) * 0.5
    prev_cdf = torch.sigmoid(est_prev * inv_s)
    next_cdf = torch.sigmoid(est_next * inv_s)
    alpha = ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).clip(0, 1)
    return alpha, sdf, grad, inv_s


def shadow_visibility(W, cfg: OracleConfig, lights, targets, cos_anneal=1.0, jitter: Optional[Tensor] = None):
    """get_visibility (models/neus_hint_model.py:373-432): march from the light to the
    estimated hit point, return the transmittance in front of the last sample."""
    n = cfg.n_shadow_samples
    o = lights
    dvec = targets - o
    L = torch.linalg.norm(dvec, ord=2, dim=-1, keepdim=True)
    sample_dist = L / n
    d = dvec / L
    z = torch.linspace(0.0, 1.0, steps=n, dtype=o.dtype, device=o.device) * L * (1.0 - cfg.shadow_ray_offset)
    if jitter is not None:                                     # stratified, one draw per sample (:388-395)
        mids = 0.5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * jitter
    z = hierarchical_z(W, cfg, o, d, z, cfg.n_shadow_importance_samples, 4)
    R, S = z.shape
    dists = torch.cat([z[..., 1:] - z[..., :-1], sample_dist.expand(R, 1)], -1)
    mid_z = z + dists * 0.5
    pts = (o[:, None, :] + d[:, None, :] * mid_z[..., :, None]).reshape(-1, 3)
    dirs = d[:, None, :].expand(R, S, 3).reshape(-1, 3)
    alpha, _, _, _ = neus_alpha(W, cfg, pts, dists, dirs, cos_anneal)
    taus = _excl_cumprod(1.0 - alpha.reshape(R, S) + 1e-7)
    return taus[..., -1:], z


def sphere_trace(W, cfg: OracleConfig, o, d, num_iterations=2000, threshold=1e-4, far=100.0):
    """sphere_trace (models/neus_hint_model.py:359-371): march p += sdf * d until |sdf| < threshold or depth > far."""
    pts = o
    depths = torch.zeros((o.shape[0], 1), dtype=o.dtype, device=o.device)
    for _ in range(num_iterations):
        sdf = sdf_mlp(W, pts, cfg)["sdf"]
        converged = (torch.abs(sdf) < threshold) | (depths > far)
        pts = torch.where(converged, pts, pts + sdf * d)
        depths = torch.where(converged, depths, depths + sdf)
        if converged.all():
            break
    return pts, depths


def specular_cue(cfg: OracleConfig, hit_normal, lights, hits, d):
    """4-lobe Cook-Torrance cue (models/neus_hint_model.py:588-616)."""
    l = F.normalize(lights - hits, dim=-1, p=2)
    v = F.normalize(-d, dim=-1, p=2)
    h = F.normalize(l + v, dim=-1, p=2)
    n_l = (hit_normal * l).sum(-1).clip(0.0, 1.0)
    n_v = (hit_normal * v).sum(-1).clip(0.0, 1.0)
    n_h = (hit_normal * h).sum(-1).clip(0.0, 1.0)
    h_v = (h * v).sum(-1).clip(0.0, 1.0)
    n_h2 = torch.pow(n_h, 2)
    cues = []
    for rough in cfg.specular_roughness:
        k = (rough + 1.0) * (rough + 1.0) / 8.0
        g = (n_v / (n_v * (1.0 - k) + k)) * (n_l / (n_l * (1.0 - k) + k))
        a2 = rough * rough
        ndf = a2 / (torch.pi * torch.pow(n_h2 * (a2 - 1.0) + 1.0, 2))
        f = 0.04 + 0.96 * torch.pow(1.0 - h_v, 5)
        cues.append(ndf * g * f / (4.0 * n_v + 1e-3))
    return torch.stack(cues, dim=-1)


# --------------------------------------------------------------------------------------
# render
# --------------------------------------------------------------------------------------
def render_core(W, cfg: OracleConfig, o, d, pl, z, sample_dist, bg, cos_anneal, warmup, jitter_shadow,
                background_alpha=None, background_color=None):
    """render_core (models/neus_hint_model.py:475-651); background_alpha / background_color [R, S + n_outside] come from
    render_outside when the outside NeRF is on (:517-519, :630-633)."""
    R, S = z.shape
    dists = torch.cat([z[..., 1:] - z[..., :-1], torch.full((R, 1), sample_dist, dtype=z.dtype, device=z.device)], -1)
    mid_z = z + dists * 0.5
    pts = (o[:, None, :] + d[:, None, :] * mid_z[..., :, None]).reshape(-1, 3)
    dirs = d[:, None, :].expand(R, S, 3).reshape(-1, 3)
    pls = pl[:, None, :].expand(R, S, 3).reshape(-1, 3)

    feat = sdf_mlp(W, pts, cfg, want_feat=True)["feat"]
    alpha, sdf, grad, inv_s = neus_alpha(W, cfg, pts, dists, dirs, cos_anneal)
    alpha = alpha.reshape(R, S)
    pts_norm = torch.linalg.norm(pts, ord=2, dim=-1).reshape(R, S)
    inside = (pts_norm < 1.0).to(z.dtype).detach()
    if background_alpha is not None:                                # :517-519
        alpha = alpha * inside + background_alpha[:, :S] * (1.0 - inside)
        alpha = torch.cat([alpha, background_alpha[:, S:]], dim=-1)
    weights_all = alpha * _excl_cumprod(1.0 - alpha + 1e-7)
    wsum = weights_all.sum(-1, keepdim=True)
    weights = weights_all[:, :S]                                    # neus_weights (:525)
    with torch.no_grad():
        if cfg.depth_type == "sphere_tracing":                      # :528-529
            hits, depth = sphere_trace(W, cfg, o, d, 2000, 1e-4, 100.0)
        elif cfg.depth_type == "maximum_point":                     # :534-538
            depth = torch.gather(mid_z, 1, torch.argmax(weights, dim=1, keepdim=True))
            hits = o + d * depth
        else:                                                       # :530-533
            depth = (mid_z[..., None] * weights[..., None]).sum(1)
            hits = o + d * depth

    vis_map = None
    vis = None
    z_shadow = None
    if cfg.shadow_hint:
        if warmup:
            vis_map = torch.zeros((R, 1), dtype=z.dtype, device=z.device)
        else:
            with torch.no_grad():
                vis_map, z_shadow = shadow_visibility(W, cfg, pl, hits, cos_anneal, jitter_shadow)
        vis = vis_map[:, None, :].expand(R, S, 1).reshape(-1, 1)

    n_hat = F.normalize(grad, dim=-1, p=2)
    hit_n = F.normalize((n_hat.reshape(R, S, 3) * weights[..., None]).sum(1), dim=-1, p=2)
    spec = None
    if cfg.specular_hint:
        nr = len(cfg.specular_roughness)
        if warmup:
            spec_ray = torch.zeros((R, nr), dtype=z.dtype, device=z.device)
        else:
            with torch.no_grad():
                spec_ray = specular_cue(cfg, hit_n, pl, hits, d)
        spec = spec_ray[:, None, :].expand(R, S, nr).reshape(-1, nr)

    normal_in = n_hat if cfg.normalized_normals else grad
    color = reflectance_mlp(W, pts, normal_in, dirs, feat, pls, vis, spec, cfg).reshape(R, S, 3)
    if background_alpha is not None:                                # :630-633
        color = color * inside[:, :, None] + background_color[:, :S] * (1.0 - inside)[:, :, None]
        color = torch.cat([color, background_color[:, S:]], dim=1)
    rgb = (color * weights_all[..., None]).sum(1)
    if bg is not None:
        rgb = rgb + bg * (1.0 - wsum)
    out = {
        "rgb": rgb, "depth": depth, "weights": weights_all,
        "s_val": (1.0 / inv_s).expand(R, S),
        "inside_sphere": inside, "relax_inside_sphere": inside,         # quirk Q1 (:746)
        "analytic_normals": grad.reshape(R, S, 3),
        "normalized_analytic_normals": n_hat.reshape(R, S, 3),
        "visibilities": vis_map,
        "specular_cue": spec.reshape(R, S, -1) if spec is not None else None,
        # extras for debugging parity (not RenderOutput fields)
        "z_vals": z, "z_shadow": z_shadow, "sampled_color": color, "sdf": sdf.reshape(R, S),
    }
    return out


def render_forward(state: Dict[str, Tensor], cfg: OracleConfig, origins, directions, pl_positions, nears, fars,
                   is_training=False, background_rgb: Optional[Tensor] = None, cos_anneal: float = 1.0,
                   warmup: bool = False, jitter_primary: Optional[Tensor] = None,
                   jitter_shadow: Optional[Tensor] = None, dtype=torch.float32, effective: bool = False,
                   jitter_outside: Optional[Tensor] = None):
    """NeuSHintRenderer.forward (models/neus_hint_model.py:653-751).

    `state` is a renderer state_dict (or already-effective weights if effective=True).
    RNG is explicit: jitter_primary [R,1] ~ U(0,1) (:682), jitter_outside [R,n_outside] (:689, outside NeRF only) and
    jitter_shadow [R,n_shadow] (:394) are drawn by the caller in that order when is_training, else None.
    cos_anneal = min(1, global_step/anneal_end) when training (:669-671)."""
    W = state if effective else effective_weights(state, dtype)
    o, d, pl = origins.to(dtype), directions.to(dtype), pl_positions.to(dtype)
    near, far = nears.to(dtype), fars.to(dtype)
    bg = background_rgb.to(dtype) if background_rgb is not None else None
    n = cfg.n_samples
    sample_dist = 2.0 / n
    z = near + (far - near) * torch.linspace(0.0, 1.0, n, dtype=dtype, device=near.device)[None, :]
    if is_training:
        assert jitter_primary is not None
        z = z + (jitter_primary.to(dtype) - 0.5) * 2.0 / n
    with torch.no_grad():
        z = hierarchical_z(W, cfg, o, d, z, cfg.n_importance_samples, cfg.up_sample_steps)
    js = jitter_shadow.to(dtype) if (is_training and jitter_shadow is not None) else None
    bg_alpha = bg_color = None
    if cfg.use_outside_nerf:                                         # :716-724
        z_out = outside_z(cfg, far, jitter_outside if is_training else None, dtype)
        z_feed, _ = torch.sort(torch.cat([z, z_out.expand(z.shape[0], -1)], dim=-1), dim=-1)
        bg_color, bg_alpha = render_outside(W, cfg, o, d, pl, z_feed, sample_dist)
    return render_core(W, cfg, o, d, pl, z, sample_dist, bg, cos_anneal if is_training else 1.0, warmup, js,
                       background_alpha=bg_alpha, background_color=bg_color)


def training_loss(out, rgb_gt, igr_weight=0.1):
    """pipelines/base_pipeline.py:57-62."""
    rgb_loss = F.l1_loss(out["rgb"], rgb_gt, reduction="sum") / (out["rgb"].shape[0] + 1e-5)
    gerr = (torch.linalg.norm(out["analytic_normals"], ord=2, dim=-1) - 1.0) ** 2
    relax = out["relax_inside_sphere"]
    eik = (relax * gerr).sum() / (relax.sum() + 1e-5)
    return rgb_loss + eik * igr_weight


# --------------------------------------------------------------------------------------
# deterministic synthetic workload -- shared by tests and bench
# --------------------------------------------------------------------------------------
def synthetic_rays(R: int, seed: int = 3407, crop: int = 800):
    """800x800 pinhole (camera_angle_x 0.6911), camera on a radius-4 sphere at -30deg elevation,
    random azimuth per ray, random pixel per ray, light on a radius-4.5 sphere; near/far from the
    unit sphere (camera/ray_generator.py:133-139).  crop < 800 restricts the pixels to the central
    crop x crop window (tests use it so that most rays hit the radius-0.5 init sphere)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    H = W_ = 800
    fx = 0.5 * W_ / math.tan(0.5 * 0.6911)
    theta = rng.uniform(-180.0, 180.0, R) / 180.0 * math.pi
    phi = -30.0 / 180.0 * math.pi
    lo, hi = (H - crop) // 2, (H + crop) // 2
    hh = rng.integers(lo, hi, R).astype(np.float64) + 0.5
    ww = rng.integers(lo, hi, R).astype(np.float64) + 0.5
    dirs = np.stack([(ww - 400.0) / fx, -(hh - 400.0) / fx, -np.ones(R)], -1)
    # camera-to-world: camera at radius 4 looking at the origin
    cp, sp = math.cos(phi), math.sin(phi)
    cam = np.stack([4.0 * cp * np.sin(theta), np.full(R, -4.0 * sp), 4.0 * cp * np.cos(theta)], -1)
    fwd = -cam / np.linalg.norm(cam, axis=-1, keepdims=True)
    up = np.tile(np.array([0.0, 1.0, 0.0]), (R, 1))
    right = np.cross(fwd, up); right /= np.linalg.norm(right, axis=-1, keepdims=True)
    upv = np.cross(right, fwd)
    d = dirs[:, :1] * right + dirs[:, 1:2] * upv + (-dirs[:, 2:3]) * fwd
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    pl = rng.normal(size=(R, 3)); pl = 4.5 * pl / np.linalg.norm(pl, axis=-1, keepdims=True)
    o32 = torch.tensor(cam, dtype=torch.float32)
    d32 = F.normalize(torch.tensor(d, dtype=torch.float32), dim=-1, p=2)
    a = torch.sum(d32 ** 2, dim=-1, keepdim=True)
    b = 2.0 * torch.sum(o32 * d32, dim=-1, keepdim=True)
    mid = 0.5 * (-b) / a
    return {"origins": o32, "directions": d32, "pl_positions": torch.tensor(pl, dtype=torch.float32),
            "nears": mid - 1.0, "fars": mid + 1.0}


# --------------------------------------------------------------------------------------
# explicit backward of (sdf, feat, grad) = sdf_mlp(pts)  -- executable spec of the fused CUDA backward
# --------------------------------------------------------------------------------------
def _softplus100_grad2(x: Tensor) -> Tensor:
    """d^2 softplus(beta=100)/dx^2 = 100 s (1 - s), s = sigmoid(100 x); 0 beyond the linear threshold."""
    s = torch.sigmoid(x * 100.0)
    return torch.where(x * 100.0 > 20.0, torch.zeros_like(x), 100.0 * s * (1.0 - s))


def sdf_mlp_backward(W: Dict[str, Tensor], pts: Tensor, cfg: OracleConfig, d_sdf: Tensor, d_feat: Tensor, d_grad: Tensor):
    """Vector-Jacobian product of sdf_mlp(want_feat, want_grad) written out by hand (no autograd): given the adjoints
    d_sdf [N,1], d_feat [N,256], d_grad [N,3] of its three outputs, returns d_pts [N,3] and the gradients of every
    effective weight / bias.  Because `grad` is itself the result of a reverse sweep (fields/sdf_field.py:136-148 with
    create_graph=True), its adjoint needs the second-order terms.  Two chained passes per point:

      phase A (adjoint of the reverse sweep; runs in FORWARD layer order with the forward weights):
          gb_0 = 3 PE'(3x) d_grad ;   ub_l = W_l gb_l ;   gb_{l+1} = s'_l * ub_l ;   t_l = s''_l * g_{l+1} * ub_l
          dW_l += u_l (x) gb_l                      (u_l = s'_l * g_{l+1}: the reverse-sweep signal of the forward call)
      phase B (ordinary backward of the forward; REVERSE layer order with the transposed weights):
          ab_8 = w_s d_sdf / 3 + W_f^T d_feat ;   zb_l = s'_l * ab_{l+1} + t_l ;   ab_l = W_l^T zb_l
          dW_l += zb_l (x) a_l ;  db_l += zb_l

    where a_l / z_l are the forward activations / pre-activations, s' / s'' the softplus derivatives at z_l and g_l the
    reverse-sweep adjoints.  The skip connection (a_4 = [softplus(z_3), e] / sqrt 2) splits / joins both chains.
    This is the per-tile GEMM chain of the CUDA backward kernel; tests check it against torch.autograd in float64."""
    L = cfg.sdf_n_layers
    scale = cfg.sdf_scale
    inv_sqrt2 = 1.0 / math.sqrt(2.0)
    x0 = pts * scale
    e = fourier_encode(x0, cfg.sdf_multires)
    n_e = e.shape[1]
    Wl = [W[f"sdf_network.lin{l}.W"] for l in range(L)]
    bl = [W[f"sdf_network.lin{l}.b"] for l in range(L)]
    w_s, W_f = W["sdf_network.out_sdf.W"], W["sdf_network.out_feat.W"]
    # ---- forward, keeping a_l (layer inputs), s', s'' ----
    a, s1, s2 = [], [], []
    h = e
    for l in range(L):
        if l in cfg.sdf_skip_in:
            h = torch.cat([h, e], dim=1) * inv_sqrt2
        a.append(h)
        pre = F.linear(h, Wl[l], bl[l])
        s1.append(_softplus100_grad(pre)); s2.append(_softplus100_grad2(pre))
        h = _softplus100(pre)
    a.append(h)                                                   # a_8
    # ---- reverse sweep, keeping g_{l+1} (adjoint of a_{l+1}, before the softplus' gate) and u_l ----
    g_next, u = [None] * L, [None] * L
    g = (w_s / scale).expand(pts.shape[0], -1)
    ge_skip = None
    for l in reversed(range(L)):
        g_next[l] = g
        u[l] = g * s1[l]
        g = u[l] @ Wl[l]
        if l in cfg.sdf_skip_in:
            ge_skip = g[:, -n_e:] * inv_sqrt2
            g = g[:, :-n_e] * inv_sqrt2
    g_e = g + (ge_skip if ge_skip is not None else 0.0)           # adjoint of e
    # ---- phase A ----
    D, Fq = pts.shape[-1], cfg.sdf_multires
    freqs = 2 ** torch.linspace(0.0, Fq - 1, Fq, dtype=pts.dtype, device=pts.device)
    sarg = x0[..., None] * freqs                                                     # [N,3,F]
    dq = d_grad * scale                                                             # adjoint of PE'^T g_e (a 3-vector)
    gb_e = torch.cat([dq, (dq[..., None] * torch.cos(sarg) * freqs).reshape(pts.shape[0], -1),
                      (dq[..., None] * torch.cos(sarg + torch.pi / 2.0) * freqs).reshape(pts.shape[0], -1)], dim=-1)
    # second derivative of the encoding: d/dx0 of (PE'(x0)^T g_e) contracted with dq
    ge_sin = g_e[:, D:D + D * Fq].reshape(-1, D, Fq)
    ge_cos = g_e[:, D + D * Fq:].reshape(-1, D, Fq)
    d_x0 = dq * ((-ge_sin * torch.sin(sarg) * freqs * freqs).sum(-1) + (-ge_cos * torch.sin(sarg + torch.pi / 2.0) * freqs * freqs).sum(-1))
    dW = [torch.zeros_like(w) for w in Wl]
    db = [torch.zeros_like(b) for b in bl]
    t = [None] * L
    gb = gb_e
    for l in range(L):
        if l in cfg.sdf_skip_in:
            gb = torch.cat([gb, gb_e], dim=1) * inv_sqrt2          # adjoint of g_4 = [g'_4 ; ge_skip] (each was * 1/sqrt2)
        dW[l] = dW[l] + u[l].t() @ gb
        ub = gb @ Wl[l].t()
        t[l] = s2[l] * g_next[l] * ub
        gb = s1[l] * ub
    d_ws = gb.sum(0, keepdim=True) / scale                        # g_8 = w_s / scale
    # ---- phase B ----
    d_ws = d_ws + (d_sdf * a[L]).sum(0, keepdim=True) / scale
    d_bs = d_sdf.sum(0) / scale
    dW_f = d_feat.t() @ a[L]
    db_f = d_feat.sum(0)
    ab = d_sdf * (w_s / scale) + d_feat @ W_f
    eb = torch.zeros_like(e)
    for l in reversed(range(L)):
        zb = s1[l] * ab + t[l]
        dW[l] = dW[l] + zb.t() @ a[l]
        db[l] = zb.sum(0)
        ab = zb @ Wl[l]
        if l in cfg.sdf_skip_in:
            eb = eb + ab[:, -n_e:] * inv_sqrt2
            ab = ab[:, :-n_e] * inv_sqrt2
    eb = eb + ab
    d_x0 = d_x0 + fourier_encode_jvp_T(x0, Fq, eb)
    out = {"d_pts": d_x0 * scale, "sdf_network.out_sdf.W": d_ws, "sdf_network.out_sdf.b": d_bs,
           "sdf_network.out_feat.W": dW_f, "sdf_network.out_feat.b": db_f}
    for l in range(L):
        out[f"sdf_network.lin{l}.W"] = dW[l]
        out[f"sdf_network.lin{l}.b"] = db[l]
    return out
