"""Evaluation route for the non-default `n_shadow_importance_clip > 0` renderer option (models/neus_hint_model.py:554-576).

With the option set the reference does not cast ONE shadow ray per camera ray (to the alpha-blended hit point) but one per GROUP of
`S / clip` consecutive samples, to the group's first sample position, and feeds every sample of the group that visibility.  None of
the reference's presets or scripts enables it, so the fused CUDA pipeline keeps the per-ray hint; this module composes the option out
of the operators the library already has -- the fused forward for everything that does not depend on the visibility (sample
positions, NeuS weights, normals, depth, specular cue), `nrh_sdf_query` for every network evaluation of the shadow marches and for the
features, and the stand-alone reflectance network -- with the march / compositing arithmetic in plain torch on the same device.
Inference (`is_training=False`, no gradients) only; SURVEY.md section 8a allows a PyTorch route for this knob.

Everything here is written against two callables so that it can be checked on the CPU against the unmodified reference
(tests/test_hint_fallback.py):
    sdf_fn(pts [N,3], want_grad=False, want_feat=False) -> (sdf [N], grad [N,3] | None, feat [N,256] | None)
    color_fn(points, normals, view_dirs, features, point_lights, visibilities, specular_cue) -> [N,3]
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch


def importance_samples(o: torch.Tensor, d: torch.Tensor, z: torch.Tensor, sdf: torch.Tensor, n_new: int, inv_s: float) -> torch.Tensor:
    """One up-sampling step with a fixed sharpness (up_sample :269-315 + sample_pdf(det=True) :21-65): NeuS weights of the k - 1
    intervals from the section-end SDF estimates, then the inverse CDF at linspace(0, 1, n_new)."""
    radius = torch.linalg.norm(o[:, None, :] + d[:, None, :] * z[..., None], dim=-1)
    near_sphere = (radius[:, :-1] < 1.0) | (radius[:, 1:] < 1.0)
    dz = z[:, 1:] - z[:, :-1]
    slope = (sdf[:, 1:] - sdf[:, :-1]) / (dz + 1e-5)
    slope_before = torch.cat([torch.zeros_like(slope[:, :1]), slope[:, :-1]], dim=-1)
    slope = torch.minimum(slope_before, slope).clip(-1e3, 0.0) * near_sphere
    centre = (sdf[:, 1:] + sdf[:, :-1]) * 0.5
    cdf_in = torch.sigmoid((centre - slope * dz * 0.5) * inv_s)
    cdf_out = torch.sigmoid((centre + slope * dz * 0.5) * inv_s)
    alpha = (cdf_in - cdf_out + 1e-5) / (cdf_in + 1e-5)
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-7], dim=-1), dim=-1)[:, :-1]
    w = alpha * trans + 1e-5
    pdf = w / w.sum(dim=-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, dim=-1)], dim=-1)                # k entries, like z
    u = torch.linspace(0.0, 1.0, n_new, device=z.device, dtype=z.dtype).expand(z.shape[0], n_new).contiguous()
    hi = torch.searchsorted(cdf, u, right=True)
    lo = (hi - 1).clamp_min(0)
    hi = hi.clamp_max(cdf.shape[-1] - 1)
    c_lo, c_hi = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
    z_lo, z_hi = torch.gather(z, 1, lo), torch.gather(z, 1, hi)
    span = c_hi - c_lo
    span = torch.where(span < 1e-5, torch.ones_like(span), span)
    return z_lo + (u - c_lo) / span * (z_hi - z_lo)


def neus_alpha(sdf: torch.Tensor, grad: torch.Tensor, dirs: torch.Tensor, dists: torch.Tensor, inv_s: torch.Tensor,
               cos_anneal: float) -> torch.Tensor:
    """get_alpha :333-357 for [N] sdf / [N,3] gradients / [N,3] directions / [N] section lengths."""
    cosine = (dirs * grad).sum(-1)
    slope = -(torch.relu(-cosine * 0.5 + 0.5) * (1.0 - cos_anneal) + torch.relu(-cosine) * cos_anneal)
    cdf_in = torch.sigmoid((sdf - slope * dists * 0.5) * inv_s)
    cdf_out = torch.sigmoid((sdf + slope * dists * 0.5) * inv_s)
    return ((cdf_in - cdf_out + 1e-5) / (cdf_in + 1e-5)).clip(0.0, 1.0)


def shadow_visibility(sdf_fn: Callable, pls: torch.Tensor, targets: torch.Tensor, n_samples: int, n_importance: int, inv_s: torch.Tensor,
                      offset: float, cos_anneal: float = 1.0, up_sample_steps: int = 4) -> torch.Tensor:
    """get_visibility :373-432 without jitter: march from the light towards `targets`, transmittance in front of the last sample.
    pls / targets [N,3] -> [N,1]."""
    to_target = targets - pls
    length = torch.linalg.norm(to_target, dim=-1, keepdim=True)
    d = to_target / length
    z = torch.linspace(0.0, 1.0, n_samples, device=pls.device, dtype=pls.dtype) * length * (1.0 - offset)
    N = pls.shape[0]
    if n_importance > 0:
        sdf = sdf_fn((pls[:, None, :] + d[:, None, :] * z[..., None]).reshape(-1, 3))[0].reshape(N, -1)
        for i in range(up_sample_steps):
            z_new = importance_samples(pls, d, z, sdf, n_importance // up_sample_steps, 64 * 2 ** i)
            z_all, order = torch.sort(torch.cat([z, z_new], dim=-1), dim=-1)
            if i + 1 < up_sample_steps:                                                        # cat_z_vals :317-331
                sdf_new = sdf_fn((pls[:, None, :] + d[:, None, :] * z_new[..., None]).reshape(-1, 3))[0].reshape(N, -1)
                sdf = torch.gather(torch.cat([sdf, sdf_new], dim=-1), 1, order)
            z = z_all
    S = z.shape[1]
    dists = torch.cat([z[:, 1:] - z[:, :-1], (length / n_samples).expand(N, 1)], dim=-1)
    mid = z + dists * 0.5
    pts = (pls[:, None, :] + d[:, None, :] * mid[..., None]).reshape(-1, 3)
    sdf_m, grad_m, _ = sdf_fn(pts, want_grad=True)
    alpha = neus_alpha(sdf_m.reshape(-1), grad_m.reshape(-1, 3), d[:, None, :].expand(N, S, 3).reshape(-1, 3), dists.reshape(-1),
                       inv_s, cos_anneal).reshape(N, S)
    taus = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-7], dim=-1), dim=-1)[:, :-1]
    return taus[:, -1:]


def grouped_visibility(sdf_fn: Callable, o: torch.Tensor, d: torch.Tensor, pl: torch.Tensor, z_vals: torch.Tensor, weights: torch.Tensor,
                       clip: int, n_shadow_samples: int, n_shadow_importance: int, inv_s: torch.Tensor, offset: float,
                       cos_anneal: float = 1.0, chunk: int = 2048) -> Tuple[torch.Tensor, torch.Tensor]:
    """render_core :554-576: one shadow march per group of S / clip samples (towards the group's first sample), evaluated in
    chunks of `chunk` shadow rays (shadow_mini_chunk_size).  Returns (per-sample visibility [R,S,1], shadow map [R,1] = the
    visibility of the sample with the largest weight)."""
    R, S = z_vals.shape
    if clip <= 0 or S % clip != 0:
        raise ValueError(f"n_shadow_importance_clip = {clip} must divide the {S} samples of a ray")
    per_group = S // clip
    first = z_vals[:, torch.arange(0, S, per_group, device=z_vals.device)]                          # [R, clip]
    targets = (o[:, None, :] + d[:, None, :] * first[..., None]).reshape(-1, 3)
    lights = pl[:, None, :].expand(R, clip, 3).reshape(-1, 3)
    parts = [shadow_visibility(sdf_fn, lights[i:i + chunk], targets[i:i + chunk], n_shadow_samples, n_shadow_importance, inv_s, offset,
                               cos_anneal) for i in range(0, R * clip, chunk)]
    vis = torch.cat(parts, dim=0).reshape(R, clip, 1).repeat_interleave(per_group, dim=1)            # [R, S, 1]
    shadow_map = torch.gather(vis[..., 0], 1, torch.argmax(weights[:, :S], dim=1, keepdim=True))
    return vis, shadow_map


def shade_samples(sdf_fn: Callable, color_fn: Callable, o: torch.Tensor, d: torch.Tensor, pl: torch.Tensor, z_vals: torch.Tensor,
                  weights: torch.Tensor, normals: torch.Tensor, vis: Optional[torch.Tensor], specular_cue: Optional[torch.Tensor],
                  last_dist: float, background_rgb: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """render_core :489-503, :625-637 for given sample positions / weights / normals / hints: features at the section mid-points,
    reflectance network, alpha compositing over the (optional) fixed background.  Returns (rgb [R,3], sampled colour [R,S,3])."""
    R, S = z_vals.shape
    dists = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], torch.full_like(z_vals[:, :1], last_dist)], dim=-1)
    mid = z_vals + dists * 0.5
    pts = (o[:, None, :] + d[:, None, :] * mid[..., None]).reshape(-1, 3)
    feat = sdf_fn(pts, want_feat=True)[2]
    colour = color_fn(pts, normals.reshape(-1, 3), d[:, None, :].expand(R, S, 3).reshape(-1, 3), feat,
                      pl[:, None, :].expand(R, S, 3).reshape(-1, 3), vis.reshape(-1, 1) if vis is not None else None,
                      specular_cue.reshape(R * S, -1) if specular_cue is not None else None).reshape(R, S, 3)
    w = weights[:, :S]
    rgb = (colour * w[..., None]).sum(dim=1)
    if background_rgb is not None:
        rgb = rgb + background_rgb * (1.0 - w.sum(dim=-1, keepdim=True))
    return rgb, colour
