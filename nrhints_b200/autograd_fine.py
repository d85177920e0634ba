"""Differentiable fine pass as COMPOSED autograd nodes -- the second implementation of the training step.

The default route for a call that needs gradients is ONE autograd node on two library calls (fused_step.py:
nrh_render_train_forward / nrh_render_backward).  This module is the composed route the renderer takes when that node does not
apply (fp32 engine, `n_importance == 0`, layers without weight norm, `renderer.fused_training = False`) and the one the parity
tests hold against the fused node: the call is split exactly where the reference puts its `torch.no_grad()` fences:

  * everything the reference computes WITHOUT gradients -- the hierarchical sampler on both rays (7 + 6 SDF
    passes), the shadow visibility, depth / hit point and the specular cue
    (/root/reference/models/neus_hint_model.py:697-713, :379, :531-533, :589) -- runs in the CUDA kernels;
  * the differentiable remainder -- SDF + feature at the 128 section mid-points, d sdf/dx with create_graph
    (fields/sdf_field.py:136-148), NeuS alpha, weights, reflectance MLP, compositing (:504-525, :583-637) --
    is a handful of autograd nodes with hand-written CUDA forwards AND backwards on the tcgen05 engine (sdf_autograd._SdfFine,
    _ReflectanceF16 on nrh_color_train_forward / _backward + nrh_wgrad_f16, _CompositeTrain) glued by torch ops, or plain torch
    expressions of the same functions on the fp32 engine, so autograd produces the reference's gradients (parameters, ray origins /
    directions / light positions; near / far only when n_importance == 0, because the reference's up-sampling block re-assigns
    z_vals under no_grad, :696-713).

This module only uses torch and the CUDA library; it never touches the oracle and never runs on the CPU in the product path
(the renderer rejects CPU tensors).  The CPU test-suite exercises its chain rule against the oracle with library calls replaced
by fp32 stand-ins.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def _fourier(x: Tensor, n_freq: int) -> Tensor:
    freqs = 2 ** torch.linspace(0.0, n_freq - 1, n_freq, device=x.device, dtype=x.dtype)
    s = (x[..., None] * freqs).reshape(*x.shape[:-1], -1)
    return torch.cat([x, torch.sin(torch.cat([s, s + torch.pi / 2.0], dim=-1))], dim=-1)


def sdf_forward(sdf_w: List[Tensor], sdf_b: List[Tensor], head: Dict[str, Tensor], pts: Tensor, scale: float = 3.0,
                skip: int = 4, n_freq: int = 6) -> Tensor:
    """[N,3] -> [N,257] (sdf | feature), fields/sdf_field.py:106-123."""
    e = _fourier(pts * scale, n_freq)
    h = e
    for l, (w, b) in enumerate(zip(sdf_w, sdf_b)):
        if l == skip:
            h = torch.cat([h, e], dim=1) / math.sqrt(2.0)
        h = F.softplus(F.linear(h, w, b), beta=100.0)
    sdf = F.linear(h, head["sdf_w"], head["sdf_b"]) / scale
    feat = F.linear(h, head["feat_w"], head["feat_b"])
    return torch.cat([sdf, feat], dim=-1)


def attach_coarse_gradient(z_final: Tensor, z_coarse: Tensor) -> Tensor:
    """Give the coarse entries of the kernel's (detached) z their gradient path to near / far.  Only used when
    n_importance == 0: with importance sampling the reference's cat_z_vals runs under no_grad (models/neus_hint_model.py:696-713)
    and the final z_vals are detached altogether."""
    if not z_coarse.requires_grad:
        return z_final
    zc = z_coarse.detach()
    idx = torch.searchsorted(z_final.contiguous(), zc.contiguous()).clamp(max=z_final.shape[1] - 1)
    lo = (idx - 1).clamp(min=0)
    pick_lo = (z_final.gather(1, lo) - zc).abs() < (z_final.gather(1, idx) - zc).abs()
    idx = torch.where(pick_lo, lo, idx)
    return z_final + torch.zeros_like(z_final).scatter_add(1, idx, z_coarse - zc)


class _LinearF16(torch.autograd.Function):
    """y = x W^T + b with fp16 operands, fp32 accumulation and fp32 results (library GEMMs) in the forward AND in both
    backward products -- the operand precision of the CUDA reflectance kernel (single fp16 pass; SURVEY.md section 7:
    the reflectance MLP enters the pixel through a sigmoid and is insensitive to 11-bit operands).  Output adjoints are
    scaled by a power of two before the cast so that small gradients do not flush to zero in fp16."""

    @staticmethod
    def forward(ctx, x, w, b):
        xh, wh = x.to(torch.float16), w.to(torch.float16)
        ctx.save_for_backward(xh, wh)
        return torch.mm(xh, wh.t(), out_dtype=torch.float32) + b

    @staticmethod
    def backward(ctx, dy):
        xh, wh = ctx.saved_tensors
        amax = dy.abs().max().clamp_min(1e-30)
        s = torch.exp2(torch.floor(torch.log2(1024.0 / amax))).clamp(2.0 ** -40, 2.0 ** 40)
        dyh = (dy * s).to(torch.float16)
        inv = 1.0 / s
        dx = torch.mm(dyh, wh, out_dtype=torch.float32) * inv
        dw = torch.mm(dyh.t(), xh, out_dtype=torch.float32) * inv
        return dx, dw, dy.sum(0)


def _split_input_gradient(ctx, dx16: Tensor, inv: Tensor):
    """dx16 [P, Kp] fp16 (loss-scaled adjoint of the concatenated input) -> the gradients of the column blocks that need one
    (per-ray blocks: summed over the samples of the ray)."""
    G0, G1, ray_axis = ctx.ray_dims
    dparts, off = [], 0
    for i, k in enumerate(ctx.widths):
        if not ctx.needs_input_grad[3 + i]:
            dparts.append(None)
        elif ctx.per_ray[i]:
            dparts.append(dx16.view(G0, G1, -1)[:, :, off:off + k].sum(1 - ray_axis, dtype=torch.float32) * inv)
        else:
            dparts.append(dx16[:, off:off + k].float() * inv)
        off += k
    return tuple(dparts)


class _ReflectanceF16(torch.autograd.Function):
    """The reflectance MLP (fields/reflectance_network.py:84-96: 4 x (Linear 256, ReLU), Linear 3; the sigmoid stays outside)
    as ONE autograd node on fp16-operand / fp32-accumulate library GEMMs -- the operand precision of the CUDA reflectance
    kernel.  Forward: the input is padded to a multiple of 64 columns (361 -> 384: the unaligned K sends cuBLAS to a 6x slower
    kernel), bias + ReLU run in the GEMM epilogue (torch._addmm_activation) and the hidden activations are kept in fp16.
    Backward: ONE power-of-two loss scale for the whole chain (chosen on the device from max|dy|), adjoints stay fp16 between
    layers (they are GEMM operands anyway), ReLU masks come from the saved activations, weight gradients are fp32-output
    GEMMs and bias gradients one column-sum launch each (nrh_colsum_f16)."""

    @staticmethod
    def forward(ctx, renderer, n_parts, ray_dims, *args):
        """renderer: the owning NeuSHintRenderer -> forward AND backward run on the hand-written tcgen05 kernels
        (nrh_color_train_forward / _backward, csrc/color_train_tc.inc; weights come from its packed operand images) and the weight
        gradients on nrh_wgrad_f16; None -> the same arithmetic on library GEMMs (kept for the host-side chain-rule test).
        args = the n_parts column blocks of the input (concatenated in order) followed by 5 weights and 5 biases.  A block is
        either per point ([P, k_i] fp32) or per RAY ([R, k_i]: view / light / hint encodings are the same for all samples of a
        ray); ray_dims = (G0, G1, ray_axis) says how the P points factor into [G0, G1] and which axis is the ray.  The blocks are
        written straight into the padded fp16 operand (no fp32 concatenation pass, per-ray blocks by a broadcast copy) and
        receive their gradients block by block, only where needed (per-ray blocks: summed over the samples)."""
        from .train_ops import colsum_f16  # noqa: F401  (fails loudly here if the CUDA library is missing)
        parts, wb = args[:n_parts], args[n_parts:]
        n = len(wb) // 2
        ws, bs = wb[:n], wb[n:]
        G0, G1, ray_axis = ray_dims
        P, n_rays = G0 * G1, (G0, G1)[ray_axis]
        widths = [int(t.shape[1]) for t in parts]
        per_ray = [t.shape[0] != P for t in parts]
        assert all(t.shape[0] == (n_rays if pr else P) for t, pr in zip(parts, per_ray))
        K = sum(widths)
        Kp = 384 if renderer is not None else (K + 63) // 64 * 64
        h16 = torch.empty(P, Kp, dtype=torch.float16, device=parts[0].device)
        h3 = h16.view(G0, G1, Kp)
        off = 0
        for t, k, pr in zip(parts, widths, per_ray):
            if pr:
                h3[:, :, off:off + k] = t.unsqueeze(1 - ray_axis)            # broadcast over the sample axis
            else:
                h16[:, off:off + k] = t
            off += k
        if Kp > K:
            h16[:, K:] = 0
        ctx.renderer = renderer
        ctx.n, ctx.K, ctx.widths, ctx.per_ray, ctx.ray_dims = n, K, widths, per_ray, ray_dims
        if renderer is not None:
            from .train_ops import color_train_forward
            acts4, y4 = color_train_forward(renderer, h16)
            ctx.save_for_backward(h16, acts4)
            return y4[:, :ws[-1].shape[0]]
        acts, w16s = [h16], []
        for l in range(n - 1):
            w16 = ws[l].detach().to(torch.float16)
            if l == 0:
                w16 = torch.nn.functional.pad(w16, (0, Kp - K))
            w16s.append(w16)
            h16 = torch._addmm_activation(bs[l].detach().to(torch.float16), h16, w16.t())
            acts.append(h16)
        wl = torch.zeros(8, ws[-1].shape[1], dtype=torch.float16, device=h16.device)       # 3 output rows padded to 8
        wl[:ws[-1].shape[0]] = ws[-1].detach()
        w16s.append(wl)
        y = torch.mm(h16, wl.t(), out_dtype=torch.float32)[:, :ws[-1].shape[0]] + bs[-1].detach()
        ctx.save_for_backward(*acts, *w16s)
        return y

    @staticmethod
    def backward(ctx, dy):
        from .train_ops import colsum_f16
        n, K = ctx.n, ctx.K
        P = dy.shape[0]
        amax = dy.abs().max().clamp_min(1e-30)
        s = torch.exp2(torch.floor(torch.log2(256.0 / amax))).clamp(2.0 ** -40, 2.0 ** 40)
        inv = 1.0 / s
        if ctx.renderer is not None:
            # hand-written path: one tcgen05 kernel for the adjoint chain, one for all weight gradients, one column-sum launch
            from .train_ops import WgradBatch, color_train_backward
            h16, acts4 = ctx.saved_tensors
            s1, inv1 = s.reshape(1).contiguous(), inv.reshape(1).contiguous()
            dz4, dy16, dx16 = color_train_backward(ctx.renderer, dy, s1, acts4)
            z32 = lambda r, c: torch.zeros(r, c, dtype=torch.float32, device=dy.device)      # noqa: E731
            dws = [z32(256, 384)] + [z32(256, 256) for _ in range(3)] + [z32(dy.shape[1], 256)]
            wb = WgradBatch()
            wb.add(dz4[0], h16, dws[0], dev_scale=inv1, n=256).add(dz4[0], h16, dws[0][:, 256:], dev_scale=inv1, b_col0=256, n=128)
            for l in range(1, 4):
                wb.add(dz4[l], acts4[l - 1], dws[l], scale=1.0 / 16.0, dev_scale=inv1)           # the saved activations are x16
            wb.add(dy16, acts4[3], dws[4], scale=1.0 / 16.0, dev_scale=inv1, m=dy.shape[1])
            wb.run()
            dws[0] = dws[0][:, :K]
            dbs = list(colsum_f16(dz4) * inv) + [dy.sum(0)]
            return (None, None, None) + _split_input_gradient(ctx, dx16, inv) + tuple(dws) + tuple(dbs)
        acts, w16s = ctx.saved_tensors[:n], ctx.saved_tensors[n:]
        from .train_ops import WgradBatch
        dz = torch.zeros(P, 8, dtype=torch.float16, device=dy.device)
        dz[:, :dy.shape[1]] = dy * s
        dws, dbs = [None] * n, [None] * n
        inv1 = inv.reshape(1).contiguous()                                   # device scalar multiplied in by the reduction kernel
        wb = WgradBatch()                                                    # all five dW = dz^T a reductions: ONE tcgen05 launch at the end
        dws[n - 1] = torch.zeros(dy.shape[1], acts[n - 1].shape[1], dtype=torch.float32, device=dy.device)
        wb.add(dz, acts[n - 1], dws[n - 1], dev_scale=inv1, m=dy.shape[1])
        dbs[n - 1] = dy.sum(0)
        for l in range(n - 2, -1, -1):
            dh = torch.mm(dz, w16s[l + 1])                                   # fp16 out, fp32 accumulate: the next operand
            dz = torch.ops.aten.threshold_backward(dh, acts[l + 1], 0.0)     # ReLU mask from the saved activation
            kw = acts[l].shape[1]
            dw = torch.zeros(dz.shape[1], kw, dtype=torch.float32, device=dy.device)
            for c0 in range(0, kw, 256):                                     # the 384-wide input layer as a 256- and a 128-column job
                wb.add(dz, acts[l], dw[:, c0:], dev_scale=inv1, b_col0=c0, n=min(256, kw - c0))
            dws[l] = dw[:, :K] if l == 0 else dw
            dbs[l] = colsum_f16(dz) * inv
        wb.run()
        dx16 = torch.mm(dz, w16s[0])                                         # [P, Kp] fp16, loss-scaled
        return (None, None, None) + _split_input_gradient(ctx, dx16, inv) + tuple(dws) + tuple(dbs)


class _CompositeTrain(torch.autograd.Function):
    """NeuS alpha, transmittance scan / weights and colour compositing of the primary ray (models/neus_hint_model.py:339-356,
    :521-526,:635-637) as ONE autograd node on the CUDA operators nrh_composite_train_forward / _backward (one thread per ray,
    hand-derived backward, the scan is recomputed instead of taped; csrc/composite_train_math.cuh) -- it replaces ~60 small torch
    launches per step.  Inputs per point in evaluation order: sdf [N,1], grad [N,3], color [N,3]; dists [R,S] (no gradient:
    the sample positions are detached), dirs [R,3], inv_s (1 element), bg [1,3] or None.  Returns (rgb [R,3], weights [R,S])."""

    @staticmethod
    def forward(ctx, sdf, grad, color, dirs, inv_s, dists, bg, cos_anneal: float, sample_major: bool):
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        dev = sdf.device
        R, S = dists.shape
        f32 = dict(dtype=torch.float32, device=dev)
        c = lambda t: t.detach().to(torch.float32).contiguous()          # noqa: E731
        sdf_c, g_c, col_c, dirs_c, dist_c = c(sdf).reshape(-1), c(grad), c(color), c(dirs), c(dists)
        s_c = c(inv_s).reshape(1)
        bg_c = c(bg).reshape(-1) if bg is not None else None
        pr, pj = (1, R) if sample_major else (S, 1)
        w, rgb = torch.empty(R, S, **f32), torch.empty(R, 3, **f32)
        with torch.cuda.device(dev):
            _lib.check(lib.nrh_composite_train_forward(sdf_c.data_ptr(), g_c.data_ptr(), col_c.data_ptr(), pr, pj, dist_c.data_ptr(),
                                                       dirs_c.data_ptr(), s_c.data_ptr(), float(cos_anneal),
                                                       bg_c.data_ptr() if bg_c is not None else None, R, S, w.data_ptr(), rgb.data_ptr(),
                                                       torch.cuda.current_stream(dev).cuda_stream), "nrh_composite_train_forward")
        ctx.save_for_backward(sdf_c, g_c, col_c, dirs_c, dist_c, s_c, bg_c)
        ctx.meta = (R, S, pr, pj, float(cos_anneal), tuple(sdf.shape), tuple(inv_s.shape))
        return rgb, w

    @staticmethod
    def backward(ctx, d_rgb, d_w):
        from . import _lib
        lib = _lib.load()
        sdf_c, g_c, col_c, dirs_c, dist_c, s_c, bg_c = ctx.saved_tensors
        R, S, pr, pj, cos_anneal, sdf_shape, s_shape = ctx.meta
        dev = sdf_c.device
        f32 = dict(dtype=torch.float32, device=dev)
        d_rgb = (d_rgb if d_rgb is not None else torch.zeros(R, 3, **f32)).to(torch.float32).contiguous()
        d_w = d_w.to(torch.float32).contiguous() if d_w is not None else None
        d_sdf, d_g, d_c = torch.empty_like(sdf_c), torch.empty_like(g_c), torch.empty_like(col_c)
        d_dirs, d_s = torch.empty(R, 3, **f32), torch.zeros(1, **f32)
        with torch.cuda.device(dev):
            _lib.check(lib.nrh_composite_train_backward(sdf_c.data_ptr(), g_c.data_ptr(), col_c.data_ptr(), pr, pj, dist_c.data_ptr(),
                                                        dirs_c.data_ptr(), s_c.data_ptr(), cos_anneal,
                                                        bg_c.data_ptr() if bg_c is not None else None, R, S, d_rgb.data_ptr(),
                                                        d_w.data_ptr() if d_w is not None else None, d_sdf.data_ptr(), d_g.data_ptr(),
                                                        d_c.data_ptr(), d_dirs.data_ptr(), d_s.data_ptr(),
                                                        torch.cuda.current_stream(dev).cuda_stream), "nrh_composite_train_backward")
        return d_sdf.reshape(sdf_shape), d_g, d_c, d_dirs, d_s.reshape(s_shape), None, None, None, None


def _linear_f16_ok(x: Tensor) -> bool:
    if not x.is_cuda:
        return False
    if not hasattr(_linear_f16_ok, "ok"):
        try:
            a = torch.zeros(8, 8, dtype=torch.float16, device=x.device)
            torch.mm(a, a, out_dtype=torch.float32)
            _linear_f16_ok.ok = True
        except (TypeError, RuntimeError):
            _linear_f16_ok.ok = False
    return _linear_f16_ok.ok


def render_fine(weights: Dict[str, object], rays_o: Tensor, rays_d: Tensor, rays_pl: Tensor, z_vals: Tensor,
                sample_dist: float, visibilities: Optional[Tensor], specular_cue: Optional[Tensor],
                background_rgb: Optional[Tensor], cos_anneal: float, inv_s: Tensor, normalized_normals: bool,
                refl_freq: int = 4, sdf_fn=None, sample_major: bool = False, renderer=None,
                outside: Optional[Dict[str, Tensor]] = None) -> Dict[str, Tensor]:
    """render_core's differentiable part (models/neus_hint_model.py:475-651) given the sample positions and hints.

    weights: dict(sdf_w, sdf_b: lists of 8; sdf_w_head, sdf_b_head, feat_w, feat_b; col_w, col_b: lists of 5).
    visibilities [R,1] / specular_cue [R,n_rough]: per-ray hint values (no gradient, as in the reference).
    sdf_fn: optional callable pts [N,3] -> (sdf [N,1], feat [N,256], grad [N,3]) replacing the torch evaluation of the SDF
    network and its create_graph input gradient by ONE autograd node with a hand-written CUDA forward and backward
    (nrhints_b200/sdf_autograd.py, the tcgen05 engine).
    outside: the background model's terms on the merged sample set (alpha [R,S+n_out], color [R,S+n_out,3]) and the inside-sphere
    mask [R,S]: alpha / colour are blended outside the unit sphere and the appended outside samples join the compositing
    (models/neus_hint_model.py:517-524,:628-633); the returned weights / colours then have S + n_out entries per ray.
    sample_major: evaluate the points in the render pipeline's order (p = j * R + r instead of r * S + j), so that a tape captured
    by nrh_render_forward (NrhTrainCapture) lines up with them; results are returned ray-major either way."""
    R, S = z_vals.shape
    dev, dt = z_vals.device, z_vals.dtype
    dists = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], torch.full((R, 1), sample_dist, device=dev, dtype=dt)], -1)
    mid_z = z_vals + dists * 0.5
    def per_point(t):                  # [R,S,...] -> one row per point in evaluation order
        return (t.transpose(0, 1) if sample_major else t).reshape(R * S, *t.shape[2:])

    def per_ray(t):                    # rows in evaluation order -> [R,S,...]
        return t.reshape(S, R, *t.shape[1:]).transpose(0, 1) if sample_major else t.reshape(R, S, *t.shape[1:])
    pts = per_point(rays_o[:, None, :] + rays_d[:, None, :] * mid_z[..., None])
    dirs = per_point(rays_d[:, None, :].expand(R, S, 3))
    pls = per_point(rays_pl[:, None, :].expand(R, S, 3))
    head = {"sdf_w": weights["sdf_w_head"], "sdf_b": weights["sdf_b_head"], "feat_w": weights["feat_w"], "feat_b": weights["feat_b"]}

    if sdf_fn is not None:
        sdf, feat, grad = sdf_fn(pts)
    else:
        out = sdf_forward(weights["sdf_w"], weights["sdf_b"], head, pts)
        feat = out[:, 1:]
        # get_alpha re-evaluates the SDF and differentiates it w.r.t. the points (:335-336)
        with torch.enable_grad():
            x = pts if pts.requires_grad else pts.detach().requires_grad_(True)
            sdf = sdf_forward(weights["sdf_w"], weights["sdf_b"], head, x)[:, :1]
            grad = torch.autograd.grad(sdf, x, torch.ones_like(sdf), create_graph=True, retain_graph=True)[0]
    # compositing: ONE CUDA autograd node on the tcgen05 path (the sample positions carry no gradient there: importance sampling
    # detaches them, so `dists` is data), the torch expression otherwise (fp32 engine, CPU tests, n_importance == 0)
    fused_composite = (sdf_fn is not None and pts.is_cuda and not dists.requires_grad and S <= 128 and outside is None
                       and os.environ.get("NRH_COMPOSITE", "cuda") != "torch")
    w = wsum = None
    if not fused_composite:
        true_cos = (dirs * grad).sum(-1, keepdim=True)
        iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal) + F.relu(-true_cos) * cos_anneal)
        half = iter_cos * per_point(dists[..., None]) * 0.5
        prev_cdf = torch.sigmoid((sdf - half) * inv_s)
        next_cdf = torch.sigmoid((sdf + half) * inv_s)
        alpha = per_ray(((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).clip(0, 1))[..., 0]
        if outside is not None:
            ins = outside["inside"]
            alpha = alpha * ins + outside["alpha"][:, :S] * (1.0 - ins)
            alpha = torch.cat([alpha, outside["alpha"][:, S:]], dim=-1)
        trans = torch.cumprod(torch.cat([torch.ones((R, 1), device=dev, dtype=dt), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
        w = alpha * trans
        wsum = w.sum(-1, keepdim=True)

    n_hat = F.normalize(grad, dim=-1, p=2)
    n_col = len(weights["col_w"])
    lowp = sdf_fn is not None and _linear_f16_ok(pts)        # tcgen05 engine: same operand precision as its reflectance kernel
    if lowp:
        # view / light / hint encodings are per-RAY quantities: encoded once per ray and broadcast into the operand by the node
        parts = [pts, _fourier(rays_d, refl_freq), n_hat if normalized_normals else grad, _fourier(rays_pl, refl_freq), feat]
        if visibilities is not None:
            parts.append(_fourier(visibilities, refl_freq))
        if specular_cue is not None:
            parts.append(_fourier(specular_cue, refl_freq))
        ray_dims = (S, R, 1) if sample_major else (R, S, 0)
        hcol = _ReflectanceF16.apply(renderer, len(parts), ray_dims, *parts, *weights["col_w"], *weights["col_b"])
    else:
        parts = [pts, _fourier(dirs, refl_freq), n_hat if normalized_normals else grad, _fourier(pls, refl_freq), feat]
        if visibilities is not None:
            parts.append(_fourier(per_point(visibilities[:, None, :].expand(R, S, 1)), refl_freq))
        if specular_cue is not None:
            nr = specular_cue.shape[-1]
            parts.append(_fourier(per_point(specular_cue[:, None, :].expand(R, S, nr)), refl_freq))
        hcol = torch.cat(parts, dim=-1)
        for l, (cw, cb) in enumerate(zip(weights["col_w"], weights["col_b"])):
            hcol = F.linear(hcol, cw, cb)
            if l < n_col - 1:
                hcol = torch.relu(hcol)
    color_pt = torch.sigmoid(hcol)
    color = per_ray(color_pt)
    if outside is not None:
        ins = outside["inside"][..., None]
        color = torch.cat([color * ins + outside["color"][:, :S] * (1.0 - ins), outside["color"][:, S:]], dim=1)
    if fused_composite:
        rgb, w = _CompositeTrain.apply(sdf, grad, color_pt, rays_d, inv_s, dists, background_rgb, float(cos_anneal), bool(sample_major))
    else:
        rgb = (color * w[..., None]).sum(1)
        if background_rgb is not None:
            rgb = rgb + background_rgb * (1.0 - wsum)
    return {"rgb": rgb, "weights": w, "analytic_normals": per_ray(grad),
            "normalized_analytic_normals": per_ray(n_hat), "sampled_color": color}
