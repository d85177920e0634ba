"""The training step of the ray-march path as TWO library calls (include/nrhints_b200.h: nrh_render_train_forward /
nrh_render_backward), and the two ways a caller reaches them:

  * `FusedRender` -- ONE autograd node for NeuSHintRenderer.forward when gradients are requested.  The reference builds an autograd
    graph of several thousand nodes per step (/root/reference/models/neus_hint_model.py:653-751 under autograd, then loss.backward(),
    /root/reference/trainer/trainer.py:269-283); here `loss.backward()` of an unmodified trainer lands in a single backward call that
    writes the gradients of all 46 parameter tensors (and of the ray origins / directions / light positions for camera optimisation).
  * `FusedTrainStep` -- the same two calls without autograd at all, for the native pipeline (pipeline.py::NRHintPipeline.train_step):
    ray generation -> forward -> loss -> backward -> ray-generation backward, gradients written STRAIGHT into the optimizer's flat
    gradient buffer (train_ops.FlatAdam: the all-reduce operand), no torch kernel in between.

CUDA + tcgen05 engine only; there is no CPU fallback (the library call fails loudly without the extension).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _lib

_WN_SCRATCH_BYTES = 4 << 20


def wn_layers(renderer) -> list:
    """The 15 weight-normed layers in the order of NrhTrainParams: sdf lin0..7, out_sdf, out_feat, colour lin0..4."""
    sn, cn = renderer.sdf_network, renderer.color_network
    return [getattr(sn, f"lin{l}") for l in range(8)] + [sn.out_sdf, sn.out_feat] + [getattr(cn, f"lin{l}") for l in range(5)]


def param_list(renderer) -> List[torch.nn.Parameter]:
    """(v, g, bias) of the 15 layers, then the variance: the 46 parameter tensors in the order the fused node takes / returns them."""
    ps = []
    for lin in wn_layers(renderer):
        ps += [lin.weight_v, lin.weight_g, lin.bias]
    ps.append(renderer.deviation_network.variance)
    return ps


def can_pack_wn(renderer) -> bool:
    return not renderer.has_outside_nerf and all(getattr(l, "weight_normed", False) for l in wn_layers(renderer))


def eligible(renderer) -> bool:
    """The fused step covers the reference-default training configuration: tcgen05 engine, weight-normed layers, importance sampling
    (the reference then detaches the sample positions, models/neus_hint_model.py:696-713), <= 128 samples, no outside NeRF."""
    r = renderer.config.renderer
    return (renderer.mlp_impl in ("auto", "tcgen05") and not renderer.has_outside_nerf and r.n_importance_samples > 0
            and r.n_samples + r.n_importance_samples <= 128 and all(getattr(l, "weight_normed", False) for l in wn_layers(renderer))
            and renderer.color_network.d_in_total <= 384
            and getattr(r.depth_type, "value", r.depth_type) != "sphere_tracing")


def train_params(renderer, grads: Optional[List[Optional[torch.Tensor]]], keep: list) -> _lib.NrhTrainParams:
    """NrhTrainParams over the renderer's parameters; grads[i] (same order as param_list) receives the gradient of parameter i."""
    P = _lib.NrhTrainParams()
    ps = param_list(renderer)
    for p in ps:
        if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
            raise RuntimeError("the fused training step needs contiguous fp32 CUDA parameters (no CPU fallback)")
    slots = [P.sdf[l] for l in range(8)] + [P.sdf_out, P.feat_out] + [P.col[l] for l in range(5)]
    for i, s in enumerate(slots):
        s.v, s.g, s.bias = ps[3 * i].data_ptr(), ps[3 * i + 1].data_ptr(), ps[3 * i + 2].data_ptr()
        if grads is not None:
            s.d_v, s.d_g, s.d_bias = (grads[3 * i + k].data_ptr() for k in range(3))
    P.variance = ps[45].data_ptr()
    if grads is not None:
        P.d_variance = grads[45].data_ptr()
    keep += ps
    return P


def pack_weights_wn(renderer, device) -> torch.Tensor:
    """renderer._packed rebuilt from (v, g, bias) by the library itself (nrh_pack_weights_wn: weight norm of all layers in one
    launch, then the operand images) -- replaces 30 torch weight_norm launches + copies per step."""
    lib = _lib.load()
    cfg = renderer._c_config()
    nbytes = lib.nrh_packed_weights_bytes(C.byref(cfg))
    if renderer._packed is None or renderer._packed.numel() < nbytes or renderer._packed.device != device:
        renderer._packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
    if getattr(renderer, "_wn_scratch", None) is None or renderer._wn_scratch.device != device:
        renderer._wn_scratch = torch.empty(_WN_SCRATCH_BYTES, dtype=torch.uint8, device=device)
    keep: list = []
    P = train_params(renderer, None, keep)
    with torch.cuda.device(device):
        _lib.check(lib.nrh_pack_weights_wn(C.byref(cfg), C.byref(P), renderer._wn_scratch.data_ptr(), renderer._wn_scratch.numel(),
                                           renderer._packed.data_ptr(), nbytes, torch.cuda.current_stream(device).cuda_stream),
                   "nrh_pack_weights_wn")
    return renderer._packed


def _rays_struct(o, d, pl, near, far) -> _lib.NrhRays:
    return _lib.NrhRays(o.data_ptr(), d.data_ptr(), pl.data_ptr(), near.data_ptr(), far.data_ptr(), None, None)


def run_forward(renderer, cfg, packed, rays: _lib.NrhRays, R: int, bg, jit_p, jit_s, cos_anneal: float, warmup: bool,
                out: Dict[str, Optional[torch.Tensor]], train_ws: torch.Tensor):
    lib = _lib.load()
    c_out = _lib.NrhOutputs(**{k: (v.data_ptr() if v is not None else None) for k, v in out.items()})
    dev = train_ws.device
    with torch.cuda.device(dev):
        _lib.check(lib.nrh_render_train_forward(C.byref(cfg), packed.data_ptr(), C.byref(rays), R, bg.data_ptr() if bg is not None else None,
                                                jit_p.data_ptr() if jit_p is not None else None,
                                                jit_s.data_ptr() if jit_s is not None else None, float(cos_anneal), int(warmup),
                                                C.byref(c_out), train_ws.data_ptr(), train_ws.numel(),
                                                torch.cuda.current_stream(dev).cuda_stream), "nrh_render_train_forward")
    renderer.last_launch_count = lib.nrh_last_launch_count()


def run_backward(renderer, cfg, packed, P: _lib.NrhTrainParams, rays: _lib.NrhRays, R: int, bg, cos_anneal: float,
                 adj: _lib.NrhTrainAdjoints, train_ws: torch.Tensor):
    lib = _lib.load()
    dev = train_ws.device
    with torch.cuda.device(dev):
        _lib.check(lib.nrh_render_backward(C.byref(cfg), packed.data_ptr(), C.byref(P), C.byref(rays), R,
                                           bg.data_ptr() if bg is not None else None, float(cos_anneal), C.byref(adj),
                                           train_ws.data_ptr(), train_ws.numel(), torch.cuda.current_stream(dev).cuda_stream),
                   "nrh_render_backward")
    renderer.last_backward_launch_count = lib.nrh_last_launch_count()


class FusedRender(torch.autograd.Function):
    """(rgb, weights, analytic_normals, normalized_analytic_normals) = render(origins, directions, pl_positions; 46 parameters).
    `side` (a dict) carries the non-differentiable inputs in and the non-differentiable RenderOutput fields out."""

    @staticmethod
    def forward(ctx, renderer, side: dict, o, d, pl, *params):
        lib = _lib.load()
        dev = o.device
        f32 = dict(dtype=torch.float32, device=dev)
        cfg = renderer._c_config()
        r = renderer.config.renderer
        R, S = o.shape[0], r.n_samples + r.n_importance_samples
        prep = lambda t: t.detach().to(**f32).contiguous()      # noqa: E731
        o_c, d_c, pl_c = prep(o), prep(d), prep(pl)
        near, far, bg = side["near"], side["far"], side["bg"]
        packed = renderer._ensure_packed(dev)
        ws = torch.empty(lib.nrh_train_workspace_bytes(C.byref(cfg), R), dtype=torch.uint8, device=dev)
        out = dict(rgb=torch.empty(R, 3, **f32), depth=torch.empty(R, 1, **f32), weights=torch.empty(R, S, **f32),
                   inside_sphere=torch.empty(R, S, **f32), analytic_normals=torch.empty(R, S, 3, **f32),
                   normalized_normals=torch.empty(R, S, 3, **f32),
                   visibilities=torch.empty(R, 1, **f32) if r.shadow_hint else None,
                   specular_cue=torch.empty(R, S, len(r.specular_roughness), **f32) if r.specular_hint else None,
                   inv_s=torch.empty(1, **f32), z_vals=torch.empty(R, S, **f32))
        rays = _rays_struct(o_c, d_c, pl_c, near, far)
        run_forward(renderer, cfg, packed, rays, R, bg, side["jit_p"], side["jit_s"], side["cos_anneal"], side["warmup"], out, ws)
        side["out"] = out
        ctx.renderer, ctx.cfg, ctx.packed, ctx.ws = renderer, cfg, packed, ws
        ctx.rays_t = (o_c, d_c, pl_c, near, far)
        ctx.bg, ctx.cos_anneal, ctx.R, ctx.S = bg, float(side["cos_anneal"]), R, S
        ctx.param_shapes = [tuple(p.shape) for p in params]
        return out["rgb"], out["weights"], out["analytic_normals"], out["normalized_normals"]

    @staticmethod
    def backward(ctx, d_rgb, d_w, d_an, d_nn):
        dev = ctx.ws.device
        f32 = dict(dtype=torch.float32, device=dev)
        R, S = ctx.R, ctx.S
        c = lambda t: t.to(torch.float32).contiguous() if t is not None else None      # noqa: E731
        d_rgb = c(d_rgb) if d_rgb is not None else torch.zeros(R, 3, **f32)
        d_w, d_an, d_nn = c(d_w), c(d_an), c(d_nn)
        sizes = [int(torch.Size(s).numel()) for s in ctx.param_shapes]
        flat = torch.empty(sum(sizes), **f32)                    # every entry is overwritten by the backward call
        grads, off = [], 0
        for n, shp in zip(sizes, ctx.param_shapes):
            grads.append(flat[off:off + n].view(shp))
            off += n
        keep: list = []
        P = train_params(ctx.renderer, grads, keep)
        need_o, need_d, need_pl = ctx.needs_input_grad[2:5]
        g_o = torch.empty(R, 3, **f32) if need_o else None
        g_d = torch.empty(R, 3, **f32) if need_d else None
        g_pl = torch.empty(R, 3, **f32) if need_pl else None
        ptr = lambda t: t.data_ptr() if t is not None else None   # noqa: E731
        adj = _lib.NrhTrainAdjoints(ptr(d_rgb), ptr(d_an), ptr(d_nn), ptr(d_w), ptr(g_o), ptr(g_d), ptr(g_pl))
        rays = _rays_struct(*ctx.rays_t)
        run_backward(ctx.renderer, ctx.cfg, ctx.packed, P, rays, R, ctx.bg, ctx.cos_anneal, adj, ctx.ws)
        pg = [g if need else None for g, need in zip(grads, ctx.needs_input_grad[5:])]
        return (None, None, g_o, g_d, g_pl) + tuple(pg)


class FusedTrainStep:
    """forward -> loss -> backward of one batch without autograd (see the module docstring).  The gradients of the renderer's
    parameters go to their `.grad` tensors (views of FlatAdam's flat buffer, or plain tensors allocated here), OVERWRITING them;
    the adjoints of the ray fields are returned for the ray generator's backward."""

    def __init__(self, renderer):
        if not eligible(renderer):
            raise NotImplementedError("FusedTrainStep covers the reference-default training configuration on the tcgen05 engine "
                                      "(weight-normed layers, importance sampling, <= 128 samples, no outside NeRF)")
        self.renderer = renderer
        self._ws = None

    def forward_backward(self, o, d, pl, near, far, rgb_gt, bg, global_step: int, igr_weight: float, need_ray_grads: bool = False):
        """Rays [R,3] / [R,1] fp32 contiguous on the device.  Returns dict(stats = the 8 floats of nrh_train_loss (loss, rgb_loss,
        eikonal_loss, psnr, ...), rgb, d_origins / d_directions / d_pl_positions when requested)."""
        lib = _lib.load()
        rn = self.renderer
        dev = o.device
        f32 = dict(dtype=torch.float32, device=dev)
        cfg = rn._c_config()
        r = rn.config.renderer
        R, S = o.shape[0], r.n_samples + r.n_importance_samples
        warmup = bool(global_step < rn.config.geometry_warmup_end)
        cos_anneal = min(1.0, global_step / rn.config.anneal_end) if rn.config.anneal_end > 0 else 1.0
        jit_p = torch.rand([R, 1], device=dev)                                     # RNG order of the reference (:682, :394)
        jit_s = torch.rand([R, r.n_shadow_samples], device=dev) if (rn.has_shadow_hint and not warmup and r.shadow_hint) else None
        if torch.cuda.is_current_stream_capturing():
            rn._packed_key = None                     # a captured step must always carry its own weight pack
        packed = rn._ensure_packed(dev)
        need = lib.nrh_train_workspace_bytes(C.byref(cfg), R)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        if getattr(self, "_R", None) != R or getattr(self, "_dev", None) != dev:
            self._R, self._dev = R, dev
            self._buf = dict(rgb=torch.empty(R, 3, **f32), depth=torch.empty(R, 1, **f32), weights=torch.empty(R, S, **f32),
                             inside_sphere=torch.empty(R, S, **f32), analytic_normals=torch.empty(R, S, 3, **f32),
                             normalized_normals=None, visibilities=torch.empty(R, 1, **f32) if r.shadow_hint else None,
                             specular_cue=None, inv_s=torch.empty(1, **f32), z_vals=None)
            self._stats = torch.empty(8, **f32)
            self._d_rgb, self._d_n = torch.empty(R, 3, **f32), torch.empty(R, S, 3, **f32)
            self._g_rays = [torch.empty(R, 3, **f32) for _ in range(3)]
        out = self._buf
        rays = _rays_struct(o, d, pl, near, far)
        bgf = bg.reshape(-1) if bg is not None else None
        run_forward(rn, cfg, packed, rays, R, bgf, jit_p, jit_s, cos_anneal, warmup, out, self._ws)
        with torch.cuda.device(dev):
            _lib.check(lib.nrh_train_loss(out["rgb"].data_ptr(), rgb_gt.data_ptr(), out["analytic_normals"].data_ptr(),
                                          out["inside_sphere"].data_ptr(), R, S, float(igr_weight), 1.0, self._stats.data_ptr(),
                                          self._d_rgb.data_ptr(), self._d_n.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                       "nrh_train_loss")
        ps = param_list(rn)
        for p in ps:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        keep: list = []
        P = train_params(rn, [p.grad for p in ps], keep)
        g = self._g_rays if need_ray_grads else [None, None, None]
        ptr = lambda t: t.data_ptr() if t is not None else None   # noqa: E731
        adj = _lib.NrhTrainAdjoints(self._d_rgb.data_ptr(), self._d_n.data_ptr(), None, None, ptr(g[0]), ptr(g[1]), ptr(g[2]))
        run_backward(rn, cfg, packed, P, rays, R, bgf, cos_anneal, adj, self._ws)
        # owned copy (one 32-byte launch): the statistics buffer is rewritten by the next step, and callers keep loss values in lists
        res = dict(stats=self._stats.clone(), rgb=out["rgb"], s_val=(1.0 / out["inv_s"]).reshape(()))
        if need_ray_grads:
            res.update(d_origins=g[0], d_directions=g[1], d_pl_positions=g[2])
        return res
