"""Host-side mirror of the reference's renderer interface for the ray-march hot path.

`NeuSHintRenderer` keeps the reference module's constructor, attribute names, parameter names /
shapes / registration order (so reference checkpoints and optimizer states load) and its
`forward(ray_bundle, is_training, background_rgb, global_step) -> RenderOutput` signature
(/root/reference/models/neus_hint_model.py:236-267, :653-758), but every per-sample computation
runs in the hand-written sm_100a CUDA library behind the C ABI of include/nrhints_b200.h.
PyTorch only owns device memory, the stream and the parameters.  There is no CPU fallback.

Mirrored reference pieces:
  SDFNetwork            /root/reference/fields/sdf_field.py:39-148        (params, geometric init, sdf/gradient/forward)
  ReflectanceNetwork    /root/reference/fields/reflectance_network.py:25-96 (params)
  SingleVarianceNetwork /root/reference/models/neus_hint_model.py:104-110
  RenderOutput          /root/reference/models/neus_hint_model.py:216-233
  extract_fields/_geometry  /root/reference/models/neus_hint_model.py:68-93,753-758
  NeRF (outside model)  /root/reference/fields/nerf_density_field.py:30-64  (params; evaluated inside nrh_render_forward)
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch
from torch import nn

from . import _lib
from . import autograd_fine
from . import fused_step
from . import sdf_autograd
from .config import (DepthComputationType, NeRFConfig, NeuSModelConfig, NormalComputationType, ReflectanceNetConfig,
                     SDFNetConfig)


# ------------------------------------------------------------------------------------------------
# output type
# ------------------------------------------------------------------------------------------------
@dataclass
class RenderOutput:
    """Same fields / shapes as the reference RenderOutput (models/neus_hint_model.py:216-233) and the batch operations its callers
    use on it (the reference type is a nerfstudio-style TensorDataclass): `.shape` (batch shape), `len()`, indexing, `.reshape()`,
    `.flatten()`, `.to()`, `.detach()`.  The reference's own helpers work on it unchanged -- `td_concat` only needs a dataclass
    whose fields are tensors or None (utils/tensor_dataclass.py:365-384) -- so the unmodified pipeline
    (pipelines/base_pipeline.py:114-124: `.to('cpu')`, `td_concat(rets)`, `.reshape(img.shape)`) runs with this type as is.
    `RenderOutput.cast(cls)` re-wraps the reference fields in another class if a caller insists on the reference's own."""
    rgb: torch.Tensor                          # [R,3]
    depth: torch.Tensor                        # [R,1]
    weights: torch.Tensor                      # [R,S]
    s_val: torch.Tensor                        # [R,S]
    inside_sphere: torch.Tensor                # [R,S]
    relax_inside_sphere: torch.Tensor          # [R,S]
    analytic_normals: torch.Tensor             # [R,S,3]
    normalized_analytic_normals: torch.Tensor  # [R,S,3]
    visibilities: Optional[torch.Tensor] = None   # [R,1]
    specular_cue: Optional[torch.Tensor] = None   # [R,S,n_rough]
    # extras (not in the reference type): final sample positions, useful for parity debugging
    z_vals: Optional[torch.Tensor] = None
    z_shadow: Optional[torch.Tensor] = None
    sampled_color: Optional[torch.Tensor] = None

    _REFERENCE_FIELDS = ("rgb", "depth", "weights", "s_val", "inside_sphere", "relax_inside_sphere",
                         "analytic_normals", "normalized_analytic_normals", "visibilities", "specular_cue")
    # number of trailing (non-batch) dimensions per field; 1 unless listed (the reference's _field_custom_dimensions)
    _TRAILING = {"analytic_normals": 2, "normalized_analytic_normals": 2, "specular_cue": 2, "sampled_color": 2}

    def as_dict(self) -> Dict[str, Optional[torch.Tensor]]:
        return {f.name: getattr(self, f.name) for f in dataclasses.fields(self)}

    def _map(self, fn) -> "RenderOutput":
        """fn(tensor, n_trailing_dims) on every tensor field."""
        return RenderOutput(**{k: (fn(v, self._TRAILING.get(k, 1)) if v is not None else None) for k, v in self.as_dict().items()})

    def to(self, device, non_blocking: bool = False) -> "RenderOutput":
        return self._map(lambda v, t: v.to(device, non_blocking=non_blocking))

    def detach(self) -> "RenderOutput":
        return self._map(lambda v, t: v.detach())

    def cast(self, cls):
        return cls(**{k: getattr(self, k) for k in self._REFERENCE_FIELDS})

    @property
    def shape(self):
        return tuple(self.rgb.shape[:-1])

    @property
    def ndim(self) -> int:
        return self.rgb.dim() - 1

    @property
    def size(self) -> int:
        return int(np.prod(self.shape)) if self.shape else 1

    def __len__(self) -> int:
        if not self.shape:
            raise TypeError("len() of a 0-d RenderOutput")
        return self.shape[0]

    def __getitem__(self, idx) -> "RenderOutput":
        """Index the batch dimensions; the trailing dimensions of every field are kept whole."""
        if isinstance(idx, torch.Tensor):
            return self._map(lambda v, t: v[idx])
        if not isinstance(idx, tuple):
            idx = (idx,)
        return self._map(lambda v, t: v[idx + (slice(None),) * t])

    def reshape(self, shape) -> "RenderOutput":
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        return self._map(lambda v, t: v.reshape(shape + tuple(v.shape[v.dim() - t:])))

    def flatten(self) -> "RenderOutput":
        return self.reshape((-1,))


# ------------------------------------------------------------------------------------------------
# parameter containers
# ------------------------------------------------------------------------------------------------
class WNLinear(nn.Module):
    """Parameters of a weight-normalised Linear: `bias`, `weight_g` [out,1], `weight_v` [out,in],
    registered in the order nn.utils.weight_norm leaves them (bias, weight_g, weight_v)."""

    def __init__(self, lin: nn.Linear, weight_norm: bool = True):
        super().__init__()
        self.in_features, self.out_features = lin.in_features, lin.out_features
        self.weight_normed = weight_norm
        self.bias = nn.Parameter(lin.bias.detach().clone())
        if weight_norm:
            v = lin.weight.detach().clone()
            self.weight_g = nn.Parameter(torch.norm_except_dim(v, 2, 0))
            self.weight_v = nn.Parameter(v)
        else:
            self.weight = nn.Parameter(lin.weight.detach().clone())

    def effective_weight(self) -> torch.Tensor:
        if self.weight_normed:
            return torch._weight_norm(self.weight_v, self.weight_g, 0)
        return self.weight


def _check_sdf_arch(cfg: SDFNetConfig):
    ok = (cfg.d_in == 3 and cfg.d_out_feat == 256 and cfg.d_hidden == 256 and cfg.n_layers == 8
          and list(cfg.skip_in) == [4] and cfg.multi_res == 6 and float(cfg.scale) == 3.0)
    if not ok:
        raise NotImplementedError("nrhints_b200 kernels implement the reference-default SDF network "
                                  "(8x256, skip_in=[4], multi_res=6, scale=3, d_out_feat=256); got " + repr(cfg))


class SDFNetwork(nn.Module):
    """Parameter container + point-query API of the SDF field (fields/sdf_field.py:39-148)."""

    def __init__(self, config: SDFNetConfig):
        super().__init__()
        _check_sdf_arch(config)
        self.config = config
        d_pe = config.d_in * (2 * config.multi_res + 1)
        dims = [d_pe] + [config.d_hidden] * config.n_layers + [config.d_out_feat + 1]
        self.num_layers = len(dims)
        self.skip_in = list(config.skip_in)
        self.scale = config.scale
        bias = config.init_bias * config.scale
        for l in range(self.num_layers - 2):
            out_dim = dims[l + 1] - dims[0] if (l + 1) in self.skip_in else dims[l + 1]
            lin = nn.Linear(dims[l], out_dim)
            if config.geometric_init:
                std = np.sqrt(2) / np.sqrt(out_dim)
                nn.init.constant_(lin.bias, 0.0)
                if l == 0:
                    nn.init.constant_(lin.weight[:, 3:], 0.0)
                    nn.init.normal_(lin.weight[:, :3], 0.0, std)
                else:
                    nn.init.normal_(lin.weight, 0.0, std)
                    if l in self.skip_in:
                        nn.init.constant_(lin.weight[:, -(dims[0] - 3):], 0.0)
            setattr(self, f"lin{l}", WNLinear(lin, config.weight_norm))
        for name, out_dim in (("sdf", 1), ("feat", dims[-1] - 1)):
            lin = nn.Linear(dims[-2], out_dim)
            if config.geometric_init:
                sign = -1.0 if config.inside_outside else 1.0
                nn.init.normal_(lin.weight, mean=sign * np.sqrt(np.pi) / np.sqrt(dims[-1]), std=0.0001)
                nn.init.constant_(lin.bias, -sign * bias)
            setattr(self, f"out_{name}", WNLinear(lin, config.weight_norm))
        self._owner = None           # set by NeuSHintRenderer (plain attribute, not a submodule)

    def _renderer(self):
        if self._owner is None:
            raise RuntimeError("SDFNetwork point queries run through the owning NeuSHintRenderer's CUDA context")
        return self._owner()

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        sdf, _, feat = self._renderer().sdf_query(inputs, want_grad=False, want_feat=True)
        return torch.cat([sdf[:, None], feat], dim=-1)

    def sdf(self, x: torch.Tensor) -> torch.Tensor:
        return self._renderer().sdf_query(x)[0][:, None]

    def sdf_hidden_appearance(self, x: torch.Tensor) -> torch.Tensor:
        return self._renderer().sdf_query(x, want_feat=True)[2]

    def gradient(self, x: torch.Tensor) -> torch.Tensor:
        return self._renderer().sdf_query(x, want_grad=True)[1].unsqueeze(1)

    def freeze_geometry(self):
        for name, p in self.named_parameters():
            if "lin" in name or "out_sdf" in name:
                p.requires_grad = False


class ReflectanceNetwork(nn.Module):
    """Parameter container of the hint-conditioned radiance MLP (fields/reflectance_network.py:25-66).
    Evaluated only inside the fused render pipeline."""

    def __init__(self, d_feature, d_in, d_out, config: ReflectanceNetConfig, shadow_hint=True, specular_hint=True,
                 specular_hint_len=4):
        super().__init__()
        if not (config.d_hidden == 256 and config.n_layers == 4 and config.multi_res == 4 and config.squeeze_out
                and d_feature == 256 and d_out == 3):
            raise NotImplementedError("nrhints_b200 kernels implement the reference-default reflectance network "
                                      "(4x256, multi_res=4, squeeze_out); got " + repr(config))
        self.config = config
        self.shadow_hint, self.specular_hint = shadow_hint, specular_hint
        pe3 = 3 * (2 * config.multi_res + 1)
        dims = [d_in + d_feature] + [config.d_hidden] * config.n_layers + [d_out]
        dims[0] += (pe3 - 3) * 2
        if shadow_hint:
            dims[0] += (2 * config.multi_res + 1) - 1
        if specular_hint:
            dims[0] += specular_hint_len * (2 * config.multi_res + 1) - specular_hint_len
        self.num_layers = len(dims)
        self.d_in_total = dims[0]
        for l in range(self.num_layers - 1):
            setattr(self, f"lin{l}", WNLinear(nn.Linear(dims[l], dims[l + 1]), config.weight_norm))

    def forward(self, points, normals, view_dirs, feature_vectors, point_lights, visibilities=None, specular_cue=None):
        """Stand-alone evaluation with the reference's signature and input order (fields/reflectance_network.py:68-96:
        [points | PE(view) | normals | PE(light) | features | PE(visibility) | PE(specular cue)] -> 4 x (Linear, ReLU) -> Linear ->
        sigmoid).  API compatibility only: the render path never calls this -- NeuSHintRenderer.forward evaluates the network
        inside the fused CUDA pipeline (csrc/mlp_tc.cu::color_tc_kernel) -- so this accessor is plain fp32 torch on whatever
        device its inputs live."""
        nf = self.config.multi_res

        def pe(x):
            freqs = 2.0 ** torch.linspace(0.0, nf - 1, nf, device=x.device, dtype=x.dtype)
            sx = (x[..., None] * freqs).reshape(*x.shape[:-1], -1)
            return torch.cat([x, torch.sin(torch.cat([sx, sx + torch.pi / 2.0], dim=-1))], dim=-1)   # cosine as sin(x + pi/2), as the reference
        parts = [points, pe(view_dirs), normals, pe(point_lights), feature_vectors]
        if self.shadow_hint:
            parts.append(pe(visibilities))
        if self.specular_hint:
            parts.append(pe(specular_cue))
        x = torch.cat(parts, dim=-1)
        for l in range(self.num_layers - 1):
            lin = getattr(self, f"lin{l}")
            x = torch.nn.functional.linear(x, lin.effective_weight(), lin.bias)
            if l < self.num_layers - 2:
                x = torch.relu(x)
        return torch.sigmoid(x)


class OutsideNeRF(nn.Module):
    """Parameter container of the NeRF++ background model (fields/nerf_density_field.py:30-64): same submodule names,
    shapes and construction order as the reference's `NeRF`, so the seeded init and the state_dict keys
    (`outside_nerf.pts_linears.{i}.weight`, `views_linears.0`, `feature_linear`, `alpha_linear`, `rgb_linear`) match.
    Evaluated only inside the fused render pipeline (render_outside, models/neus_hint_model.py:434-473)."""

    def __init__(self, d_in: int = 4, d_in_view: int = 6, config: NeRFConfig = None):
        super().__init__()
        config = config if config is not None else NeRFConfig()
        if not (d_in == 4 and d_in_view == 6 and config.d_hidden == 256 and config.n_layers == 8 and config.multi_res == 10
                and config.multi_res_view == 4 and list(config.skips) == [4]):
            raise NotImplementedError("nrhints_b200 kernels implement the reference-default outside NeRF "
                                      "(8x256, multi_res=10, multi_res_view=4, skips=[4]); got " + repr(config))
        self.config = config
        W = config.d_hidden
        self.input_ch = d_in * (2 * config.multi_res + 1)
        self.input_ch_view = d_in_view * (2 * config.multi_res_view + 1)
        self.skips = list(config.skips)
        self.pts_linears = nn.ModuleList(
            [nn.Linear(self.input_ch, W)] +
            [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + self.input_ch, W) for i in range(config.n_layers - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(self.input_ch_view + W, W // 2)])
        self.feature_linear = nn.Linear(W, W)
        self.alpha_linear = nn.Linear(W, 1)
        self.rgb_linear = nn.Linear(W // 2, 3)

    def forward(self, input_pts, input_views, input_pls):
        """NeRF.forward with the reference's signature (fields/nerf_density_field.py:66-89: PE(10) of the 4-D inverted-sphere point,
        8 x 256 ReLU trunk with the encoded point re-concatenated IN FRONT after layer 4, density head, feature layer, one 128-wide
        layer on [feature | PE(4) of (view, light)], rgb head) in plain fp32 torch -> (density [N,1], rgb logits [N,3]).
        The render path never calls this -- NeuSHintRenderer.forward evaluates the network inside the fused CUDA pipeline
        (csrc/mlp_simt.cu::nerf_mlp_kernel); it is the DIFFERENTIABLE route of the outside model when a call needs gradients."""
        F = torch.nn.functional
        e = autograd_fine._fourier(input_pts, self.config.multi_res)
        ev = autograd_fine._fourier(torch.cat([input_views, input_pls], dim=-1), self.config.multi_res_view)
        h = e
        for i, lin in enumerate(self.pts_linears):
            h = F.relu(lin(h))
            if i in self.skips:
                h = torch.cat([e, h], -1)
        density = self.alpha_linear(h)
        h = torch.cat([self.feature_linear(h), ev], -1)
        h = F.relu(self.views_linears[0](h))
        return density, self.rgb_linear(h)


class SingleVarianceNetwork(nn.Module):
    def __init__(self, init_val):
        super().__init__()
        self.register_parameter("variance", nn.Parameter(torch.tensor(init_val)))

    def forward(self, x):
        return torch.ones([len(x), 1], device=x.device) * torch.exp(self.variance * 10.0)


# ------------------------------------------------------------------------------------------------
# the renderer
# ------------------------------------------------------------------------------------------------
class NeuSHintRenderer(nn.Module):
    def __init__(self, config: NeuSModelConfig = None, mlp_impl: str = "auto"):
        super().__init__()
        config = config if config is not None else NeuSModelConfig()
        r = config.renderer
        if r.use_outside_nerf and not (1 <= r.n_outside_samples <= _lib.NRH_MAX_OUTSIDE):
            raise NotImplementedError(f"n_outside_samples must be in [1, {_lib.NRH_MAX_OUTSIDE}]")
        if r.n_shadow_importance_clip != -1:
            # one shadow march per GROUP of samples (models/neus_hint_model.py:554-576; off in every preset of the reference): served by
            # the composed evaluation route of hint_fallback.py (inference only), not by the fused pipeline
            S_all = r.n_samples + r.n_importance_samples
            if not (r.n_shadow_importance_clip > 0 and r.shadow_hint and not r.use_outside_nerf and S_all % r.n_shadow_importance_clip == 0):
                raise NotImplementedError("n_shadow_importance_clip must be -1, or a positive divisor of n_samples + n_importance_samples "
                                          "with the shadow hint on and the outside NeRF off; got " + repr(r.n_shadow_importance_clip))
        if r.shadow_hint_gradient or r.specular_hint_gradient:
            raise NotImplementedError("hint gradients are not implemented (reference default is off)")
        if (r.force_shadow_map and not r.shadow_hint) or (r.force_specular_cue and not r.specular_hint):
            raise NotImplementedError("force_shadow_map / force_specular_cue without the hint crash in the reference "
                                      "(width mismatch, SURVEY.md section 8a) and are rejected here")
        self.has_shadow_hint = r.shadow_hint or r.force_shadow_map
        self.has_specular_hint = r.specular_hint or r.force_specular_cue
        self.sdf_network = SDFNetwork(config.sdf_network)
        self.deviation_network = SingleVarianceNetwork(init_val=config.deviation_network.init_val)
        color_d_in = 12 + (1 if self.has_shadow_hint else 0) + (len(r.specular_roughness) if self.has_specular_hint else 0)
        self.color_network = ReflectanceNetwork(
            d_feature=config.sdf_network.d_out_feat, d_in=color_d_in, d_out=3, config=config.reflectance_network,
            shadow_hint=r.shadow_hint, specular_hint=r.specular_hint, specular_hint_len=len(r.specular_roughness))
        self.has_outside_nerf = bool(r.use_outside_nerf)
        if self.has_outside_nerf:
            self.outside_nerf = OutsideNeRF(d_in=4, d_in_view=6, config=getattr(config, "outside_nerf", None))
        self.config = config
        self.mlp_impl = mlp_impl
        import weakref
        self.sdf_network._owner = weakref.ref(self)
        # kernel-side state (not parameters / buffers; rebuilt lazily per device)
        self._packed = None
        self._packed_key = None
        self._workspace = None
        self.last_launch_count = 0
        # gradients through ONE fused autograd node (fused_step.py) when the configuration allows it; False keeps the composed
        # autograd path (autograd_fine.py / sdf_autograd.py), which the parity tests use as a second implementation
        self.fused_training = True

    # -- C-ABI plumbing ---------------------------------------------------------------------------
    def _c_config(self) -> _lib.NrhConfig:
        r = self.config.renderer
        c = _lib.NrhConfig()
        c.n_samples, c.n_importance, c.up_sample_steps = r.n_samples, r.n_importance_samples, r.up_sample_steps
        c.n_shadow_samples, c.n_shadow_importance = r.n_shadow_samples, r.n_shadow_importance_samples
        c.shadow_hint, c.specular_hint = int(r.shadow_hint), int(r.specular_hint)
        rough = list(r.specular_roughness)
        if r.specular_hint and len(rough) > _lib.NRH_MAX_ROUGHNESS:
            raise NotImplementedError(f"at most {_lib.NRH_MAX_ROUGHNESS} specular roughness lobes are supported")
        c.n_roughness = len(rough) if r.specular_hint else 0
        for i, v in enumerate(rough[:_lib.NRH_MAX_ROUGHNESS]):
            c.roughness[i] = float(v)
        c.shadow_ray_offset = float(r.shadow_ray_offset)
        c.normalized_normals = int(getattr(r.normal_type, "value", r.normal_type) == NormalComputationType.NormalizedAnalytic.value)
        c.mlp_impl = _lib.MLP_IMPLS[self.mlp_impl]
        c.depth_type = _lib.DEPTH_TYPES[getattr(r.depth_type, "value", r.depth_type)]
        c.use_outside_nerf, c.n_outside = int(self.has_outside_nerf), int(r.n_outside_samples)
        return c

    def _weight_tensors(self) -> List[torch.Tensor]:
        ws = []
        for l in range(8):
            lin = getattr(self.sdf_network, f"lin{l}")
            ws += [lin.effective_weight(), lin.bias]
        ws += [self.sdf_network.out_sdf.effective_weight(), self.sdf_network.out_sdf.bias,
               self.sdf_network.out_feat.effective_weight(), self.sdf_network.out_feat.bias]
        for l in range(5):
            lin = getattr(self.color_network, f"lin{l}")
            ws += [lin.effective_weight(), lin.bias]
        ws.append(self.deviation_network.variance)
        if self.has_outside_nerf:                  # 31..: 8 x (W, b), alpha, feature, view, rgb
            on = self.outside_nerf
            for lin in list(on.pts_linears) + [on.alpha_linear, on.feature_linear, on.views_linears[0], on.rgb_linear]:
                ws += [lin.weight, lin.bias]
        return ws

    def _ensure_packed(self, device: torch.device) -> torch.Tensor:
        if device.type != "cuda":
            raise RuntimeError("nrhints_b200 runs on CUDA devices only (there is no CPU fallback); "
                               f"got tensors on {device}")
        key = (device, self.mlp_impl, tuple((p.data_ptr(), p._version) for p in self.parameters()))
        if self._packed is not None and self._packed_key == key:
            return self._packed
        lib = _lib.load()
        cfg = self._c_config()
        if fused_step.can_pack_wn(self) and all(p.device == device and p.dtype == torch.float32 and p.is_contiguous()
                                                for p in self.parameters()):
            # weight norm of all layers + the operand images by the library itself (nrh_pack_weights_wn): 4 launches instead of ~90
            fused_step.pack_weights_wn(self, device)
            self._packed_key = key
            return self._packed
        with torch.no_grad():
            ws = [w.detach().to(device=device, dtype=torch.float32).contiguous() for w in self._weight_tensors()]
        raw = _lib.NrhRawWeights()
        for l in range(8):
            raw.sdf_W[l], raw.sdf_b[l] = ws[2 * l].data_ptr(), ws[2 * l + 1].data_ptr()
        raw.sdf_out_W, raw.sdf_out_b, raw.feat_W, raw.feat_b = (w.data_ptr() for w in ws[16:20])
        for l in range(5):
            raw.col_W[l], raw.col_b[l] = ws[20 + 2 * l].data_ptr(), ws[21 + 2 * l].data_ptr()
        raw.variance = ws[30].data_ptr()
        if self.has_outside_nerf:
            for l in range(8):
                raw.nerf_W[l], raw.nerf_b[l] = ws[31 + 2 * l].data_ptr(), ws[32 + 2 * l].data_ptr()
            (raw.nerf_alpha_W, raw.nerf_alpha_b, raw.nerf_feat_W, raw.nerf_feat_b, raw.nerf_view_W, raw.nerf_view_b,
             raw.nerf_rgb_W, raw.nerf_rgb_b) = (w.data_ptr() for w in ws[47:55])
        nbytes = lib.nrh_packed_weights_bytes(C.byref(cfg))
        if self._packed is None or self._packed.numel() < nbytes or self._packed.device != device:
            self._packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        with torch.cuda.device(device):
            _lib.check(lib.nrh_pack_weights(C.byref(cfg), C.byref(raw), self._packed.data_ptr(), nbytes, stream),
                       "nrh_pack_weights")
        self._packed_key = key
        self._pack_keepalive = ws            # keep sources alive until the stream has consumed them
        return self._packed

    def _ensure_workspace(self, nbytes: int, device: torch.device) -> torch.Tensor:
        if self._workspace is None or self._workspace.numel() < nbytes or self._workspace.device != device:
            self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self._workspace

    # -- point queries (SDFNetwork.sdf / .gradient / .forward, extract_fields) --------------------------
    @torch.no_grad()
    def sdf_query(self, pts: torch.Tensor, want_grad: bool = False, want_feat: bool = False):
        lib = _lib.load()
        device = pts.device
        packed = self._ensure_packed(device)
        cfg = self._c_config()
        x = pts.detach().to(torch.float32).reshape(-1, 3).contiguous()
        N = x.shape[0]
        sdf = torch.empty(N, dtype=torch.float32, device=device)
        grad = torch.empty(N, 3, dtype=torch.float32, device=device) if want_grad else None
        feat = torch.empty(N, 256, dtype=torch.float32, device=device) if want_feat else None
        wsb = lib.nrh_query_workspace_bytes(C.byref(cfg), N)
        ws = self._ensure_workspace(wsb, device)
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            _lib.check(lib.nrh_sdf_query(C.byref(cfg), packed.data_ptr(), x.data_ptr(), N, sdf.data_ptr(),
                                         grad.data_ptr() if want_grad else None, feat.data_ptr() if want_feat else None,
                                         ws.data_ptr(), ws.numel(), stream), "nrh_sdf_query")
        self.last_launch_count = lib.nrh_last_launch_count()
        return sdf, grad, feat

    @torch.no_grad()
    def sphere_trace(self, rays_o: torch.Tensor, rays_d: torch.Tensor, num_iterations: int, convergence_threshold: float,
                     far: float, check_every: int = 16):
        """NeuSHintRenderer.sphere_trace (models/neus_hint_model.py:359-371) on the SDF kernel -> (points [R,3], depths [R,1])."""
        lib = _lib.load()
        device = rays_o.device
        packed = self._ensure_packed(device)
        cfg = self._c_config()
        o = rays_o.detach().to(torch.float32).contiguous()
        d = rays_d.detach().to(torch.float32).contiguous()
        R = o.shape[0]
        pts = torch.empty(R, 3, dtype=torch.float32, device=device)
        dep = torch.empty(R, 1, dtype=torch.float32, device=device)
        wsb = lib.nrh_query_workspace_bytes(C.byref(cfg), R) + 32 * R
        ws = self._ensure_workspace(wsb, device)
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            _lib.check(lib.nrh_sphere_trace(C.byref(cfg), packed.data_ptr(), o.data_ptr(), d.data_ptr(), R, int(num_iterations),
                                            float(convergence_threshold), float(far), int(check_every), pts.data_ptr(),
                                            dep.data_ptr(), ws.data_ptr(), ws.numel(), stream), "nrh_sphere_trace")
        return pts, dep

    # -- full-image evaluation ------------------------------------------------------------------
    @torch.no_grad()
    def render_maps(self, ray_bundle, background_rgb: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """Per-ray maps of an evaluation render, reduced on the device: rgb [R,3], depth [R,1], shadow_map [R,1],
        specular_hint [R,n_rough] and the weighted normal maps `einsum('...ij,...i,...i->...j', normals, weights, inside)`
        that pipelines/base_pipeline.py:126-133 computes on the host from the per-sample tensors (before its camera
        rotation).  Nothing per-sample is written or copied: 60 B/ray instead of 7 KB/ray leave the device."""
        lib = _lib.load()
        device = ray_bundle.origins.device
        packed = self._ensure_packed(device)
        cfg = self._c_config()
        r = self.config.renderer
        R = ray_bundle.origins.shape[0]
        f32 = dict(dtype=torch.float32, device=device)

        def prep(t):
            return t.detach().to(**f32).contiguous()
        o, d, pl = prep(ray_bundle.origins), prep(ray_bundle.directions), prep(ray_bundle.pl_positions)
        near, far = prep(ray_bundle.nears), prep(ray_bundle.fars)
        bg = prep(background_rgb).reshape(-1) if background_rgb is not None else None
        maps = dict(rgb=torch.empty(R, 3, **f32), depth=torch.empty(R, 1, **f32),
                    analytic_normals=torch.empty(R, 3, **f32), normalized_analytic_normals=torch.empty(R, 3, **f32))
        if r.shadow_hint:
            maps["shadow_map"] = torch.empty(R, 1, **f32)
        if r.specular_hint:
            maps["specular_hint"] = torch.empty(R, len(r.specular_roughness), **f32)
        hit_pts = hit_dep = None
        if cfg.depth_type == _lib.DEPTH_TYPES["sphere_tracing"] and R > 0:
            hit_pts, hit_dep = self.sphere_trace(o, d, 2000, 1e-4, 100.0)
        c_rays = _lib.NrhRays(o.data_ptr(), d.data_ptr(), pl.data_ptr(), near.data_ptr(), far.data_ptr(),
                              hit_pts.data_ptr() if hit_pts is not None else None, hit_dep.data_ptr() if hit_dep is not None else None)
        c_out = _lib.NrhOutputs(rgb=maps["rgb"].data_ptr(), depth=maps["depth"].data_ptr(),
                                visibilities=maps["shadow_map"].data_ptr() if r.shadow_hint else None,
                                normal_map=maps["analytic_normals"].data_ptr(),
                                normalized_normal_map=maps["normalized_analytic_normals"].data_ptr(),
                                specular_cue_ray=maps["specular_hint"].data_ptr() if r.specular_hint else None)
        ws = self._ensure_workspace(lib.nrh_workspace_bytes(C.byref(cfg), R), device)
        if R > 0:
            with torch.cuda.device(device):
                stream = torch.cuda.current_stream(device).cuda_stream
                _lib.check(lib.nrh_render_forward(C.byref(cfg), packed.data_ptr(), C.byref(c_rays), R,
                                                  bg.data_ptr() if bg is not None else None, None, None, None, 1.0, 0,
                                                  C.byref(c_out), ws.data_ptr(), ws.numel(), stream), "nrh_render_forward")
            self.last_launch_count = lib.nrh_last_launch_count()
        return maps

    @torch.no_grad()
    def render_image(self, ray_bundle, background_rgb: Optional[torch.Tensor] = None, chunk_rays: int = 16384,
                     device=None, to_host: bool = True) -> Dict[str, torch.Tensor]:
        """Full-image evaluation: all rays of a view in one call.  The reference renders a view as
        H*W/512 separate forward calls, each followed by `.to('cpu')` of the 7 KB/ray RenderOutput, concatenates them on
        the host and reduces the normal maps there (pipelines/base_pipeline.py:107-133).  Here the (host or device)
        RayBundle is walked in `chunk_rays` slices through render_maps -- per-ray maps reduced on the device, 60 B/ray --
        the slices are written straight into full-size device maps, and ONE pinned device->host copy per map ends the
        call.  Host->device copies of slice i+1 are issued on a second stream while slice i renders.
        Returns rgb [N,3], depth [N,1], analytic_normals [N,3], normalized_analytic_normals [N,3] (weighted normal maps
        before the camera rotation of :127-133), shadow_map [N,1], specular_hint [N,n_rough]."""
        device = torch.device(device) if device is not None else next(self.parameters()).device
        N = ray_bundle.origins.shape[0]
        main = torch.cuda.current_stream(device)
        if getattr(self, "_copy_stream", None) is None or self._copy_stream.device != device:
            self._copy_stream = torch.cuda.Stream(device=device)
            self._early_event = torch.cuda.Event()
        bg = background_rgb.to(device, non_blocking=True) if background_rgb is not None else None
        fields = ("origins", "directions", "pl_positions", "nears", "fars")
        on_device = ray_bundle.origins.device == device

        def stage(i0):
            sl = {k: getattr(ray_bundle, k)[i0:i0 + chunk_rays] for k in fields}
            if on_device:
                return type(ray_bundle)(**sl), None
            with torch.cuda.stream(self._copy_stream):
                dev = type(ray_bundle)(**{k: v.to(device, non_blocking=True) for k, v in sl.items()})
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            return dev, ev

        full: Dict[str, torch.Tensor] = {}
        nxt = stage(0) if N > 0 else None
        for i0 in range(0, N, chunk_rays):
            cur, ev = nxt
            if ev is not None:
                main.wait_event(ev)
            nxt = stage(i0 + chunk_rays) if i0 + chunk_rays < N else None
            maps = self.render_maps(cur, background_rgb=bg)
            for k, v in maps.items():
                if k not in full:
                    full[k] = torch.empty((N,) + tuple(v.shape[1:]), dtype=v.dtype, device=device)
                full[k][i0:i0 + v.shape[0]].copy_(v)
            if ev is not None:
                for t in (cur.origins, cur.directions, cur.pl_positions, cur.nears, cur.fars):
                    t.record_stream(main)                    # allocated on the copy stream, consumed on the main stream
        if not to_host:
            return full
        host = {}
        for k, v in full.items():
            buf = self._pinned_like("image::" + k, v)
            buf.copy_(v, non_blocking=True)
            host[k] = buf
        main.synchronize()
        return host

    # -- the hot path ------------------------------------------------------------------------------------
    def forward(self, ray_bundle, is_training: bool = False, background_rgb: Optional[torch.Tensor] = None,
                global_step: int = 0, return_extras: bool = False, _early_event=None, _fine_events=None) -> RenderOutput:
        if self.config.renderer.n_shadow_importance_clip > 0:
            return self._forward_grouped_shadow(ray_bundle, is_training, background_rgb, global_step, return_extras, _early_event)
        return self._forward_fused(ray_bundle, is_training, background_rgb, global_step, return_extras, _early_event, _fine_events)

    def _forward_grouped_shadow(self, ray_bundle, is_training, background_rgb, global_step, return_extras, _early_event=None) -> RenderOutput:
        """`n_shadow_importance_clip > 0` (models/neus_hint_model.py:554-576): evaluation route composed of the fused forward (sample
        positions, weights, normals, depth, specular cue: nothing of that depends on the visibility), nrh_sdf_query for the grouped
        shadow marches and the features, and the stand-alone reflectance network (nrhints_b200/hint_fallback.py).  Inference only."""
        from . import hint_fallback as hf
        fields = (ray_bundle.origins, ray_bundle.directions, ray_bundle.pl_positions)
        needs_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or any(t.requires_grad for t in fields))
        if is_training or needs_grad:
            raise NotImplementedError("n_shadow_importance_clip > 0 is served by an evaluation-only route: call it with is_training=False "
                                      "under torch.no_grad() (training with this option stays with the reference renderer)")
        r = self.config.renderer
        with torch.no_grad():
            base = self._forward_fused(ray_bundle, False, background_rgb, global_step, True, None, None)
            o, d, pl = (t.detach().to(torch.float32).contiguous() for t in fields)
            inv_s = torch.exp(self.deviation_network.variance * 10.0).clip(1e-6, 1e6).to(o.device).reshape(())
            vis, shadow_map = hf.grouped_visibility(self.sdf_query, o, d, pl, base.z_vals, base.weights, r.n_shadow_importance_clip,
                                                    r.n_shadow_samples, r.n_shadow_importance_samples, inv_s, r.shadow_ray_offset,
                                                    chunk=int(getattr(self.config, "shadow_mini_chunk_size", 2048)))
            normalized = getattr(r.normal_type, "value", r.normal_type) == NormalComputationType.NormalizedAnalytic.value
            normals = base.normalized_analytic_normals if normalized else base.analytic_normals
            bg = background_rgb.detach().to(o.device, torch.float32) if background_rgb is not None else None
            rgb, colour = hf.shade_samples(self.sdf_query, self.color_network, o, d, pl, base.z_vals, base.weights, normals, vis,
                                           base.specular_cue if self.has_specular_hint else None, 2.0 / r.n_samples, bg)
        if _early_event is not None:                     # render_to_host: every field is final only here
            _early_event.record(torch.cuda.current_stream(o.device))
        return dataclasses.replace(base, rgb=rgb, visibilities=shadow_map, sampled_color=colour if return_extras else None,
                                   z_vals=base.z_vals if return_extras else None, z_shadow=None)

    def _forward_fused(self, ray_bundle, is_training: bool = False, background_rgb: Optional[torch.Tensor] = None,
                       global_step: int = 0, return_extras: bool = False, _early_event=None, _fine_events=None) -> RenderOutput:
        lib = _lib.load()
        rays_o = ray_bundle.origins
        device = rays_o.device
        packed = self._ensure_packed(device)
        cfg = self._c_config()
        r = self.config.renderer
        R = rays_o.shape[0]
        S = r.n_samples + r.n_importance_samples
        St = S + (r.n_outside_samples if self.has_outside_nerf else 0)    # weights carry the appended outside samples (:521-524)
        Ss = r.n_shadow_samples + r.n_shadow_importance_samples
        f32 = dict(dtype=torch.float32, device=device)

        # Gradients requested?  (training, or camera registration at eval time, pipelines/base_pipeline.py:71-91)
        ray_fields = (rays_o, ray_bundle.directions, ray_bundle.pl_positions, ray_bundle.nears, ray_bundle.fars)
        needs_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters())
                                                  or any(t.requires_grad for t in ray_fields))
        want_z = return_extras or needs_grad

        def prep(t):
            return t.detach().to(**f32).contiguous()
        o, d, pl = prep(rays_o), prep(ray_bundle.directions), prep(ray_bundle.pl_positions)
        near, far = prep(ray_bundle.nears), prep(ray_bundle.fars)
        bg = prep(background_rgb).reshape(-1) if background_rgb is not None else None

        warmup = bool(is_training and global_step < self.config.geometry_warmup_end)
        cos_anneal = 1.0
        if is_training and self.config.anneal_end > 0:
            cos_anneal = min(1.0, global_step / self.config.anneal_end)
        # RNG draws in the reference's order (:682, :689 (outside NeRF only), then :394)
        jit_p = jit_s = jit_o = None
        if is_training:
            jit_p = torch.rand([R, 1], device=device)
            if self.has_outside_nerf:
                jit_o = torch.rand([R, r.n_outside_samples], device=device)
            if self.has_shadow_hint and not warmup and r.shadow_hint:
                jit_s = torch.rand([R, r.n_shadow_samples], device=device)

        if needs_grad and R > 0 and not return_extras and self.fused_training and fused_step.eligible(self):
            # the whole differentiable step as ONE autograd node on two library calls (nrh_render_train_forward / nrh_render_backward)
            side = dict(near=near, far=far, bg=bg, jit_p=jit_p, jit_s=jit_s, cos_anneal=cos_anneal, warmup=warmup)
            rgb, w, an, nn_ = fused_step.FusedRender.apply(self, side, rays_o, ray_bundle.directions, ray_bundle.pl_positions,
                                                           *fused_step.param_list(self))
            fo = side["out"]
            inv_s = torch.exp(self.deviation_network.variance * 10.0).clip(1e-6, 1e6).to(device)
            return RenderOutput(
                rgb=rgb, depth=fo["depth"], weights=w, s_val=(1.0 / inv_s).reshape(1, 1).expand(R, S),
                inside_sphere=fo["inside_sphere"], relax_inside_sphere=fo["inside_sphere"], analytic_normals=an,
                normalized_analytic_normals=nn_, visibilities=fo["visibilities"] if self.has_shadow_hint else None,
                specular_cue=fo["specular_cue"] if self.has_specular_hint else None, z_vals=fo["z_vals"], z_shadow=None,
                sampled_color=None)

        out = dict(
            rgb=torch.empty(R, 3, **f32), depth=torch.empty(R, 1, **f32), weights=torch.empty(R, St, **f32),
            inside_sphere=torch.empty(R, S, **f32), analytic_normals=torch.empty(R, S, 3, **f32),
            normalized_normals=torch.empty(R, S, 3, **f32),
            visibilities=torch.empty(R, 1, **f32) if r.shadow_hint else None,
            specular_cue=torch.empty(R, S, len(r.specular_roughness), **f32) if r.specular_hint else None,
            inv_s=torch.empty(1, **f32),
            z_vals=torch.empty(R, S, **f32) if want_z else None,
            z_shadow=torch.zeros(R, Ss, **f32) if (return_extras and r.shadow_hint) else None,
            sampled_color=torch.empty(R, St, 3, **f32) if return_extras else None)
        hit_pts = hit_dep = None
        if cfg.depth_type == _lib.DEPTH_TYPES["sphere_tracing"] and R > 0:
            hit_pts, hit_dep = self.sphere_trace(o, d, 2000, 1e-4, 100.0)        # models/neus_hint_model.py:529
        c_rays = _lib.NrhRays(o.data_ptr(), d.data_ptr(), pl.data_ptr(), near.data_ptr(), far.data_ptr(),
                              hit_pts.data_ptr() if hit_pts is not None else None,
                              hit_dep.data_ptr() if hit_dep is not None else None)
        c_out = _lib.NrhOutputs(**{k: (v.data_ptr() if v is not None else None) for k, v in out.items()})
        if _early_event is not None:
            c_out.early_event = _early_event.cuda_event
        if _fine_events is not None:               # (begin, end) torch.cuda.Event pair around the primary fine-pass kernel
            c_out.fine_begin_event, c_out.fine_end_event = _fine_events[0].cuda_event, _fine_events[1].cuda_event
        # training capture (tcgen05 engine): the primary fine pass of the render call doubles as the forward of the SDF autograd
        # node (it writes the tape), so the 128 fine samples per ray are evaluated once per step instead of twice
        captured = c_cap = None
        if needs_grad and R > 0 and self.mlp_impl in ("auto", "tcgen05") and not self.has_outside_nerf:
            N = R * S
            lay = _lib.NrhTrainLayout()
            _lib.check(lib.nrh_sdf_train_layout(C.byref(cfg), N, C.byref(lay)), "nrh_sdf_train_layout")
            captured = dict(tape=torch.empty(int(lay.tape_bytes), dtype=torch.uint8, device=device), sdf=torch.empty(N, **f32),
                            grad_soa=torch.empty(3, N, **f32), feat=torch.empty(N, 256, **f32), pts_soa=torch.empty(3, N, **f32))
            c_cap = _lib.NrhTrainCapture(captured["tape"].data_ptr(), captured["tape"].numel(), captured["sdf"].data_ptr(),
                                         captured["grad_soa"].data_ptr(), captured["feat"].data_ptr(), captured["pts_soa"].data_ptr())
            c_out.train_capture = C.addressof(c_cap)
        wsb = lib.nrh_workspace_bytes(C.byref(cfg), R)
        ws = self._ensure_workspace(wsb, device)
        if R > 0:
            with torch.cuda.device(device):
                stream = torch.cuda.current_stream(device).cuda_stream
                _lib.check(lib.nrh_render_forward(
                    C.byref(cfg), packed.data_ptr(), C.byref(c_rays), R, bg.data_ptr() if bg is not None else None,
                    jit_p.data_ptr() if jit_p is not None else None, jit_o.data_ptr() if jit_o is not None else None,
                    jit_s.data_ptr() if jit_s is not None else None,
                    float(cos_anneal), int(warmup), C.byref(c_out), ws.data_ptr(), ws.numel(), stream),
                    "nrh_render_forward")
            self.last_launch_count = lib.nrh_last_launch_count()

        inv_s = torch.exp(self.deviation_network.variance * 10.0).clip(1e-6, 1e6).to(device)
        s_val = (1.0 / inv_s).reshape(1, 1).expand(R, S)
        if needs_grad and R > 0:
            # composed autograd route (nrhints_b200/autograd_fine.py; the fused node above did not apply): the no_grad parts of the
            # reference ran in the CUDA kernels above; the differentiable fine pass is a handful of autograd nodes on the same device
            fine = self._differentiable_fine(ray_fields, out, jit_p, background_rgb, cos_anneal, inv_s, f32, captured, jit_o)
            out["rgb"], out["weights"] = fine["rgb"], fine["weights"]
            out["analytic_normals"], out["normalized_normals"] = fine["analytic_normals"], fine["normalized_analytic_normals"]
            if out["sampled_color"] is not None:
                out["sampled_color"] = fine["sampled_color"]
        return RenderOutput(
            rgb=out["rgb"], depth=out["depth"], weights=out["weights"], s_val=s_val,
            inside_sphere=out["inside_sphere"], relax_inside_sphere=out["inside_sphere"],      # reference quirk (:746)
            analytic_normals=out["analytic_normals"], normalized_analytic_normals=out["normalized_normals"],
            visibilities=out["visibilities"] if self.has_shadow_hint else None,
            specular_cue=out["specular_cue"] if self.has_specular_hint else None,
            z_vals=out["z_vals"], z_shadow=out["z_shadow"], sampled_color=out["sampled_color"])

    def _autograd_weights(self) -> Dict[str, object]:
        sn, cn = self.sdf_network, self.color_network
        return {
            "sdf_w": [getattr(sn, f"lin{l}").effective_weight() for l in range(8)],
            "sdf_b": [getattr(sn, f"lin{l}").bias for l in range(8)],
            "sdf_w_head": sn.out_sdf.effective_weight(), "sdf_b_head": sn.out_sdf.bias,
            "feat_w": sn.out_feat.effective_weight(), "feat_b": sn.out_feat.bias,
            "col_w": [getattr(cn, f"lin{l}").effective_weight() for l in range(5)],
            "col_b": [getattr(cn, f"lin{l}").bias for l in range(5)],
        }

    def _outside_terms(self, rays_o, rays_d, rays_pl, fars, z_inner, jit_o, f32):
        """Differentiable background model of a call that needs gradients (models/neus_hint_model.py:677-694,:716-724,:434-473):
        the outside sample positions (inverse-depth spacing beyond `far`, stratified jitter = the same torch.rand draw the CUDA forward
        received), merged and sorted with the inner positions, evaluated by the torch NeRF -> per-section alpha and colour of the
        merged set.  The fused forward computes the same quantities in nerf_mlp_kernel; this route exists for autograd."""
        r = self.config.renderer
        R, S = z_inner.shape
        n, n_out = r.n_samples, r.n_outside_samples
        zo = torch.linspace(1e-3, 1.0 - 1.0 / (n_out + 1.0), n_out, **f32)
        if jit_o is not None:
            mids = 0.5 * (zo[1:] + zo[:-1])
            upper, lower = torch.cat([mids, zo[-1:]]), torch.cat([zo[:1], mids])
            zo = lower[None, :] + (upper - lower)[None, :] * jit_o
        zo = (fars / torch.flip(zo, dims=[-1]) + 1.0 / n).expand(R, n_out)
        z_feed, _ = torch.sort(torch.cat([z_inner, zo], dim=-1), dim=-1)
        dists = torch.cat([z_feed[:, 1:] - z_feed[:, :-1], torch.full((R, 1), 2.0 / n, **f32)], -1)
        pts = rays_o[:, None, :] + rays_d[:, None, :] * (z_feed + dists * 0.5)[..., None]
        dis = torch.linalg.norm(pts, ord=2, dim=-1, keepdim=True).clip(1.0, 1e10)
        pts4 = torch.cat([pts / dis, 1.0 / dis], dim=-1).reshape(-1, 4)
        St = S + n_out
        density, col = self.outside_nerf(pts4, rays_d[:, None, :].expand(R, St, 3).reshape(-1, 3),
                                         rays_pl[:, None, :].expand(R, St, 3).reshape(-1, 3))
        alpha = 1.0 - torch.exp(-torch.nn.functional.softplus(density.reshape(R, St)) * dists)
        return {"alpha": alpha, "color": torch.sigmoid(col).reshape(R, St, 3)}

    def _differentiable_fine(self, ray_fields, out, jit_p, background_rgb, cos_anneal, inv_s, f32, captured=None, jit_o=None):
        r = self.config.renderer
        rays_o, rays_d, rays_pl, nears, fars = (t.to(**f32) for t in ray_fields)
        n = r.n_samples
        z_coarse = nears + (fars - nears) * torch.linspace(0.0, 1.0, n, **f32)[None, :]
        if jit_p is not None:
            z_coarse = z_coarse + (jit_p - 0.5) * 2.0 / n
        # The reference re-assigns z_vals inside its `with torch.no_grad()` up-sampling block (models/neus_hint_model.py:696-713),
        # so with n_importance > 0 the final sample positions carry NO gradient (near / far get none); only without importance
        # sampling do the coarse positions keep their path to near / far.
        z = out["z_vals"] if r.n_importance_samples > 0 else autograd_fine.attach_coarse_gradient(out["z_vals"], z_coarse)
        vis = out["visibilities"] if r.shadow_hint else None
        spec = out["specular_cue"][:, 0, :] if r.specular_hint else None
        bg = background_rgb.to(**f32) if background_rgb is not None else None
        normalized = getattr(r.normal_type, "value", r.normal_type) == NormalComputationType.NormalizedAnalytic.value
        w = self._autograd_weights()
        # tcgen05 engine: the SDF network + its input gradient are one autograd node with a fused CUDA forward AND backward
        # (sdf_autograd.py / csrc/mlp_tc_bwd.inc); the fp32 engine keeps the torch expression of the same function
        cap = None
        if captured is not None:
            cap = dict(tape=captured["tape"], sdf=captured["sdf"], feat=captured["feat"], grad=captured["grad_soa"].t().contiguous(),
                       pts=captured["pts_soa"].t().contiguous())
        sdf_fn = (lambda pts: sdf_autograd.sdf_fine(self, pts, w, cap)) if self.mlp_impl in ("auto", "tcgen05") else None
        outside = None
        if self.has_outside_nerf:
            outside = self._outside_terms(rays_o, rays_d, rays_pl, fars, out["z_vals"], jit_o, f32)
            outside["inside"] = out["inside_sphere"]
        return autograd_fine.render_fine(w, rays_o, rays_d, rays_pl, z, 2.0 / n, vis, spec, bg,
                                         float(cos_anneal), inv_s, normalized, refl_freq=self.config.reflectance_network.multi_res,
                                         sdf_fn=sdf_fn, sample_major=captured is not None,
                                         renderer=self if (sdf_fn is not None and self.color_network.d_in_total <= 384) else None,
                                         outside=outside)

    # -- device -> host hand-off of a RenderOutput (the reference does `rendering_res.to('cpu')` per 512-ray chunk,
    #    pipelines/base_pipeline.py:120, through pageable memory; here: pinned buffers owned by the returned object (recycled by
    #    torch's caching host allocator once dropped), one async copy per field on the current stream, one sync) -------------
    @staticmethod
    def _pinned_like(name: str, v: torch.Tensor) -> torch.Tensor:
        """A pinned host tensor OWNED BY THE CALLER.  Every call hands out fresh memory from torch's caching host allocator
        (page-locked blocks are recycled only after the tensor that held them is gone), so results of consecutive calls never
        alias -- the reference's evaluation loop appends `rendering_res.to('cpu')` of every chunk to a list before `td_concat`
        (pipelines/base_pipeline.py:112-123) and must be able to do the same with these."""
        del name
        return torch.empty(v.shape, dtype=v.dtype, pin_memory=True)

    def to_host(self, out: RenderOutput, _prefilled=None) -> RenderOutput:
        """RenderOutput in pinned host memory owned by the returned object (never aliased by later calls)."""
        host = {}
        for k, v in out.as_dict().items():
            if v is None:
                host[k] = None
                continue
            if k == "relax_inside_sphere":
                continue
            if _prefilled is not None and k in _prefilled:
                host[k] = _prefilled[k]
                continue
            buf = self._pinned_like(k, v)
            buf.copy_(v.detach(), non_blocking=True)
            host[k] = buf
        host["relax_inside_sphere"] = host["inside_sphere"]
        torch.cuda.current_stream(out.rgb.device).synchronize()
        return RenderOutput(**host)

    # the per-sample geometry block is final before the shadow march and the reflectance network run (NrhOutputs.early_event)
    _EARLY_FIELDS = ("weights", "inside_sphere", "analytic_normals", "normalized_analytic_normals", "specular_cue", "z_vals")

    @torch.no_grad()
    def render_to_host(self, ray_bundle, background_rgb: Optional[torch.Tensor] = None, device=None,
                       return_extras: bool = False) -> RenderOutput:
        """Inference render of a (host or device) RayBundle with the whole RenderOutput delivered in pinned host memory --
        the `renderer(rays); rendering_res.to('cpu')` pattern of the reference's evaluation loop
        (pipelines/base_pipeline.py:114-120) as one call.  The 7 KB/ray of per-sample geometry are copied on a second
        stream while the shadow march and the reflectance network still run, so the call costs about one forward."""
        device = torch.device(device) if device is not None else next(self.parameters()).device
        main = torch.cuda.current_stream(device)
        if getattr(self, "_copy_stream", None) is None or self._copy_stream.device != device:
            self._copy_stream = torch.cuda.Stream(device=device)
            self._early_event = torch.cuda.Event()
        rays = ray_bundle.to(device, non_blocking=True)
        bg = background_rgb.to(device, non_blocking=True) if background_rgb is not None else None
        self._early_event.record(main)                       # materialises the CUDA event; the library re-records it
        out = self.forward(rays, is_training=False, background_rgb=bg, return_extras=return_extras,
                           _early_event=self._early_event)
        early = {k: getattr(out, k) for k in self._EARLY_FIELDS if getattr(out, k, None) is not None}
        self._copy_stream.wait_event(self._early_event)
        early_host = {}
        with torch.cuda.stream(self._copy_stream):
            for k, v in early.items():
                early_host[k] = self._pinned_like(k, v)
                early_host[k].copy_(v, non_blocking=True)
                v.record_stream(self._copy_stream)               # allocated on the main stream, read on the copy stream
        host = self.to_host(out, _prefilled=early_host)         # the remaining small fields, then a sync of the main stream
        self._copy_stream.synchronize()
        return host

    # -- meshing helpers (models/neus_hint_model.py:68-93,753-758) -------------------------------------------
    @torch.no_grad()
    def extract_fields(self, bound_min, bound_max, resolution: int, chunk: int = 1 << 22) -> np.ndarray:
        """-sdf on a resolution^3 grid (the reference walks 64^3 blocks; here the grid is streamed in
        `chunk`-point slabs straight through the SDF kernel)."""
        device = next(self.parameters()).device
        xs = [torch.linspace(float(bound_min[i]), float(bound_max[i]), resolution, device=device) for i in range(3)]
        u = np.zeros([resolution] * 3, dtype=np.float32)
        rows = max(1, chunk // (resolution * resolution))
        for x0 in range(0, resolution, rows):
            xx, yy, zz = torch.meshgrid(xs[0][x0:x0 + rows], xs[1], xs[2], indexing="ij")
            pts = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=-1)
            val = -self.sdf_query(pts)[0]
            u[x0:x0 + rows] = val.reshape(xx.shape).cpu().numpy()
        return u

    def extract_geometry(self, bound_min, bound_max, resolution, threshold=0.0):
        try:
            import mcubes
        except ImportError as e:          # same dependency as the reference (models/neus_hint_model.py:6)
            raise ImportError("extract_geometry needs PyMCubes (`mcubes`), as the reference does") from e
        u = self.extract_fields(bound_min, bound_max, resolution)
        vertices, triangles = mcubes.marching_cubes(u, threshold)
        b_max = torch.as_tensor(bound_max).detach().cpu().numpy()
        b_min = torch.as_tensor(bound_min).detach().cpu().numpy()
        vertices = vertices / (resolution - 1.0) * (b_max - b_min)[None, :] + b_min[None, :]
        return vertices, triangles
