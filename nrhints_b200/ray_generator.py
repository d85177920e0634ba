"""Host-side mirror of the reference ray generator (/root/reference/camera/ray_generator.py:14-150) on the CUDA
operators nrh_raygen_forward / nrh_raygen_backward (csrc/raygen.cu).

Same constructor, config fields, parameter / buffer names (`cam_pose_adjustment` [N,6], `pl_adjustment` [N,3],
`cam_pose_noise` [N,3,4], `pl_noise` [N,3]) and forward signature as the reference module, so checkpoints and the optimizer's
second parameter group (pipelines/base_pipeline.py:36) carry over.  One kernel launch forward and one backward instead of
~25 ATen launches and an autograd graph per batch.  There is no CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Literal, Optional

import torch
from torch import nn

from . import _lib
from .rays import RayBundle


@dataclass(frozen=True)
class RayGeneratorConfig:
    """camera/ray_generator.py:14-39 (same fields and defaults)."""
    override_near_far_from_sphere: bool = True
    cam_opt_mode: Literal["off", "SO3xR3", "SE3"] = "off"
    pl_opt: bool = False
    opt_lr: float = 3e-5
    cam_position_noise_std: float = 0.0
    cam_orientation_noise_std: float = 0.0
    pl_position_noise_std: float = 0.0


@dataclass
class CameraModel:
    """camera/camera_model.py:5-24."""
    H: int
    W: int
    cx: float
    cy: float
    fx: float
    fy: float
    zn: float
    zf: float


def exp_map_SE3_host(tangent: torch.Tensor) -> torch.Tensor:
    """exp_map_SE3 (camera/lie_groups.py:65-116) for the ONE-OFF construction of the `cam_pose_noise` buffer at module
    init (ray_generator.py:62-68) -- tiny, not on the per-step path."""
    v = tangent[:, :3].reshape(-1, 3, 1)
    w = tangent[:, 3:].reshape(-1, 3, 1)
    th = torch.linalg.norm(w, dim=1).unsqueeze(1)
    th2, th3 = th ** 2, th ** 3
    nz = th < 1e-2
    one = torch.ones(1, dtype=tangent.dtype, device=tangent.device)
    sine = th.sin()
    cosine = torch.where(nz, 8 / (4 + th2) - 1, th.cos())
    s1 = torch.where(nz, 0.5 * cosine + 0.5, sine / torch.where(nz, one, th))
    c2 = torch.where(nz, 0.5 * s1, (1 - cosine) / torch.where(nz, one, th2))
    ret = torch.zeros(tangent.shape[0], 3, 4, dtype=tangent.dtype, device=tangent.device)
    ret[:, :3, :3] = c2 * w @ w.transpose(1, 2)
    for i in range(3):
        ret[:, i, i] += cosine.view(-1)
    t = s1.view(-1, 1) * w.view(-1, 3)
    ret[:, 0, 1] -= t[:, 2]; ret[:, 1, 0] += t[:, 2]; ret[:, 0, 2] += t[:, 1]
    ret[:, 2, 0] -= t[:, 1]; ret[:, 1, 2] -= t[:, 0]; ret[:, 2, 1] += t[:, 0]
    s1t = torch.where(nz, 1 - th2 / 6, s1)
    c2t = torch.where(nz, 0.5 - th2 / 24, c2)
    c3 = torch.where(nz, 1.0 / 6 - th2 / 120, (th - sine) / torch.where(nz, one, th3))
    ret[:, :, 3:] = s1t * v + c2t * torch.cross(w, v, dim=1) + c3 * (w @ (w.transpose(1, 2) @ v))
    return ret


def _ptr(t: Optional[torch.Tensor]):
    return t.data_ptr() if t is not None else None


class _RayGenFn(torch.autograd.Function):
    """(cam_pose_adjustment, pl_adjustment) -> RayBundle fields; everything else is data."""

    @staticmethod
    def forward(ctx, gen: "RayGenerator", cam_adj, pl_adj, w_idx, h_idx, img_idx, poses, pls):
        lib = _lib.load()
        dev = poses.device
        R = poses.shape[0]
        f32 = dict(dtype=torch.float32, device=dev)
        o, d, pl = torch.empty(R, 3, **f32), torch.empty(R, 3, **f32), torch.empty(R, 3, **f32)
        near, far = torch.empty(R, 1, **f32), torch.empty(R, 1, **f32)
        ins = gen._inputs(w_idx, h_idx, img_idx, poses, pls, cam_adj, pl_adj)
        with torch.cuda.device(dev):
            _lib.check(lib.nrh_raygen_forward(C.byref(gen._camera()), gen._mode, int(gen.config.override_near_far_from_sphere),
                                              C.byref(ins), R, o.data_ptr(), d.data_ptr(), pl.data_ptr(), near.data_ptr(),
                                              far.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "nrh_raygen_forward")
        ctx.gen = gen
        ctx.save_for_backward(cam_adj, pl_adj, w_idx, h_idx, img_idx, poses, pls)
        return o, d, pl, near, far

    @staticmethod
    def backward(ctx, g_o, g_d, g_pl, g_near, g_far):
        gen = ctx.gen
        cam_adj, pl_adj, w_idx, h_idx, img_idx, poses, pls = ctx.saved_tensors
        lib = _lib.load()
        dev = poses.device
        R = poses.shape[0]
        want_cam = cam_adj is not None and ctx.needs_input_grad[1]
        want_pl = pl_adj is not None and ctx.needs_input_grad[2]
        d_cam = torch.zeros_like(cam_adj, dtype=torch.float32) if want_cam else None
        d_pl = torch.zeros_like(pl_adj, dtype=torch.float32) if want_pl else None
        if img_idx is not None and (want_cam or want_pl):
            c = lambda g: g.contiguous().to(torch.float32) if g is not None else None      # noqa: E731
            g_o, g_d, g_pl, g_near, g_far = c(g_o), c(g_d), c(g_pl), c(g_near), c(g_far)
            ins = gen._inputs(w_idx, h_idx, img_idx, poses, pls, cam_adj, pl_adj)
            with torch.cuda.device(dev):
                _lib.check(lib.nrh_raygen_backward(C.byref(gen._camera()), gen._mode, int(gen.config.override_near_far_from_sphere),
                                                   C.byref(ins), R, _ptr(g_o), _ptr(g_d), _ptr(g_pl), _ptr(g_near), _ptr(g_far),
                                                   _ptr(d_cam), _ptr(d_pl), torch.cuda.current_stream(dev).cuda_stream),
                           "nrh_raygen_backward")
        return None, d_cam, d_pl, None, None, None, None, None


class RayGenerator(nn.Module):
    """Drop-in for camera/ray_generator.py::RayGenerator."""

    def __init__(self, camera: CameraModel, num_cameras: int, config: RayGeneratorConfig):
        super().__init__()
        self.camera = camera
        self.config = config
        self.num_cameras = int(num_cameras)
        # the reference indexes its per-image tables with torch indexing: negative indices wrap, anything outside [-n, n) raises an
        # IndexError.  Here the range is asserted on the device (torch._assert_async: no host sync, the error surfaces at the next
        # synchronisation); switch off to save the three tiny launches once a data loader is trusted.
        self.validate_indices = True
        if config.cam_opt_mode not in _lib.CAM_OPT_MODES:
            raise ValueError(f"Unknown camera pose optimization mode: {config.cam_opt_mode}")
        self._mode = _lib.CAM_OPT_MODES[config.cam_opt_mode]
        if config.cam_opt_mode != "off":
            self.cam_pose_adjustment = nn.Parameter(torch.zeros((num_cameras, 6)))
        if config.pl_opt:
            self.pl_adjustment = nn.Parameter(torch.zeros((num_cameras, 3)))
        # same RNG draws, in the same order, as the reference constructor (ray_generator.py:62-73)
        if config.cam_position_noise_std != 0.0 or config.cam_orientation_noise_std != 0.0:
            assert config.cam_position_noise_std >= 0.0 and config.cam_orientation_noise_std >= 0.0
            std = torch.tensor([[config.cam_position_noise_std] * 3 + [config.cam_orientation_noise_std] * 3], dtype=torch.float32)
            self.register_buffer("cam_pose_noise", exp_map_SE3_host(torch.normal(torch.zeros((num_cameras, 6)), std)), persistent=True)
        if config.pl_position_noise_std != 0.0:
            assert config.pl_position_noise_std >= 0.0
            self.register_buffer("pl_noise", torch.normal(torch.zeros((num_cameras, 3)), config.pl_position_noise_std), persistent=True)

    def _camera(self) -> _lib.NrhCamera:
        c = self.camera
        return _lib.NrhCamera(float(c.fx), float(c.fy), float(c.cx), float(c.cy), float(c.zn), float(c.zf))

    def _inputs(self, w_idx, h_idx, img_idx, poses, pls, cam_adj, pl_adj) -> _lib.NrhRayGenInputs:
        return _lib.NrhRayGenInputs(
            w_indices=w_idx.data_ptr(), h_indices=h_idx.data_ptr(), img_indices=_ptr(img_idx), poses=poses.data_ptr(),
            pls=pls.data_ptr(), cam_pose_noise=_ptr(getattr(self, "cam_pose_noise", None)), pl_noise=_ptr(getattr(self, "pl_noise", None)),
            cam_pose_adjustment=_ptr(cam_adj), pl_adjustment=_ptr(pl_adj), n_cameras=self.num_cameras)

    def forward(self, pixel_bundle) -> RayBundle:
        """pixel_bundle: anything with the RawPixelBundle fields (data/data_loader.py:80-89): img_indices [R,1] int or None,
        h_indices / w_indices [R,1], poses [R,4,4], pls [R,3]."""
        poses = pixel_bundle.poses
        if not poses.is_cuda:
            raise RuntimeError("nrhints_b200.RayGenerator runs on CUDA tensors only (no CPU fallback)")
        dev = poses.device
        f32 = dict(dtype=torch.float32, device=dev)
        w_idx = pixel_bundle.w_indices[..., 0].to(**f32).contiguous()
        h_idx = pixel_bundle.h_indices[..., 0].to(**f32).contiguous()
        img_idx = None
        if pixel_bundle.img_indices is not None:
            img_idx = pixel_bundle.img_indices[..., 0].to(device=dev, dtype=torch.int64).contiguous()
            if self.validate_indices and img_idx.numel() > 0:
                lo, hi = torch.aminmax(img_idx)
                torch._assert_async((lo >= -self.num_cameras) & (hi < self.num_cameras),
                                    f"RayGenerator: img_indices outside [-{self.num_cameras}, {self.num_cameras})")
        poses = poses.detach().to(**f32).contiguous()
        pls = pixel_bundle.pls.detach().to(**f32).contiguous()
        for name in ("cam_pose_adjustment", "pl_adjustment", "cam_pose_noise", "pl_noise"):
            t = getattr(self, name, None)
            if t is not None and t.device != dev:
                raise RuntimeError(f"RayGenerator.{name} lives on {t.device}, the pixel bundle on {dev}")
        cam_adj = getattr(self, "cam_pose_adjustment", None)
        pl_adj = getattr(self, "pl_adjustment", None)
        o, d, pl, near, far = _RayGenFn.apply(self, cam_adj, pl_adj, w_idx, h_idx, img_idx, poses, pls)
        return RayBundle(origins=o, directions=d, pl_positions=pl, nears=near, fars=far)
