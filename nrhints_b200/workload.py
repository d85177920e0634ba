"""Deterministic synthetic workload of BASELINE.json config #2: rays of an 800x800
pinhole camera (camera_angle_x = 0.6911, /root/reference/data/data_parser.py:60-66) placed on a radius-4
sphere at -30 deg elevation with a random azimuth per ray, one random pixel per ray, a point light on a
radius-4.5 sphere, near/far from the unit sphere (/root/reference/camera/ray_generator.py:133-139)."""
import math

import numpy as np
import torch


def synthetic_rays(R: int, seed: int = 3407, crop: int = 800):
    rng = np.random.default_rng(seed)
    H = 800
    fx = 0.5 * H / math.tan(0.5 * 0.6911)
    theta = rng.uniform(-180.0, 180.0, R) / 180.0 * math.pi
    phi = -30.0 / 180.0 * math.pi
    lo, hi = (H - crop) // 2, (H + crop) // 2
    hh = rng.integers(lo, hi, R).astype(np.float64) + 0.5
    ww = rng.integers(lo, hi, R).astype(np.float64) + 0.5
    dirs = np.stack([(ww - 400.0) / fx, -(hh - 400.0) / fx, -np.ones(R)], -1)
    cp, sp = math.cos(phi), math.sin(phi)
    cam = np.stack([4.0 * cp * np.sin(theta), np.full(R, -4.0 * sp), 4.0 * cp * np.cos(theta)], -1)
    fwd = -cam / np.linalg.norm(cam, axis=-1, keepdims=True)
    up = np.tile(np.array([0.0, 1.0, 0.0]), (R, 1))
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right, axis=-1, keepdims=True)
    upv = np.cross(right, fwd)
    d = dirs[:, :1] * right + dirs[:, 1:2] * upv + (-dirs[:, 2:3]) * fwd
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    pl = rng.normal(size=(R, 3))
    pl = 4.5 * pl / np.linalg.norm(pl, axis=-1, keepdims=True)
    o32 = torch.tensor(cam, dtype=torch.float32)
    d32 = torch.nn.functional.normalize(torch.tensor(d, dtype=torch.float32), dim=-1, p=2)
    a = torch.sum(d32 ** 2, dim=-1, keepdim=True)
    b = 2.0 * torch.sum(o32 * d32, dim=-1, keepdim=True)
    mid = 0.5 * (-b) / a
    return {"origins": o32, "directions": d32, "pl_positions": torch.tensor(pl, dtype=torch.float32),
            "nears": mid - 1.0, "fars": mid + 1.0}


def shard_rays(rays: dict, rank: int, world: int) -> dict:
    """Contiguous ray slice of one rank (the reference's per-rank batch = batch_size // world_size,
    /root/reference/trainer/trainer.py:118).  Rays are independent: no data-path collective is needed."""
    R = rays["origins"].shape[0]
    per = R // world
    lo = rank * per
    hi = R if rank == world - 1 else lo + per
    return {k: v[lo:hi] for k, v in rays.items()}


def synthetic_pixel_bundle(R: int, seed: int = 3407, n_cameras: int = 64, H: int = 800):
    """The same synthetic scene as synthetic_rays, one level up: a RawPixelBundle-like batch (pixel indices, per-ray
    camera-to-world pose, light position, image index, ground-truth colour; /root/reference/data/data_loader.py:80-89) of an
    ALL_IMAGES sampler (trainer/trainer.py:118-125) over `n_cameras` views on the radius-4 sphere at -30 deg elevation
    (camera/video_pose_utils.py:28-34 style poses), plus the camera intrinsics.  Returns (fields dict, camera dict)."""
    from types import SimpleNamespace
    rng = np.random.default_rng(seed)
    fx = 0.5 * H / math.tan(0.5 * 0.6911)
    theta = rng.uniform(-180.0, 180.0, n_cameras) / 180.0 * math.pi
    phi = -30.0 / 180.0 * math.pi
    cp, sp = math.cos(phi), math.sin(phi)
    cam = np.stack([4.0 * cp * np.sin(theta), np.full(n_cameras, -4.0 * sp), 4.0 * cp * np.cos(theta)], -1)
    fwd = -cam / np.linalg.norm(cam, axis=-1, keepdims=True)
    up = np.tile(np.array([0.0, 1.0, 0.0]), (n_cameras, 1))
    right = np.cross(fwd, up); right /= np.linalg.norm(right, axis=-1, keepdims=True)
    upv = np.cross(right, fwd)
    c2w = np.tile(np.eye(4), (n_cameras, 1, 1))
    c2w[:, :3, 0], c2w[:, :3, 1], c2w[:, :3, 2], c2w[:, :3, 3] = right, upv, -fwd, cam
    pl_cam = rng.normal(size=(n_cameras, 3)); pl_cam = 4.5 * pl_cam / np.linalg.norm(pl_cam, axis=-1, keepdims=True)
    img = rng.integers(0, n_cameras, R)
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)        # noqa: E731
    fields = dict(img_indices=torch.tensor(img, dtype=torch.int64)[:, None], h_indices=f32(rng.integers(0, H, R))[:, None],
                  w_indices=f32(rng.integers(0, H, R))[:, None], poses=f32(c2w[img]), pls=f32(pl_cam[img]),
                  rgb_gt=f32(rng.uniform(0.0, 1.0, (R, 3))))
    camera = dict(H=H, W=H, cx=H / 2.0, cy=H / 2.0, fx=fx, fy=fx, zn=2.0, zf=6.0)
    return SimpleNamespace(**fields), camera
