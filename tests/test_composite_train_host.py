"""csrc/composite_train_math.cuh (what k_composite_train_fwd / _bwd execute per ray) compiled for the host, against the torch
expression of get_alpha / weights / compositing (models/neus_hint_model.py:339-356,:521-526,:635-637) and its autograd in fp64."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

HERE = Path(__file__).resolve().parent
FP = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def hlib():
    out = HERE / "_build" / "libnrh_hostcheck.so"
    out.parent.mkdir(exist_ok=True)
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-o", str(out), str(HERE / "host_harness.cpp")], check=True)
    lib = C.CDLL(str(out))
    lib.h_composite_train_backward.restype = C.c_float
    return lib


def _p(a):
    return a.ctypes.data_as(FP) if a is not None else None


def torch_composite(sdf, g, c, dist, d, inv_s, r, bg):
    """One ray, the expressions of nrhints_b200/autograd_fine.render_fine (themselves pinned to the reference by the gradient fixture)."""
    tc = (d[None, :] * g).sum(-1)
    it = -(F.relu(-tc * 0.5 + 0.5) * (1.0 - r) + F.relu(-tc) * r)
    half = it * dist * 0.5
    p, n = torch.sigmoid((sdf - half) * inv_s), torch.sigmoid((sdf + half) * inv_s)
    alpha = ((p - n + 1e-5) / (p + 1e-5)).clip(0, 1)
    T = torch.cumprod(torch.cat([torch.ones(1, dtype=sdf.dtype), 1.0 - alpha + 1e-7]), 0)[:-1]
    w = alpha * T
    rgb = (c * w[:, None]).sum(0)
    if bg is not None:
        rgb = rgb + bg * (1.0 - w.sum())
    return w, rgb


@pytest.mark.parametrize("S,r,with_bg,inv_s,seed", [(32, 1.0, True, 20.0, 0), (128, 0.5, True, 64.0, 1), (16, 0.0, False, 403.0, 2),
                                                     (128, 1.0, True, 1097.0, 3), (1, 1.0, True, 20.0, 4)])
def test_forward_and_backward_match_autograd(hlib, S, r, with_bg, inv_s, seed):
    g_ = torch.Generator().manual_seed(seed)
    z = torch.sort(torch.rand(S, generator=g_) * 2.0).values
    sdf = (0.6 - z) * 0.8 + 0.02 * torch.randn(S, generator=g_)          # crosses zero inside the ray: a surface
    d = F.normalize(torch.randn(3, generator=g_), dim=0)
    g = -d[None, :] * (0.8 + 0.1 * torch.randn(S, 1, generator=g_)) + 0.2 * torch.randn(S, 3, generator=g_)
    c = torch.rand(S, 3, generator=g_)
    dist = torch.rand(S, generator=g_) * 0.03 + 0.002
    bg = torch.rand(3, generator=g_) if with_bg else None
    d_rgb, d_w = torch.randn(3, generator=g_), torch.randn(S, generator=g_) * 0.3

    leaves = [t.double().requires_grad_(True) for t in (sdf, g, c, d)]
    s64 = torch.tensor(inv_s, dtype=torch.float64, requires_grad=True)
    w64, rgb64 = torch_composite(leaves[0], leaves[1], leaves[2], dist.double(), leaves[3], s64, r, bg.double() if with_bg else None)
    loss = (rgb64 * d_rgb.double()).sum() + (w64 * d_w.double()).sum()
    want = torch.autograd.grad(loss, leaves + [s64])

    f = lambda t: np.ascontiguousarray(t.numpy(), dtype=np.float32)          # noqa: E731
    a = dict(sdf=f(sdf), g=f(g), c=f(c), dist=f(dist), d=f(d), bg=f(bg) if with_bg else None, d_rgb=f(d_rgb), d_w=f(d_w))
    w, rgb = np.zeros(S, np.float32), np.zeros(3, np.float32)
    hlib.h_composite_train_forward(S, _p(a["sdf"]), _p(a["g"]), _p(a["c"]), _p(a["dist"]), _p(a["d"]), C.c_float(inv_s), C.c_float(r),
                                   _p(a["bg"]), _p(w), _p(rgb))
    np.testing.assert_allclose(w, w64.detach().numpy(), rtol=0, atol=3e-5)
    np.testing.assert_allclose(rgb, rgb64.detach().numpy(), rtol=0, atol=3e-5)
    d_sdf, d_g, d_c, d_dir = np.zeros(S, np.float32), np.zeros((S, 3), np.float32), np.zeros((S, 3), np.float32), np.zeros(3, np.float32)
    d_s = hlib.h_composite_train_backward(S, _p(a["sdf"]), _p(a["g"]), _p(a["c"]), _p(a["dist"]), _p(a["d"]), C.c_float(inv_s), C.c_float(r),
                                          _p(a["bg"]), _p(a["d_rgb"]), _p(a["d_w"]), _p(d_sdf), _p(d_g), _p(d_c), _p(d_dir))
    for got, ref, name in ((d_sdf, want[0], "d_sdf"), (d_g, want[1], "d_grad"), (d_c, want[2], "d_color"), (d_dir, want[3], "d_dir"),
                           (np.float32(d_s), want[4], "d_inv_s")):
        ref = ref.numpy()
        scale = max(float(np.abs(ref).max()), 1e-6)
        assert float(np.abs(np.asarray(got, np.float64) - ref).max()) < 2e-3 * scale + 1e-7, (name, got, ref)
