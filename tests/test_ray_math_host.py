"""nrhints_b200/csrc/ray_math.cuh compiled for the host (tests/host_harness.cpp) against the oracle:
the sampler / compositor functions the CUDA kernels call, checked without a GPU."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

import nrh_testlib as T
from oracle import nrh_oracle as orc

HERE = Path(__file__).resolve().parent
F = C.POINTER(C.c_float)


def _p(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(F)


@pytest.fixture(scope="module")
def hlib():
    out = HERE / "_build" / "libnrh_hostcheck.so"
    out.parent.mkdir(exist_ok=True)
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-o", str(out), str(HERE / "host_harness.cpp")], check=True)
    lib = C.CDLL(str(out))
    lib.h_shadow_init.restype = C.c_float
    lib.h_shadow_transmittance.restype = C.c_float
    return lib


def f32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


@pytest.mark.parametrize("n", [2, 3, 15, 16, 32, 64])
def test_linspace_bitwise(hlib, n):
    out = np.empty(n, np.float32)
    hlib.h_linspace01(n, _p(out))
    np.testing.assert_array_equal(out, torch.linspace(0.0, 1.0, n).numpy())


def test_coarse_z(hlib):
    near, far = np.float32(2.9), np.float32(4.9)
    z = np.empty(64, np.float32)
    hlib.h_coarse_z(C.c_float(near), C.c_float(far), 64, 0, C.c_float(0), _p(z))
    want = torch.tensor([[near]]) + (torch.tensor([[far]]) - torch.tensor([[near]])) * torch.linspace(0.0, 1.0, 64)[None, :]
    np.testing.assert_allclose(z, want[0].numpy(), rtol=0, atol=5e-7)
    hlib.h_coarse_z(C.c_float(near), C.c_float(far), 64, 1, C.c_float(0.8), _p(z))
    want2 = want + (torch.tensor([[0.8]]) - 0.5) * 2.0 / 64
    np.testing.assert_allclose(z, want2[0].numpy(), rtol=0, atol=5e-7)


@pytest.mark.parametrize("D,Fq", [(3, 6), (3, 4), (1, 4), (4, 4)])
def test_fourier_encode(hlib, D, Fq):
    g = torch.Generator().manual_seed(D * 10 + Fq)
    x = (torch.rand(D, generator=g) - 0.5) * 27.0            # up to 13.5 * 32 = 432 rad at the top frequency
    out = np.empty(D * (2 * Fq + 1), np.float32)
    hlib.h_fourier_encode(_p(f32(x.numpy())), D, Fq, _p(out))
    want = orc.fourier_encode(x[None], Fq)[0].numpy()
    np.testing.assert_allclose(out, want, rtol=0, atol=2e-6)


def _march_inputs(seed, k):
    """a plausible sorted z / sdf pair: a ray through the init sphere"""
    g = torch.Generator().manual_seed(seed)
    o = torch.tensor([[0.2, -0.1, 3.9]])
    d = torch.nn.functional.normalize(torch.tensor([[-0.05, 0.03, -1.0]]), dim=-1)
    mid = -(o * d).sum(-1, keepdim=True)
    z, _ = torch.sort(mid - 1.0 + 2.0 * torch.rand(1, k, generator=g), dim=-1)
    pts = o[:, None, :] + d[:, None, :] * z[..., None]
    sdf = torch.linalg.norm(pts, dim=-1) - 0.5 + 0.01 * torch.randn(1, k, generator=g)
    return o, d, z, sdf


@pytest.mark.parametrize("k,inv_s", [(64, 64.0), (80, 128.0), (96, 256.0), (112, 512.0), (16, 64.0)])
def test_upsample_step(hlib, k, inv_s):
    o, d, z, sdf = _march_inputs(k, k)
    n_new = 16
    want = orc.up_sample(o, d, z, sdf, n_new, inv_s)[0].numpy()
    wbuf = np.empty(k, np.float32)
    got = np.empty(n_new, np.float32)
    hlib.h_upsample(_p(f32(o[0])), _p(f32(d[0])), k, _p(f32(z[0])), _p(f32(sdf[0])), C.c_float(inv_s), n_new, _p(wbuf), _p(got))
    # the far-end sample may legitimately sit one bin earlier (cumsum rounding, see compare_outputs)
    bad = np.abs(got - want) > 2e-5
    assert bad.sum() <= 1 and not bad[:-1].any(), (got, want)
    assert np.all(np.diff(got) >= 0)


def test_merge(hlib):
    o, d, z, sdf = _march_inputs(3, 64)
    g = torch.Generator().manual_seed(4)
    zn, _ = torch.sort(z[0, 0] + (z[0, -1] - z[0, 0]) * torch.rand(16, generator=g))
    sn = torch.randn(16, generator=g)
    zo, so = np.empty(80, np.float32), np.empty(80, np.float32)
    hlib.h_merge(64, _p(f32(z[0])), _p(f32(sdf[0])), 16, _p(f32(zn)), _p(f32(sn)), _p(zo), _p(so), 1)
    zc, idx = torch.sort(torch.cat([z[0], zn]))
    np.testing.assert_array_equal(zo, zc.numpy())
    np.testing.assert_array_equal(so, torch.cat([sdf[0], sn])[idx].numpy())


def test_merge_backward_equals_forward(hlib):
    """the in-place backward merge used by the shared-memory importance kernel == the forward merge, ties included"""
    g = torch.Generator().manual_seed(12)
    for trial in range(20):
        k, n = 64 + 16 * (trial % 4), 16
        zo, _ = torch.sort(torch.rand(k, generator=g))
        zn, _ = torch.sort(torch.rand(n, generator=g))
        if trial % 2:                       # force ties
            zn[3] = zo[10]; zn[4] = zo[10]; zn[-1] = zo[-1]; zn[0] = zo[0]
            zn, _ = torch.sort(zn)
        so, sn = torch.randn(k, generator=g), torch.randn(n, generator=g)
        zf, sf = np.empty(k + n, np.float32), np.empty(k + n, np.float32)
        hlib.h_merge(k, _p(f32(zo)), _p(f32(so)), n, _p(f32(zn)), _p(f32(sn)), _p(zf), _p(sf), 1)
        zb = np.concatenate([f32(zo), np.zeros(n, np.float32)]); sb = np.concatenate([f32(so), np.zeros(n, np.float32)])
        hlib.h_merge_backward(k, _p(zb), _p(sb), n, _p(f32(zn)), _p(f32(sn)), 1)
        np.testing.assert_array_equal(zb, zf)
        np.testing.assert_array_equal(sb, sf)


def test_merge_rank_equals_forward(hlib):
    """the rank form of the merge (every entry computes its own slot; k_importance_step runs it on all threads) == the sequential
    stable merge, ties between and inside the runs included"""
    g = torch.Generator().manual_seed(13)
    for trial in range(40):
        k, n = 64 + 16 * (trial % 4), (16 if trial % 3 else 5)
        zo, _ = torch.sort(torch.rand(k, generator=g))
        zn, _ = torch.sort(torch.rand(n, generator=g))
        if trial % 2:                       # force ties: new == old, duplicates inside both runs, both ends
            zn[1] = zo[10]; zn[2] = zo[10]; zn[-1] = zo[-1]; zn[0] = zo[0]; zo[20] = zo[21]
            zn, _ = torch.sort(zn); zo, _ = torch.sort(zo)
        if trial % 5 == 0:                  # all new entries below / above every old one
            zn = zn * 0.0 - 1.0 if trial % 10 == 0 else zn * 0.0 + 2.0
        so, sn = torch.randn(k, generator=g), torch.randn(n, generator=g)
        zf, sf = np.empty(k + n, np.float32), np.empty(k + n, np.float32)
        hlib.h_merge(k, _p(f32(zo)), _p(f32(so)), n, _p(f32(zn)), _p(f32(sn)), _p(zf), _p(sf), 1)
        zr, sr = np.full(k + n, np.nan, np.float32), np.full(k + n, np.nan, np.float32)
        hlib.h_merge_rank(k, _p(f32(zo)), _p(f32(so)), n, _p(f32(zn)), _p(f32(sn)), _p(zr), _p(sr))
        np.testing.assert_array_equal(zr, zf)
        np.testing.assert_array_equal(sr, sf)


@pytest.mark.parametrize("cos_anneal", [1.0, 0.5])
def test_composite_primary(hlib, cos_anneal):
    S = 128
    o, d, z, sdf = _march_inputs(7, S)
    g = torch.Generator().manual_seed(8)
    last = 2.0 / 64
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full((1, 1), last)], -1)
    mid = z + dists * 0.5
    pts = (o[:, None, :] + d[:, None, :] * mid[..., None]).reshape(-1, 3)
    grad = torch.nn.functional.normalize(pts, dim=-1) * (1.0 + 0.1 * torch.randn(S, 1, generator=g))
    inv_s = 403.0
    sdf_mid = (torch.linalg.norm(pts, dim=-1) - 0.5)
    # oracle-side alpha / weights
    true_cos = (d.expand(S, 3) * grad).sum(-1, keepdim=True)
    iter_cos = -(torch.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal) + torch.relu(-true_cos) * cos_anneal)
    e_next = sdf_mid[:, None] + iter_cos * dists.reshape(-1, 1) * 0.5
    e_prev = sdf_mid[:, None] - iter_cos * dists.reshape(-1, 1) * 0.5
    pc, nc = torch.sigmoid(e_prev * inv_s), torch.sigmoid(e_next * inv_s)
    alpha = ((pc - nc + 1e-5) / (pc + 1e-5)).clip(0, 1).reshape(1, S)
    w = alpha * orc._excl_cumprod(1.0 - alpha + 1e-7)
    depth = (mid * w).sum()
    nhat = torch.nn.functional.normalize(grad, dim=-1)
    wout, inside = np.empty(S, np.float32), np.empty(S, np.float32)
    nx, ny, nz = (np.empty(S, np.float32) for _ in range(3))
    res = np.empty(5, np.float32)
    hlib.h_composite_primary(_p(f32(o[0])), _p(f32(d[0])), S, _p(f32(z[0])), C.c_float(last), _p(f32(sdf_mid)),
                             _p(f32(grad[:, 0])), _p(f32(grad[:, 1])), _p(f32(grad[:, 2])), C.c_float(inv_s),
                             C.c_float(cos_anneal), _p(wout), _p(inside), _p(nx), _p(ny), _p(nz), _p(res))
    np.testing.assert_allclose(wout, w[0].numpy(), atol=2e-6)
    np.testing.assert_allclose(res[0], w.sum().item(), atol=2e-6)
    np.testing.assert_allclose(res[1], depth.item(), atol=1e-5)
    np.testing.assert_allclose(np.stack([nx, ny, nz], -1), nhat.numpy(), atol=1e-6)
    np.testing.assert_array_equal(inside, (torch.linalg.norm(pts, dim=-1) < 1.0).float().numpy())
    np.testing.assert_allclose(res[2:], (nhat * w[0, :, None]).sum(0).numpy(), atol=2e-6)


@pytest.mark.parametrize("jitter", [False, True])
def test_shadow_ray_init(hlib, jitter):
    pl = torch.tensor([[1.0, 4.2, -1.2]])
    hit = torch.tensor([[0.1, 0.3, -0.35]])
    n, off = 64, 1e-2
    g = torch.Generator().manual_seed(3)
    jit = torch.rand(1, n, generator=g)
    dvec = hit - pl
    L = torch.linalg.norm(dvec, dim=-1, keepdim=True)
    z = torch.linspace(0.0, 1.0, n) * L * (1.0 - off)
    if jitter:
        mids = 0.5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * jit
    dir_out, zo = np.empty(3, np.float32), np.empty(n, np.float32)
    Lg = hlib.h_shadow_init(_p(f32(pl[0])), _p(f32(hit[0])), n, C.c_float(off), int(jitter), _p(f32(jit[0])), _p(dir_out), _p(zo))
    np.testing.assert_allclose(Lg, L.item(), rtol=1e-6)
    np.testing.assert_allclose(dir_out, (dvec / L)[0].numpy(), atol=1e-6)
    np.testing.assert_allclose(zo, z[0].numpy(), atol=2e-6)


def test_specular_cue(hlib):
    g = torch.Generator().manual_seed(5)
    for _ in range(20):
        hn = torch.nn.functional.normalize(torch.randn(1, 3, generator=g), dim=-1)
        pl = 4.5 * torch.nn.functional.normalize(torch.randn(1, 3, generator=g), dim=-1)
        hit = 0.5 * torch.nn.functional.normalize(torch.randn(1, 3, generator=g), dim=-1)
        d = torch.nn.functional.normalize(torch.randn(1, 3, generator=g), dim=-1)
        cfg = orc.OracleConfig()
        want = orc.specular_cue(cfg, hn, pl, hit, d)[0].numpy()
        rough = f32(cfg.specular_roughness)
        cue = np.empty(4, np.float32)
        hlib.h_specular_cue(_p(f32(hn[0])), _p(f32(pl[0])), _p(f32(hit[0])), _p(f32(d[0])), 4, _p(rough), _p(cue))
        np.testing.assert_allclose(cue, want, rtol=2e-5, atol=1e-7)


# ---- outside (NeRF++) model: models/neus_hint_model.py:677-694 (sample positions), :434-473 (render_outside) ----------
@pytest.mark.parametrize("jitter", [False, True])
@pytest.mark.parametrize("n_out", [16, 32])
def test_outside_z_and_sections(hlib, jitter, n_out):
    hlib.h_outside_alpha.restype = C.c_float
    cfg = orc.OracleConfig(n_samples=64, n_outside_samples=n_out, use_outside_nerf=True)
    far = torch.tensor([[4.93]])
    g = torch.Generator().manual_seed(n_out)
    jit = torch.rand(1, n_out, generator=g)
    want = orc.outside_z(cfg, far, jit if jitter else None, torch.float32)
    zo = np.empty(n_out, np.float32)
    hlib.h_outside_z(C.c_float(far.item()), 64, n_out, int(jitter), _p(f32(jit[0])), _p(zo))
    np.testing.assert_allclose(zo, want.reshape(-1).numpy(), rtol=3e-7)
    assert np.all(np.diff(zo) > 0)
    # merged sections against torch.sort(cat(...))
    S = 40
    z = torch.sort(2.9 + 2.0 * torch.rand(1, S, generator=g), dim=-1)[0]
    z_feed, _ = torch.sort(torch.cat([z, torch.from_numpy(zo)[None, :]], -1), -1)
    sd = 2.0 / 64
    dists = torch.cat([z_feed[:, 1:] - z_feed[:, :-1], torch.full((1, 1), sd)], -1)
    mids = z_feed + dists * 0.5
    dist, mid = np.empty(S + n_out, np.float32), np.empty(S + n_out, np.float32)
    hlib.h_outside_sections(S, _p(f32(z[0])), n_out, _p(zo), C.c_float(sd), _p(dist), _p(mid))
    np.testing.assert_array_equal(dist, dists[0].numpy())
    np.testing.assert_array_equal(mid, mids[0].numpy())


def test_outside_point_and_alpha(hlib):
    hlib.h_outside_alpha.restype = C.c_float
    o = torch.tensor([[0.3, 3.2, -2.1]]); d = torch.nn.functional.normalize(torch.tensor([[-0.1, -0.8, 0.55]]), dim=-1)
    for mid in (0.2, 3.0, 4.7, 40.0, 4000.0):
        pts = o + d * mid
        dis = torch.linalg.norm(pts, dim=-1, keepdim=True).clip(1.0, 1e10)
        want = torch.cat([pts / dis, 1.0 / dis], -1)[0].numpy()
        p4 = np.empty(4, np.float32)
        hlib.h_outside_point(_p(f32(o[0])), _p(f32(d[0])), C.c_float(mid), _p(p4))
        np.testing.assert_allclose(p4, want, rtol=1e-6, atol=1e-7)
    for dens, dist in ((-3.0, 0.1), (0.0, 0.03), (2.5, 1.7), (25.0, 0.2), (-30.0, 5.0)):
        want = float(1.0 - torch.exp(-torch.nn.functional.softplus(torch.tensor(dens)) * dist))
        got = hlib.h_outside_alpha(C.c_float(dens), C.c_float(dist))
        assert abs(got - want) < 2e-7, (dens, dist, got, want)


def test_composite_primary_with_background(hlib):
    """alpha blending with the background alpha outside the unit sphere + appended far samples (render_core :517-525)."""
    g = torch.Generator().manual_seed(8)
    S, n_out = 24, 8
    o = torch.tensor([[0.0, 0.0, -2.0]]); d = torch.tensor([[0.05, 0.0, 1.0]]); d = d / d.norm()
    z = torch.sort(0.6 + 2.4 * torch.rand(1, S, generator=g), -1)[0]             # some mid-points outside the unit sphere
    last, inv_s, ca = 2.0 / 64, 55.0, 1.0
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full((1, 1), last)], -1)
    mid = z + dists * 0.5
    pts = o + d * mid[0, :, None]
    grad = torch.nn.functional.normalize(pts, dim=-1)
    sdf_mid = torch.linalg.norm(pts, dim=-1) - 0.5
    true_cos = (d.expand(S, 3) * grad).sum(-1, keepdim=True)
    iter_cos = -(torch.relu(-true_cos * 0.5 + 0.5) * (1.0 - ca) + torch.relu(-true_cos) * ca)
    pc = torch.sigmoid((sdf_mid[:, None] - iter_cos * dists.reshape(-1, 1) * 0.5) * inv_s)
    nc = torch.sigmoid((sdf_mid[:, None] + iter_cos * dists.reshape(-1, 1) * 0.5) * inv_s)
    alpha = ((pc - nc + 1e-5) / (pc + 1e-5)).clip(0, 1).reshape(1, S)
    inside = (torch.linalg.norm(pts, dim=-1) < 1.0).float()[None, :]
    assert 0 < inside.sum() < S
    dens = torch.randn(1, S + n_out, generator=g) * 2.0
    bdist = 0.02 + torch.rand(1, S + n_out, generator=g)
    bg_alpha = 1.0 - torch.exp(-torch.nn.functional.softplus(dens) * bdist)
    a = torch.cat([alpha * inside + bg_alpha[:, :S] * (1.0 - inside), bg_alpha[:, S:]], -1)
    w = a * orc._excl_cumprod(1.0 - a + 1e-7)
    wout = np.empty(S + n_out, np.float32); ins = np.empty(S, np.float32)
    nx, ny, nz = (np.empty(S, np.float32) for _ in range(3))
    res = np.empty(5, np.float32)
    hlib.h_composite_primary_bg(_p(f32(o[0])), _p(f32(d[0])), S, _p(f32(z[0])), C.c_float(last), _p(f32(sdf_mid)),
                                _p(f32(grad[:, 0])), _p(f32(grad[:, 1])), _p(f32(grad[:, 2])), C.c_float(inv_s), C.c_float(ca),
                                _p(wout), _p(ins), _p(nx), _p(ny), _p(nz), _p(res), n_out, _p(f32(dens[0])), _p(f32(bdist[0])))
    np.testing.assert_allclose(wout, w[0].numpy(), atol=2e-6)
    np.testing.assert_allclose(res[0], w.sum().item(), atol=3e-6)
    np.testing.assert_allclose(res[1], (mid * w[:, :S]).sum().item(), atol=1e-5)
    np.testing.assert_array_equal(ins, inside[0].numpy())
