"""Grid SDF query for meshing: NeuSHintRenderer.extract_fields / extract_geometry
(/root/reference/models/neus_hint_model.py:68-93, :753-758) on the CUDA SDF kernel against the oracle and -- when baseline/_ref is
installed -- against the unmodified reference's own extract_fields / extract_geometry driven by its own network on the same GPU:
grid order ('ij', x slowest), sign (u = -sdf), a resolution that is not a multiple of the reference's 64-point blocks, a
non-cubic bounding box, and the vertex transform / threshold hand-off to marching cubes."""
import sys
import types
from pathlib import Path

import numpy as np
import pytest
import torch

import nrh_testlib as T
import nrhints_b200 as nb
from oracle import nrh_oracle as orc

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "baseline"))
import ref_loader  # noqa: E402

pytestmark = pytest.mark.gpu
BMIN, BMAX = torch.tensor([-1.0, -0.9, -1.1]), torch.tensor([1.0, 1.1, 0.8])


def _module(kind="sharp"):
    cfg = nb.NeuSModelConfig()
    sd = T.make_state(kind, cfg)
    m = nb.NeuSHintRenderer(cfg); m.load_state_dict(sd); m.cuda()
    return m, cfg, sd


@pytest.mark.parametrize("res", [64, 70])
def test_extract_fields_matches_oracle(res):
    m, cfg, sd = _module()
    u = m.extract_fields(BMIN, BMAX, res, chunk=50_000)            # several slabs, the last one ragged
    assert u.shape == (res, res, res) and u.dtype == np.float32
    xs = [torch.linspace(float(BMIN[i]), float(BMAX[i]), res) for i in range(3)]
    xx, yy, zz = torch.meshgrid(*xs, indexing="ij")
    pts = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], -1)
    W = orc.effective_weights(sd)
    want = -orc.sdf_mlp(W, pts, orc.OracleConfig.from_model_config(cfg))["sdf"].reshape(res, res, res).numpy()
    assert float(np.abs(u - want).max()) < 1e-4                   # measured 2.5e-5 (tcgen05 engine, see test_sdf_query_matches_oracle)
    # sign / order spot checks: the init-like field is a ~0.5 sphere => u = -sdf is positive at the centre, negative at the corners,
    # and the x index is the slowest axis
    c = res // 2
    assert u[c, c, c] > 0 and u[0, 0, 0] < 0 and u[-1, -1, -1] < 0
    i, j, k = 3, res - 5, res // 3
    p = torch.tensor([[xs[0][i], xs[1][j], xs[2][k]]])
    assert abs(float(u[i, j, k]) + float(orc.sdf_mlp(W, p, orc.OracleConfig.from_model_config(cfg))["sdf"])) < 1e-4


@pytest.mark.skipif(not ref_loader.available(), reason="baseline/_ref is not installed")
def test_extract_geometry_matches_the_reference(monkeypatch):
    """Same grid through the unmodified reference (its 64^3 block walk, its network on CUDA) and through ours; marching cubes is
    replaced on both sides by the same recording stand-in (PyMCubes is not in this image), which checks what is handed to it
    (field, threshold) and the vertex transform applied to what it returns."""
    calls = []

    def marching_cubes(u, threshold):
        calls.append((np.array(u, copy=True), float(threshold)))
        return np.array([[0.0, 0.0, 0.0], [69.0, 69.0, 69.0], [10.0, 20.0, 30.0]]), np.array([[0, 1, 2]])
    fake = types.ModuleType("mcubes"); fake.marching_cubes = marching_cubes
    monkeypatch.setitem(sys.modules, "mcubes", fake)
    ns = ref_loader.load()
    monkeypatch.setattr(ns.model, "mcubes", fake)
    m, cfg, sd = _module()
    torch.manual_seed(3407)
    ref = ns.NeuSHintRenderer(ns.NeuSModelConfig())
    ref.load_state_dict(sd, strict=True)
    ref = ref.cuda()
    res = 70
    with torch.device("cuda"):                       # the reference builds its grid with device-less constructors (trainer default: CUDA)
        v_ref, t_ref = ref.extract_geometry(BMIN.cuda(), BMAX.cuda(), res, threshold=0.02)
    v_new, t_new = m.extract_geometry(BMIN, BMAX, res, threshold=0.02)
    (u_ref, th_ref), (u_new, th_new) = calls
    assert th_ref == th_new == 0.02 and u_ref.shape == u_new.shape == (res, res, res)
    assert float(np.abs(u_ref - u_new).max()) < 1e-4
    assert np.allclose(v_ref, v_new, atol=1e-6) and np.array_equal(t_ref, t_new)
    assert np.allclose(v_new[1], BMAX.numpy(), atol=1e-6) and np.allclose(v_new[0], BMIN.numpy(), atol=1e-6)
