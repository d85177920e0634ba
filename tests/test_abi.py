"""The C-ABI library loads and exports every symbol include/nrhints_b200.h declares (no compute calls,
no GPU needed), and the ctypes structs agree with the header."""
import ctypes as C
import re
from pathlib import Path

import pytest

import nrh_testlib as T
from nrhints_b200 import _lib

HEADER = (T.ROOT / "include" / "nrhints_b200.h").read_text()


def declared_functions():
    body = HEADER[HEADER.index("int nrh_version"):]
    return sorted(set(re.findall(r"\b(nrh_[a-z_0-9]+)\s*\(", body)))


def test_library_builds_and_exports_every_declared_symbol():
    path = _lib.build()
    lib = C.CDLL(str(path))
    names = declared_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nrhints_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "ctypes binding table out of sync with the header"


def test_version_and_config_validation_without_gpu():
    lib = _lib.load()
    assert lib.nrh_version() == _lib.NRH_ABI_VERSION == 7
    import nrhints_b200 as nb
    m = nb.NeuSHintRenderer(nb.NeuSModelConfig())
    cfg = m._c_config()
    assert lib.nrh_check_config(C.byref(cfg)) == 0
    assert lib.nrh_packed_weights_bytes(C.byref(cfg)) > 820923 * 4
    assert lib.nrh_workspace_bytes(C.byref(cfg), 4096) > 4096 * 128 * 256 * 4
    cfg.n_importance = 63                 # not a multiple of up_sample_steps
    assert lib.nrh_check_config(C.byref(cfg)) == -2
    assert b"multiple" in lib.nrh_last_error()
    cfg = m._c_config(); cfg.n_samples = 100; cfg.n_importance = 64
    assert lib.nrh_check_config(C.byref(cfg)) == -2


def test_struct_layout_matches_header():
    assert C.sizeof(_lib.NrhConfig) == 8 * 4 + 4 * 4 + 4 + 3 * 4 + 2 * 4
    assert C.sizeof(_lib.NrhRawWeights) == (8 + 8 + 4 + 5 + 5 + 1 + 8 + 8 + 8) * 8
    assert _lib.NRH_ABI_VERSION == int(re.search(r"#define NRH_ABI_VERSION (\d+)", HEADER).group(1))
    assert C.sizeof(_lib.NrhRays) == 7 * 8
    assert C.sizeof(_lib.NrhOutputs) == 19 * 8
    assert C.sizeof(_lib.NrhTrainCapture) == 8 * 8 and C.sizeof(_lib.NrhRayGenInputs) == 10 * 8 and C.sizeof(_lib.NrhCamera) == 6 * 4
    fields = re.findall(r"(?:float|void|const NrhTrainCapture)\*\s+(\w+);", HEADER[HEADER.index("typedef struct NrhOutputs"):HEADER.index("} NrhOutputs;")])
    assert fields == [n for n, _ in _lib.NrhOutputs._fields_]


def test_cpu_tensors_are_rejected_loudly():
    import torch
    import nrhints_b200 as nb
    from oracle import nrh_oracle as orc
    m = nb.NeuSHintRenderer(nb.NeuSModelConfig())
    rays = orc.synthetic_rays(4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(nb.RayBundle(**rays))


def test_unsupported_configs_raise():
    import nrhints_b200 as nb
    with pytest.raises(NotImplementedError):
        nb.NeuSHintRenderer(nb.NeuSModelConfig(renderer=nb.NeuSRendererConfig(use_outside_nerf=True, n_outside_samples=100)))
    with pytest.raises(NotImplementedError):
        nb.NeuSHintRenderer(nb.NeuSModelConfig(renderer=nb.NeuSRendererConfig(use_outside_nerf=True),
                                               outside_nerf=nb.NeRFConfig(multi_res=6)))
    with pytest.raises(NotImplementedError):
        nb.NeuSHintRenderer(nb.NeuSModelConfig(sdf_network=nb.SDFNetConfig(d_hidden=128)))
    with pytest.raises(NotImplementedError):                   # must divide the 128 samples of a ray (hint_fallback.py serves the rest)
        nb.NeuSHintRenderer(nb.NeuSModelConfig(renderer=nb.NeuSRendererConfig(n_shadow_importance_clip=5)))
    with pytest.raises(NotImplementedError):
        nb.NeuSHintRenderer(nb.NeuSModelConfig(renderer=nb.NeuSRendererConfig(n_shadow_importance_clip=4, shadow_hint=False)))
    nb.NeuSHintRenderer(nb.NeuSModelConfig(renderer=nb.NeuSRendererConfig(n_shadow_importance_clip=4)))


def test_state_dict_layout_matches_reference_inventory():
    """820 923 scalars, key pattern sdf_network.lin{i}.{bias,weight_g,weight_v} ..."""
    import nrhints_b200 as nb
    m = nb.NeuSHintRenderer(nb.NeuSModelConfig())
    sd = m.state_dict()
    assert sum(v.numel() for v in sd.values()) == 820923 and len(sd) == 46
    assert list(sd)[:3] == ["sdf_network.lin0.bias", "sdf_network.lin0.weight_g", "sdf_network.lin0.weight_v"]
    assert tuple(sd["sdf_network.lin3.weight_v"].shape) == (217, 256)
    assert tuple(sd["color_network.lin0.weight_v"].shape) == (256, 361)
    assert tuple(sd["deviation_network.variance"].shape) == ()
    # with the outside NeRF: the reference's `outside_nerf.*` keys (fields/nerf_density_field.py:57-64), registered last
    m2 = nb.NeuSHintRenderer(nb.NeuSModelConfig(renderer=nb.NeuSRendererConfig(use_outside_nerf=True)))
    sd2 = m2.state_dict()
    assert list(sd2)[:46] == list(sd) and len(sd2) == 46 + 24
    assert tuple(sd2["outside_nerf.pts_linears.5.weight"].shape) == (256, 340)
    assert tuple(sd2["outside_nerf.views_linears.0.weight"].shape) == (128, 310)
    assert list(sd2)[-2:] == ["outside_nerf.rgb_linear.weight", "outside_nerf.rgb_linear.bias"]
    cfg = m2._c_config()
    from nrhints_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    assert lib.nrh_check_config(C.byref(cfg)) == 0
    assert lib.nrh_packed_weights_bytes(C.byref(cfg)) > lib.nrh_packed_weights_bytes(C.byref(m._c_config())) + 600000 * 4
