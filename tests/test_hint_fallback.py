"""nrhints_b200/hint_fallback.py (the torch route of the non-default `n_shadow_importance_clip > 0` option) against the UNMODIFIED
reference on the CPU: the march / grouping / shading arithmetic is fed the reference's own networks through the two callables the
module is written against, so that everything it adds is compared without a GPU.  (On the GPU the callables are nrh_sdf_query and the
product's reflectance network, both covered by tests/test_gpu_parity.py.)"""
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "baseline"))
import ref_loader  # noqa: E402

import nrh_testlib as T  # noqa: E402
from nrhints_b200 import hint_fallback as hf  # noqa: E402
from oracle import nrh_oracle as orc  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="baseline/_ref is not installed")

RENDERER = dict(n_samples=32, n_importance_samples=32, n_shadow_samples=16, n_shadow_importance_samples=16)


def _reference(clip, weights="sharp", **extra):
    ns = ref_loader.load_pipeline()
    M = ns.model
    cfg = M.NeuSModelConfig(renderer=M.NeuSRendererConfig(n_shadow_importance_clip=clip, **RENDERER, **extra))
    m = M.NeuSHintRenderer(cfg)
    import nrhints_b200 as nb
    m.load_state_dict(T.make_state(weights, nb.NeuSModelConfig(renderer=nb.NeuSRendererConfig(**RENDERER, **extra))), strict=True)
    return ns, m, cfg


def _callables(m):
    def sdf_fn(pts, want_grad=False, want_feat=False):
        with torch.no_grad():
            out = m.sdf_network(pts)
        grad = m.sdf_network.gradient(pts).reshape(-1, 3).detach() if want_grad else None
        return out[:, 0], grad, (out[:, 1:] if want_feat else None)

    def color_fn(*a):
        with torch.no_grad():
            return m.color_network(*a)
    return sdf_fn, color_fn


@pytest.mark.parametrize("weights", ["init", "sharp"])
def test_shadow_visibility_matches_reference_get_visibility(weights):
    ns, m, cfg = _reference(-1, weights)
    sdf_fn, _ = _callables(m)
    g = torch.Generator().manual_seed(5)
    pls = 4.5 * torch.nn.functional.normalize(torch.randn(40, 3, generator=g), dim=-1)
    targets = (torch.rand(40, 3, generator=g) - 0.5) * 1.2
    want = m.get_visibility(pls, targets, offset=cfg.renderer.shadow_ray_offset)
    inv_s = m.deviation_network(torch.zeros(1, 3))[:, :1].clip(1e-6, 1e6).detach().reshape(())
    got = hf.shadow_visibility(sdf_fn, pls, targets, 16, 16, inv_s, cfg.renderer.shadow_ray_offset)
    assert got.shape == want.shape == (40, 1)
    assert float((got - want).abs().max()) < 2e-6, float((got - want).abs().max())
    assert 0.05 < float(want.mean()) < 0.999                      # the marches see both lit and shadowed targets


@pytest.mark.parametrize("clip,normal_type", [(4, "normalized"), (16, "analytic")])
def test_grouped_visibility_render_matches_reference(clip, normal_type):
    import nrhints_b200 as nb
    extra = {}
    ns, m, cfg = _reference(clip, "sharp")
    if normal_type == "analytic":
        extra = dict(normal_type=ns.model.NormalComputationType.Analytic)
        ns, m, cfg = _reference(clip, "sharp", **extra)
    R = 24
    rays = orc.synthetic_rays(R, seed=17, crop=300)
    from camera.ray_utils import RayBundle
    bg = torch.ones(1, 3)
    with torch.no_grad():
        want = m(RayBundle(**rays), background_rgb=bg)
    # everything that does not depend on the visibility, from the oracle (pinned to the reference by tests/test_oracle_golden.py)
    pcfg = nb.NeuSModelConfig(renderer=nb.NeuSRendererConfig(**RENDERER, **({"normal_type": nb.NormalComputationType.Analytic} if extra else {})))
    sd = T.make_state("sharp", pcfg)
    with torch.no_grad():
        base = orc.render_forward(sd, orc.OracleConfig.from_model_config(pcfg), rays["origins"], rays["directions"], rays["pl_positions"],
                                  rays["nears"], rays["fars"], background_rgb=bg)
    assert float((base["weights"] - want.weights).abs().max()) < 5e-4
    sdf_fn, color_fn = _callables(m)
    inv_s = m.deviation_network(torch.zeros(1, 3))[:, :1].clip(1e-6, 1e6).detach().reshape(())
    r = cfg.renderer
    vis, shadow_map = hf.grouped_visibility(sdf_fn, rays["origins"], rays["directions"], rays["pl_positions"], base["z_vals"], base["weights"],
                                            clip, r.n_shadow_samples, r.n_shadow_importance_samples, inv_s, r.shadow_ray_offset,
                                            chunk=cfg.shadow_mini_chunk_size)
    assert vis.shape == (R, 64, 1) and shadow_map.shape == (R, 1)
    assert torch.equal(vis[:, 0], vis[:, 64 // clip - 1])         # constant inside a group
    normals = base["analytic_normals"] if extra else base["normalized_analytic_normals"]
    rgb, colour = hf.shade_samples(sdf_fn, color_fn, rays["origins"], rays["directions"], rays["pl_positions"], base["z_vals"], base["weights"],
                                   normals, vis, base["specular_cue"], 2.0 / r.n_samples, bg)
    assert colour.shape == (R, 64, 3)
    # the oracle's sample positions differ from the reference's by fp32 noise (far-end importance samples): the shadow map of a ray
    # whose largest weight sits at a group boundary may come from the neighbouring group
    d_map = (shadow_map - want.visibilities).abs().reshape(-1)
    assert float(d_map.median()) < 1e-5 and float((d_map > 1e-3).float().mean()) <= 0.1, d_map
    assert float((rgb - want.rgb).abs().max()) < 1e-3, float((rgb - want.rgb).abs().max())
    assert float((rgb - want.rgb).abs().median()) < 2e-5


def test_grouping_needs_a_divisor():
    with pytest.raises(ValueError):
        hf.grouped_visibility(None, torch.zeros(2, 3), torch.zeros(2, 3), torch.zeros(2, 3), torch.zeros(2, 64), torch.zeros(2, 64), 5,
                              16, 16, torch.tensor(1.0), 0.01)
