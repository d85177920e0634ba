"""`n_shadow_importance_clip > 0` on the GPU (NeuSHintRenderer._forward_grouped_shadow: fused forward + nrh_sdf_query + the
stand-alone reflectance network, composed by nrhints_b200/hint_fallback.py) against fixtures of the UNMODIFIED reference
(tests/golden/make_clip_golden.py)."""
import numpy as np
import pytest
import torch

import nrh_testlib as T
import nrhints_b200 as nb

pytestmark = pytest.mark.gpu

CASES = {"clip4_sharp_40x64": ("sharp", dict(n_samples=32, n_importance_samples=32, n_shadow_samples=32, n_shadow_importance_samples=32)),
         "clip16_init_24x128": ("init", dict())}


@pytest.mark.parametrize("name", list(CASES))
def test_grouped_shadow_render_matches_reference_fixture(name):
    weights, kw = CASES[name]
    fx = np.load(T.GOLDEN_DIR / f"{name}.npz")
    clip = int(fx["clip"])
    cfg = nb.NeuSModelConfig(renderer=nb.NeuSRendererConfig(n_shadow_importance_clip=clip, **kw))
    sd = T.make_state(weights, nb.NeuSModelConfig(renderer=nb.NeuSRendererConfig(**kw)))
    assert T.state_digest(sd) == str(fx["digest"])
    m = nb.NeuSHintRenderer(cfg); m.load_state_dict(sd); m.cuda()
    rays = nb.RayBundle(**{k[3:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("in_")}).to("cuda")
    with torch.no_grad():
        out = m(rays, background_rgb=torch.ones(1, 3, device="cuda"))
    torch.cuda.synchronize()
    want = {k[4:]: fx[k] for k in fx.files if k.startswith("out_")}
    d_rgb = np.abs(out.rgb.cpu().numpy() - want["rgb"])
    assert d_rgb.max() < 1e-3, f"max |d rgb| = {d_rgb.max():.2e} (BASELINE.json gate 1e-3)"
    assert np.abs(out.depth.cpu().numpy() - want["depth"]).max() < 1e-3
    # the shadow map is the visibility of the group holding the largest weight: a ray whose maximum sits at a group boundary may
    # report the neighbouring group under fp32 noise (as between the reference's own CPU and GPU runs)
    d_vis = np.abs(out.visibilities.cpu().numpy() - want["visibilities"]).reshape(-1)
    assert np.median(d_vis) < 2e-4 and (d_vis > 2e-3).mean() <= 0.1, d_vis
    assert out.visibilities.shape == (rays.origins.shape[0], 1)
    print(name, "max |d rgb|", float(d_rgb.max()), "median |d vis|", float(np.median(d_vis)), "rays with |d vis| > 2e-3:", int((d_vis > 2e-3).sum()))


def test_grouped_shadow_is_evaluation_only():
    cfg = nb.NeuSModelConfig(renderer=nb.NeuSRendererConfig(n_shadow_importance_clip=4))
    m = nb.NeuSHintRenderer(cfg).cuda()
    rays = nb.RayBundle(**T.case_inputs(T.CASES["cfg2_32x128"])[0]).to("cuda")
    with pytest.raises(NotImplementedError):
        m(rays, is_training=True, background_rgb=torch.ones(1, 3, device="cuda"))
    with pytest.raises(NotImplementedError):
        m(rays, background_rgb=torch.ones(1, 3, device="cuda"))          # gradients enabled, parameters require grad
    with torch.no_grad():
        host = m.render_to_host(rays, background_rgb=torch.ones(1, 3, device="cuda"))
        dev = m(rays, background_rgb=torch.ones(1, 3, device="cuda"))
    assert torch.equal(host.rgb, dev.rgb.cpu()) and torch.equal(host.weights, dev.weights.cpu())
