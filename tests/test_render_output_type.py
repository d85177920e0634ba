"""RenderOutput behaves like the reference's TensorDataclass for the operations its callers use (pipelines/base_pipeline.py:
114-133: `.to('cpu')`, `td_concat`, `.reshape(img.shape)`, field access), and ReflectanceNetwork is callable with the
reference's signature (fields/reflectance_network.py:68-96).  CPU only."""
import dataclasses

import torch

import nrhints_b200 as nb
from nrhints_b200.renderer import RenderOutput


def _make(R, S=5, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)      # noqa: E731
    return RenderOutput(rgb=r(R, 3), depth=r(R, 1), weights=r(R, S), s_val=r(R, S), inside_sphere=r(R, S), relax_inside_sphere=r(R, S),
                        analytic_normals=r(R, S, 3), normalized_analytic_normals=r(R, S, 3), visibilities=r(R, 1), specular_cue=r(R, S, 4))


def _td_concat(tds):
    """What the reference's td_concat does (utils/tensor_dataclass.py:365-384), restated: concatenate every tensor field."""
    data = {f.name: [getattr(t, f.name) for t in tds] for f in dataclasses.fields(tds[0])}
    return tds[0].__class__(**{k: (None if v[0] is None else torch.concat(v, dim=0)) for k, v in data.items()})


def test_batch_operations():
    a, b = _make(6, seed=1), _make(6, seed=2)
    assert a.shape == (6,) and len(a) == 6 and a.ndim == 1 and a.size == 6
    cat = _td_concat([a.to("cpu"), b.to("cpu")])
    assert cat.shape == (12,) and cat.z_vals is None
    img = cat.reshape((3, 4))
    assert img.shape == (3, 4) and img.rgb.shape == (3, 4, 3) and img.analytic_normals.shape == (3, 4, 5, 3)
    assert img.specular_cue.shape == (3, 4, 5, 4) and img.weights.shape == (3, 4, 5)
    assert torch.equal(img.flatten().rgb, cat.rgb)
    row = img[1]
    assert row.shape == (4,) and torch.equal(row.analytic_normals, cat.analytic_normals[4:8])
    assert torch.equal(img[:, 0].depth, cat.depth[0::4])
    assert torch.equal(img[..., 1:3].weights, img.weights[:, 1:3])
    mask = torch.tensor([True, False] * 6)
    assert torch.equal(cat[mask].rgb, cat.rgb[mask])
    d = cat.detach()
    assert d.rgb.data_ptr() == cat.rgb.data_ptr()


def test_reflectance_network_is_callable_with_the_reference_signature():
    cfg = nb.NeuSModelConfig()
    torch.manual_seed(3407)
    m = nb.NeuSHintRenderer(cfg)
    P = 7
    g = torch.Generator().manual_seed(3)
    pts, nrm, view, light = (torch.randn(P, 3, generator=g) for _ in range(4))
    feat, vis, spec = torch.randn(P, 256, generator=g), torch.rand(P, 1, generator=g), torch.rand(P, 4, generator=g)
    out = m.color_network(pts, nrm, view, feat, light, vis, spec)
    assert out.shape == (P, 3) and float(out.min()) >= 0 and float(out.max()) <= 1
    # against the oracle's restatement of the same network
    from oracle import nrh_oracle as orc
    W = orc.effective_weights({k: v.detach() for k, v in m.state_dict().items()})
    want = orc.reflectance_mlp(W, pts, nrm, view, feat, light, vis, spec, orc.OracleConfig.from_model_config(cfg))
    assert torch.allclose(out, want, atol=1e-6)
