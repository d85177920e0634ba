"""The hand-written backward of the SDF network (oracle.sdf_mlp_backward: phase A / phase B chains with the second-order
terms of the analytic normals) against torch.autograd through the oracle's forward + explicit reverse sweep, float64."""
import torch

import nrh_testlib as T
from oracle import nrh_oracle as orc


def test_explicit_backward_matches_autograd():
    cfg = T.make_config(T.CASES["cfg2_32x128"])
    ocfg = orc.OracleConfig.from_model_config(cfg)
    W = {k: v.clone().requires_grad_(True) for k, v in orc.effective_weights(T.make_state("sharp", cfg), torch.float64).items()}
    g = torch.Generator().manual_seed(0)
    N = 193
    pts = ((torch.rand(N, 3, generator=g, dtype=torch.float64) - 0.5) * 2.2).requires_grad_(True)
    d_sdf = torch.randn(N, 1, generator=g, dtype=torch.float64)
    d_feat = torch.randn(N, 256, generator=g, dtype=torch.float64) * 0.1
    d_grad = torch.randn(N, 3, generator=g, dtype=torch.float64)
    out = orc.sdf_mlp(W, pts, ocfg, want_feat=True, want_grad=True)
    loss = (out["sdf"] * d_sdf).sum() + (out["feat"] * d_feat).sum() + (out["grad"] * d_grad).sum()
    names = [k for k in W if k.startswith("sdf_network.")]
    auto = torch.autograd.grad(loss, [pts] + [W[k] for k in names])
    with torch.no_grad():
        man = orc.sdf_mlp_backward({k: v.detach() for k, v in W.items()}, pts.detach(), ocfg, d_sdf, d_feat, d_grad)
    assert torch.allclose(man["d_pts"], auto[0], rtol=1e-8, atol=1e-10)
    for k, ga in zip(names, auto[1:]):
        gm = man[k].reshape(ga.shape)
        assert torch.allclose(gm, ga, rtol=1e-8, atol=1e-10), (k, float((gm - ga).abs().max()), float(ga.abs().max()))
