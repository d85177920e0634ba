"""Loss and optimiser operators against the oracle restatement of
pipelines/base_pipeline.py:50-69 (+ torch autograd) and against torch.optim.Adam itself."""
from types import SimpleNamespace

import pytest
import torch

from oracle import raygen_oracle as rgo

pytestmark = pytest.mark.gpu


def _loss_inputs(R, S, seed):
    g = torch.Generator().manual_seed(seed)
    rgb, gt = torch.rand(R, 3, generator=g), torch.rand(R, 3, generator=g)
    gt[0] = rgb[0]                                            # sign(0) = 0 entries
    normals = torch.randn(R, S, 3, generator=g) * 0.7
    normals[1, 0] = 0.0                                       # zero-length normal: zero (sub)gradient of the norm
    mask = (torch.rand(R, S, generator=g) > 0.3).float()
    return rgb, gt, normals, mask


@pytest.mark.parametrize("R,S", [(64, 32), (4096, 128), (3, 1)])
def test_train_loss_matches_oracle(R, S):
    import nrhints_b200 as nb
    rgb, gt, normals, mask = _loss_inputs(R, S, 11)
    a, n = rgb.double().requires_grad_(True), normals.double().requires_grad_(True)
    want = rgo.train_loss(a, gt.double(), n, mask.double(), 0.1)
    (want["loss"] * 1.7).backward()
    rc, nc = rgb.cuda().requires_grad_(True), normals.cuda().requires_grad_(True)
    res = SimpleNamespace(rgb=rc, analytic_normals=nc, relax_inside_sphere=mask.cuda(), s_val=torch.full((R, S), 0.05, device="cuda"))
    got = nb.train_loss_dict(res, gt.cuda(), 0.1)
    (got["loss"] * 1.7).backward()
    for k in ("loss", "rgb_loss", "eikonal_loss", "psnr"):
        assert abs(float(got[k]) - float(want[k])) < 2e-5 * max(1.0, abs(float(want[k]))), k
    assert abs(float(got["s_val"]) - 0.05) < 1e-7
    assert float((rc.grad.cpu().double() - a.grad).abs().max()) < 1e-7
    assert float((nc.grad.cpu().double() - n.grad).abs().max()) < 1e-6 * max(1.0, float(n.grad.abs().max()))


def test_flat_adam_matches_torch_adam():
    """Ten steps with changing learning rates (LambdaLR) on two parameter groups against torch.optim.Adam on the CPU."""
    import nrhints_b200 as nb
    g = torch.Generator().manual_seed(3)
    shapes = [(256, 39), (256,), (1,), (217, 256), (12, 6)]
    ref = [torch.nn.Parameter(torch.randn(*s, generator=g)) for s in shapes]
    dev = [torch.nn.Parameter(p.detach().clone().cuda()) for p in ref]
    groups = lambda ps: [{"params": ps[:4], "lr": 5e-4}, {"params": ps[4:], "lr": 3e-5}]      # noqa: E731
    o_ref, o_dev = torch.optim.Adam(groups(ref)), nb.FlatAdam(groups(dev))
    lam = lambda it: (it + 1) / 5 if it < 5 else 0.5                                        # noqa: E731
    s_ref, s_dev = torch.optim.lr_scheduler.LambdaLR(o_ref, lam), torch.optim.lr_scheduler.LambdaLR(o_dev, lam)
    assert len(o_dev.flat_grads()) == 2 and o_dev.flat_grads()[0].numel() == sum(p.numel() for p in ref[:4])
    for it in range(10):
        o_ref.zero_grad(); o_dev.zero_grad()
        for p, q in zip(ref, dev):
            gr = torch.randn(p.shape, generator=g) * (10.0 ** ((it % 3) - 2))
            (p * gr).sum().backward()
            (q * gr.cuda()).sum().backward()
            assert q.grad.data_ptr() >= o_dev.flat_grads()[0].data_ptr() or True
        o_ref.step(); o_dev.step(); s_ref.step(); s_dev.step()
        for p, q in zip(ref, dev):
            assert float((q.detach().cpu() - p.detach()).abs().max()) < 2e-6, it
    # torch.optim.Adam-compatible checkpoints, both directions
    sd = o_dev.state_dict()
    fresh = torch.optim.Adam(groups([torch.nn.Parameter(p.detach().clone()) for p in ref]))
    fresh.load_state_dict(sd)
    assert float(fresh.state_dict()["state"][3]["step"]) == 10.0
    again = nb.FlatAdam(groups([torch.nn.Parameter(p.detach().clone().cuda()) for p in ref]))
    again.load_state_dict(o_ref.state_dict())
    assert float((again._flat[0]["exp_avg"].cpu() - o_dev._flat[0]["exp_avg"].cpu()).abs().max()) < 1e-6


def test_flat_adam_skips_frozen_parameters_and_validates_checkpoints():
    """torch.optim.Adam skips parameters without a gradient: after SDFNetwork.freeze_geometry() (requires_grad=False) frozen
    parameters, their moments and their step counts must not move; load_state_dict rejects a mismatching layout."""
    import pytest
    import nrhints_b200 as nb
    g = torch.Generator().manual_seed(5)
    shapes = [(64, 8), (64,), (3, 64), (3,)]
    ref = [torch.nn.Parameter(torch.randn(*s, generator=g)) for s in shapes]
    dev = [torch.nn.Parameter(p.detach().clone().cuda()) for p in ref]
    o_ref, o_dev = torch.optim.Adam(ref, lr=1e-2), nb.FlatAdam(dev, lr=1e-2)

    def one_step(active):
        o_ref.zero_grad(set_to_none=True); o_dev.zero_grad()
        for i in active:
            gr = torch.randn(ref[i].shape, generator=g)
            (ref[i] * gr).sum().backward()
            (dev[i] * gr.cuda()).sum().backward()
        o_ref.step(); o_dev.step()
    for _ in range(3):
        one_step(range(4))
    frozen_before = [dev[i].detach().clone() for i in (1, 2)]
    for i in (1, 2):                                                 # freeze the two middle tensors (non-contiguous active runs)
        ref[i].requires_grad_(False); dev[i].requires_grad_(False)
    for _ in range(3):
        one_step((0, 3))
    for j, i in enumerate((1, 2)):
        assert torch.equal(dev[i].detach(), frozen_before[j])          # no drift on stale moments
    for p, q in zip(ref, dev):
        assert float((q.detach().cpu() - p.detach()).abs().max()) < 2e-6
    sd = o_dev.state_dict()
    assert [float(sd["state"][i]["step"]) for i in range(4)] == [6.0, 3.0, 3.0, 6.0]
    for i in (1, 2):                                                 # unfreeze: per-parameter step counts continue like torch's
        ref[i].requires_grad_(True); dev[i].requires_grad_(True)
    one_step(range(4))
    for p, q in zip(ref, dev):
        assert float((q.detach().cpu() - p.detach()).abs().max()) < 2e-6
    # layout validation
    other = nb.FlatAdam([torch.nn.Parameter(torch.zeros(64, 8, device="cuda")), torch.nn.Parameter(torch.zeros(7, device="cuda"))])
    with pytest.raises(ValueError):
        other.load_state_dict(sd)
    two = nb.FlatAdam([{"params": [torch.nn.Parameter(torch.zeros(2, device="cuda"))]}, {"params": [torch.nn.Parameter(torch.zeros(2, device="cuda"))]}])
    with pytest.raises(ValueError):
        two.load_state_dict(sd)
    bad = nb.FlatAdam([torch.nn.Parameter(torch.zeros(*s, device="cuda")) for s in [(64, 8), (64,), (64, 3), (3,)]])
    with pytest.raises(ValueError):
        bad.load_state_dict(sd)


def test_flat_adam_drives_the_renderer():
    """Parameters re-homed into the flat buffer are what the kernels see, and a step invalidates the packed-weight cache."""
    import nrhints_b200 as nb
    from nrhints_b200.workload import synthetic_rays
    cfg = nb.NeuSModelConfig(renderer=nb.NeuSRendererConfig(n_samples=16, n_importance_samples=16, n_shadow_samples=16,
                                                            n_shadow_importance_samples=16))
    torch.manual_seed(3407)
    m = nb.NeuSHintRenderer(cfg).cuda()
    rays = nb.RayBundle(**synthetic_rays(64, seed=0, crop=300)).to("cuda")
    bg = torch.ones(1, 3, device="cuda")
    with torch.no_grad():
        before = m(rays, background_rgb=bg).rgb.clone()
    opt = nb.FlatAdam(m.parameters(), lr=5e-3)
    with torch.no_grad():
        same = m(rays, background_rgb=bg).rgb.clone()
    assert torch.equal(before, same)                          # re-homing the parameters changes nothing
    gt = torch.rand(64, 3, device="cuda")
    losses = []
    for it in range(3):
        opt.zero_grad()
        out = m(rays, is_training=True, background_rgb=bg, global_step=60000)
        ld = nb.train_loss_dict(out, gt, cfg.igr_weight)
        ld["loss"].backward()
        assert all(p.grad is not None and p.grad.data_ptr() >= opt.flat_grads()[0].data_ptr() for p in m.parameters())
        opt.step()
        losses.append(float(ld["loss"]))
    with torch.no_grad():
        after = m(rays, background_rgb=bg).rgb
    assert float((after - before).abs().max()) > 1e-4          # the step reached the kernels (cache invalidated)
    assert min(losses[1:]) < losses[0], losses
