"""The fused training step (nrh_render_train_forward / nrh_render_backward; nrhints_b200/fused_step.py) against the composed
autograd path of the same library and through the native pipeline (NRHintPipeline.train_step).  The comparison with the oracle's
autograd gradients (pinned to the reference's own loss.backward() fixture) is tests/test_gpu_parity.py::
test_training_gradients_match_oracle, which runs the fused node by default."""
import ctypes as C

import numpy as np
import pytest
import torch

import nrh_testlib as T
import nrhints_b200 as nb
from nrhints_b200 import _lib, fused_step
from nrhints_b200.workload import synthetic_pixel_bundle

pytestmark = pytest.mark.gpu


def _module(kind="sharp", fused=True):
    cfg = nb.NeuSModelConfig()
    m = nb.NeuSHintRenderer(cfg, mlp_impl="auto")
    m.load_state_dict(T.make_state(kind, cfg))
    m.fused_training = fused
    return m.cuda(), cfg


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12)), float((a - b).norm() / b.norm().clamp_min(1e-12))


@pytest.mark.parametrize("kind", ["init", "sharp"])
def test_pack_weights_wn_is_bitwise_torch_weight_norm(kind):
    """nrh_pack_weights_wn applies the weight norm inside the library (one launch for all 15 layers).  Its effective weights are
    BITWISE those of torch._weight_norm on the same device -- what the reference's weight-normed layers evaluate
    (fields/sdf_field.py:97-98) -- so the packed buffer equals nrh_pack_weights on torch's effective weights byte for byte."""
    m, cfg = _module(kind)
    dev = torch.device("cuda", torch.cuda.current_device())
    assert fused_step.can_pack_wn(m)
    a = m._ensure_packed(dev).clone()
    w = m._wn_scratch.view(torch.float32)
    off = 0
    for lin in fused_step.wn_layers(m):
        e = lin.effective_weight().detach().reshape(-1)
        assert torch.equal(w[off:off + e.numel()], e), tuple(lin.weight_v.shape)
        off += e.numel()
    m._packed_key = None
    orig = fused_step.can_pack_wn
    fused_step.can_pack_wn = lambda r: False          # force the torch-side weight norm + nrh_pack_weights
    try:
        b = m._ensure_packed(dev).clone()
    finally:
        fused_step.can_pack_wn = orig
    assert torch.equal(a, b)


@pytest.mark.parametrize("kind,step", [("init", 25000), ("sharp", 60000)])
def test_fused_node_matches_composed_autograd_path(kind, step):
    """Same seeds, same jitters: outputs equal to fp32 rounding, gradients of all 46 parameters and of the rays within the fp16
    backward's noise (both paths run the same tensor-core kernels; the glue differs)."""
    from oracle import nrh_oracle as orc
    R = 512
    rays = orc.synthetic_rays(R, seed=11, crop=500)
    gt = torch.rand(R, 3, generator=torch.Generator().manual_seed(3)).cuda()
    res = {}
    for fused in (True, False):
        m, cfg = _module(kind, fused)
        dev = {k: v.cuda() for k, v in rays.items()}
        for k in ("origins", "directions", "pl_positions"):
            dev[k].requires_grad_(True)
        torch.manual_seed(123)
        out = m(nb.RayBundle(**dev), is_training=True, background_rgb=torch.ones(1, 3).cuda(), global_step=step)
        loss = orc.training_loss({"rgb": out.rgb, "analytic_normals": out.analytic_normals,
                                  "relax_inside_sphere": out.relax_inside_sphere}, gt)
        loss.backward()
        g = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
        g.update({"ray::" + k: dev[k].grad.detach().clone() for k in ("origins", "directions", "pl_positions")})
        res[fused] = (out, float(loss), g)
    (oa, la, ga), (ob, lb, gb) = res[True], res[False]
    assert abs(la - lb) < 2e-6 * max(1.0, abs(lb)), (la, lb)
    for f in ("rgb", "weights", "depth", "analytic_normals", "normalized_analytic_normals", "visibilities", "inside_sphere"):
        assert float((getattr(oa, f) - getattr(ob, f)).abs().max()) < 2e-5, f
    worst = ("", 0.0, 0.0)
    for n in gb:
        e_max, e_l2 = _rel(ga[n], gb[n])
        if e_max > worst[1]:
            worst = (n, e_max, e_l2)
        assert e_max < 2e-2 and e_l2 < 1e-2, f"{n}: max {e_max:.2e} l2 {e_l2:.2e}"
    print("FUSED-vs-COMPOSED", kind, "worst", worst)


def test_pipeline_train_step_equals_autograd_step():
    """NRHintPipeline.train_step (no autograd graph, gradients written into FlatAdam's flat buffer) produces the gradients of
    forward -> get_train_loss_dict -> backward, incl. the camera / light parameters of the ray generator, and optimises."""
    R = 1024
    pb, cam = synthetic_pixel_bundle(R, seed=21)
    from types import SimpleNamespace
    px = SimpleNamespace(**{k: v.cuda() for k, v in vars(pb).items()})
    torch.manual_seed(0)
    cfg = nb.NeuSModelConfig()
    pipe = nb.NRHintPipeline(cfg, nb.RayGeneratorConfig(cam_opt_mode="SO3xR3", pl_opt=True, cam_position_noise_std=1e-3,
                                                        cam_orientation_noise_std=1e-3), nb.CameraModel(**cam), 64)
    pipe.renderer.load_state_dict(T.make_state("sharp", cfg))
    pipe = pipe.cuda()
    opt = pipe.make_optimizer()
    grads = {}
    for direct in (True, False):
        torch.manual_seed(100)
        opt.zero_grad()
        if direct:
            d = pipe.train_step(px, global_step=60000, optimizer=None)
        else:
            d = pipe.get_train_loss_dict(pipe(px, global_step=60000), px)
            d["loss"].backward()
        grads[direct] = ({n: p.grad.detach().clone() for n, p in pipe.named_parameters()}, {k: float(v) for k, v in d.items()})
    (ga, da), (gb, db) = grads[True], grads[False]
    for k in db:
        assert abs(da[k] - db[k]) < 2e-4 * max(1.0, abs(db[k])), (k, da[k], db[k])
    for n in gb:
        e_max, e_l2 = _rel(ga[n], gb[n])
        assert e_max < 2e-2 and e_l2 < 1e-2, f"{n}: max {e_max:.2e} l2 {e_l2:.2e}"
    assert float(gb["ray_generator.cam_pose_adjustment"].abs().max()) > 0 and float(gb["ray_generator.pl_adjustment"].abs().max()) > 0
    # and it optimises: a few steps on a fixed batch lower the loss
    before = {n: p.detach().clone() for n, p in pipe.named_parameters()}
    losses = []
    for it in range(8):
        losses.append(float(pipe.train_step(px, global_step=60000 + it, optimizer=opt)["loss"]))
    assert all(np.isfinite(losses)) and min(losses[4:]) < losses[0], losses
    assert all(float((p.detach() - before[n]).abs().max()) > 0 for n, p in pipe.named_parameters() if p.requires_grad)


def test_backward_rejects_missing_gradient_pointers():
    m, cfg = _module("init")
    lib = _lib.load()
    c = m._c_config()
    P = _lib.NrhTrainParams()
    rays = _lib.NrhRays()
    adj = _lib.NrhTrainAdjoints()
    ws = torch.empty(1024, dtype=torch.uint8, device="cuda")
    rc = lib.nrh_render_backward(C.byref(c), ws.data_ptr(), C.byref(P), C.byref(rays), 16, None, 1.0, C.byref(adj), ws.data_ptr(), 1024, None)
    assert rc != 0 and b"null" in lib.nrh_last_error()


def test_flat_adam_capturable_matches_host_scalars():
    """nrh_adam_step_dev (step count and learning rate on the device) takes the same steps as nrh_adam_step, incl. a changing lr."""
    from nrhints_b200.train_ops import FlatAdam
    torch.manual_seed(0)
    ps = [[torch.nn.Parameter(torch.randn(257, 33, device="cuda")), torch.nn.Parameter(torch.randn(19, device="cuda"))] for _ in range(2)]
    for a, b in zip(*ps):
        b.data.copy_(a.data)
    opts = [FlatAdam([{"params": ps[0], "lr": 1e-2}]), FlatAdam([{"params": ps[1], "lr": 1e-2}], capturable=True)]
    for it in range(5):
        gs = [torch.randn_like(p) for p in ps[0]]
        for o, pl in zip(opts, ps):
            o.zero_grad()
            for p, g in zip(pl, gs):
                p.grad.copy_(g)
            o.param_groups[0]["lr"] = 1e-2 * (0.9 ** it)
            o.step()
    for a, b in zip(*ps):
        assert float((a - b).abs().max()) < 2e-7 * float(a.abs().max())
    sd = opts[1].state_dict()
    assert int(sd["state"][0]["step"]) == 5


@pytest.mark.parametrize("with_sync", [False, True])
def test_graphed_train_step_matches_eager_train_step(with_sync):
    """NRHintPipeline.capture_train_step: the CUDA-graph replay of the whole step (weight pack, ray generation, forward, loss, backward,
    Adam with device-side step count) walks the same trajectory as eager train_step calls on the same batches and seeds."""
    from types import SimpleNamespace
    R = 512
    batches = []
    for i in range(4):
        pb, cam = synthetic_pixel_bundle(R, seed=40 + i)
        batches.append(SimpleNamespace(**{k: v.cuda() for k, v in vars(pb).items()}))
    traj = {}
    for graphed in (True, False):
        torch.manual_seed(0)
        cfg = nb.NeuSModelConfig()
        pipe = nb.NRHintPipeline(cfg, nb.RayGeneratorConfig(), nb.CameraModel(**cam), 64)
        pipe.renderer.load_state_dict(T.make_state("sharp", cfg))
        pipe = pipe.cuda()
        opt = pipe.make_optimizer(capturable=True)
        losses = []
        torch.manual_seed(77)
        if graphed:
            # with a gradient exchange the graph ends after the backward; the collective (here: a stand-in) and Adam follow eagerly
            g = pipe.capture_train_step(batches[0], opt, global_step=60000, grad_sync=(lambda: 1.0) if with_sync else None)
            losses.append(float(g.first["loss"]))
            for i in range(1, 4):
                losses.append(float(g(batches[i], 60000 + i)["loss"]))
        else:
            for i in range(4):
                losses.append(float(pipe.train_step(batches[i], global_step=60000 + i, optimizer=opt)["loss"]))
        assert int(opt._flat[0]["step_dev"].item()) == 4                 # the replays advanced the device-side step count
        traj[graphed] = (losses, {n: p.detach().clone() for n, p in pipe.named_parameters()})
    (la, pa), (lb, pb_) = traj[True], traj[False]
    # the jitters differ (a graph replays its own Philox offsets), so the comparison is statistical: same loss level, same step sizes
    assert all(np.isfinite(la)) and all(np.isfinite(lb))
    assert abs(la[0] - lb[0]) < 2e-4 * max(1.0, abs(lb[0]))            # the first step is eager in both
    assert abs(np.mean(la) - np.mean(lb)) < 0.05 * abs(np.mean(lb))
    moved_a = max(float((pa[n] - pb_[n]).abs().max()) for n in pa)
    assert moved_a < 4 * 4 * 5e-4                                         # both took 4 Adam steps of size <= lr from the same start
