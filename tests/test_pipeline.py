"""The composed callers of the path (nrhints_b200/pipeline.py = pipelines/base_pipeline.py:16-91 on the CUDA operators):
training iterations over both parameter groups, and camera registration with the renderer frozen."""
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu


def _pipeline(cam_opt="SO3xR3", pl_opt=True, R=128, seed=5):
    import nrhints_b200 as nb
    from nrhints_b200.workload import synthetic_pixel_bundle
    bundle, cam = synthetic_pixel_bundle(R, seed=seed, n_cameras=6)
    # keep the pixels near the image centre so that most rays hit the radius-0.5 sphere of the geometric init
    g = torch.Generator().manual_seed(seed)
    bundle.h_indices = torch.randint(300, 500, (R, 1), generator=g).float()
    bundle.w_indices = torch.randint(300, 500, (R, 1), generator=g).float()
    mcfg = nb.NeuSModelConfig(renderer=nb.NeuSRendererConfig(n_samples=16, n_importance_samples=16, n_shadow_samples=16,
                                                             n_shadow_importance_samples=16))
    torch.manual_seed(3407)
    pipe = nb.NRHintPipeline(mcfg, nb.RayGeneratorConfig(cam_opt_mode=cam_opt, pl_opt=pl_opt), nb.CameraModel(**cam), 6).cuda()
    dev = SimpleNamespace(**{k: v.cuda() for k, v in vars(bundle).items()})
    return pipe, dev


def test_training_iterations_update_both_parameter_groups():
    """trainer/trainer.py:269-283 (train_iter) on the pipeline: forward, loss dict, zero_grad, backward, FlatAdam step."""
    pipe, bundle = _pipeline()
    opt = pipe.make_optimizer()
    assert len(opt.param_groups) == 2 and opt.param_groups[1]["lr"] == pipe.ray_generator.config.opt_lr
    opt.param_groups[1]["lr"] = 1e-3                        # visible movement of the pose parameters within a few steps
    before = [p.detach().clone() for p in pipe.parameters()]
    losses = []
    for it in range(4):
        res = pipe(bundle, global_step=60000 + it)
        ld = pipe.get_train_loss_dict(res, bundle)
        assert set(ld) == {"loss", "rgb_loss", "eikonal_loss", "s_val", "psnr"}
        opt.zero_grad()
        ld["loss"].backward()
        for name, p in pipe.named_parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all(), name
        opt.step()
        losses.append(float(ld["loss"]))
    assert all(l == l for l in losses) and min(losses[1:]) < losses[0], losses
    moved = [float((p.detach() - b).abs().max()) for p, b in zip(pipe.parameters(), before)]
    assert all(m > 0 for m in moved), "a parameter did not move"
    assert float(pipe.ray_generator.cam_pose_adjustment.abs().max()) > 0 and float(pipe.ray_generator.pl_adjustment.abs().max()) > 0


def test_register_view_recovers_a_pose_offset():
    """pipelines/base_pipeline.py:71-91: with the renderer frozen, fitting cam_pose_adjustment to images rendered from offset
    poses reduces the image loss; the renderer's parameters and their requires_grad flags are left untouched."""
    pipe, bundle = _pipeline(pl_opt=False, R=256)
    with torch.no_grad():
        pipe.ray_generator.cam_pose_adjustment[:, :3] = 0.03            # the "true" poses: every camera shifted
        rays = pipe.ray_generator(bundle)
        bundle.rgb_gt = pipe.renderer(rays, background_rgb=torch.ones(1, 3, device="cuda")).rgb.clone()
        pipe.ray_generator.cam_pose_adjustment.zero_()                  # start the registration from the unshifted poses
    w0 = [p.detach().clone() for p in pipe.renderer.parameters()]
    opt = torch.optim.Adam(pipe.ray_generator.parameters(), lr=3e-3)
    losses = pipe.register_view(iter(lambda: bundle, None), steps=25, optimizer=opt)
    assert losses.shape == (25,) and float(losses[-5:].mean()) < 0.9 * float(losses[:3].mean()), losses
    assert all(torch.equal(a, p.detach()) and p.requires_grad for a, p in zip(w0, pipe.renderer.parameters()))
