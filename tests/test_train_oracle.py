"""The oracle restatements of the loss (pipelines/base_pipeline.py:50-69) and of the Adam step (trainer/trainer.py:99) against
torch itself -- these are what the GPU tests of nrh_train_loss / nrh_adam_step compare with (tests/test_train_ops.py)."""
import torch

from oracle import raygen_oracle as rgo


def test_adam_oracle_equals_torch_adam():
    g = torch.Generator().manual_seed(0)
    p_ref = torch.nn.Parameter(torch.randn(50, 7, generator=g, dtype=torch.float64))
    opt = torch.optim.Adam([p_ref], lr=5e-4)
    p, m, v = p_ref.detach().clone(), torch.zeros(50, 7, dtype=torch.float64), torch.zeros(50, 7, dtype=torch.float64)
    for step in range(1, 8):
        grad = torch.randn(50, 7, generator=g, dtype=torch.float64) * (10.0 ** (step % 3 - 1))
        p_ref.grad = grad.clone()
        opt.step()
        p, m, v = rgo.adam_step(p, grad, m, v, 5e-4, 0.9, 0.999, 1e-8, step)
        assert float((p - p_ref.detach()).abs().max()) < 1e-12, step


def test_loss_oracle_matches_the_reference_expression_and_torchmetrics_psnr():
    g = torch.Generator().manual_seed(1)
    R, S = 40, 16
    rgb, gt = torch.rand(R, 3, generator=g, dtype=torch.float64), torch.rand(R, 3, generator=g, dtype=torch.float64)
    normals = torch.randn(R, S, 3, generator=g, dtype=torch.float64)
    mask = (torch.rand(R, S, generator=g) > 0.4).double()
    out = rgo.train_loss(rgb, gt, normals, mask, 0.1)
    # the reference's lines, verbatim semantics (pipelines/base_pipeline.py:57-62)
    rgb_loss = torch.nn.functional.l1_loss(rgb, gt, reduction="sum") / (rgb.size(0) + 1e-5)
    gradient_error = (torch.linalg.norm(normals, ord=2, dim=-1) - 1.0) ** 2
    eikonal = (mask * gradient_error).sum() / (mask.sum() + 1e-5)
    assert float((out["loss"] - (rgb_loss + eikonal * 0.1)).abs()) < 1e-14
    assert float((out["rgb_loss"] - rgb_loss).abs()) < 1e-14 and float((out["eikonal_loss"] - eikonal).abs()) < 1e-14
    # torchmetrics.functional.image.peak_signal_noise_ratio(data_range=1.0): 10 log10(1 / mse), mse over every element
    mse = ((rgb - gt) ** 2).sum() / rgb.numel()
    assert float((out["psnr"] - 10.0 * torch.log10(1.0 / mse)).abs()) < 1e-12
