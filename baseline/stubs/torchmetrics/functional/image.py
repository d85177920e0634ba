import torch
import torch.nn.functional as F


def peak_signal_noise_ratio(preds, target, data_range=1.0):
    """10 log10(data_range^2 / mse) over all elements (torchmetrics' default reduction)."""
    mse = torch.mean((preds.float() - target.float()) ** 2)
    return 10.0 * torch.log10(torch.as_tensor(float(data_range) ** 2, device=mse.device) / mse)


def structural_similarity_index_measure(preds, target, data_range=1.0, kernel_size=11, sigma=1.5, k1=0.01, k2=0.03):
    """Gaussian-window SSIM on [N,C,H,W] images (torchmetrics defaults)."""
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    ax = torch.arange(kernel_size, dtype=torch.float32, device=preds.device) - (kernel_size - 1) / 2.0
    g = torch.exp(-(ax ** 2) / (2 * sigma ** 2)); g = g / g.sum()
    C = preds.shape[1]
    w = (g[:, None] * g[None, :]).expand(C, 1, kernel_size, kernel_size).contiguous()
    pad = (kernel_size - 1) // 2
    x, y = F.pad(preds.float(), (pad,) * 4, mode="reflect"), F.pad(target.float(), (pad,) * 4, mode="reflect")
    mu_x, mu_y = F.conv2d(x, w, groups=C), F.conv2d(y, w, groups=C)
    sxx = F.conv2d(x * x, w, groups=C) - mu_x ** 2
    syy = F.conv2d(y * y, w, groups=C) - mu_y ** 2
    sxy = F.conv2d(x * y, w, groups=C) - mu_x * mu_y
    ssim = ((2 * mu_x * mu_y + c1) * (2 * sxy + c2)) / ((mu_x ** 2 + mu_y ** 2 + c1) * (sxx + syy + c2))
    return ssim.mean()
