"""Shim for lpips (needs downloaded AlexNet weights: no network).  Returns NaN so that nobody mistakes it for a measurement."""
import torch


class LPIPS(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, a, b, normalize=False):
        return torch.full((1,), float("nan"))
