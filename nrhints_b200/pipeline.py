"""The callers either side of the ray-march path composed the way the reference composes them:
BaseNRHintPipeline (/root/reference/pipelines/base_pipeline.py:16-91) = RayGenerator -> NeuSHintRenderer -> loss, on the CUDA
operators of this package.  Same attribute names (`ray_generator`, `renderer`), the same two optimizer parameter groups, the
same `forward` / `get_train_loss_dict` / `register_view` signatures; the evaluation / metric / video parts of the reference
pipeline (SSIM, LPIPS, image dumps) are outside the hot path and stay with the reference.
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional, Union

import torch
from torch import nn

from .config import NeuSModelConfig
from .ray_generator import CameraModel, RayGenerator, RayGeneratorConfig
from .renderer import NeuSHintRenderer, RenderOutput
from .train_ops import FlatAdam, train_loss_dict


class NRHintPipeline(nn.Module):
    def __init__(self, model: NeuSModelConfig, ray_generator: RayGeneratorConfig, camera: CameraModel, total_image_num: int,
                 white_background: bool = True, mlp_impl: str = "auto"):
        super().__init__()
        self.model_config, self.white_background = model, white_background
        self.ray_generator = RayGenerator(camera=camera, num_cameras=total_image_num, config=ray_generator)      # :25-29
        self.renderer = NeuSHintRenderer(model, mlp_impl=mlp_impl)                                             # :30

    def get_param_groups(self) -> List[Dict[str, Union[Iterator[nn.Parameter], float]]]:
        """base_pipeline.py:32-37."""
        return [{"params": list(self.renderer.parameters()), "lr": self.model_config.lr},
                {"params": list(self.ray_generator.parameters()), "lr": self.ray_generator.config.opt_lr}]

    def _background(self, device) -> torch.Tensor:
        return (torch.ones if self.white_background else torch.zeros)([1, 3], device=device)

    def forward(self, pixel_bundle, global_step: int = 0) -> RenderOutput:
        """base_pipeline.py:39-48 (always is_training=True)."""
        ray_bundle = self.ray_generator(pixel_bundle)
        return self.renderer(ray_bundle, background_rgb=self._background(pixel_bundle.pls.device), is_training=True,
                             global_step=global_step)

    def get_train_loss_dict(self, rendering_res: RenderOutput, pixel_bundle) -> Dict[str, torch.Tensor]:
        """base_pipeline.py:50-69; the values are 0-d device tensors (no host sync)."""
        return train_loss_dict(rendering_res, pixel_bundle.rgb_gt, self.model_config.igr_weight)

    def make_optimizer(self) -> FlatAdam:
        """trainer/trainer.py:99 with flat buffers (one Adam launch per parameter group; empty groups are skipped)."""
        return FlatAdam([g for g in self.get_param_groups() if len(g["params"]) > 0])

    @torch.enable_grad()
    def register_view(self, pixel_bundles: Iterator, steps: int = 500, optimizer: Optional[torch.optim.Optimizer] = None):
        """base_pipeline.py:71-91: fit the ray generator's per-image parameters to a view with the renderer frozen
        (`is_training=False`).  `pixel_bundles` yields one device pixel bundle per step (the reference draws batch_size random
        pixels of the view per step on the host); returns the per-step losses as a device tensor."""
        opt = optimizer or torch.optim.Adam(self.ray_generator.parameters(), lr=self.ray_generator.config.opt_lr)
        frozen = [(p, p.requires_grad) for p in self.renderer.parameters()]
        for p, _ in frozen:
            p.requires_grad_(False)
        losses = []
        try:
            for _, pixel_bundle in zip(range(steps), pixel_bundles):
                ray_bundle = self.ray_generator(pixel_bundle)
                res = self.renderer(ray_bundle, background_rgb=self._background(pixel_bundle.pls.device), is_training=False)
                loss = torch.nn.functional.l1_loss(res.rgb, pixel_bundle.rgb_gt, reduction="sum") / (res.rgb.size(0) + 1e-5)
                opt.zero_grad()
                loss.backward()
                opt.step()
                losses.append(loss.detach())
        finally:
            for p, flag in frozen:
                p.requires_grad_(flag)
        return torch.stack(losses) if losses else torch.empty(0)
