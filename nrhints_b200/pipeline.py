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
from . import fused_step
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

    def make_optimizer(self, capturable: bool = False) -> FlatAdam:
        """trainer/trainer.py:99 with flat buffers (one Adam launch per parameter group; empty groups are skipped).
        capturable=True: step count / learning rate on the device, for `capture_train_step`."""
        return FlatAdam([g for g in self.get_param_groups() if len(g["params"]) > 0], capturable=capturable)

    def capture_train_step(self, pixel_bundle, optimizer: FlatAdam, grad_sync=None, global_step: int = 0) -> "GraphedTrainStep":
        """`train_step` (ray generation -> forward -> loss -> backward -> [gradient all-reduce] -> Adam) captured ONCE as a CUDA graph and
        replayed per step: at the reference's data-parallel batch split (trainer/trainer.py:118: batch_size // world_size rays per
        rank) the ~75 launches of a step are launch-latency bound otherwise.  The optimizer must be FlatAdam(capturable=True)."""
        return GraphedTrainStep(self, pixel_bundle, optimizer, grad_sync, global_step)

    def train_step(self, pixel_bundle, global_step: int = 0, optimizer: Optional[torch.optim.Optimizer] = None,
                   grad_sync=None) -> Dict[str, torch.Tensor]:
        """The reference's train_iter (trainer/trainer.py:269-283: forward, get_train_loss_dict, zero_grad, backward, step) without an
        autograd graph: ray generation -> nrh_render_train_forward -> nrh_train_loss -> nrh_render_backward, the gradients of the
        renderer's parameters written straight into their `.grad` tensors (with FlatAdam: views of the flat gradient buffer, i.e. the
        all-reduce operand), the adjoints of the rays handed to the ray generator's backward when it has parameters to optimise.
        `grad_sync`: optional callable run between backward and the optimizer step (grad_sync.allreduce_flat); its return value is
        passed to FlatAdam.step as grad_scale.  Returns the loss dict of get_train_loss_dict (0-d device tensors, no host sync)."""
        if getattr(self, "_fused", None) is None or self._fused.renderer is not self.renderer:
            self._fused = fused_step.FusedTrainStep(self.renderer)
        need_ray = any(p.requires_grad for p in self.ray_generator.parameters())
        if optimizer is not None:
            optimizer.zero_grad()
        with torch.set_grad_enabled(need_ray):
            rays = self.ray_generator(pixel_bundle)
        det = lambda t: t.detach().to(torch.float32).contiguous()      # noqa: E731
        res = self._fused.forward_backward(det(rays.origins), det(rays.directions), det(rays.pl_positions), det(rays.nears),
                                           det(rays.fars), det(pixel_bundle.rgb_gt), self._background(pixel_bundle.pls.device),
                                           global_step, self.model_config.igr_weight, need_ray_grads=need_ray)
        if need_ray:
            torch.autograd.backward([rays.origins, rays.directions, rays.pl_positions],
                                    [res["d_origins"], res["d_directions"], res["d_pl_positions"]])
        if optimizer is not None:
            scale = grad_sync() if grad_sync is not None else None
            if scale is not None:
                optimizer.step(grad_scale=scale)
            else:
                optimizer.step()
        st = res["stats"]
        return {"loss": st[0], "rgb_loss": st[1], "eikonal_loss": st[2], "s_val": res["s_val"], "psnr": st[3]}

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


class GraphedTrainStep:
    """A captured NRHintPipeline.train_step.  Calling it copies the pixel bundle into the static input buffers, pushes the current
    learning rates to the device and replays the graph.  Without a gradient exchange the graph holds the whole step (weight pack, ray
    generation, forward, loss, backward, Adam with its step count on the device); with `grad_sync` (data parallel) it ends after the
    backward and the collective + Adam launches follow eagerly on the same stream -- three launches, and no NCCL kernel inside a graph.
    The graph bakes the two host scalars of a step that depend on `global_step` -- cos_anneal = min(1, step / anneal_end) and the
    geometry warm-up flag -- so it is re-captured when they change and the step runs eagerly while cos_anneal still moves
    (global_step < anneal_end)."""

    def __init__(self, pipe: NRHintPipeline, pixel_bundle, optimizer: FlatAdam, grad_sync, global_step: int):
        if grad_sync is None and not getattr(optimizer, "capturable", False):
            raise ValueError("capture_train_step needs FlatAdam(capturable=True) (make_optimizer(capturable=True))")
        self.pipe, self.opt, self.grad_sync = pipe, optimizer, grad_sync
        self.fields = [k for k, v in vars(pixel_bundle).items() if isinstance(v, torch.Tensor)]
        self.static = type(pixel_bundle)(**{k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in vars(pixel_bundle).items()})
        self.graph, self.key, self.out = None, None, None
        self.first = self._capture(global_step)            # loss dict of the step taken while capturing

    def _key(self, global_step: int):
        m = self.pipe.model_config
        cos = min(1.0, global_step / m.anneal_end) if m.anneal_end > 0 else 1.0
        return (cos, bool(global_step < m.geometry_warmup_end))

    def _finish(self):
        """data-parallel tail of a step: the one collective of the loop and the optimizer launch(es), eager."""
        if self.grad_sync is not None:
            scale = self.grad_sync()
            self.opt.step(grad_scale=scale) if scale is not None else self.opt.step()

    def _capture(self, global_step: int):
        dev = self.static.poses.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                     # one REAL step outside the capture (lazy allocations, library loading)
            res = self.pipe.train_step(self.static, global_step=global_step, optimizer=self.opt, grad_sync=self.grad_sync)
            res = {k: v.clone() for k, v in res.items()}
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):                 # nothing executes here: the launches are recorded
            if self.grad_sync is None:
                self.out = self.pipe.train_step(self.static, global_step=global_step, optimizer=self.opt)
            else:
                self.opt.zero_grad()
                self.out = self.pipe.train_step(self.static, global_step=global_step, optimizer=None)
        self.key = self._key(global_step)
        return res

    def __call__(self, pixel_bundle, global_step: int):
        for k in self.fields:
            getattr(self.static, k).copy_(getattr(pixel_bundle, k), non_blocking=True)
        key = self._key(global_step)
        if key != self.key:
            if key[0] < 1.0:                               # still annealing: a new scalar every step -> eager
                return self.pipe.train_step(self.static, global_step=global_step, optimizer=self.opt, grad_sync=self.grad_sync)
            return self._capture(global_step)              # performs this step for real during its warm-up
        self.opt.sync_lr()
        self.graph.replay()
        self._finish()
        # the step updated the parameters through raw pointers: version-keyed caches (the renderer's packed weights) must see it
        torch.autograd.graph.increment_version([p for p in self.pipe.parameters() if p.requires_grad])
        return self.out
