"""Gradient synchronisation of the data-parallel training step: ONE collective per step.

Rays are independent, every rank owns a full 3.3 MB weight replica and renders its own contiguous slice of the batch
(trainer/trainer.py:118); the only exchange on the path is the gradient all-reduce the reference gets from
DistributedDataParallel's bucketed NCCL reducer (trainer/trainer.py:88-93).  Here all 46 parameter gradients
(820 923 fp32 values) travel as one flat buffer through a single all-reduce on the compute stream (NCCL over
NVLink / NVSwitch on the GPUs, gloo in the CPU tests), issued right after backward.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.distributed as dist


class FlatGradSync:
    """Flat-buffer gradient all-reduce for a module (or any iterable of parameters).

    sync() leaves `p.grad` holding the mean (or sum) over ranks for every parameter that requires grad.  A parameter whose
    grad is None contributes zeros and receives the reduced value (the reference's DDP runs with
    find_unused_parameters=False, so every parameter has a grad every step; this keeps ranks consistent if one did not).
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None, average: bool = True):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.average = average
        self.numel = sum(p.numel() for p in self.params)
        self._flat: Optional[torch.Tensor] = None

    def _buffer(self) -> torch.Tensor:
        p0 = self.params[0]
        if self._flat is None or self._flat.device != p0.device:
            self._flat = torch.empty(self.numel, dtype=torch.float32, device=p0.device)
        return self._flat

    @torch.no_grad()
    def sync(self) -> torch.Tensor:
        if not self.params:
            return torch.empty(0)
        flat = self._buffer()
        off = 0
        for p in self.params:                           # pack (one fused copy per parameter, same stream as backward)
            n = p.numel()
            if p.grad is None:
                flat[off:off + n].zero_()
            else:
                flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)          # the one collective of the step
            if self.average:
                flat.div_(dist.get_world_size(self.group))
        off = 0
        for p in self.params:                           # unpack as views into fresh storage-independent grads
            n = p.numel()
            g = flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
        return flat


def allreduce_gradients(module: torch.nn.Module, group: Optional[dist.ProcessGroup] = None, average: bool = True) -> torch.Tensor:
    """Convenience wrapper: one flat all-reduce over all gradients of `module` (caches the buffer on the module)."""
    sync = getattr(module, "_flat_grad_sync", None)
    if sync is None or sync.group is not group or sync.average != average:
        sync = FlatGradSync(module.parameters(), group=group, average=average)
        module._flat_grad_sync = sync
    return sync.sync()


def allreduce_flat(flat_grads, group: Optional[dist.ProcessGroup] = None) -> float:
    """All-reduce (sum) the flat gradient buffers of a FlatAdam optimiser IN PLACE -- the parameters' `.grad` are views of them,
    so there is nothing to pack or unpack -- and return the factor that turns the sum into the mean over ranks (hand it to
    FlatAdam.step(grad_scale=...), which folds it into the optimiser launch)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 1.0
    for g in flat_grads:
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / dist.get_world_size(group)
