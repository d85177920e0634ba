"""Loss and optimiser behind the ray-march path on the CUDA operators of csrc/train_ops.cu.

  train_loss_dict  -- BaseNRHintPipeline.get_train_loss_dict (/root/reference/pipelines/base_pipeline.py:50-69): same keys; the
                      loss tensor is autograd-connected to `rgb` and `analytic_normals` through ONE node whose backward is a
                      scaled copy of gradients the forward launch already produced.  No host sync ('psnr' stays a device scalar;
                      the reference's float(...) of it is a sync per step).
  FlatAdam         -- torch.optim.Adam as the reference configures it (/root/reference/trainer/trainer.py:99: default betas/eps,
                      no weight decay, one lr per parameter group) over FLAT buffers: parameters, gradients and both moments of
                      a group live in one allocation each, `p.data` / `p.grad` are views into them, so zero_grad is one memset,
                      the data-parallel gradient all-reduce is one collective on the buffer itself (no packing copies) and the
                      step is one kernel launch per group.  state_dict()/load_state_dict() use torch.optim.Adam's layout.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from . import _lib


def colsum_f16(mats: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """fp32 column sums of fp16 matrices: [n,P,W] (or [P,W]) -> [n,W] ([W]) = scale * sum over P, one launch (nrh_colsum_f16)."""
    lib = _lib.load()
    if not (mats.is_cuda and mats.dtype == torch.float16 and mats.is_contiguous()):
        raise RuntimeError("colsum_f16 needs a contiguous CUDA fp16 tensor (no CPU fallback)")
    m3 = mats if mats.dim() == 3 else mats[None]
    n, P, W = m3.shape
    out = torch.empty(n, W, dtype=torch.float32, device=mats.device)
    with torch.cuda.device(mats.device):
        _lib.check(lib.nrh_colsum_f16(m3.data_ptr(), n, P, W, P * W, float(scale), out.data_ptr(),
                                      torch.cuda.current_stream(mats.device).cuda_stream), "nrh_colsum_f16")
    return out if mats.dim() == 3 else out[0]


class WgradBatch:
    """All weight-gradient reductions of a training step in ONE launch (nrh_wgrad_f16, csrc/wgrad_tc.cu: tcgen05 + TMA tensor maps):
        out[m, n] += scale * dev_scale * sum_p A[p, a_col0 + m] * B[p, b_col0 + n]
    `add()` queues a job (A [P, a_ld] / B [P, b_ld]: fp16 row-major CUDA tensors or column windows of them; out: fp32 [m, >= n],
    accumulated); `run()` launches them and keeps the operands alive until then."""

    def __init__(self):
        self.jobs, self.keep = [], []

    def add(self, a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, scale: float = 1.0, dev_scale: torch.Tensor = None,
            m: int = None, n: int = None, a_col0: int = 0, b_col0: int = 0, rows_valid: int = 0, cols_valid: int = 0):
        for t in (a, b):
            if not (t.is_cuda and t.dtype == torch.float16 and t.dim() == 2 and t.stride(1) == 1):
                raise RuntimeError("WgradBatch operands must be 2-D fp16 CUDA tensors with unit column stride (no CPU fallback)")
        if not (out.is_cuda and out.dtype == torch.float32 and out.dim() == 2 and out.stride(1) == 1):
            raise RuntimeError("WgradBatch output must be a 2-D fp32 CUDA tensor with unit column stride")
        if a.shape[0] != b.shape[0]:
            raise RuntimeError("WgradBatch operands need the same number of rows")
        j = _lib.NrhWgradJob()
        j.a, j.a_ld, j.a_col0 = a.data_ptr(), a.stride(0), a_col0
        j.b, j.b_ld, j.b_col0 = b.data_ptr(), b.stride(0), b_col0
        j.rows = a.shape[0]
        j.m = m if m is not None else a.shape[1] - a_col0
        j.n = n if n is not None else b.shape[1] - b_col0
        j.rows_valid, j.cols_valid = rows_valid, cols_valid
        j.scale = float(scale)
        j.dev_scale = dev_scale.data_ptr() if dev_scale is not None else None
        j.out, j.ld_out = out.data_ptr(), out.stride(0)
        self.jobs.append(j)
        self.keep += [a, b, out, dev_scale]
        return self

    def run(self):
        if not self.jobs:
            return
        import ctypes as C
        lib = _lib.load()
        dev = self.keep[0].device
        arr = (_lib.NrhWgradJob * len(self.jobs))(*self.jobs)
        with torch.cuda.device(dev):
            _lib.check(lib.nrh_wgrad_f16(arr, len(self.jobs), torch.cuda.current_stream(dev).cuda_stream), "nrh_wgrad_f16")
        self.jobs, self.keep = [], []


def color_train_forward(renderer, x16: torch.Tensor):
    """Reflectance network forward of a training step on tcgen05 (nrh_color_train_forward): x16 [P,384] fp16 (the reference's input
    order, zero padded) -> (acts [4,P,256] fp16 = post-ReLU hidden activations x16, y [P,4] fp32 pre-sigmoid outputs)."""
    import ctypes as C
    lib = _lib.load()
    if not (x16.is_cuda and x16.dtype == torch.float16 and x16.dim() == 2 and x16.shape[1] == 384 and x16.is_contiguous()):
        raise RuntimeError("color_train_forward needs a contiguous CUDA fp16 [P,384] input (no CPU fallback)")
    dev, P = x16.device, x16.shape[0]
    packed, cfg = renderer._ensure_packed(dev), renderer._c_config()
    acts = torch.empty(4, P, 256, dtype=torch.float16, device=dev)
    y = torch.empty(P, 4, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nrh_color_train_forward(C.byref(cfg), packed.data_ptr(), x16.data_ptr(), P, acts.data_ptr(), y.data_ptr(),
                                               torch.cuda.current_stream(dev).cuda_stream), "nrh_color_train_forward")
    return acts, y


def color_train_backward(renderer, dy: torch.Tensor, scale: torch.Tensor, acts: torch.Tensor):
    """Backward chain of the reflectance network on tcgen05 (nrh_color_train_backward): dy [P,3] fp32, scale = device scalar S ->
    (dz [4,P,256] fp16, dy16 [P,8] fp16, dx [P,384] fp16), all in units of S."""
    import ctypes as C
    lib = _lib.load()
    dev, P = dy.device, dy.shape[0]
    packed, cfg = renderer._ensure_packed(dev), renderer._c_config()
    dy = dy.to(torch.float32).contiguous()
    dz = torch.empty(4, P, 256, dtype=torch.float16, device=dev)
    dy16 = torch.empty(P, 8, dtype=torch.float16, device=dev)
    dx = torch.empty(P, 384, dtype=torch.float16, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nrh_color_train_backward(C.byref(cfg), packed.data_ptr(), dy.data_ptr(), scale.data_ptr(), acts.data_ptr(), P,
                                                dz.data_ptr(), dy16.data_ptr(), dx.data_ptr(),
                                                torch.cuda.current_stream(dev).cuda_stream), "nrh_color_train_backward")
    return dz, dy16, dx


class _TrainLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, rgb_gt, normals, mask, igr_weight: float):
        lib = _lib.load()
        if not rgb.is_cuda:
            raise RuntimeError("nrhints_b200.train_loss_dict runs on CUDA tensors only (no CPU fallback)")
        dev = rgb.device
        c = lambda t: t.detach().to(torch.float32).contiguous()      # noqa: E731
        rgb_c, gt_c, n_c, m_c = c(rgb), c(rgb_gt), c(normals), c(mask)
        R, S = m_c.shape[0], m_c.shape[1]
        stats = torch.empty(8, dtype=torch.float32, device=dev)
        need_rgb, need_n = ctx.needs_input_grad[0], ctx.needs_input_grad[2]
        d_rgb = torch.empty_like(rgb_c) if need_rgb else None
        d_n = torch.empty_like(n_c) if need_n else None
        with torch.cuda.device(dev):
            _lib.check(lib.nrh_train_loss(rgb_c.data_ptr(), gt_c.data_ptr(), n_c.data_ptr(), m_c.data_ptr(), R, S, float(igr_weight), 1.0,
                                          stats.data_ptr(), d_rgb.data_ptr() if need_rgb else None, d_n.data_ptr() if need_n else None,
                                          torch.cuda.current_stream(dev).cuda_stream), "nrh_train_loss")
        ctx.save_for_backward(d_rgb, d_n)
        ctx.mark_non_differentiable(stats)
        return stats[0], stats

    @staticmethod
    def backward(ctx, g_loss, _g_stats):
        d_rgb, d_n = ctx.saved_tensors
        return (d_rgb * g_loss if d_rgb is not None else None, None, d_n * g_loss if d_n is not None else None, None, None)


def train_loss_dict(rendering_res, rgb_gt: torch.Tensor, igr_weight: float) -> Dict[str, torch.Tensor]:
    """get_train_loss_dict (pipelines/base_pipeline.py:50-69).  `rendering_res` needs .rgb, .analytic_normals,
    .relax_inside_sphere and .s_val; returns loss / rgb_loss / eikonal_loss / s_val / psnr as 0-d device tensors."""
    loss, stats = _TrainLossFn.apply(rendering_res.rgb, rgb_gt, rendering_res.analytic_normals, rendering_res.relax_inside_sphere,
                                     float(igr_weight))
    return {"loss": loss, "rgb_loss": stats[1], "eikonal_loss": stats[2], "s_val": rendering_res.s_val.mean(), "psnr": stats[3]}


class FlatAdam(torch.optim.Optimizer):
    """Adam over flat per-group buffers (see the module docstring).  Accepts torch.optim.Adam's param-group dicts
    ({'params': ..., 'lr': ...}); betas / eps / lr can be changed per group like in torch (LambdaLR works unchanged)."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, capturable: bool = False):
        """capturable=True keeps the step count and the learning rate of every group on the device (nrh_adam_step_dev), so that a
        CUDA graph holding the step stays valid from replay to replay; call `sync_lr()` before a replay when a scheduler changed
        `group['lr']`.  All parameters of a group must then take part in every step (no frozen subsets)."""
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.capturable = bool(capturable)
        self._flat: List[Dict[str, torch.Tensor]] = []
        for group in self.param_groups:
            ps = [p for p in group["params"]]
            if not ps:
                self._flat.append({})
                continue
            dev = ps[0].device
            if not all(p.is_cuda and p.device == dev and p.dtype == torch.float32 for p in ps):
                raise RuntimeError("FlatAdam needs every parameter of a group as fp32 on one CUDA device (no CPU fallback)")
            n = sum(p.numel() for p in ps)
            buf = {k: torch.zeros(n, dtype=torch.float32, device=dev) for k in ("param", "grad", "exp_avg", "exp_avg_sq")}
            off = 0
            with torch.no_grad():
                for p in ps:
                    k = p.numel()
                    buf["param"][off:off + k].copy_(p.detach().reshape(-1))
                    p.data = buf["param"][off:off + k].view_as(p)
                    p.grad = buf["grad"][off:off + k].view_as(p)
                    off += k
            buf["steps"] = [0] * len(ps)                       # per parameter, as torch.optim.Adam counts them
            if self.capturable:
                buf["step_dev"] = torch.zeros(1, dtype=torch.int64, device=dev)
                buf["lr_dev"] = torch.full((1,), float(group["lr"]), dtype=torch.float32, device=dev)
                buf["coef"] = torch.zeros(2, dtype=torch.float32, device=dev)
                buf["lr_host"] = float(group["lr"])
            self._flat.append(buf)

    def sync_lr(self):
        """capturable mode: push every group's current `lr` to its device scalar (one tiny fill per group whose lr changed)."""
        if not self.capturable:
            return
        for group, buf in zip(self.param_groups, self._flat):
            if buf and buf["lr_host"] != float(group["lr"]):
                buf["lr_dev"].fill_(float(group["lr"]))
                buf["lr_host"] = float(group["lr"])

    def flat_grads(self) -> List[torch.Tensor]:
        """One flat gradient tensor per parameter group: the all-reduce operand (nrhints_b200/grad_sync.py)."""
        return [b["grad"] for b in self._flat if b]

    def zero_grad(self, set_to_none: bool = False):
        """One memset per group.  The gradients must stay views of the flat buffer, so they are never set to None."""
        for group, buf in zip(self.param_groups, self._flat):
            if not buf:
                continue
            buf["grad"].zero_()
            off = 0
            for p in group["params"]:
                k = p.numel()
                if p.grad is None or p.grad.data_ptr() != buf["grad"].data_ptr() + 4 * off:
                    p.grad = buf["grad"][off:off + k].view_as(p)           # someone replaced it: re-attach
                off += k

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        """One launch per group in the usual case.  Like torch.optim.Adam, parameters that take no part in the optimisation are
        skipped: a parameter with requires_grad=False (SDFNetwork.freeze_geometry) keeps its value, its moments and its step
        count; the launch is then split into the contiguous runs of active parameters that share a step count."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group, buf in zip(self.param_groups, self._flat):
            if not buf:
                continue
            off = 0
            runs = []                                                        # [offset, numel, step] of contiguous active parameters
            for i, p in enumerate(group["params"]):                          # autograd may have re-bound .grad (first backward after set_to_none)
                k = p.numel()
                if p.grad is not None and p.grad.data_ptr() != buf["grad"].data_ptr() + 4 * off:
                    buf["grad"][off:off + k].copy_(p.grad.reshape(-1))
                    p.grad = buf["grad"][off:off + k].view_as(p)
                if p.requires_grad:
                    buf["steps"][i] += 1
                    if runs and runs[-1][0] + runs[-1][1] == off and runs[-1][2] == buf["steps"][i]:
                        runs[-1][1] += k
                    else:
                        runs.append([off, k, buf["steps"][i]])
                off += k
            b1, b2 = group["betas"]
            dev = buf["param"].device
            if self.capturable:
                if len(runs) != 1 or runs[0][1] != buf["param"].numel():
                    raise RuntimeError("FlatAdam(capturable=True) steps whole groups only (a parameter of the group is frozen)")
                if not torch.cuda.is_current_stream_capturing():
                    self.sync_lr()
                with torch.cuda.device(dev):
                    _lib.check(lib.nrh_adam_step_dev(buf["param"].data_ptr(), buf["grad"].data_ptr(), buf["exp_avg"].data_ptr(),
                                                     buf["exp_avg_sq"].data_ptr(), buf["param"].numel(), buf["lr_dev"].data_ptr(),
                                                     float(b1), float(b2), float(group["eps"]), buf["step_dev"].data_ptr(),
                                                     buf["coef"].data_ptr(), float(grad_scale),
                                                     torch.cuda.current_stream(dev).cuda_stream), "nrh_adam_step_dev")
                torch.autograd.graph.increment_version([p for p in group["params"] if p.requires_grad])
                continue
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream(dev).cuda_stream
                for o, k, st in runs:
                    _lib.check(lib.nrh_adam_step(buf["param"].data_ptr() + 4 * o, buf["grad"].data_ptr() + 4 * o,
                                                 buf["exp_avg"].data_ptr() + 4 * o, buf["exp_avg_sq"].data_ptr() + 4 * o, k,
                                                 float(group["lr"]), float(b1), float(b2), float(group["eps"]), int(st),
                                                 float(grad_scale), stream), "nrh_adam_step")
            # the kernel wrote through raw pointers: tell autograd / version-keyed caches (the renderer's packed weights)
            torch.autograd.graph.increment_version([p for p in group["params"] if p.requires_grad])
        return loss

    # torch.optim.Adam-compatible checkpoints (trainer/trainer.py:156,223)
    def _pull_steps(self):
        """capturable mode: the per-parameter step counts follow the device counter (graph replays advance it without Python)."""
        for buf in self._flat:
            if buf and self.capturable:
                buf["steps"] = [int(buf["step_dev"].item())] * len(buf["steps"])

    def state_dict(self):
        self._pull_steps()
        state, groups, idx = {}, [], 0
        for group, buf in zip(self.param_groups, self._flat):
            ids, off = [], 0
            for i, p in enumerate(group["params"]):
                k = p.numel()
                if buf and buf["steps"][i] > 0:
                    state[idx] = {"step": torch.tensor(float(buf["steps"][i])),
                                  "exp_avg": buf["exp_avg"][off:off + k].view_as(p).clone(),
                                  "exp_avg_sq": buf["exp_avg_sq"][off:off + k].view_as(p).clone()}
                ids.append(idx); idx += 1; off += k
            g = {k: v for k, v in group.items() if k != "params"}
            g["params"] = ids
            groups.append(g)
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        """Accepts torch.optim.Adam's layout; the layout is validated first (group count, parameters per group, moment shapes) so a
        mismatching checkpoint raises instead of filling the flat buffers with shifted data."""
        saved_groups = sd["param_groups"]
        if len(saved_groups) != len(self.param_groups):
            raise ValueError(f"loaded state dict has {len(saved_groups)} parameter groups, the optimizer has {len(self.param_groups)}")
        idx = 0
        for gi, (group, saved) in enumerate(zip(self.param_groups, saved_groups)):
            if len(saved["params"]) != len(group["params"]):
                raise ValueError(f"parameter group {gi}: loaded state dict holds {len(saved['params'])} parameters, "
                                 f"the optimizer {len(group['params'])}")
            for p in group["params"]:
                st = sd["state"].get(idx)
                if st is not None:
                    for k in ("exp_avg", "exp_avg_sq"):
                        if tuple(st[k].shape) != tuple(p.shape):
                            raise ValueError(f"parameter {idx}: {k} has shape {tuple(st[k].shape)}, the parameter {tuple(p.shape)}")
                idx += 1
        idx = 0
        for group, buf, saved in zip(self.param_groups, self._flat, saved_groups):
            for k, v in saved.items():
                if k != "params":
                    group[k] = v
            off = 0
            for i, p in enumerate(group["params"]):
                k = p.numel()
                st = sd["state"].get(idx)
                if st is not None:
                    buf["exp_avg"][off:off + k].copy_(st["exp_avg"].reshape(-1))
                    buf["exp_avg_sq"][off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                    buf["steps"][i] = int(st["step"])
                else:
                    buf["exp_avg"][off:off + k].zero_()
                    buf["exp_avg_sq"][off:off + k].zero_()
                    buf["steps"][i] = 0
                idx += 1; off += k
            if buf and self.capturable:
                buf["step_dev"].fill_(max(buf["steps"]) if buf["steps"] else 0)
                buf["lr_dev"].fill_(float(group["lr"])); buf["lr_host"] = float(group["lr"])
