"""SDF fine pass with a hand-written CUDA backward (training step, BASELINE config #3).

`sdf_fine(renderer, pts, weights)` returns (sdf [N,1], feat [N,256], grad [N,3]) -- what the reference computes with
`sdf_network(pts)` and `sdf_network.gradient(pts)` (create_graph=True) in render_core / get_alpha
(/root/reference/models/neus_hint_model.py:504-508, :335-336; /root/reference/fields/sdf_field.py:106-148) -- as ONE
autograd node:

  forward : nrh_sdf_train_forward  (tcgen05 fused MLP: forward + feature head + reverse sweep, writing a tape)
  backward: nrh_sdf_train_backward (tcgen05, phase A / phase B chains incl. the second-order terms; csrc/mlp_tc_bwd.inc)
            -> d_pts and fp16 operand dumps; the weight gradients are point-reductions over the dumps: all 19 of them in ONE
            hand-written tcgen05 launch (nrh_wgrad_f16, csrc/wgrad_tc.cu: TMA tensor-map loads, accumulators in tensor memory).

The effective (weight-normed) weights enter as autograd inputs so that torch differentiates the weight-norm
re-parametrisation itself (as in the reference); their VALUES are taken from the renderer's packed weight buffer.
Math spec: oracle/nrh_oracle.py::sdf_mlp_backward.  CUDA + tcgen05 engine only; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _lib
from .train_ops import WgradBatch, colsum_f16

_ACT_SCALE = 16.0
_G_SCALE = 1024.0
_SDF_SCALE = 3.0


def _fourier(x: torch.Tensor, n_freq: int) -> torch.Tensor:
    freqs = 2 ** torch.linspace(0.0, n_freq - 1, n_freq, device=x.device, dtype=x.dtype)
    s = (x[..., None] * freqs).reshape(*x.shape[:-1], -1)
    return torch.cat([x, torch.sin(torch.cat([s, s + torch.pi / 2.0], dim=-1))], dim=-1)


class _SdfFine(torch.autograd.Function):
    @staticmethod
    def forward(ctx, renderer, pts, captured, *weights):
        lib = _lib.load()
        device = pts.device
        packed = renderer._ensure_packed(device)
        cfg = renderer._c_config()
        x = pts.detach().to(torch.float32).contiguous()
        N = x.shape[0]
        lay = _lib.NrhTrainLayout()
        _lib.check(lib.nrh_sdf_train_layout(C.byref(cfg), N, C.byref(lay)), "nrh_sdf_train_layout")
        if captured is not None:
            # nrh_render_forward already ran this forward as its primary fine pass (NrhTrainCapture): adopt tape and results
            tape, sdf, grad, feat = captured["tape"], captured["sdf"], captured["grad"], captured["feat"]
            assert tape.numel() >= int(lay.tape_bytes) and sdf.numel() == N
            x = captured["pts"]                                # exactly the points the forward kernel evaluated
        else:
            tape = torch.empty(int(lay.tape_bytes), dtype=torch.uint8, device=device)
            sdf = torch.empty(N, dtype=torch.float32, device=device)
            grad = torch.empty(N, 3, dtype=torch.float32, device=device)
            feat = torch.empty(N, 256, dtype=torch.float32, device=device)
            ws = renderer._ensure_workspace(lib.nrh_query_workspace_bytes(C.byref(cfg), N), device)
            with torch.cuda.device(device):
                stream = torch.cuda.current_stream(device).cuda_stream
                _lib.check(lib.nrh_sdf_train_forward(C.byref(cfg), packed.data_ptr(), x.data_ptr(), N, sdf.data_ptr(), grad.data_ptr(),
                                                     feat.data_ptr(), tape.data_ptr(), tape.numel(), ws.data_ptr(), ws.numel(), stream),
                           "nrh_sdf_train_forward")
        ctx.renderer, ctx.lay, ctx.N = renderer, lay, N
        ctx.packed = packed
        ctx.save_for_backward(x, tape)
        return sdf[:, None], feat, grad

    @staticmethod
    def backward(ctx, d_sdf, d_feat, d_grad):
        x, tape = ctx.saved_tensors
        renderer, lay, N = ctx.renderer, ctx.lay, ctx.N
        lib = _lib.load()
        device = x.device
        cfg = renderer._c_config()
        f32 = dict(dtype=torch.float32, device=device)
        d_sdf = (d_sdf if d_sdf is not None else torch.zeros(N, 1, **f32)).to(torch.float32).contiguous()
        d_feat = (d_feat if d_feat is not None else torch.zeros(N, 256, **f32)).to(torch.float32).contiguous()
        d_grad = (d_grad if d_grad is not None else torch.zeros(N, 3, **f32)).to(torch.float32).contiguous()
        # power-of-two loss scale, computed on the device (no host sync): largest adjoint -> ~2^9
        inf = float("inf")                                     # max |x| in one reduction pass, without the 512 MB |d_feat| temporary
        amax = torch.stack([torch.linalg.vector_norm(d_sdf, ord=inf), torch.linalg.vector_norm(d_feat, ord=inf),
                            torch.linalg.vector_norm(d_grad, ord=inf) * _SDF_SCALE]).max().clamp_min(1e-30)
        scale = torch.exp2(torch.floor(torch.log2(512.0 / amax))).clamp(2.0 ** -40, 2.0 ** 40).reshape(1).contiguous()
        bwd = torch.empty(int(lay.bwd_bytes), dtype=torch.uint8, device=device)
        d_pts = torch.empty(N, 3, **f32)
        ws = renderer._ensure_workspace(int(lay.bwd_workspace_bytes), device)
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            _lib.check(lib.nrh_sdf_train_backward(C.byref(cfg), ctx.packed.data_ptr(), x.data_ptr(), N, tape.data_ptr(), tape.numel(),
                                                  d_sdf.data_ptr(), d_feat.data_ptr(), d_grad.data_ptr(), scale.data_ptr(),
                                                  bwd.data_ptr(), bwd.numel(), d_pts.data_ptr(), ws.data_ptr(), ws.numel(), stream),
                       "nrh_sdf_train_backward")
        P = int(lay.p_pad)

        def view(buf, off, n, width=256):
            return buf[off:off + n * P * width * 2].view(torch.float16).view(n, P, width)
        act = view(tape, int(lay.tape_act_off), 8)          # a_1 .. a_8, x16
        u = view(tape, int(lay.tape_u_off), 8)              # u_0 .. u_7, x1024
        gb0 = view(bwd, int(lay.bwd_gb0_off), 1, 64)[0]     # gb_0 (S units)
        gb = view(bwd, int(lay.bwd_gb_off), 8)              # gb_1 .. gb_8
        zb = view(bwd, int(lay.bwd_zb_off), 8)              # zb_0 .. zb_7
        inv_s = (1.0 / scale).reshape(1).contiguous()        # device scalar: the reductions below multiply by it (no host sync)
        # bias gradients: column sums of the fp16 dumps, all eight layers in ONE launch (nrh_colsum_f16) + the sdf head's
        db_all = colsum_f16(zb) * inv_s                      # [8,256]
        gb8_sum = colsum_f16(gb[7]) * inv_s                  # [256]
        # weight gradients: dW_l = u_l^T gb_{l-1} / (1024 S) + zb_l^T a_{l-1} / (16 S), the feature head and the d_sdf row of the sdf
        # head -- 19 point-reductions over the dumps, ALL in one tcgen05 launch (nrh_wgrad_f16, csrc/wgrad_tc.cu)
        z32 = lambda r, c: torch.zeros(r, c, dtype=torch.float32, device=device)        # noqa: E731
        wb = WgradBatch()
        # a_0 = PE(3 pts) enters as an fp16 operand like every other activation (|PE| <= 1.5: no scaling needed)
        e = torch.zeros(P, 64, dtype=torch.float16, device=device)
        e[:N, :39] = _fourier(x * _SDF_SCALE, 6).to(torch.float16)
        dWs = [z32(256, 64)] + [z32(256, 256) for _ in range(7)]
        wb.add(u[0], gb0, dWs[0], scale=1.0 / _G_SCALE, dev_scale=inv_s).add(zb[0], e, dWs[0], dev_scale=inv_s)
        for l in range(1, 8):
            rv = 217 if l == 3 else 0
            wb.add(u[l], gb[l - 1], dWs[l], scale=1.0 / _G_SCALE, dev_scale=inv_s, rows_valid=rv)
            wb.add(zb[l], act[l - 1], dWs[l], scale=1.0 / _ACT_SCALE, dev_scale=inv_s, rows_valid=rv)
        a8 = act[7]
        # head operands: d_feat * S as fp16 [P,256] (feature head) and d_sdf * S as column 0 of an fp16 [P,8] matrix (sdf head row)
        df16 = torch.zeros(P, 256, dtype=torch.float16, device=device)
        df16[:N] = d_feat * scale
        ds16 = torch.zeros(P, 8, dtype=torch.float16, device=device)
        ds16[:N, 0] = d_sdf.reshape(N) * scale
        dW_f, head_s = z32(256, 256), z32(1, 256)
        wb.add(df16, a8, dW_f, scale=1.0 / _ACT_SCALE, dev_scale=inv_s)
        wb.add(ds16, a8, head_s, scale=1.0 / _ACT_SCALE, dev_scale=inv_s, m=1)
        wb.run()
        grads: List[torch.Tensor] = []
        for l in range(8):
            dW, db = dWs[l], db_all[l]
            if l == 0:
                dW = dW[:, :39]
            if l == 3:
                dW, db = dW[:217], db[:217]
            grads += [dW, db]
        d_ws = (gb8_sum + head_s[0]) / _SDF_SCALE
        d_bs = d_sdf.sum().reshape(1) / _SDF_SCALE
        db_f = d_feat.sum(0)
        grads += [d_ws.reshape(1, 256), d_bs, dW_f, db_f]
        return (None, d_pts, None) + tuple(grads)


def sdf_fine(renderer, pts: torch.Tensor, weights: dict, captured=None):
    """weights: dict(sdf_w, sdf_b: lists of 8; sdf_w_head, sdf_b_head, feat_w, feat_b) of EFFECTIVE weights (autograd-connected).
    captured: dict(tape, sdf [N], grad [N,3], feat [N,256]) from a nrh_render_forward call with NrhTrainCapture over the SAME
    points in the same order -- the forward kernel is then not launched again."""
    flat = []
    for l in range(8):
        flat += [weights["sdf_w"][l], weights["sdf_b"][l]]
    flat += [weights["sdf_w_head"], weights["sdf_b_head"], weights["feat_w"], weights["feat_b"]]
    return _SdfFine.apply(renderer, pts, captured, *flat)
