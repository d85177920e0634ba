"""Host-side mirror of the reference's multi-resolution hash encoding.

`HashEncoding` keeps the constructor, attributes (`hash_table` parameter, `scalings`, `hash_offset`) and output layout of
/root/reference/fields/encodings.py:237-371 (torch fallback `pytorch_fwd`, :324-366), evaluated by the CUDA operator
`nrh_hash_encode` (nrhints_b200/csrc/hash_encode.cu) behind the C ABI.  The reference never instantiates this encoder
(the SDF / reflectance / outside networks hard-wire the Fourier encoding), so it is a standalone
operator here as well; the Fourier encoding that IS on the hot path lives inside the fused MLP kernels.

There is no CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
from torch import nn

from . import _lib


class _HashEncodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pts, table, enc):
        lib = _lib.load()
        N = pts.shape[0]
        out = torch.empty(N, enc.num_levels * enc.features_per_level, dtype=torch.float32, device=pts.device)
        with torch.cuda.device(pts.device):
            stream = torch.cuda.current_stream(pts.device).cuda_stream
            _lib.check(lib.nrh_hash_encode(pts.data_ptr(), N, table.data_ptr(), enc._c_scalings, enc.num_levels,
                                           enc.log2_hashmap_size, enc.features_per_level, out.data_ptr(), stream),
                       "nrh_hash_encode")
        ctx.save_for_backward(pts)
        ctx.enc, ctx.table_shape = enc, table.shape
        return out

    @staticmethod
    def backward(ctx, d_out):
        (pts,) = ctx.saved_tensors
        enc = ctx.enc
        lib = _lib.load()
        d_table = torch.zeros(ctx.table_shape, dtype=torch.float32, device=pts.device)
        g = d_out.contiguous().to(torch.float32)
        with torch.cuda.device(pts.device):
            stream = torch.cuda.current_stream(pts.device).cuda_stream
            _lib.check(lib.nrh_hash_encode_backward(pts.data_ptr(), pts.shape[0], g.data_ptr(), enc._c_scalings, enc.num_levels,
                                                    enc.log2_hashmap_size, enc.features_per_level, d_table.data_ptr(), stream),
                       "nrh_hash_encode_backward")
        return None, d_table, None


class HashEncoding(nn.Module):
    """Instant-NGP style hash encoding with the reference's torch-fallback semantics (fields/encodings.py:237-371):
    16 levels from `min_res` to `max_res` (floor of a geometric progression), table of 2^log2_hashmap_size entries per
    level, int64 hash without 32-bit wrap, trilinear weights toward the ceil corner.  forward: [..., 3] in [0, 1] ->
    [..., num_levels * features_per_level].  Gradients flow to `hash_table` (fp32 atomic scatter-add); gradients w.r.t. the
    input positions are not implemented and raise."""

    def __init__(self, num_levels: int = 16, min_res: int = 16, max_res: int = 1024, log2_hashmap_size: int = 19,
                 features_per_level: int = 2, hash_init_scale: float = 0.001, implementation: str = "torch",
                 interpolation: Optional[str] = None) -> None:
        super().__init__()
        assert interpolation is None or interpolation == "Linear", \
            f"interpolation '{interpolation}' is not supported (the reference's torch backend has the same limit)"
        if features_per_level not in (1, 2, 4, 8) or num_levels > 32:
            raise NotImplementedError("nrh_hash_encode supports features_per_level in {1,2,4,8} and at most 32 levels")
        self.in_dim = 3
        self.num_levels = num_levels
        self.features_per_level = features_per_level
        self.log2_hashmap_size = log2_hashmap_size
        self.hash_table_size = 2 ** log2_hashmap_size
        levels = torch.arange(num_levels)
        growth_factor = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1.0
        self.scalings = torch.floor(min_res * growth_factor ** levels)              # fp32, as in the reference
        self.hash_offset = levels * self.hash_table_size
        table = torch.rand(size=(self.hash_table_size * num_levels, features_per_level)) * 2 - 1
        table *= hash_init_scale
        self.hash_table = nn.Parameter(table)
        self._c_scalings = (C.c_float * num_levels)(*[float(v) for v in self.scalings.to(torch.float32)])

    def get_out_dim(self) -> int:
        return self.num_levels * self.features_per_level

    def forward(self, in_tensor: torch.Tensor) -> torch.Tensor:
        assert in_tensor.shape[-1] == 3
        if in_tensor.device.type != "cuda" or self.hash_table.device != in_tensor.device:
            raise RuntimeError("nrhints_b200.HashEncoding runs on CUDA devices only (there is no CPU fallback); got input on "
                               f"{in_tensor.device}, table on {self.hash_table.device}")
        if in_tensor.requires_grad and torch.is_grad_enabled():
            raise NotImplementedError("HashEncoding: gradients w.r.t. the input positions are not implemented")
        pts = in_tensor.detach().to(torch.float32).reshape(-1, 3).contiguous()
        out = _HashEncodeFn.apply(pts, self.hash_table, self)
        return out.reshape(*in_tensor.shape[:-1], self.get_out_dim())
