"""Minimal stand-in for the reference RayBundle (/root/reference/camera/ray_utils.py:214-235).

The renderer is duck-typed on `.origins .directions .pl_positions .nears .fars`, so the reference's own
RayBundle (a nerfstudio TensorDataclass) can be passed unchanged; this class exists so the package is
usable without the reference tree."""
from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class RayBundle:
    origins: torch.Tensor        # [R,3]
    directions: torch.Tensor     # [R,3] unit
    pl_positions: torch.Tensor   # [R,3]
    nears: Optional[torch.Tensor] = None   # [R,1]
    fars: Optional[torch.Tensor] = None    # [R,1]

    def to(self, device, non_blocking: bool = False) -> "RayBundle":
        return RayBundle(**{k: (v.to(device, non_blocking=non_blocking) if v is not None else None)
                            for k, v in self.__dict__.items()})

    def pin_memory(self) -> "RayBundle":
        return RayBundle(**{k: (v.pin_memory() if v is not None else None) for k, v in self.__dict__.items()})

    def __len__(self):
        return self.origins.shape[0]
