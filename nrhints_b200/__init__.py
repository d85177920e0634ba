"""nrhints_b200 -- B200-native (sm_100a) implementation of the NRHints ray-march hot path.

Public surface (mirrors /root/reference/models/neus_hint_model.py for this path):
    NeuSHintRenderer, RenderOutput, NeuSModelConfig (+ sub-configs), RayBundle
The compute lives in nrhints_b200/csrc (hand-written CUDA behind the C ABI in include/nrhints_b200.h).
"""
from .config import (DepthComputationType, NeRFConfig, NeuSModelConfig, NeuSRendererConfig, NormalComputationType,
                     ReflectanceNetConfig, SDFNetConfig, SingleVarianceNetConfig)
from .ray_generator import CameraModel, RayGenerator, RayGeneratorConfig
from .rays import RayBundle
from .train_ops import FlatAdam, train_loss_dict
from .pipeline import NRHintPipeline
from .renderer import NeuSHintRenderer, ReflectanceNetwork, RenderOutput, SDFNetwork, SingleVarianceNetwork

__all__ = ["NeuSHintRenderer", "RenderOutput", "RayBundle", "NeuSModelConfig", "NeuSRendererConfig", "SDFNetConfig",
           "ReflectanceNetConfig", "SingleVarianceNetConfig", "NeRFConfig", "DepthComputationType", "NormalComputationType",
           "SDFNetwork", "ReflectanceNetwork", "SingleVarianceNetwork", "RayGenerator", "RayGeneratorConfig", "CameraModel",
           "FlatAdam", "train_loss_dict", "NRHintPipeline"]
