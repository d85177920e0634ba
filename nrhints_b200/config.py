"""Configuration surface of the hot path.

Field names, defaults and enum values mirror the reference so that a reference
`NeuSModelConfig` can be passed unchanged (duck-typed) and tyro/yaml dumps stay
interchangeable:
  NeuSRendererConfig / NeuSModelConfig   /root/reference/models/neus_hint_model.py:133-213
  SDFNetConfig                           /root/reference/fields/sdf_field.py:11-36
  ReflectanceNetConfig                   /root/reference/fields/reflectance_network.py:9-22
  SingleVarianceNetConfig                /root/reference/models/neus_hint_model.py:96-101
  NeRFConfig (outside model)             /root/reference/fields/nerf_density_field.py:12-26
"""
from dataclasses import dataclass, field
from enum import Enum
from typing import List


class DepthComputationType(Enum):
    AlphaBlend = 'alpha_blending'
    MaximalWeightPoint = 'maximum_point'
    SphereTracing = 'sphere_tracing'


class NormalComputationType(Enum):
    Analytic = 'analytic'
    NormalizedAnalytic = 'normalized_analytic'


@dataclass(frozen=True)
class SDFNetConfig:
    d_in: int = 3
    d_out_feat: int = 256
    d_hidden: int = 256
    n_layers: int = 8
    skip_in: List[int] = field(default_factory=lambda: [4])
    multi_res: int = 6
    init_bias: float = 0.5
    scale: float = 3.0
    geometric_init: bool = True
    weight_norm: bool = True
    inside_outside: bool = False


@dataclass(frozen=True)
class ReflectanceNetConfig:
    d_hidden: int = 256
    n_layers: int = 4
    weight_norm: bool = True
    multi_res: int = 4
    squeeze_out: bool = True


@dataclass(frozen=True)
class NeRFConfig:
    d_hidden: int = 256
    n_layers: int = 8
    multi_res: int = 10
    multi_res_view: int = 4
    skips: List[int] = field(default_factory=lambda: [4])


@dataclass(frozen=True)
class SingleVarianceNetConfig:
    init_val: float = 0.3


@dataclass(frozen=True)
class NeuSRendererConfig:
    use_outside_nerf: bool = False
    n_samples: int = 64
    n_importance_samples: int = 64
    n_outside_samples: int = 32
    normal_type: NormalComputationType = NormalComputationType.NormalizedAnalytic
    up_sample_steps: int = 4
    depth_type: DepthComputationType = DepthComputationType.AlphaBlend
    shadow_hint: bool = True
    force_shadow_map: bool = False
    specular_hint: bool = True
    force_specular_cue: bool = False
    shadow_ray_offset: float = 1e-2
    specular_roughness: List[float] = field(default_factory=lambda: [0.02, 0.05, 0.13, 0.34])
    shadow_hint_gradient: bool = False
    specular_hint_gradient: bool = False
    n_shadow_importance_clip: int = -1
    n_shadow_samples: int = 64
    n_shadow_importance_samples: int = 64
    override_near_far_to_sphere: bool = True


@dataclass(frozen=True)
class NeuSModelConfig:
    sdf_network: SDFNetConfig = field(default_factory=SDFNetConfig)
    outside_nerf: NeRFConfig = field(default_factory=NeRFConfig)
    deviation_network: SingleVarianceNetConfig = field(default_factory=SingleVarianceNetConfig)
    reflectance_network: ReflectanceNetConfig = field(default_factory=ReflectanceNetConfig)
    renderer: NeuSRendererConfig = field(default_factory=NeuSRendererConfig)

    igr_weight: float = 0.1
    lr: float = 5e-4
    lr_alpha: float = 0.05
    warm_up_end: int = 5_000
    end_iter: int = 1_000_000
    anneal_end: int = 50_000
    geometry_warmup_end: int = 0

    batch_size: int = 512
    shadow_mini_chunk_size: int = 2048
    training_chunk_size: int = 512
    inference_chunk_size: int = 512
