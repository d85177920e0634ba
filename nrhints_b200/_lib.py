"""ctypes binding of libnrhints_b200.so (the C ABI declared in include/nrhints_b200.h).

The library is plain CUDA-runtime code with no torch types in its signatures; PyTorch is used by
the callers only to own device memory and streams.  There is NO CPU fallback: if the shared
library is missing the import of the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

_CSRC = Path(__file__).resolve().parent / "csrc"
# NRH_DEV_LIB=1 (developer tools under tests/tc_*.py only) selects the developer build: the same sources compiled with -DNRH_DEV,
# which adds kernel timelines / ablation switches behind the explicit nrh_dev_configure() call.  The production library has no
# such hooks and reads nothing from the environment.
DEV_BUILD = os.environ.get("NRH_DEV_LIB", "0") == "1"
_LIB_PATH = _CSRC / ("libnrhints_b200_dev.so" if DEV_BUILD else "libnrhints_b200.so")
_SOURCES = ["api.cu", "mlp_simt.cu", "sampler_kernels.cu", "mlp_tc.cu", "hash_encode.cu", "raygen.cu", "train_ops.cu", "wgrad_tc.cu", "train_step.cu"]
_HEADERS = ["nrh_common.cuh", "ray_math.cuh", "sampler_kernels.cuh", "mlp_tc.cuh", "tc_primitives.cuh", "mlp_tc_bwd.inc", "mlp_tc2.inc", "color_train_tc.inc", "raygen_math.cuh", "composite_train_math.cuh", "train_step.cuh",
            "../../include/nrhints_b200.h"]

NRH_ABI_VERSION = 7
NRH_MAX_ROUGHNESS = 4
NRH_MAX_OUTSIDE = 64
NRH_MLP_AUTO, NRH_MLP_FP32_SIMT, NRH_MLP_TCGEN05 = 0, 1, 2
DEPTH_TYPES = {"alpha_blending": 0, "maximum_point": 1, "sphere_tracing": 2}
MLP_IMPLS = {"auto": NRH_MLP_AUTO, "fp32": NRH_MLP_FP32_SIMT, "tcgen05": NRH_MLP_TCGEN05}


class NrhConfig(C.Structure):
    _fields_ = [
        ("n_samples", C.c_int32), ("n_importance", C.c_int32), ("up_sample_steps", C.c_int32),
        ("n_shadow_samples", C.c_int32), ("n_shadow_importance", C.c_int32),
        ("shadow_hint", C.c_int32), ("specular_hint", C.c_int32), ("n_roughness", C.c_int32),
        ("roughness", C.c_float * NRH_MAX_ROUGHNESS), ("shadow_ray_offset", C.c_float),
        ("normalized_normals", C.c_int32), ("mlp_impl", C.c_int32), ("depth_type", C.c_int32),
        ("use_outside_nerf", C.c_int32), ("n_outside", C.c_int32),
    ]


class NrhRawWeights(C.Structure):
    _fields_ = [
        ("sdf_W", C.c_void_p * 8), ("sdf_b", C.c_void_p * 8),
        ("sdf_out_W", C.c_void_p), ("sdf_out_b", C.c_void_p),
        ("feat_W", C.c_void_p), ("feat_b", C.c_void_p),
        ("col_W", C.c_void_p * 5), ("col_b", C.c_void_p * 5),
        ("variance", C.c_void_p),
        ("nerf_W", C.c_void_p * 8), ("nerf_b", C.c_void_p * 8),
        ("nerf_alpha_W", C.c_void_p), ("nerf_alpha_b", C.c_void_p),
        ("nerf_feat_W", C.c_void_p), ("nerf_feat_b", C.c_void_p),
        ("nerf_view_W", C.c_void_p), ("nerf_view_b", C.c_void_p),
        ("nerf_rgb_W", C.c_void_p), ("nerf_rgb_b", C.c_void_p),
    ]


class NrhRays(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("origins", "directions", "pl_positions", "nears", "fars", "hit_points", "hit_depths")]


class NrhOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "rgb", "depth", "weights", "inside_sphere", "analytic_normals", "normalized_normals",
        "visibilities", "specular_cue", "inv_s", "z_vals", "z_shadow", "sampled_color",
        "normal_map", "normalized_normal_map", "specular_cue_ray", "early_event", "train_capture", "fine_begin_event", "fine_end_event")]


class NrhTrainCapture(C.Structure):
    _fields_ = [("tape", C.c_void_p), ("tape_bytes", C.c_size_t), ("sdf", C.c_void_p), ("grad_soa", C.c_void_p),
                ("feat", C.c_void_p), ("pts_soa", C.c_void_p), ("feat16", C.c_void_p), ("feat16_ld", C.c_int64)]


class NrhTrainLayout(C.Structure):
    _fields_ = [("p_pad", C.c_int64)] + [(n, C.c_uint64) for n in (
        "tape_tiles_off", "tape_act_off", "tape_u_off", "tape_bytes", "bwd_gb0_off", "bwd_gb_off", "bwd_zb_off", "bwd_bytes",
        "bwd_workspace_bytes")]


class NrhCamera(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("fx", "fy", "cx", "cy", "zn", "zf")]


class NrhRayGenInputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w_indices", "h_indices", "img_indices", "poses", "pls", "cam_pose_noise", "pl_noise",
                                          "cam_pose_adjustment", "pl_adjustment")] + [("n_cameras", C.c_int64)]


class NrhWgradJob(C.Structure):
    _fields_ = [("a", C.c_void_p), ("a_ld", C.c_int64), ("a_col0", C.c_int32), ("b", C.c_void_p), ("b_ld", C.c_int64), ("b_col0", C.c_int32),
                ("rows", C.c_int64), ("m", C.c_int32), ("n", C.c_int32), ("rows_valid", C.c_int32), ("cols_valid", C.c_int32),
                ("scale", C.c_float), ("dev_scale", C.c_void_p), ("out", C.c_void_p), ("ld_out", C.c_int64)]


class NrhLayerParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("v", "g", "bias", "d_v", "d_g", "d_bias")]


class NrhTrainParams(C.Structure):
    _fields_ = [("sdf", NrhLayerParams * 8), ("sdf_out", NrhLayerParams), ("feat_out", NrhLayerParams), ("col", NrhLayerParams * 5),
                ("variance", C.c_void_p), ("d_variance", C.c_void_p)]


class NrhTrainAdjoints(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d_rgb", "d_analytic_normals", "d_normalized_normals", "d_weights", "d_origins",
                                          "d_directions", "d_pl_positions")]


CAM_OPT_MODES = {"off": 0, "SO3xR3": 1, "SE3": 2}

EXPORTS = {
    "nrh_version": (C.c_int, []),
    "nrh_last_error": (C.c_char_p, []),
    "nrh_check_config": (C.c_int, [C.POINTER(NrhConfig)]),
    "nrh_packed_weights_bytes": (C.c_size_t, [C.POINTER(NrhConfig)]),
    "nrh_pack_weights": (C.c_int, [C.POINTER(NrhConfig), C.POINTER(NrhRawWeights), C.c_void_p, C.c_size_t, C.c_void_p]),
    "nrh_workspace_bytes": (C.c_size_t, [C.POINTER(NrhConfig), C.c_int64]),
    "nrh_query_workspace_bytes": (C.c_size_t, [C.POINTER(NrhConfig), C.c_int64]),
    "nrh_render_forward": (C.c_int, [C.POINTER(NrhConfig), C.c_void_p, C.POINTER(NrhRays), C.c_int64,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int,
                                     C.POINTER(NrhOutputs), C.c_void_p, C.c_size_t, C.c_void_p]),
    "nrh_sdf_query": (C.c_int, [C.POINTER(NrhConfig), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nrh_sphere_trace": (C.c_int, [C.POINTER(NrhConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_float,
                                   C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nrh_sdf_train_layout": (C.c_int, [C.POINTER(NrhConfig), C.c_int64, C.POINTER(NrhTrainLayout)]),
    "nrh_sdf_train_forward": (C.c_int, [C.POINTER(NrhConfig), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nrh_sdf_train_backward": (C.c_int, [C.POINTER(NrhConfig), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                         C.c_void_p, C.c_size_t, C.c_void_p]),
    "nrh_hash_encode": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p]),
    "nrh_hash_encode_backward": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p]),
    "nrh_raygen_forward": (C.c_int, [C.POINTER(NrhCamera), C.c_int, C.c_int, C.POINTER(NrhRayGenInputs), C.c_int64,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrh_raygen_backward": (C.c_int, [C.POINTER(NrhCamera), C.c_int, C.c_int, C.POINTER(NrhRayGenInputs), C.c_int64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrh_train_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_float,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrh_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double,
                                C.c_double, C.c_int64, C.c_float, C.c_void_p]),
    "nrh_adam_step_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_double, C.c_double,
                                    C.c_double, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "nrh_colsum_f16": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_float, C.c_void_p, C.c_void_p]),
    "nrh_composite_train_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_float, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrh_composite_train_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_float, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrh_wgrad_f16": (C.c_int, [C.POINTER(NrhWgradJob), C.c_int, C.c_void_p]),
    "nrh_color_train_forward": (C.c_int, [C.POINTER(NrhConfig), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrh_color_train_backward": (C.c_int, [C.POINTER(NrhConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrh_train_workspace_bytes": (C.c_size_t, [C.POINTER(NrhConfig), C.c_int64]),
    "nrh_pack_weights_wn": (C.c_int, [C.POINTER(NrhConfig), C.POINTER(NrhTrainParams), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                      C.c_void_p]),
    "nrh_render_train_forward": (C.c_int, [C.POINTER(NrhConfig), C.c_void_p, C.POINTER(NrhRays), C.c_int64, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_float, C.c_int, C.POINTER(NrhOutputs), C.c_void_p, C.c_size_t, C.c_void_p]),
    "nrh_render_backward": (C.c_int, [C.POINTER(NrhConfig), C.c_void_p, C.POINTER(NrhTrainParams), C.POINTER(NrhRays), C.c_int64,
                                      C.c_void_p, C.c_float, C.POINTER(NrhTrainAdjoints), C.c_void_p, C.c_size_t, C.c_void_p]),
    "nrh_last_launch_count": (C.c_int, []),
}


_OBJ_DIR = _CSRC / ("_obj_dev" if DEV_BUILD else "_obj")
_NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"] + (
    ["-DNRH_DEV"] if DEV_BUILD else [])


def nvcc_command(out: Path = _LIB_PATH):
    """The one-shot equivalent of build(): every source straight into the shared library."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    return [nvcc] + _NVCC_FLAGS + ["-shared", "-o", str(out)] + [str(_CSRC / s) for s in _SOURCES]


def needs_build() -> bool:
    if not _LIB_PATH.exists():
        return True
    t = _LIB_PATH.stat().st_mtime
    return any((_CSRC / s).resolve().stat().st_mtime > t for s in _SOURCES + _HEADERS)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU): one object per source, compiled
    in parallel and re-used while neither the source nor any header is newer, then linked into libnrhints_b200.so."""
    if not (force or needs_build()):
        return _LIB_PATH
    import fcntl
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    _OBJ_DIR.mkdir(exist_ok=True)
    # one builder at a time: the ranks of a torchrun launch may all find the library stale at once
    lock = open(_OBJ_DIR / ".build.lock", "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and not needs_build():            # another process built it while we waited
            return _LIB_PATH
        return _build_locked(nvcc, force, verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(nvcc: str, force: bool, verbose: bool) -> Path:
    from concurrent.futures import ThreadPoolExecutor
    hdr_t = max((_CSRC / h).resolve().stat().st_mtime for h in _HEADERS)

    def compile_one(src: str) -> Path:
        obj = _OBJ_DIR / (src[:-3] + ".o")
        stale = force or not obj.exists() or obj.stat().st_mtime < max(hdr_t, (_CSRC / src).stat().st_mtime)
        if stale:
            cmd = [nvcc] + _NVCC_FLAGS + ["-c", "-o", str(obj), str(_CSRC / src)]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=len(_SOURCES)) as ex:
        objs = list(ex.map(compile_one, _SOURCES))
    tmp = _LIB_PATH.with_suffix(".so.tmp")
    cmd = [nvcc, "-shared", "-o", str(tmp)] + [str(o) for o in objs]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    os.replace(tmp, _LIB_PATH)                          # readers never see a half-written library
    return _LIB_PATH


_lib = None


def load():
    """Load the shared library (building it first if nvcc is around and sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    if needs_build():
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        if os.path.exists(nvcc):
            build()
    if not _LIB_PATH.exists():
        raise ImportError(f"nrhints_b200: CUDA library {_LIB_PATH} is missing and could not be built; "
                          "run `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)          # AttributeError here = ABI drift: fail loudly
        fn.restype, fn.argtypes = res, args
    if DEV_BUILD:
        lib.nrh_dev_configure.restype = C.c_int
        lib.nrh_dev_configure.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_void_p]
    _lib = lib
    return lib


class NrhError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != 0:
        msg = load().nrh_last_error()
        raise NrhError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
