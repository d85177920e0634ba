"""Import the unmodified reference from baseline/_ref/reference (see install_ref.py).  Used by bench.py's reference / GPU-baseline
arms and by the drop-in tests (tests/test_reference_dropin.py); never by the product path."""
import importlib
import sys
from pathlib import Path
from types import SimpleNamespace

HERE = Path(__file__).resolve().parent
REF = HERE / "_ref" / "reference"
STUBS = HERE / "stubs"


def available() -> bool:
    return (REF / "models" / "neus_hint_model.py").exists()


def load(verify: bool = True) -> SimpleNamespace:
    """-> namespace with the reference modules (model, ray_utils, tensor_dataclass; pipeline pieces are imported lazily by
    `load_pipeline`).  Shims from baseline/stubs are appended to sys.path, so a real installation of a module always wins."""
    if not available():
        raise ImportError("baseline/_ref/reference is missing: run `python baseline/install_ref.py` where /root/reference exists")
    if verify:
        sys.path.insert(0, str(HERE))
        import install_ref
        if not install_ref.verify():
            raise ImportError("baseline/_ref/reference does not match its manifest (edited after installation?)")
    for p in (str(STUBS), str(REF)):
        if p not in sys.path:
            sys.path.append(p) if p == str(STUBS) else sys.path.insert(0, p)
    M = importlib.import_module("models.neus_hint_model")
    RU = importlib.import_module("camera.ray_utils")
    TD = importlib.import_module("utils.tensor_dataclass")
    return SimpleNamespace(model=M, ray_utils=RU, tensor_dataclass=TD, NeuSHintRenderer=M.NeuSHintRenderer,
                           NeuSModelConfig=M.NeuSModelConfig, NeuSRendererConfig=M.NeuSRendererConfig, RayBundle=RU.RayBundle, root=REF)


def load_pipeline() -> SimpleNamespace:
    ns = load()
    P = importlib.import_module("pipelines.base_pipeline")
    C = importlib.import_module("configs.main_config")
    RG = importlib.import_module("camera.ray_generator")
    CM = importlib.import_module("camera.camera_model")
    DL = importlib.import_module("data.data_loader")
    ns.pipeline, ns.configs, ns.ray_generator, ns.camera_model, ns.data_loader = P, C, RG, CM, DL
    return ns
