"""Developer tooling helper (GPU box): selects the developer build of the CUDA library (-DNRH_DEV, libnrhints_b200_dev.so) BEFORE
nrhints_b200 is imported and wraps its explicit hook API.  The production library has no hooks and reads no environment."""
import os
import sys

os.environ["NRH_DEV_LIB"] = "1"
sys.path.insert(0, "tests"); sys.path.insert(0, ".")

import torch                              # noqa: E402
from nrhints_b200 import _lib             # noqa: E402

assert _lib.DEV_BUILD


def configure(gen: int = 1, dbg: int = 0, fmask: int = 0xFF, tlog: torch.Tensor = None):
    """gen: engine generation (1 = shipped, 2 = TMEM operand / N-split, 3 = the same with the two-team epilogue);
    dbg: ablation code (see SdfTcParams::dbg); fmask: hand-off mask of generation 1; tlog: int64 CUDA tensor (>= 1024) or None."""
    lib = _lib.load()
    lib.nrh_dev_configure(int(gen), int(dbg), int(fmask), tlog.data_ptr() if tlog is not None else None)
