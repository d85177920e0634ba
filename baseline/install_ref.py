#!/usr/bin/env python
"""Install the UNMODIFIED reference (iamNCJ/NRHints) into baseline/_ref/ so that it can be timed and used as a live oracle on the
GPU box, where /root/reference does not exist.

    python baseline/install_ref.py [--src /root/reference]

The reference is a plain Python source tree (no setup.py / pyproject.toml, nothing to compile), so "installing" it is a byte-for-
byte copy of its .py files into baseline/_ref/reference/ plus a manifest (baseline/_ref/MANIFEST.json: sha256 of every copied
file, so `verify()` can prove at run time that what is timed is the reference as shipped).  baseline/_ref/ is git-ignored -- the
reference's sources never enter this repository's history -- but it is NOT gpurun-ignored, so it travels to the GPU box with the
snapshot.  Third-party modules the reference imports that are missing from this image (mcubes, torchmetrics, lpips, imageio,
trimesh, torchinfo: no network) are provided by the small shims under baseline/stubs/, which are this repository's own code.
"""
import argparse
import hashlib
import json
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
DEST = HERE / "_ref"


def sha256(p: Path) -> str:
    return hashlib.sha256(p.read_bytes()).hexdigest()


def install(src: Path = Path("/root/reference"), quiet: bool = False) -> Path:
    if not (src / "models" / "neus_hint_model.py").exists():
        raise FileNotFoundError(f"{src} does not look like the NRHints reference tree")
    dst = DEST / "reference"
    if dst.exists():
        shutil.rmtree(dst)
    manifest = {}
    for f in sorted(src.rglob("*")):
        rel = f.relative_to(src)
        if not f.is_file() or ".git" in rel.parts or f.suffix not in (".py", ".txt", ".md", ".sh", ".cff") and f.name != "LICENSE":
            continue
        out = dst / rel
        out.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(f, out)
        manifest[str(rel)] = sha256(out)
    (DEST / "MANIFEST.json").write_text(json.dumps({"source": str(src), "files": manifest}, indent=1))
    if not quiet:
        print(f"installed {len(manifest)} reference files into {dst}")
    return dst


def verify() -> bool:
    """True iff baseline/_ref/reference matches its manifest (nothing edited after the copy)."""
    mf = DEST / "MANIFEST.json"
    if not mf.exists():
        return False
    files = json.loads(mf.read_text())["files"]
    return all((DEST / "reference" / rel).exists() and sha256(DEST / "reference" / rel) == h for rel, h in files.items())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    install(Path(a.src))
    sys.exit(0 if verify() else 1)
