"""Copy the reference checkout's python sources into ``baseline/_ref`` (git-ignored, shipped by gpurun).

    python baseline/vendor.py            # needs /root/reference; no-op when it is absent

Only ``src/`` and ``configs/detrpose/`` are taken, byte for byte: the GPU box has no ``/root/reference``,
and the whole-model parity tests and the ``model_e2e`` leg of ``bench.py`` run the reference's own model
code with this package's kernels dropped in.  Nothing is modified, nothing from here is
committed.
"""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def vendor(verbose: bool = True) -> bool:
    if not os.path.isdir(os.path.join(SRC, "src")):
        if verbose:
            print(f"{SRC} not present: keeping {DST} as it is")
        return os.path.isdir(os.path.join(DST, "src"))
    os.makedirs(DST, exist_ok=True)
    for sub in ("src", os.path.join("configs", "detrpose")):
        dst = os.path.join(DST, sub)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, sub), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for name in ("LICENSE", "README.md"):
        if os.path.exists(os.path.join(SRC, name)):
            shutil.copy2(os.path.join(SRC, name), os.path.join(DST, name))
    if verbose:
        print(f"vendored {SRC}/{{src,configs/detrpose}} -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if vendor() else 1)
