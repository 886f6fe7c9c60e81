"""In-tree build of libusaug.so (sm_100a only) with plain nvcc.

The shared library is a C-ABI object (include/usaug.h) with no libtorch / Python dependency,
so the build is one nvcc invocation.  The .so lands in ``ultrasoundaugmentation_b200/lib/`` (git-ignored,
but it travels to the GPU box with the gpurun snapshot).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
LIB_PATH = LIB_DIR / "libusaug.so"
STAMP = LIB_DIR / "libusaug.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    # no --use_fast_math: the label coordinate path must stay IEEE (bit-exact parity)
]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [ROOT / "include" / "usaug.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu into lib/libusaug.so unless an up-to-date build exists."""
    fp = _fingerprint()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == fp:
        return LIB_PATH
    LIB_DIR.mkdir(parents=True, exist_ok=True)
    cmd = [find_nvcc(), *NVCC_FLAGS, f"-I{ROOT / 'include'}", f"-I{CSRC}", "-o", str(LIB_PATH)]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [str(s) for s in _sources()]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    STAMP.write_text(fp)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
