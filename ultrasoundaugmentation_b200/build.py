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
from concurrent.futures import ThreadPoolExecutor
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
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",   # only the USAUG_API entry points are exported
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


def build(force: bool = False, verbose: bool = False, extra_flags=(), out_path: Path | None = None) -> Path:
    """Compile csrc/*.cu into lib/libusaug.so unless an up-to-date build exists.

    ``extra_flags`` / ``out_path``: a differently configured library next to the product one (e.g. the bounds-checked
    debug build ``-DUSAUG_BOUNDS_CHECK`` -> lib/libusaug_dbg.so, selected at run time with USAUG_LIB)."""
    fp = _fingerprint()
    out_path = LIB_PATH if out_path is None else Path(out_path)
    if out_path == LIB_PATH and not extra_flags:
        if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == fp:
            return LIB_PATH
    LIB_DIR.mkdir(parents=True, exist_ok=True)
    obj_dir = LIB_DIR / ("obj" if out_path == LIB_PATH else "obj_" + out_path.stem)   # one object cache per configuration
    obj_dir.mkdir(exist_ok=True)
    nvcc = find_nvcc()
    extra = list(extra_flags) + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src: Path):
        # one translation unit per nvcc process, all in parallel; an object is reused while its own fingerprint
        # (source + every header + flags) is unchanged
        obj = obj_dir / (src.stem + ".o")
        tag = obj_dir / (src.stem + ".stamp")
        h = hashlib.sha256(src.read_bytes())
        for p in sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "usaug.h"]:
            h.update(p.read_bytes())
        h.update(" ".join(NVCC_FLAGS + extra).encode())
        if not force and obj.exists() and tag.exists() and tag.read_text() == h.hexdigest():
            return obj, ""
        cmd = [nvcc, *NVCC_FLAGS, *extra, f"-I{ROOT / 'include'}", f"-I{CSRC}", "-c", "-o", str(obj), str(src)]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        tag.write_text(h.hexdigest())
        return obj, proc.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, _sources()))
    if verbose:
        print("".join(log for _, log in results))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
           "-o", str(out_path), *[str(o) for o, _ in results]]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if out_path == LIB_PATH:
        STAMP.write_text(fp)
    return out_path


DEBUG_LIB_PATH = LIB_DIR / "libusaug_dbg.so"


def build_debug(force: bool = False) -> Path:
    """The bounds-checked library (csrc/common.cuh: red zones behind every workspace plane, device-side index asserts).
    Select it at run time with USAUG_LIB=<this path>; it synchronises at the end of every call."""
    return build(force=force, extra_flags=("-DUSAUG_BOUNDS_CHECK",), out_path=DEBUG_LIB_PATH)


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--debug" in sys.argv:
        print(build_debug(force="--force" in sys.argv))
