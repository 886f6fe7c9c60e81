"""The sanitizer substitute (compute-sanitizer is closed on the GPU pool): lib/libusaug_dbg.so is the same library built
with -DUSAUG_BOUNDS_CHECK -- a red zone behind every workspace plane, verified at the end of each call, and device-side
asserts on the kernels' index computations (csrc/common.cuh).  The odd-shape, fuzz and solver-variant parity tests are run
on it in a child process (the library is chosen per process by USAUG_LIB)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def debug_lib():
    from ultrasoundaugmentation_b200 import build
    return build.build_debug()


def test_debug_library_builds_and_exports_the_same_abi(debug_lib):
    import ctypes
    from ultrasoundaugmentation_b200 import _ffi
    lib = ctypes.CDLL(str(debug_lib))
    for name in _ffi.SYMBOLS:
        assert hasattr(lib, name), name
    # the red zones are part of the workspace: the debug build asks for more
    prod = _ffi.load()
    lib.usaug_warp_workspace_bytes.restype = ctypes.c_size_t
    lib.usaug_warp_workspace_bytes.argtypes = [ctypes.c_int] * 3
    assert lib.usaug_warp_workspace_bytes(2, 64, 64) == prod.usaug_warp_workspace_bytes(2, 64, 64) + 3 * 256


def _child(debug_lib, args, extra_env=None, timeout=900):
    env = dict(os.environ, USAUG_LIB=str(debug_lib), PYTHONPATH=str(ROOT))
    env.update(extra_env or {})
    return subprocess.run([sys.executable, *args], cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


@pytest.mark.gpu
def test_parity_suites_pass_on_the_bounds_checked_library(debug_lib):
    """Odd shapes (W = 1, H = 3, widths that are not multiples of 4 / 32 / 128), hypothesis fuzz, every pinned solver
    variant, sink modes, the 8-bit pipeline and the chain: no red zone touched, no index assert fired, results in parity."""
    sel = ("(small or constant or sink or solver_variants_small or coarse_grid or determinism or legacy or "
           "warp or snr or reverb or u8 or chain or hypothesis or fuzz or local_energy or scan) "
           "and not graphed")      # (the debug library synchronises at the end of a call: it cannot be graph-captured)
    r = _child(debug_lib, ["-m", "pytest", "tests/test_parity_warp.py", "tests/test_parity_snr_reverb.py",
                           "tests/test_parity_confmap.py", "tests/test_parity_chain.py", "tests/test_parity_u8_pipeline.py",
                           "tests/test_properties_hypothesis.py", "tests/test_parity_local_energy.py",
                           "tests/test_parity_scanconv.py", "-m", "gpu", "-q", "-x", "-k", sel, "-p", "no:cacheprovider"])
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail


@pytest.mark.gpu
def test_red_zone_violation_is_reported(debug_lib):
    """Self-test of the mechanism: with USAUG_DBG_FAULT the debug library lets one stray write land right behind the scan
    partials of usaug_bone_box; the call must come back as USAUG_E_BOUNDS (-4) naming the red zone."""
    code = ("import torch; from ultrasoundaugmentation_b200 import ops, _ffi\n"
            "lab = torch.zeros((2, 64, 64), dtype=torch.uint8, device='cuda')\n"
            "try:\n    ops.bone_box(lab); print('NO ERROR')\n"
            "except _ffi.UsaugError as e:\n    print('ERR', e)\n")
    r = _child(debug_lib, ["-c", code], {"USAUG_DBG_FAULT": "1"}, timeout=300)
    assert "ERR" in r.stdout and "code -4" in r.stdout and "red zone" in r.stdout, r.stdout + r.stderr
    r = _child(debug_lib, ["-c", code], timeout=300)
    assert "NO ERROR" in r.stdout, r.stdout + r.stderr
