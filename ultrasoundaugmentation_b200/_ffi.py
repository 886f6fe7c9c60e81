"""ctypes binding of libusaug.so (include/usaug.h).  The ONLY caller of the C ABI.

There is no fallback of any kind: if the shared library is missing or a call fails, the
product raises.  PyTorch is used for device memory and streams only; the signatures carry raw
device pointers (``tensor.data_ptr()``) and the current ``cudaStream_t``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_size_t, c_void_p
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_DEFAULT_LIB = _PKG / "lib" / "libusaug.so"

USAUG_OK = 0
PRECOND_JACOBI = 0
PRECOND_LINE = 1
CM_FP64_VECTORS = 1
CM_ZERO_GUESS = 2
CM_TWIST_ON = 256
CM_TWIST_OFF = 512
CM_FORCE_RECOMPUTE = 1024
CM_FORCE_PLANES = 2048
CM_NO_LEVEL = 4096
CM_SINKS = 8192
SINK_MODES = {"bottom_row": 0, "bottom_and_sides": 1, "mask": 2}

# every symbol include/usaug.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "usaug_version": (c_int, []),
    "usaug_last_error": (c_char_p, []),
    "usaug_upload": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "usaug_warp_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "usaug_fused_tgc_deform": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                       c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                       c_size_t, c_void_p]),
    "usaug_fused_tgc_deform_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                          c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                          c_size_t, c_void_p]),
    "usaug_confmap_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "usaug_confmap_pcg": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_float,
                                  c_double, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                  c_size_t, c_void_p]),
    "usaug_confmap_pcg_sinks": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_float,
                                        c_double, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_size_t, c_void_p]),
    "usaug_confmap_set_timing_events": (None, [c_void_p, c_void_p]),
    "usaug_snr_reverb_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "usaug_snr_reverb": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_size_t,
                                 c_void_p]),
    "usaug_snr_reverb_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_size_t,
                                    c_void_p]),
    "usaug_snr_reverb_signal": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                        c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                        c_void_p, c_size_t, c_void_p]),
    "usaug_local_energy_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "usaug_local_energy": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_int,
                                   c_float, c_void_p, c_size_t, c_void_p]),
    "usaug_scan_convert": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_float, c_float, c_float, c_float, c_float, c_float, c_float, c_void_p, c_void_p,
                                   c_void_p]),
    "usaug_bone_box_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "usaug_bone_box": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
}

_lib = None


def lib_path() -> Path:
    return Path(os.environ.get("USAUG_LIB", str(_DEFAULT_LIB)))


def load():
    """dlopen libusaug.so (once per process; DataLoader workers re-open lazily)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not path.exists():
        raise RuntimeError(
            f"{path} not found: the CUDA library is not built and there is no CPU fallback. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `python -m ultrasoundaugmentation_b200.build`).")
    lib = ctypes.CDLL(str(path))
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the ABI is incomplete
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


class UsaugError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != USAUG_OK:
        msg = load().usaug_last_error()
        raise UsaugError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")
