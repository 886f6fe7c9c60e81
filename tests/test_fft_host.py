"""The block FFT of csrc/fft_stockham.cuh on the CPU: its index, swizzle and butterfly functions are
__host__ __device__, and tests/manual/fft_host_check.cu runs them (all transform lengths 8 .. 2048, forward and inverse)
against a direct fp64 DFT.  Needs only nvcc's host compiler path -- no GPU."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_stockham_fft_functions_on_the_host(tmp_path):
    exe = tmp_path / "fft_host_check"
    cc = subprocess.run(["nvcc", "-O2", "-Wno-deprecated-gpu-targets", "-o", str(exe), str(ROOT / "tests/manual/fft_host_check.cu")],
                        capture_output=True, text=True, timeout=600)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and run.stdout.strip().endswith("ok"), run.stdout + run.stderr
