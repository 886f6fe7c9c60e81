"""CPU-only tests of the host side: parameter sampling, sharding, shape/dtype/device validation,
pickling, the C-ABI surface (library loads and exports every symbol include/usaug.h declares; no
compute calls without a GPU), and the rule that the product never touches the oracle."""
import ctypes
import pickle
import re
from pathlib import Path

import numpy as np
import pytest
import torch

import ultrasoundaugmentation_b200 as us
from oracle import deform as odeform
from ultrasoundaugmentation_b200 import _ffi, ops, params, sharding

ROOT = Path(__file__).resolve().parent.parent


# ---------------------------------------------------------------- C ABI surface
def header_functions():
    txt = (ROOT / "include" / "usaug.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(usaug_[a-z0-9_]+)\s*\(", txt)))


def test_library_loads_and_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    names = header_functions()
    assert len(names) >= 11
    for n in names:
        assert hasattr(lib, n), f"{n} declared in usaug.h but not exported by libusaug.so"
    assert set(_ffi.SYMBOLS) == set(names), "ctypes table and header out of sync"
    assert _ffi.load().usaug_version() == 10200
    assert _ffi.load().usaug_last_error() is not None


def test_library_exports_nothing_but_the_c_abi(built_lib):
    """-fvisibility=hidden: the dynamic symbol table holds the usaug_* entry points of the header and no C++ symbol."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", str(built_lib)], capture_output=True, text=True, check=True).stdout
    defined = sorted(line.split()[-1] for line in out.splitlines() if " T " in line)
    assert defined == header_functions(), set(defined) ^ set(header_functions())


def test_workspace_queries_are_pure_host_functions(built_lib):
    lib = _ffi.load()
    assert lib.usaug_warp_workspace_bytes(4, 64, 64) > 0
    assert lib.usaug_warp_workspace_bytes(0, 64, 64) == 0
    small = lib.usaug_confmap_workspace_bytes(1, 64, 64, 1, 0)
    assert lib.usaug_confmap_workspace_bytes(4, 64, 64, 1, 0) > 3 * small // 2
    assert lib.usaug_confmap_workspace_bytes(1, 64, 64, 1, _ffi.CM_FP64_VECTORS) > small
    assert lib.usaug_confmap_workspace_bytes(1, 2, 64, 1, 0) == 0            # H < 3 is invalid
    assert lib.usaug_confmap_workspace_bytes(1, 64, 64, 1, _ffi.CM_SINKS) >= small + 64 * 64   # room for the sink plane
    assert lib.usaug_snr_reverb_workspace_bytes(2, 32, 32) >= lib.usaug_bone_box_workspace_bytes(2, 32, 32) > 0


def test_invalid_arguments_return_error_codes_without_a_gpu(built_lib):
    lib = _ffi.load()
    rc = lib.usaug_fused_tgc_deform(None, None, None, None, 1, 8, 8, None, None, 1, None, 0, None, 0, None)
    assert rc == -1 and b"null" in lib.usaug_last_error()
    with pytest.raises(_ffi.UsaugError):
        _ffi.check(rc, "usaug_fused_tgc_deform")
    assert lib.usaug_confmap_pcg(None, None, 1, 8, 8, 2.0, 90.0, 0.05, 1e-6, 10, 1, 0, None, None, None, 0, None) == -1
    assert lib.usaug_upload(None, None, 16, None) == -1 and lib.usaug_upload(None, None, 0, None) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setenv("USAUG_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_ffi, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _ffi.load()


def test_product_never_imports_the_oracle():
    for p in (ROOT / "ultrasoundaugmentation_b200").rglob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{p} imports the oracle"


# ---------------------------------------------------------------- parameters
def test_counter_rng_is_a_pure_function_of_seed_index_stream():
    a = params.counter_uniform(7, np.arange(10), 3, 4)
    b = params.counter_uniform(7, np.arange(10), 3, 4)
    assert np.array_equal(a, b) and a.shape == (10, 4)
    assert not np.array_equal(a, params.counter_uniform(8, np.arange(10), 3, 4))
    assert not np.array_equal(a, params.counter_uniform(7, np.arange(10), 2, 4))
    assert np.array_equal(a[3:7], params.counter_uniform(7, np.arange(3, 7), 3, 4))
    big = params.counter_uniform(1, np.arange(20000), 1, 2)
    assert 0 <= big.min() and big.max() < 1 and abs(big.mean() - 0.5) < 0.01 and abs(big.var() - 1 / 12) < 0.005


def test_get_params_shapes_ranges_and_p():
    ch = us.PhysicsChain.full(seed=3)
    p = ch.get_params(64, image_size=(512, 512))
    assert p["d"].shape == (64,) and p["tgc"].shape == (64, 8) and p["n_ghost"].dtype == torch.int32
    assert (p["d"] >= -0.05 * 512).all() and (p["d"] <= 0.15 * 512).all()
    assert (p["tgc"] >= 0.6).all() and (p["tgc"] <= 1.4).all()
    assert (p["a_snr"] >= 0.7).all() and (p["a_snr"] <= 1.4).all()
    assert set(p["n_ghost"].tolist()) <= {1, 2} and (p["rho"] >= 0.3).all() and (p["rho"] <= 0.7).all()
    off = us.PhysicsChain([us.ProbePressureDeformation(p=0.0), us.TimeGainCompensation(p=0.0),
                           us.ConfidenceSNR(p=0.0), us.Reverberation(p=0.0)]).get_params(8, image_size=(64, 64))
    assert (off["d"] == 0).all() and (off["tgc"] == 1).all() and (off["a_snr"] == 1).all() and (off["n_ghost"] == 0).all()
    half = us.Reverberation(p=0.5, seed=1).get_params(4000)
    assert 0.4 < (half["n_ghost"] > 0).float().mean() < 0.6


def test_get_params_generators_and_errors():
    t = us.TimeGainCompensation(seed=0)
    g1 = t.get_params(5, generator=np.random.default_rng(1))["tgc"]
    g2 = t.get_params(5, generator=np.random.default_rng(1))["tgc"]
    assert torch.equal(g1, g2)
    tg = torch.Generator().manual_seed(4)
    assert t.get_params(5, generator=tg)["tgc"].shape == (5, 8)
    with pytest.raises(ValueError):
        us.ProbePressureDeformation().get_params(4)                      # needs image_size
    with pytest.raises(TypeError):
        t.get_params(4, generator="nope")
    with pytest.raises(ValueError):
        us.TimeGainCompensation(p=1.5)
    with pytest.raises(ValueError):
        us.PhysicsChain([])
    with pytest.raises(ValueError):
        us.PhysicsChain([us.Reverberation(), us.Reverberation()])
    with pytest.raises(TypeError):
        us.PhysicsChain([object()])


def test_shard_params_equal_rows_of_the_global_draw():
    ch = us.PhysicsChain.full(seed=11)
    full = ch.get_params(37, image_size=(128, 128))
    for world in (1, 2, 4, 8):
        rows = {k: [] for k in full}
        for r in range(world):
            p = sharding.shard_params(ch, 37, r, world, (128, 128))
            for k in full:
                rows[k].append(p[k])
        for k in full:
            assert torch.equal(torch.cat(rows[k]), full[k]), (world, k)


def test_shard_bounds_cover_and_balance():
    for total in (0, 1, 7, 4096):
        for world in (1, 2, 3, 8):
            b = [sharding.shard_bounds(total, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 4, 4)


def test_taps_match_the_oracle_bit_for_bit():
    for sigma in (0.0, 0.5, 2.0, 5.12, 10.24):
        assert np.array_equal(ops.gaussian_taps(sigma).numpy(), odeform.gaussian_taps(sigma))


# ---------------------------------------------------------------- validation (no GPU needed: all raise before any launch)
def test_cpu_tensors_are_rejected_not_silently_computed(built_lib):
    img, lab = torch.zeros(2, 16, 16), torch.zeros(2, 16, 16, dtype=torch.uint8)
    ch = us.PhysicsChain.full()
    with pytest.raises(ValueError, match="CUDA"):
        ch(img, lab)
    with pytest.raises(ValueError, match="CUDA"):
        ops.confidence_map(img)
    with pytest.raises(ValueError, match="CUDA"):
        ops.snr_reverb(img, None, lab, 1.0, 0.5, 1, ops.gaussian_taps(2.0))
    with pytest.raises(ValueError, match="CUDA"):
        ops.bone_box(lab)


def test_shape_and_type_errors():
    ch = us.PhysicsChain([us.TimeGainCompensation()])
    with pytest.raises(ValueError):
        ch(torch.zeros(2, 3, 8, 8), torch.zeros(2, 3, 8, 8, dtype=torch.uint8))     # C != 1
    with pytest.raises(ValueError):
        ch(torch.zeros(8, 8), torch.zeros(8, 9, dtype=torch.uint8))
    with pytest.raises(TypeError):
        ch(np.zeros((8, 8)), np.zeros((8, 8)))
    with pytest.raises(ValueError):
        us.ConfidenceSNR(precond="ilu")
    with pytest.raises(ValueError):
        us.Reverberation(sigma_feather=20.0)


def test_transforms_pickle_config_only():
    ch = us.PhysicsChain.full(seed=9, rtol=1e-7)
    ch.snr.last_info = ("not", "picklable-safe")
    ch2 = pickle.loads(pickle.dumps(ch))
    assert ch2.snr.rtol == 1e-7 and ch2.snr.last_info is None and ch2.deform.seed == 9
    a = ch.get_params(6, image_size=(64, 64)); b = ch2.get_params(6, image_size=(64, 64))
    assert all(torch.equal(a[k], b[k]) for k in a)
