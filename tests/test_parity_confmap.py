"""K2 parity: CUDA matrix-free PCG confidence map vs the oracle's direct fp64 solve (A.3).

Tolerance (stated, residual-derived): max|C_gpu - C_oracle| <= 5e4 * rtol + 1e-6
(amplification K <= 2.9e4 measured at beta = 90, Appendix B).
"""
import numpy as np
import pytest
import torch

from oracle import confmap as ocm
from tests import util

pytestmark = pytest.mark.gpu
RTOL = 1e-8


def gpu_cm(dev, img, **kw):
    from ultrasoundaugmentation_b200 import ops
    cm, it, rr = ops.confidence_map(util.cuda(img, dev), return_info=True, **kw)
    torch.cuda.synchronize()
    return cm.cpu().numpy(), it.cpu().numpy(), rr.cpu().numpy()


@pytest.mark.parametrize("H,W", [(33, 17), (64, 64), (50, 3)])
@pytest.mark.parametrize("fp64", [False, True])
def test_constant_image_is_linear_ramp(cuda_device, H, W, fp64):
    """KAT independent of any implementation: constant image => C[y,x] = 1 - y/(H-1) exactly."""
    img = np.full((2, H, W), 0.3, np.float32)
    cm, it, rr = gpu_cm(cuda_device, img, rtol=1e-10, fp64_vectors=fp64, zero_guess=True)
    ramp = (1.0 - np.arange(H) / (H - 1.0))[None, :, None]
    assert np.abs(cm - ramp).max() <= 2e-6
    assert (rr <= 1e-10).all() and (it > 0).all()


@pytest.mark.parametrize("B,H,W", [(2, 64, 48), (3, 96, 80), (2, 128, 128), (1, 67, 131), (2, 3, 40), (1, 40, 1)])
@pytest.mark.parametrize("precond", ["line", "jacobi"])
def test_cm_parity_small(cuda_device, B, H, W, precond):
    img, _ = util.synth_batch(B, H, W)
    ref = ocm.confidence_map_batch(img)
    cm, it, rr = gpu_cm(cuda_device, img, rtol=RTOL, precond=precond, maxiter=60000)
    assert (rr <= RTOL).all(), rr
    assert np.abs(cm - ref).max() <= util.tol_cm(RTOL)
    assert np.all(cm[:, 0] == 1.0) and np.all(cm[:, -1] == 0.0)
    assert cm.min() >= -1e-6 and cm.max() <= 1.0 + 1e-6


@pytest.mark.parametrize("variant", ["fused", "fused-fine", "fused-standard", "legacy", "fp64"])
def test_cm_parity_256(cuda_device, variant):
    img, _ = util.synth_batch(2, 256, 256)
    ref = ocm.confidence_map_batch(img)
    coarse = {"fused-fine": "fine", "fused-standard": "standard"}.get(variant, True)
    cm, it, rr = gpu_cm(cuda_device, img, rtol=RTOL, fp64_vectors=(variant == "fp64"),
                        legacy_kernel=(variant == "legacy"), coarse=coarse)
    assert (rr <= RTOL).all(), rr
    assert np.abs(cm - ref).max() <= util.tol_cm(RTOL)


@pytest.mark.parametrize("twist", [False, True])
@pytest.mark.parametrize("up_variant", ["planes", "recompute"])
@pytest.mark.parametrize("coarse", ["fine", "standard", False])
@pytest.mark.parametrize("H,W", [(40, 128), (67, 132), (96, 260), (131, 64), (50, 1100)])
def test_cm_solver_variants_small(cuda_device, H, W, coarse, up_variant, twist):
    """Every pinned variant of the fused kernel (twisted / plain line factorisation x weights recomputed / read x coarse
    grid) against the oracle, on shapes with one, two and three strips, odd heights and a 4-column last strip."""
    img, _ = util.synth_batch(2, H, W, seed0=77 + H)
    ref = ocm.confidence_map_batch(img)
    cm, it, rr = gpu_cm(cuda_device, img, rtol=RTOL, coarse=coarse, up_variant=up_variant, twist=twist, maxiter=60000)
    assert (rr <= RTOL).all(), rr
    assert np.abs(cm - ref).max() <= util.tol_cm(RTOL)
    assert np.all(cm[:, 0] == 1.0) and np.all(cm[:, -1] == 0.0)


@pytest.mark.parametrize("twist", [False, True])
@pytest.mark.parametrize("up_variant", ["planes", "recompute"])
@pytest.mark.parametrize("level", [True, False])
def test_cm_solver_variants_256(cuda_device, level, up_variant, twist):
    img, _ = util.synth_batch(2, 256, 256)
    ref = ocm.confidence_map_batch(img)
    cm, it, rr = gpu_cm(cuda_device, img, rtol=RTOL, up_variant=up_variant, twist=twist, level=level)
    assert (rr <= RTOL).all(), rr
    assert np.abs(cm - ref).max() <= util.tol_cm(RTOL)
    # the twisted factorisation is the same preconditioner: the iteration count must not move
    _, it0, _ = gpu_cm(cuda_device, img, rtol=RTOL, up_variant=up_variant, twist=False, level=level)
    assert np.abs(it.astype(int) - it0.astype(int)).max() <= 0.05 * it0.max() + 8
    if level:       # the level preconditioner (line solve on 1 x 4 aggregates) must pay for itself
        _, it1, _ = gpu_cm(cuda_device, img, rtol=RTOL, up_variant=up_variant, twist=twist, level=False)
        assert it.max() <= 0.9 * it1.max()


def test_cm_batch_independence_and_determinism(cuda_device):
    """An image's map does not depend on its batch mates; repeated runs are bitwise identical."""
    img, _ = util.synth_batch(5, 96, 64)
    a, ita, _ = gpu_cm(cuda_device, img, rtol=1e-7)
    b, itb, _ = gpu_cm(cuda_device, img, rtol=1e-7)
    assert np.array_equal(a, b) and np.array_equal(ita, itb)
    c, itc, _ = gpu_cm(cuda_device, img[2:3], rtol=1e-7)
    assert np.array_equal(c[0], a[2]) and itc[0] == ita[2]
    d, _, _ = gpu_cm(cuda_device, img, rtol=1e-7, max_wave=2)
    assert np.array_equal(d, a)


def test_cm_fused_and_legacy_kernels_agree(cuda_device):
    """The fused fp32 kernel (W % 4 == 0) and the two-slot kernel solve the same systems."""
    img, _ = util.synth_batch(3, 96, 132)
    a, ita, rra = gpu_cm(cuda_device, img, rtol=1e-9)
    b, itb, rrb = gpu_cm(cuda_device, img, rtol=1e-9, legacy_kernel=True)
    assert (rra <= 1e-9).all() and (rrb <= 1e-9).all()
    assert np.abs(a - b).max() <= 2 * util.tol_cm(1e-9)
    assert (ita <= itb + 5).all()          # the coarse-grid correction of the fused kernel never costs iterations
    c, itc, rrc = gpu_cm(cuda_device, img, rtol=1e-9, coarse=False, level=False)
    assert (rrc <= 1e-9).all() and np.abs(a - c).max() <= 2 * util.tol_cm(1e-9)
    assert np.abs(itc.astype(int) - itb.astype(int)).max() <= 0.1 * itb.max() + 5   # same preconditioner as the two-slot kernel


def test_cm_maxiter_reports_nonconvergence(cuda_device):
    img, _ = util.synth_batch(1, 128, 128)
    cm, it, rr = gpu_cm(cuda_device, img, rtol=1e-12, maxiter=5)
    assert it[0] == 5 and rr[0] > 1e-12 and np.isfinite(cm).all()


def test_cm_monotone_for_constant_columns(cuda_device):
    """Discrete maximum principle on a depth-only image: monotone non-increasing in depth."""
    H, W = 64, 32
    img = np.tile(np.linspace(0.2, 0.9, H, dtype=np.float32)[None, :, None], (1, 1, W))
    cm, _, _ = gpu_cm(cuda_device, img, rtol=1e-9)
    assert (np.diff(cm[0], axis=0) <= 1e-6).all()


@pytest.mark.parametrize("H,W,what", [(1500, 64, "tall: 375 / 750 aggregate rows of two 8-wide blocks"),
                                      (2100, 128, "very tall: more than 1024 aggregate rows at the smallest height, so it doubles"),
                                      (72, 1000, "wide: 32 aggregate columns, 32 x 32 blocks"),
                                      (40, 2016, "64-column aggregates"),
                                      (40, 4000, "128-column aggregates: one per strip"),
                                      (9, 64, "fewer than 8 unknown rows: coarse grid disabled")])
def test_cm_coarse_grid_geometries(cuda_device, H, W, what):
    """Every coarse-grid geometry of the fused kernel against the ORACLE, with 4-row and with 2-row aggregates."""
    from ultrasoundaugmentation_b200 import synth
    img, _ = synth.make_batch(2, H, W, seed=H + W, device=cuda_device)
    img = img.cpu().numpy()
    ref = ocm.confidence_map_batch(img)
    for coarse in ("standard", "fine"):
        a, ita, rra = gpu_cm(cuda_device, img, rtol=1e-9, maxiter=100000, coarse=coarse)
        assert (rra <= 1e-9).all(), (what, coarse)
        assert np.abs(a - ref).max() <= util.tol_cm(1e-9), (what, coarse)
        assert np.all(a[:, 0] == 1.0) and np.all(a[:, -1] == 0.0)
    # and the generic (line-only, grid-synchronous) kernel on the same systems
    b, itb, rrb = gpu_cm(cuda_device, img, rtol=1e-9, maxiter=100000, legacy_kernel=True)
    assert (rrb <= 1e-9).all() and np.abs(b - ref).max() <= util.tol_cm(1e-9), what


@pytest.mark.parametrize("B,H,W,what", [(1800, 40, 128, ">= 12 tasks per SM: weight-recomputing kernel, 2-row aggregates"),
                                         (900, 48, 128, ">= 6 tasks per SM: weight-recomputing kernel, 4-row aggregates"),
                                         (1000, 36, 132, "two strips per image, the second one 4 columns wide")])
def test_cm_large_batch_kernel_variants(cuda_device, B, H, W, what):
    """The kernel variants that only large batches select (launch_fused / cm_layout), on many tiny images so that
    the oracle stays affordable: every image converges, a sample of them is checked against the direct solve."""
    img = np.stack([util.synth_batch(1, H, W, seed0=3000 + (b % 97))[0][0] for b in range(97)])
    img = np.concatenate([img] * (B // 97 + 1))[:B]
    rng = np.random.default_rng(1)
    img = np.clip(img + 0.02 * rng.standard_normal(img.shape).astype(np.float32), 0, 1).astype(np.float32)
    cm, it, rr = gpu_cm(cuda_device, img, rtol=RTOL)
    assert (rr <= RTOL).all(), what
    assert (it > 0).all() and np.all(cm[:, 0] == 1.0) and np.all(cm[:, -1] == 0.0)
    for b in range(0, B, max(B // 12, 1)):
        assert np.abs(cm[b] - ocm.confidence_map(img[b])).max() <= util.tol_cm(RTOL), (what, b)
    cm2, it2, _ = gpu_cm(cuda_device, img, rtol=RTOL)
    assert np.array_equal(cm, cm2) and np.array_equal(it, it2)           # repeatable under dynamic scheduling


def test_cm_watchdog_ignores_starved_workers(cuda_device, monkeypatch):
    """The scheduler's watchdog must watch global progress, not a worker's own luck: with one slowly converging image
    left among many trivial ones, most workers find no task for far longer than the (here: 20 ms) watchdog period while
    the solve is healthy.  Every image must still come back converged."""
    monkeypatch.setenv("USAUG_CM_WATCHDOG_MS", "20")
    H = W = 512
    img = np.full((64, H, W), 0.3, np.float32)                 # constant images: done after the first residual
    img[17] = util.synth_batch(1, H, W)[0][0]                  # one real image: ~500 iterations, ~0.25 s
    cm, it, rr = gpu_cm(cuda_device, img, rtol=RTOL)
    assert (it >= 0).all() and np.isfinite(rr).all(), "watchdog fired on a healthy solve"
    assert (rr <= RTOL).all() and it[17] > 100
    assert np.abs(cm[17] - ocm.confidence_map(img[17])).max() <= util.tol_cm(RTOL)


# ---- sink modes -------------------------------------------------------------------------------
def _sink_case(B, H, W, mode, seed0=500):
    img, lab = util.synth_batch(B, H, W, seed0=seed0)
    mask = None
    if mode == "mask":
        rng = np.random.default_rng(H * W)
        mask = (lab != 0).astype(np.uint8)                          # the bone arc ...
        mask |= (rng.random(lab.shape) < 0.01).astype(np.uint8)      # ... and scattered single pixels
        mask[:, H // 3, : W // 2] = 1                               # ... and half a row (cuts strips and aggregates)
        mask[:, :, W - 1] = 1                                       # ... and the last column
    ref = ocm.confidence_map_batch(img, sink=mode, masks=mask)
    return img, mask, ref


@pytest.mark.parametrize("twist", [False, True])
@pytest.mark.parametrize("up_variant", ["planes", "recompute"])
@pytest.mark.parametrize("coarse", ["fine", "standard", False])
@pytest.mark.parametrize("mode", ["bottom_and_sides", "mask"])
def test_cm_sink_modes_fused_variants(cuda_device, mode, coarse, up_variant, twist):
    """Sinks other than the bottom row, every pinned variant of the fused kernel against the oracle's direct solve."""
    for H, W in ((67, 132), (96, 260)):
        img, mask, ref = _sink_case(2, H, W, mode)
        cm, it, rr = gpu_cm(cuda_device, img, rtol=RTOL, coarse=coarse, up_variant=up_variant, twist=twist, sink=mode,
                            sink_mask=None if mask is None else util.cuda(mask, cuda_device), maxiter=60000)
        assert (rr <= RTOL).all(), rr
        assert np.abs(cm - ref).max() <= util.tol_cm(RTOL)
        sinks = np.stack([ocm.sink_mask(H, W, mode, None if mask is None else mask[b]) for b in range(2)])
        assert (cm[sinks] == 0.0).all() and np.all(cm[:, 0] == 1.0) and np.all(cm[:, -1] == 0.0)


@pytest.mark.parametrize("mode", ["bottom_and_sides", "mask"])
@pytest.mark.parametrize("kw", [dict(legacy_kernel=True), dict(fp64_vectors=True), dict(precond="jacobi"), dict(level=False)])
def test_cm_sink_modes_other_kernels(cuda_device, mode, kw):
    """... and the generic kernel (any width, fp64 vectors), Jacobi, and the fused kernel without the level solve."""
    H, W = (64, 61) if ("legacy_kernel" in kw or "fp64_vectors" in kw) else (64, 64)
    img, mask, ref = _sink_case(2, H, W, mode, seed0=900)
    cm, it, rr = gpu_cm(cuda_device, img, rtol=RTOL, sink=mode, maxiter=60000,
                        sink_mask=None if mask is None else util.cuda(mask, cuda_device), **kw)
    assert (rr <= RTOL).all(), rr
    assert np.abs(cm - ref).max() <= util.tol_cm(RTOL)


def test_cm_sink_row_ramp_kat(cuda_device):
    """Closed form: constant image with a full row r of sinks => C = 1 - y/r above the row, 0 from it downwards."""
    H, W, r = 96, 128, 40
    img = np.full((2, H, W), 0.3, np.float32)
    mask = np.zeros((2, H, W), np.uint8); mask[:, r] = 1
    cm, it, rr = gpu_cm(cuda_device, img, rtol=1e-10, sink="mask", sink_mask=util.cuda(mask, cuda_device), zero_guess=True)
    ref = np.clip(1.0 - np.arange(H) / r, 0.0, None)[None, :, None]
    assert np.abs(cm - ref).max() <= 2e-6 and (cm[:, r:] == 0.0).all()


def test_cm_sink_256_iterations_and_plain_equivalence(cuda_device):
    """256 x 256: parity in both modes; an all-zero mask reproduces the plain solve bit for bit; the two-level
    preconditioner still pays with sinks (fewer iterations than the line preconditioner alone)."""
    img, mask, ref = _sink_case(2, 256, 256, "mask")
    dmask = util.cuda(mask, cuda_device)
    cm, it, rr = gpu_cm(cuda_device, img, rtol=RTOL, sink="mask", sink_mask=dmask)
    assert (rr <= RTOL).all() and np.abs(cm - ref).max() <= util.tol_cm(RTOL)
    _, it_line, _ = gpu_cm(cuda_device, img, rtol=RTOL, sink="mask", sink_mask=dmask, coarse=False, level=False)
    assert it.max() <= 0.85 * it_line.max(), (it, it_line)
    plain, itp, _ = gpu_cm(cuda_device, img, rtol=RTOL)
    zero, itz, _ = gpu_cm(cuda_device, img, rtol=RTOL, sink="mask", sink_mask=torch.zeros_like(dmask))
    assert np.array_equal(plain, zero) and np.array_equal(itp, itz)
    _, _, ref_s = _sink_case(2, 256, 256, "bottom_and_sides")
    cm_s, _, rr_s = gpu_cm(cuda_device, img, rtol=RTOL, sink="bottom_and_sides")
    assert (rr_s <= RTOL).all() and np.abs(cm_s - ref_s).max() <= util.tol_cm(RTOL)


def test_cm_sink_argument_errors(cuda_device):
    from ultrasoundaugmentation_b200 import ops
    img = util.cuda(util.synth_batch(1, 32, 32)[0], cuda_device)
    with pytest.raises(ValueError):
        ops.confidence_map(img, sink="mask")
    with pytest.raises(ValueError):
        ops.confidence_map(img, sink="top")
    with pytest.raises(ValueError):
        ops.confidence_map(img, sink="bottom_row", sink_mask=torch.zeros((1, 32, 32), dtype=torch.uint8, device=cuda_device))


def test_cm_sharding_and_solver_variants(cuda_device):
    """What sharding a batch can change.  The library picks the solver variant from the size of a launch (cm_layout):
    * F_UP reads the weight planes (few strip tasks) or recomputes the weights (many): the same device function produces
      the weights either way, so the two are BITWISE equal -- 1000 tiny images in one launch == two launches of 500;
    * images taller than 514 rows use 2-row aggregates only in deeply bandwidth-bound launches: 1800 images of 600 x 128 in
      one launch vs two launches of 900 differ in the preconditioner, i.e. in the low bits -- same systems, same stopping
      rule, so they agree within the confidence-map tolerance, and pinning the variant makes them bitwise equal again."""
    def batch(B, H, W):
        img = np.stack([util.synth_batch(1, H, W, seed0=5000 + (b % 61))[0][0] for b in range(61)])
        return np.concatenate([img] * (B // 61 + 1))[:B]
    img = batch(1000, 40, 128)
    whole, itw, rrw = gpu_cm(cuda_device, img, rtol=RTOL)
    parts = np.concatenate([gpu_cm(cuda_device, img[lo:lo + 500], rtol=RTOL)[0] for lo in (0, 500)])
    assert (rrw <= RTOL).all() and np.array_equal(whole, parts)
    a, _, _ = gpu_cm(cuda_device, img[:64], rtol=RTOL, up_variant="planes")
    b, _, _ = gpu_cm(cuda_device, img[:64], rtol=RTOL, up_variant="recompute")
    assert np.array_equal(a, b) and np.array_equal(a, whole[:64])

    img = batch(1800, 600, 128)
    whole, itw, rrw = gpu_cm(cuda_device, img, rtol=RTOL)
    halves = [gpu_cm(cuda_device, img[lo:lo + 900], rtol=RTOL) for lo in (0, 900)]
    parts = np.concatenate([h[0] for h in halves])
    assert (rrw <= RTOL).all() and all((h[2] <= RTOL).all() for h in halves)
    assert not np.array_equal(itw, np.concatenate([h[1] for h in halves]))      # another preconditioner: other iteration counts
    assert np.abs(whole - parts).max() <= 2 * util.tol_cm(RTOL)
    parts_p = np.concatenate([gpu_cm(cuda_device, img[lo:lo + 900], rtol=RTOL, coarse="fine")[0] for lo in (0, 900)])
    assert np.array_equal(whole, parts_p)
