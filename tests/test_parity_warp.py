"""K1 parity: CUDA fused TGC + deformation + remap vs the oracle.

Labels bit-exact (nearest sampling); image within 1e-5 relative error (tolerance from north_star).
"""
import numpy as np
import pytest
import torch

from oracle import deform as odeform
from tests import util

pytestmark = pytest.mark.gpu


def run_gpu(dev, img, lab, d, tgc, taps):
    from ultrasoundaugmentation_b200 import ops
    o, l = ops.fused_tgc_deform(util.cuda(img, dev), util.cuda(lab, dev), torch.from_numpy(d),
                                torch.from_numpy(tgc), torch.from_numpy(taps))
    torch.cuda.synchronize()
    return o.cpu().numpy(), l.cpu().numpy()


def check(dev, img, lab, d, tgc, taps):
    ref_i, ref_l = odeform.fused_tgc_deform_batch(img, lab, d, tgc, taps)
    got_i, got_l = run_gpu(dev, img, lab, d, tgc, taps)
    assert np.array_equal(got_l, ref_l), f"label mismatch at {np.argwhere(got_l != ref_l)[:5]}"
    assert util.rel_err(got_i, ref_i) <= util.REL_TOL
    return got_i, got_l


@pytest.mark.parametrize("B,H,W", [(1, 256, 256), (4, 256, 256), (3, 512, 512), (2, 97, 131),
                                    (2, 64, 48), (1, 33, 1030), (2, 3, 7), (1, 40, 1)])
def test_warp_parity_shapes(cuda_device, B, H, W):
    img, lab = util.synth_batch(B, H, W)
    p = util.sample_params(B, H)
    check(cuda_device, img, lab, p["d"], p["tgc"], util.warp_taps(W))


def test_config1_and_2_batch256(cuda_device):
    """BASELINE.json configs[0]/[1]: 256x256 with bone mask; a 32-image slice of the batch of 256
    is checked against the oracle, the full batch for determinism and range."""
    B, H, W = 256, 256, 256
    img, lab = util.synth_batch(32, H, W)
    img = np.tile(img, (8, 1, 1)); lab = np.tile(lab, (8, 1, 1))
    p = util.sample_params(B, H)
    taps = util.warp_taps(W)
    g1 = run_gpu(cuda_device, img, lab, p["d"], p["tgc"], taps)
    g2 = run_gpu(cuda_device, img, lab, p["d"], p["tgc"], taps)
    assert np.array_equal(g1[0], g2[0]) and np.array_equal(g1[1], g2[1])      # bitwise repeatable
    sel = np.arange(0, B, 8)
    ref_i, ref_l = odeform.fused_tgc_deform_batch(img[sel], lab[sel], p["d"][sel], p["tgc"][sel], taps)
    assert np.array_equal(g1[1][sel], ref_l)
    assert util.rel_err(g1[0][sel], ref_i) <= util.REL_TOL
    assert g1[0].min() >= 0.0 and g1[0].max() <= 1.0
    assert set(np.unique(g1[1])) <= {0, 1}


def test_identity_when_d0_gain1(cuda_device):
    img, lab = util.synth_batch(2, 128, 96)
    d = np.zeros(2, np.float32); tgc = np.ones((2, 8), np.float32)
    gi, gl = run_gpu(cuda_device, img, lab, d, tgc, util.warp_taps(96))
    assert np.array_equal(gl, lab)
    assert np.array_equal(gi, img)


@pytest.mark.parametrize("case", ["empty", "full", "top_row", "bottom_row", "left_col", "single_px"])
def test_mask_edge_cases(cuda_device, case):
    H, W = 80, 72
    img, _ = util.synth_batch(2, H, W)
    lab = np.zeros((2, H, W), np.uint8)
    if case == "full":
        lab[:] = 3
    elif case == "top_row":
        lab[:, 0, :] = 1
    elif case == "bottom_row":
        lab[:, H - 1, :] = 2
    elif case == "left_col":
        lab[:, 30:, 0] = 1
    elif case == "single_px":
        lab[:, 41, 17] = 255
    p = util.sample_params(2, H, seed=11)
    check(cuda_device, img, lab, p["d"], p["tgc"], util.warp_taps(W))


@pytest.mark.parametrize("d", [-30.0, -3.5, 0.25, 19.99, 1e4, -1e4])
def test_displacement_extremes(cuda_device, d):
    H, W = 128, 128
    img, lab = util.synth_batch(1, H, W)
    p = util.sample_params(1, H, K=5)
    check(cuda_device, img, lab, np.array([d], np.float32), p["tgc"], util.warp_taps(W))


@pytest.mark.parametrize("K", [1, 2, 3, 8, 17])
def test_tgc_control_points(cuda_device, K):
    H, W = 100, 64
    img, lab = util.synth_batch(2, H, W)
    p = util.sample_params(2, H, K=K)
    check(cuda_device, img, lab, p["d"], p["tgc"], util.warp_taps(W))


def test_multilabel_values_subset(cuda_device):
    """Label output values are a subset of the input values plus 0 (nearest sampling never blends)."""
    H, W = 96, 96
    img, lab = util.synth_batch(1, H, W)
    lab = (lab * 7).astype(np.uint8); lab[0, 70:80, 10:30] = 200
    p = util.sample_params(1, H)
    _, gl = check(cuda_device, img, lab, p["d"], p["tgc"], util.warp_taps(W))
    assert set(np.unique(gl)) <= set(np.unique(lab)) | {0}


def test_rejects_cpu_tensors(built_lib):
    from ultrasoundaugmentation_b200 import ops
    with pytest.raises(ValueError):
        ops.fused_tgc_deform(torch.zeros(1, 8, 8), torch.zeros(1, 8, 8, dtype=torch.uint8),
                             torch.zeros(1), torch.ones(1, 2), torch.ones(1))
