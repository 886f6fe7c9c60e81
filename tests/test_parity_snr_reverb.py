"""K3 parity: fused SNR change + reverberation vs the oracle (A.4, A.5); 1e-5 relative."""
import numpy as np
import pytest
import torch

from oracle import confmap as ocm
from oracle import deform as odeform
from oracle import snr_reverb as osr
from tests import util

pytestmark = pytest.mark.gpu


def run_gpu(dev, img, cm, lab, a, rho, n, taps):
    from ultrasoundaugmentation_b200 import ops
    out = ops.snr_reverb(util.cuda(img, dev), None if cm is None else util.cuda(cm, dev), util.cuda(lab, dev),
                         torch.from_numpy(a), torch.from_numpy(rho), torch.from_numpy(n), torch.from_numpy(taps))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def smooth_cm(B, H, W, seed=3):
    rng = np.random.default_rng(seed)
    ramp = (1.0 - np.arange(H) / (H - 1.0))[None, :, None]
    return np.clip(ramp * rng.uniform(0.5, 1.0, (B, H, W)), 0, 1).astype(np.float32)


@pytest.mark.parametrize("B,H,W", [(2, 128, 128), (3, 256, 256), (2, 97, 131), (1, 64, 5)])
@pytest.mark.parametrize("with_cm", [True, False])
def test_snr_reverb_parity(cuda_device, B, H, W, with_cm):
    img, lab = util.synth_batch(B, H, W)
    p = util.sample_params(B, H)
    cm = smooth_cm(B, H, W) if with_cm else None
    taps = odeform.gaussian_taps(2.0)
    ref = osr.snr_reverb_batch(img, cm, lab, p["a_snr"], p["rho"], p["n_ghost"], taps)
    got = run_gpu(cuda_device, img, cm, lab, p["a_snr"], p["rho"], p["n_ghost"], taps)
    assert util.rel_err(got, ref) <= util.REL_TOL
    assert np.abs(ref - img).max() > 1e-3          # the artifact is actually there


def test_bone_box_matches_oracle(cuda_device):
    from ultrasoundaugmentation_b200 import ops
    _, lab = util.synth_batch(3, 90, 70)
    lab[1] = 0
    lab[2] = 0; lab[2, 5, 69] = 1; lab[2, 88, 0] = 1
    box = ops.bone_box(util.cuda(lab, cuda_device)).cpu().numpy()
    for b in range(3):
        ys, xs = np.nonzero(lab[b])
        exp = [-1, -1, -1, -1] if len(ys) == 0 else [ys.min(), ys.max(), xs.min(), xs.max()]
        assert box[b].tolist() == [int(v) for v in exp]
        _, y_top, _ = odeform.bone_surface(lab[b])
        assert box[b, 0] == y_top


@pytest.mark.parametrize("case", ["empty_mask", "rho0", "n0", "bone_row0", "many_ghosts", "a1"])
def test_reverb_edge_cases(cuda_device, case):
    H, W = 120, 96
    img, lab = util.synth_batch(2, H, W)
    p = util.sample_params(2, H)
    cm = smooth_cm(2, H, W)
    a, rho, n = p["a_snr"].copy(), p["rho"].copy(), p["n_ghost"].copy()
    if case == "empty_mask":
        lab[:] = 0
    elif case == "rho0":
        rho[:] = 0
    elif case == "n0":
        n[:] = 0
    elif case == "bone_row0":
        lab[:, 0, 10:20] = 1
    elif case == "many_ghosts":
        n[:] = 6; lab[:] = 0; lab[:, 12:15, 20:60] = 1
    elif case == "a1":
        a[:] = 1.0
    taps = odeform.gaussian_taps(2.0)
    ref = osr.snr_reverb_batch(img, cm, lab, a, rho, n, taps)
    got = run_gpu(cuda_device, img, cm, lab, a, rho, n, taps)
    assert util.rel_err(got, ref) <= util.REL_TOL
    if case in ("empty_mask", "n0"):
        only_snr = np.stack([osr.snr_apply(img[b], cm[b], a[b]) for b in range(2)])
        assert np.array_equal(ref, only_snr)


def test_identity_without_cm_and_ghosts(cuda_device):
    img, lab = util.synth_batch(1, 64, 64)
    z = np.zeros(1, np.float32)
    got = run_gpu(cuda_device, img, None, lab, z + 1, z, np.zeros(1, np.int32), odeform.gaussian_taps(2.0))
    assert np.array_equal(got, img)


def test_real_cm_512(cuda_device):
    """BASELINE.json configs[2] shape at one image: oracle CM feeds both sides of K3."""
    img, lab = util.synth_batch(1, 256, 256)
    cm = ocm.confidence_map_batch(img).astype(np.float32)
    p = util.sample_params(1, 256)
    taps = odeform.gaussian_taps(2.0)
    ref = osr.snr_reverb_batch(img, cm, lab, p["a_snr"], p["rho"], p["n_ghost"], taps)
    got = run_gpu(cuda_device, img, cm, lab, p["a_snr"], p["rho"], p["n_ghost"], taps)
    assert util.rel_err(got, ref) <= util.REL_TOL
