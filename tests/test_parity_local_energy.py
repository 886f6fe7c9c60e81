"""Monogenic-signal local energy: hand-written batched 2-D FFT vs the fp64 NumPy oracle.

Tolerance (floating point, stated here): the CUDA path is fp32 end to end; an fp32 radix-2 FFT of n points carries
a relative error of about log2(n) * 6e-8 of the LARGEST coefficient, so the normalised map E_hat in [0,1] is
compared ABSOLUTELY: |E_hat_gpu - E_hat_oracle| <= 2e-5.  Through A.4 the image then differs by at most
|a - 1| * C * 2e-5 <= 8e-6 on top of the 1e-5 relative bar of the other image transforms.
"""
import numpy as np
import pytest
import torch

from oracle import chain as ochain
from oracle import confmap as ocm
from oracle import deform as odeform
from oracle import local_energy as ole
from oracle import snr_reverb as osr
from tests import util

pytestmark = pytest.mark.gpu
TOL_E = 2e-5


def gpu_energy(dev, img, **kw):
    from ultrasoundaugmentation_b200 import ops
    e, s = ops.local_energy(util.cuda(img, dev), **kw)
    torch.cuda.synchronize()
    return e.cpu().numpy(), s.cpu().numpy()


@pytest.mark.parametrize("B,H,W", [(2, 64, 64), (3, 128, 256), (2, 512, 512), (1, 1024, 1024), (2, 8, 8), (1, 2048, 16),
                                    (1, 16, 2048),
                                    # not powers of two: mirror-extended to the next one
                                    (2, 97, 131), (2, 100, 64), (1, 480, 640), (2, 5, 9), (1, 4, 4), (1, 1030, 33),
                                    # largest transform, odd batch, long thin shapes in both directions
                                    (1, 2048, 2048), (1, 1500, 700), (3, 260, 8), (2, 8, 260), (1, 4, 2048), (5, 32, 32)])
def test_local_energy_parity(cuda_device, B, H, W):
    img, _ = util.synth_batch(B, H, W)
    e, s = gpu_energy(cuda_device, img)
    ref_raw = np.stack([ole.local_energy_raw(im) for im in img])
    ref = ole.local_energy_batch(img)
    got = e * s[:, None, None]
    assert np.abs(got - ref).max() <= TOL_E
    assert np.allclose(s, 1.0 / ref_raw.reshape(B, -1).max(axis=1), rtol=1e-5)
    assert got.max() <= 1.0 + 1e-6 and got.min() >= 0.0


@pytest.mark.parametrize("kw", [dict(lambda_min=4.0, mult=1.6, scales=4, sigma_f=0.65), dict(lambda_min=16.0, mult=3.0, scales=1, sigma_f=0.4)])
def test_local_energy_filter_parameters(cuda_device, kw):
    img, _ = util.synth_batch(2, 128, 128)
    e, s = gpu_energy(cuda_device, img, **kw)
    assert np.abs(e * s[:, None, None] - ole.local_energy_batch(img, **kw)).max() <= TOL_E


def test_local_energy_known_answers_on_gpu(cuda_device):
    """The oracle's implementation-independent pins, on the CUDA path: constant -> 0 (scale 0), plane wave -> constant."""
    H, W = 128, 256
    const = np.full((1, H, W), 0.37, np.float32)
    e, s = gpu_energy(cuda_device, const)
    assert e.max() <= 1e-6 and (s[0] == 0.0 or e.max() * s[0] <= 1.0)
    for k0, axis in [(8, 1), (5, 0), (32, 1)]:
        n = W if axis == 1 else H
        t = (0.5 + 0.4 * np.cos(2 * np.pi * k0 * np.arange(n) / n + 0.3)).astype(np.float32)
        img = np.tile(t[None, :], (H, 1)) if axis == 1 else np.tile(t[:, None], (1, W))
        e, s = gpu_energy(cuda_device, img[None])
        G = ole.log_gabor_sum(H, W)[0]
        want = 0.4 * (G[0, k0] if axis == 1 else G[k0, 0])
        assert np.abs(e - want).max() <= 5e-6, (k0, axis)
        assert np.abs(e * s[0] - 1.0).max() <= TOL_E


def test_local_energy_flip_and_batch_independence(cuda_device):
    img, _ = util.synth_batch(3, 256, 128)
    e, s = gpu_energy(cuda_device, img)
    ef, sf = gpu_energy(cuda_device, img[:, :, ::-1].copy())
    assert np.abs(ef[:, :, ::-1] - e).max() <= 2e-6 * e.max()
    e1, s1 = gpu_energy(cuda_device, img[1:2])
    assert np.array_equal(e1[0], e[1]) and s1[0] == s[1]          # bitwise independent of batch mates
    e2, s2 = gpu_energy(cuda_device, img)
    assert np.array_equal(e2, e) and np.array_equal(s2, s)        # bitwise repeatable


def test_local_energy_rejects_bad_shapes(cuda_device):
    from ultrasoundaugmentation_b200 import ops
    for shape in [(1, 3, 128), (1, 128, 2), (1, 4096, 8), (1, 8, 2049)]:
        with pytest.raises(ValueError):
            ops.local_energy(torch.zeros(shape, device=cuda_device))
    with pytest.raises(ValueError):
        ops.local_energy(torch.zeros(1, 64, 64))                  # host tensor: no CPU fallback


def test_snr_with_local_energy_signal(cuda_device):
    """K3 general form against the oracle, signal and confidence map given."""
    from ultrasoundaugmentation_b200 import ops
    B, H, W = 2, 128, 256
    img, lab = util.synth_batch(B, H, W)
    p = util.sample_params(B, H)
    rng = np.random.default_rng(2)
    cm = rng.random((B, H, W), dtype=np.float32)
    sig = rng.random((B, H, W), dtype=np.float32) * 3.0
    scale = np.array([0.3, 0.25], np.float32)
    taps = odeform.gaussian_taps(2.0)
    ref = np.stack([osr.reverb(ole.snr_apply_signal(img[b], cm[b], (sig[b] * scale[b]).astype(np.float32), p["a_snr"][b]),
                               lab[b], p["rho"][b], p["n_ghost"][b], taps) for b in range(B)])
    got = ops.snr_reverb(util.cuda(img, cuda_device), util.cuda(cm, cuda_device), util.cuda(lab, cuda_device),
                         torch.from_numpy(p["a_snr"]), torch.from_numpy(p["rho"]), torch.from_numpy(p["n_ghost"]),
                         torch.from_numpy(taps), signal=util.cuda(sig, cuda_device), signal_scale=util.cuda(scale, cuda_device))
    assert util.rel_err(got.cpu().numpy(), ref) <= util.REL_TOL


@pytest.mark.parametrize("B,H,W", [(2, 128, 128), (1, 97, 132)])
def test_full_chain_with_local_energy(cuda_device, B, H, W):
    import ultrasoundaugmentation_b200 as us
    RTOL = 1e-8
    img, lab = util.synth_batch(B, H, W)
    chain = us.PhysicsChain.full(seed=5, rtol=RTOL, signal="local_energy")
    params = chain.get_params(B, image_size=(H, W))
    gi, gl = chain(util.cuda(img, cuda_device), util.cuda(lab, cuda_device), params)
    torch.cuda.synchronize()
    p = {k: v.numpy() for k, v in params.items()}
    wt, ft = odeform.gaussian_taps(0.02 * W), odeform.gaussian_taps(2.0)
    for b in range(B):
        i1, l1, _ = odeform.fused_tgc_deform(img[b], lab[b], p["d"][b], p["tgc"][b], wt)
        cm = ocm.confidence_map(i1)
        sig = ole.local_energy(i1)
        ref = osr.reverb(ole.snr_apply_signal(i1, cm.astype(np.float32), sig.astype(np.float32), p["a_snr"][b]), l1,
                         p["rho"][b], p["n_ghost"][b], ft)
        assert np.array_equal(gl[b].cpu().numpy(), l1)
        err = np.abs(gi[b].cpu().numpy().astype(np.float64) - ref)
        tol = 0.4 * (util.tol_cm(RTOL) + TOL_E)
        assert (err <= tol + util.REL_TOL * np.maximum(np.abs(ref), util.REL_DEN_FLOOR)).all(), err.max()
    # the default chain is untouched by the new option
    base = us.PhysicsChain.full(seed=5, rtol=RTOL)
    bi, _ = base(util.cuda(img, cuda_device), util.cuda(lab, cuda_device), params)
    ri, _ = ochain.full_chain_batch(img, lab, p, wt, ft)
    assert np.abs(bi.cpu().numpy() - ri).max() <= 0.4 * util.tol_cm(RTOL) + 2e-5
