"""8-bit ingest / egress and the double-buffered host pipeline.

Ingest (v/255, correctly rounded) is exact, so K1 keeps its bars: labels bit-exact, image <= 1e-5 relative.
Egress quantises floor(255 v + 0.5): a result within the fp32 tolerance of the oracle's can land on the other
side of a rounding boundary, so the bar is <= 1 LSB everywhere and a bounded fraction of flipped pixels.
"""
import numpy as np
import pytest
import torch

from oracle import chain as ochain
from oracle import deform as odeform
from oracle import snr_reverb as osr
from tests import util

pytestmark = pytest.mark.gpu
RTOL = 1e-8


def to_u8(img):
    return np.floor(img * 255.0 + 0.5).astype(np.uint8)


@pytest.mark.parametrize("B,H,W", [(2, 256, 256), (2, 97, 131), (1, 33, 1030), (2, 3, 7), (1, 40, 1)])
def test_u8_ingest_warp_parity(cuda_device, B, H, W):
    from ultrasoundaugmentation_b200 import ops
    img, lab = util.synth_batch(B, H, W)
    img8 = to_u8(img)
    p = util.sample_params(B, H)
    taps = util.warp_taps(W)
    ref_i, ref_l = odeform.fused_tgc_deform_batch(img8, lab, p["d"], p["tgc"], taps)
    o, l = ops.fused_tgc_deform(util.cuda(img8, cuda_device), util.cuda(lab, cuda_device), torch.from_numpy(p["d"]),
                                torch.from_numpy(p["tgc"]), torch.from_numpy(taps))
    torch.cuda.synchronize()
    assert o.dtype == torch.float32
    assert np.array_equal(l.cpu().numpy(), ref_l)
    assert util.rel_err(o.cpu().numpy(), ref_i) <= util.REL_TOL
    # the same pixels given as fp32 v/255 produce the same bits: ingest is exact
    o2, l2 = ops.fused_tgc_deform(util.cuda(odeform.ingest_u8(img8), cuda_device), util.cuda(lab, cuda_device),
                                  torch.from_numpy(p["d"]), torch.from_numpy(p["tgc"]), torch.from_numpy(taps))
    assert torch.equal(o, o2) and torch.equal(l, l2)


def test_u8_all_values_round_trip(cuda_device):
    """Every 8-bit value through ingest -> identity chain -> egress comes back unchanged (size-independent property)."""
    import ultrasoundaugmentation_b200 as us
    H, W = 64, 256
    img8 = np.tile(np.arange(256, dtype=np.uint8)[None, None, :], (2, H, 1))
    img8[1] = img8[1, :, ::-1]
    lab = np.zeros((2, H, W), np.uint8); lab[:, 40:44, 30:200] = 1
    chain = us.PhysicsChain([us.TimeGainCompensation(p=0.0)], output_dtype=torch.uint8)
    oi, ol = chain(util.cuda(img8, cuda_device), util.cuda(lab, cuda_device))
    assert oi.dtype == torch.uint8
    assert np.array_equal(oi.cpu().numpy(), img8) and np.array_equal(ol.cpu().numpy(), lab)
    # float32 in, uint8 out == the oracle's quantiser on the same values
    chain_f = us.PhysicsChain([us.TimeGainCompensation(p=0.0)], output_dtype=torch.uint8)
    f = np.random.default_rng(3).random((2, H, W), dtype=np.float32)
    of, _ = chain_f(util.cuda(f, cuda_device), util.cuda(lab, cuda_device))
    assert np.array_equal(of.cpu().numpy(), osr.quantize_u8(f))


@pytest.mark.parametrize("with_cm", [True, False])
def test_u8_egress_snr_reverb(cuda_device, with_cm):
    from ultrasoundaugmentation_b200 import ops
    B, H, W = 3, 160, 200
    img, lab = util.synth_batch(B, H, W)
    p = util.sample_params(B, H)
    cm = np.random.default_rng(1).random((B, H, W), dtype=np.float32) if with_cm else None
    taps = odeform.gaussian_taps(2.0)
    ref = osr.quantize_u8(osr.snr_reverb_batch(img, cm if with_cm else [None] * B, lab, p["a_snr"], p["rho"], p["n_ghost"], taps))
    got = ops.snr_reverb(util.cuda(img, cuda_device), util.cuda(cm, cuda_device) if with_cm else None,
                         util.cuda(lab, cuda_device), torch.from_numpy(p["a_snr"]), torch.from_numpy(p["rho"]),
                         torch.from_numpy(p["n_ghost"]), torch.from_numpy(taps), out_dtype=torch.uint8)
    diff = np.abs(got.cpu().numpy().astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= 1
    assert (diff != 0).mean() <= 1e-3          # fp32 agreement 1e-5 relative => ~255e-5 of the pixels can flip at most


def test_u8_full_chain_parity(cuda_device):
    import ultrasoundaugmentation_b200 as us
    B, H, W = 2, 128, 128
    img, lab = util.synth_batch(B, H, W)
    img8 = to_u8(img)
    chain = us.PhysicsChain.full(seed=5, rtol=RTOL, output_dtype=torch.uint8)
    params = chain.get_params(B, image_size=(H, W))
    gi, gl = chain(util.cuda(img8, cuda_device), util.cuda(lab, cuda_device), params)
    torch.cuda.synchronize()
    assert gi.dtype == torch.uint8
    p = {k: v.numpy() for k, v in params.items()}
    ri, rl = ochain.full_chain_batch(img8, lab, p, odeform.gaussian_taps(0.02 * W), odeform.gaussian_taps(2.0))
    assert np.array_equal(gl.cpu().numpy(), rl)
    diff = np.abs(gi.cpu().numpy().astype(np.int32) - osr.quantize_u8(ri).astype(np.int32))
    assert diff.max() <= 1
    # image tolerance before quantisation is 0.4 * tol_cm + 1e-5 relative (~2e-4 => 0.05 LSB): few pixels may flip
    assert (diff != 0).mean() <= 2e-2


@pytest.mark.parametrize("u8", [False, True])
def test_host_stream_matches_augment_host(cuda_device, u8):
    """The double-buffered pipeline returns, batch by batch and in order, the bits augment_host returns."""
    import ultrasoundaugmentation_b200 as us
    chain = us.PhysicsChain.full(seed=9, rtol=1e-6, output_dtype=torch.uint8 if u8 else torch.float32)
    shapes = [(3, 64, 64), (2, 96, 64), (3, 64, 64), (1, 64, 128), (3, 64, 64)]
    batches, lo = [], 0
    for (B, H, W) in shapes:
        img, lab = util.synth_batch(B, H, W, seed0=100 + lo)
        ti = torch.from_numpy(to_u8(img) if u8 else img).pin_memory()
        batches.append((ti, torch.from_numpy(lab).pin_memory(), torch.arange(lo, lo + B)))
        lo += B
    want = [tuple(t.clone() for t in chain.augment_host(i, l, sample_index=s, device=cuda_device)) for i, l, s in batches]
    got = []
    for oi, ol in chain.augment_host_stream(iter(batches), device=cuda_device, depth=2):
        got.append((oi.clone(), ol.clone()))           # ring buffers are reused: keep copies
    assert len(got) == len(want)
    for (gi, gl), (wi, wl) in zip(got, want):
        assert gi.dtype == wi.dtype and gi.shape == wi.shape
        assert torch.equal(gi, wi) and torch.equal(gl, wl)


def test_host_stream_rejects_cuda_and_bad_depth(cuda_device):
    import ultrasoundaugmentation_b200 as us
    chain = us.PhysicsChain([us.TimeGainCompensation()])
    with pytest.raises(ValueError):
        list(chain.augment_host_stream([], device=cuda_device, depth=0))
    t = torch.zeros(1, 8, 8, device=cuda_device)
    with pytest.raises(ValueError):
        list(chain.augment_host_stream([(t, t.to(torch.uint8))], device=cuda_device))
