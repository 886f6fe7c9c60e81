"""Full chain parity through the public transform API (deform+TGC -> CM -> SNR -> reverb)."""
import numpy as np
import pytest
import torch

from oracle import chain as ochain
from oracle import deform as odeform
from tests import util

pytestmark = pytest.mark.gpu
RTOL = 1e-8


def make_chain(**kw):
    import ultrasoundaugmentation_b200 as us
    return us.PhysicsChain.full(seed=5, rtol=RTOL, **kw)


def oracle_chain(img, lab, params, W):
    p = {k: v.numpy() for k, v in params.items()}
    return ochain.full_chain_batch(img, lab, p, odeform.gaussian_taps(0.02 * W), odeform.gaussian_taps(2.0))


@pytest.mark.parametrize("B,H,W", [(2, 128, 128), (2, 96, 160), (1, 256, 256)])
def test_full_chain_parity(cuda_device, B, H, W):
    img, lab = util.synth_batch(B, H, W)
    chain = make_chain()
    params = chain.get_params(B, image_size=(H, W))
    gi, gl = chain(util.cuda(img, cuda_device).unsqueeze(1), util.cuda(lab, cuda_device).unsqueeze(1), params)
    torch.cuda.synchronize()
    assert gi.shape == (B, 1, H, W) and gl.shape == (B, 1, H, W)
    ri, rl = oracle_chain(img, lab, params, W)
    assert np.array_equal(gl[:, 0].cpu().numpy(), rl)
    # image tolerance: SNR term multiplies C (tol_cm) by |a-1| * I <= 0.4, plus fp32 1e-5 relative
    tol = 0.4 * util.tol_cm(RTOL)
    err = np.abs(gi[:, 0].cpu().numpy().astype(np.float64) - ri)
    assert (err <= tol + util.REL_TOL * np.maximum(np.abs(ri), util.REL_DEN_FLOOR)).all(), err.max()


def test_shapes_and_single_transforms(cuda_device):
    import ultrasoundaugmentation_b200 as us
    img, lab = util.synth_batch(1, 64, 64)
    ti, tl = util.cuda(img[0], cuda_device), util.cuda(lab[0], cuda_device)
    for t in (us.ProbePressureDeformation(seed=1), us.TimeGainCompensation(seed=1), us.Reverberation(seed=1),
              us.ConfidenceSNR(seed=1, rtol=1e-6)):
        oi, ol = t(ti, tl)
        assert oi.shape == ti.shape and ol.shape == tl.shape and oi.is_cuda
        oi3, _ = t(ti[None], tl[None])
        assert torch.equal(oi3[0], oi)
    # TGC alone leaves labels untouched; reverb alone leaves labels untouched
    _, ol = us.TimeGainCompensation(seed=1)(ti, tl)
    assert torch.equal(ol, tl)


def test_probability_zero_is_identity(cuda_device):
    import ultrasoundaugmentation_b200 as us
    img, lab = util.synth_batch(2, 64, 64)
    ti, tl = util.cuda(img, cuda_device), util.cuda(lab, cuda_device)
    chain = us.PhysicsChain([us.ProbePressureDeformation(p=0.0), us.TimeGainCompensation(p=0.0),
                             us.Reverberation(p=0.0)])
    oi, ol = chain(ti, tl)
    assert torch.equal(oi, ti) and torch.equal(ol, tl)


def test_host_path_equals_device_path(cuda_device):
    img, lab = util.synth_batch(2, 96, 96)
    chain = make_chain()
    params = chain.get_params(2, image_size=(96, 96))
    di, dl = chain(util.cuda(img, cuda_device), util.cuda(lab, cuda_device), params)
    hi, hl = chain.augment_host(torch.from_numpy(img).pin_memory(), torch.from_numpy(lab).pin_memory(), params)
    assert not hi.is_cuda and torch.equal(hi, di.cpu()) and torch.equal(hl, dl.cpu())


def test_sharded_equals_unsharded(cuda_device):
    """N-way sharding (emulated on one GPU) reproduces the single-shard result bit for bit."""
    from ultrasoundaugmentation_b200 import sharding
    B, H, W = 6, 64, 64
    img, lab = util.synth_batch(B, H, W)
    chain = make_chain()
    full_i, full_l = chain(util.cuda(img, cuda_device), util.cuda(lab, cuda_device),
                           chain.get_params(B, image_size=(H, W)))
    for world in (2, 4):
        parts_i, parts_l = [], []
        for r in range(world):
            lo, hi = sharding.shard_bounds(B, r, world)
            p = sharding.shard_params(chain, B, r, world, (H, W))
            oi, ol = chain(util.cuda(img[lo:hi], cuda_device), util.cuda(lab[lo:hi], cuda_device), p)
            parts_i.append(oi); parts_l.append(ol)
        assert torch.equal(torch.cat(parts_i), full_i) and torch.equal(torch.cat(parts_l), full_l)


def test_graphed_chain_is_bitwise_the_eager_chain(cuda_device):
    """PhysicsChain.capture: the whole chain as one CUDA-graph launch (parameter block uploaded from pinned memory by
    a graph node, cooperative PCG kernel inside the graph).  Replays with new inputs and parameters must reproduce the
    eager results bit for bit."""
    B, H, W = 3, 96, 128
    img, lab = util.synth_batch(B, H, W)
    img2, lab2 = util.synth_batch(B, H, W, seed0=99)
    chain = make_chain()
    g = chain.capture(util.cuda(img, cuda_device), util.cuda(lab, cuda_device))
    for k, (i_, l_) in enumerate(((img, lab), (img2, lab2), (img, lab))):
        params = chain.get_params(B, sample_index=np.arange(B) + 10 * k, image_size=(H, W))
        ti, tl = util.cuda(i_, cuda_device), util.cuda(l_, cuda_device)
        ei, el = chain(ti, tl, params)
        gi, gl = g(ti, tl, params)
        torch.cuda.synchronize()
        assert torch.equal(gi, ei) and torch.equal(gl, el), k
    ri, rl = oracle_chain(img, lab, params, W)                   # and the last replay against the oracle
    assert np.array_equal(gl.cpu().numpy(), rl)
    err = np.abs(gi.cpu().numpy().astype(np.float64) - ri)
    assert (err <= 0.4 * util.tol_cm(RTOL) + util.REL_TOL * np.maximum(np.abs(ri), util.REL_DEN_FLOOR)).all(), err.max()


def test_graphed_chain_with_label_sinks(cuda_device):
    """The captured chain with the bone pixels of the deformed label as sinks: the sink mask is produced inside the
    graph (K1 writes the label the PCG setup reads), so a replay with other labels must follow them."""
    B, H, W = 2, 96, 128
    chain = make_chain(sink="label")
    batches = [util.synth_batch(B, H, W), util.synth_batch(B, H, W, seed0=7)]
    g = chain.capture(util.cuda(batches[0][0], cuda_device), util.cuda(batches[0][1], cuda_device))
    for k, (i_, l_) in enumerate(batches + batches[:1]):
        params = chain.get_params(B, sample_index=np.arange(B) + 5 * k, image_size=(H, W))
        ti, tl = util.cuda(i_, cuda_device), util.cuda(l_, cuda_device)
        ei, el = chain(ti, tl, params)
        gi, gl = g(ti, tl, params)
        torch.cuda.synchronize()
        assert torch.equal(gi, ei) and torch.equal(gl, el), k


@pytest.mark.parametrize("sink", ["bottom_and_sides", "label"])
def test_full_chain_with_sink_modes(cuda_device, sink):
    """SURVEY.md 8(f) row 3 through the transform API: the confidence map of the chain absorbs walks at the image sides,
    or at the bone pixels of the DEFORMED label."""
    B, H, W = 2, 128, 128
    img, lab = util.synth_batch(B, H, W)
    chain = make_chain(sink=sink)
    params = chain.get_params(B, image_size=(H, W))
    gi, gl = chain(util.cuda(img, cuda_device), util.cuda(lab, cuda_device), params)
    torch.cuda.synchronize()
    p = {k: v.numpy() for k, v in params.items()}
    ri, rl = ochain.full_chain_batch(img, lab, p, odeform.gaussian_taps(0.02 * W), odeform.gaussian_taps(2.0), sink=sink)
    assert np.array_equal(gl.cpu().numpy(), rl)
    err = np.abs(gi.cpu().numpy().astype(np.float64) - ri)
    assert (err <= 0.4 * util.tol_cm(RTOL) + util.REL_TOL * np.maximum(np.abs(ri), util.REL_DEN_FLOOR)).all(), err.max()
    plain, _ = make_chain()(util.cuda(img, cuda_device), util.cuda(lab, cuda_device), params)
    assert not torch.equal(plain, gi)                     # the sink mode does change the result
