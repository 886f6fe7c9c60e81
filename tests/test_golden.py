"""Golden fixtures (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py from the oracle).

CPU: the oracle still reproduces them bit for bit (drift guard).  GPU: the CUDA path hits them
through the C ABI (labels bit-exact, images 1e-5 relative, CM within the stated tolerance).
"""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import chain as ochain, confmap as ocm, deform as odeform, snr_reverb as osr
from tests import util

G = np.load(Path(__file__).parent / "golden" / "golden_v1.npz")


def test_oracle_reproduces_golden_k1():
    i, l = odeform.fused_tgc_deform_batch(G["img"], G["lab"], G["d"], G["tgc"], G["warp_taps"])
    assert np.array_equal(l, G["k1_lab"])
    assert util.rel_err(i, G["k1_img"]) <= 1e-6          # fp32 libm differences only


def test_oracle_reproduces_golden_cm():
    cm = ocm.confidence_map_batch(G["k1_img"])
    assert np.abs(cm - G["cm"]).max() <= 1e-7


def test_oracle_reproduces_golden_chain():
    p = {k: G[k] for k in ("d", "tgc", "a_snr", "rho", "n_ghost")}
    i, l = ochain.full_chain_batch(G["img"], G["lab"], p, G["warp_taps"], G["feather_taps"])
    assert np.array_equal(l, G["k1_lab"])
    assert util.rel_err(i, G["out_img"]) <= 1e-5
    k3 = osr.snr_reverb_batch(G["k1_img"], G["cm"].astype(np.float32), G["k1_lab"], G["a_snr"], G["rho"],
                              G["n_ghost"], G["feather_taps"])
    assert util.rel_err(k3, G["out_img"]) <= 1e-6


@pytest.mark.gpu
def test_cuda_hits_golden(cuda_device):
    from ultrasoundaugmentation_b200 import ops
    dev = cuda_device
    t = lambda k: torch.from_numpy(G[k])
    i, l = ops.fused_tgc_deform(t("img").to(dev), t("lab").to(dev), t("d"), t("tgc"), t("warp_taps"))
    assert np.array_equal(l.cpu().numpy(), G["k1_lab"])
    assert util.rel_err(i.cpu().numpy(), G["k1_img"]) <= util.REL_TOL
    cm, it, rr = ops.confidence_map(t("k1_img").to(dev), rtol=1e-8, return_info=True)
    assert np.abs(cm.cpu().numpy() - G["cm"]).max() <= util.tol_cm(1e-8)
    out = ops.snr_reverb(t("k1_img").to(dev), torch.from_numpy(G["cm"].astype(np.float32)).to(dev),
                         t("k1_lab").to(dev), t("a_snr"), t("rho"), t("n_ghost"), t("feather_taps"))
    assert util.rel_err(out.cpu().numpy(), G["out_img"]) <= util.REL_TOL


# ---- golden_v2: 8-bit ingest / egress and the local-energy signal map ----------------
G2 = np.load(Path(__file__).parent / "golden" / "golden_v2.npz")


def test_oracle_reproduces_golden_v2():
    from oracle import local_energy as ole
    i, l = odeform.fused_tgc_deform_batch(G2["img8"], G2["lab"], G2["d"], G2["tgc"], G2["warp_taps"])
    assert i.dtype == np.float32 and np.array_equal(l, G2["k1_lab"])
    assert util.rel_err(i, G2["k1_img"]) <= 1e-6
    assert np.abs(ole.local_energy_batch(G2["k1_img"]) - G2["energy"]).max() <= 1e-9
    assert np.abs(ocm.confidence_map_batch(G2["k1_img"]) - G2["cm"]).max() <= 1e-7
    assert np.array_equal(osr.quantize_u8(G2["out_img"]), G2["out_u8"])


@pytest.mark.gpu
def test_cuda_hits_golden_v2(cuda_device):
    from ultrasoundaugmentation_b200 import ops
    dev = cuda_device
    t = lambda k: torch.from_numpy(G2[k])
    i, l = ops.fused_tgc_deform(t("img8").to(dev), t("lab").to(dev), t("d"), t("tgc"), t("warp_taps"))
    assert np.array_equal(l.cpu().numpy(), G2["k1_lab"])
    assert util.rel_err(i.cpu().numpy(), G2["k1_img"]) <= util.REL_TOL
    e, s = ops.local_energy(t("k1_img").to(dev))
    assert np.abs((e * s[:, None, None]).cpu().numpy() - G2["energy"]).max() <= 2e-5
    cm32 = torch.from_numpy(G2["cm"].astype(np.float32)).to(dev)
    args = (t("k1_img").to(dev), cm32, t("k1_lab").to(dev), t("a_snr"), t("rho"), t("n_ghost"), t("feather_taps"))
    sig = torch.from_numpy(G2["energy"].astype(np.float32)).to(dev)
    one = torch.ones(sig.shape[0], device=dev)
    out = ops.snr_reverb(*args, signal=sig, signal_scale=one)
    assert util.rel_err(out.cpu().numpy(), G2["out_img"]) <= util.REL_TOL
    out8 = ops.snr_reverb(*args, signal=sig, signal_scale=one, out_dtype=torch.uint8)
    diff = np.abs(out8.cpu().numpy().astype(np.int32) - G2["out_u8"].astype(np.int32))
    assert diff.max() <= 1 and (diff != 0).mean() <= 1e-3
