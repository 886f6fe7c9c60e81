"""Generates tests/golden/golden_v1.npz from the oracle (run from the repo root):

    python tests/golden/make_golden.py

The reference mount holds no code or vectors (README.md only), so these fixtures are outputs of
the SELF-WRITTEN oracle; they pin the oracle against accidental drift and give the CUDA path a
committed, implementation-independent-of-runtime target.  Closed-form KATs live in
tests/test_oracle_kat.py.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import chain, confmap, deform, snr_reverb, synth  # noqa: E402

H, W, B = 48, 40, 2


def build():
    img, lab = synth.make_batch(B, H, W, seed0=4321)
    rng = np.random.default_rng(99)
    d = rng.uniform(-0.05 * H, 0.15 * H, B).astype(np.float32)
    tgc = rng.uniform(0.6, 1.4, (B, 5)).astype(np.float32)
    a_snr = rng.uniform(0.7, 1.4, B).astype(np.float32)
    rho = rng.uniform(0.3, 0.7, B).astype(np.float32)
    n_ghost = np.array([1, 2], np.int32)
    wt = deform.gaussian_taps(0.02 * W)
    ft = deform.gaussian_taps(2.0)
    k1_img, k1_lab = deform.fused_tgc_deform_batch(img, lab, d, tgc, wt)
    cm = confmap.confidence_map_batch(k1_img)
    k3 = snr_reverb.snr_reverb_batch(k1_img, cm.astype(np.float32), k1_lab, a_snr, rho, n_ghost, ft)
    params = dict(d=d, tgc=tgc, a_snr=a_snr, rho=rho, n_ghost=n_ghost)
    ch_img, ch_lab = chain.full_chain_batch(img, lab, params, wt, ft)
    assert np.array_equal(ch_img, k3) and np.array_equal(ch_lab, k1_lab)
    return dict(img=img, lab=lab, d=d, tgc=tgc, a_snr=a_snr, rho=rho, n_ghost=n_ghost, warp_taps=wt,
                feather_taps=ft, k1_img=k1_img, k1_lab=k1_lab, cm=cm, out_img=k3)


if __name__ == "__main__":
    out = Path(__file__).with_name("golden_v1.npz")
    np.savez_compressed(out, **build())
    print("wrote", out, out.stat().st_size, "bytes")
