"""Generates tests/golden/golden_v2.npz from the oracle (run from the repo root):

    python tests/golden/make_golden_v2.py

Fixtures for the SURVEY.md 8(f) rows built after golden_v1: 8-bit ingest / egress and the monogenic-signal
local energy as the SNR signal map.  Like golden_v1 these are outputs of the SELF-WRITTEN oracle (the reference
mount holds no code or vectors); they pin the oracle against drift and give the CUDA path a committed target.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import confmap, deform, local_energy, snr_reverb, synth  # noqa: E402

H, W, B = 64, 32, 2            # powers of two: the local-energy FFT needs them


def build():
    img, lab = synth.make_batch(B, H, W, seed0=8765)
    img8 = np.floor(img * 255.0 + 0.5).astype(np.uint8)
    rng = np.random.default_rng(77)
    d = rng.uniform(-0.05 * H, 0.15 * H, B).astype(np.float32)
    tgc = rng.uniform(0.6, 1.4, (B, 4)).astype(np.float32)
    a_snr = rng.uniform(0.7, 1.4, B).astype(np.float32)
    rho = rng.uniform(0.3, 0.7, B).astype(np.float32)
    n_ghost = np.array([2, 1], np.int32)
    wt = deform.gaussian_taps(0.02 * W)
    ft = deform.gaussian_taps(2.0)
    k1_img, k1_lab = deform.fused_tgc_deform_batch(img8, lab, d, tgc, wt)          # 8-bit ingest
    cm = confmap.confidence_map_batch(k1_img)
    energy = local_energy.local_energy_batch(k1_img)                                # E_hat, fp64
    out = np.stack([snr_reverb.reverb(local_energy.snr_apply_signal(k1_img[b], cm[b].astype(np.float32),
                                                                     energy[b].astype(np.float32), a_snr[b]),
                                      k1_lab[b], rho[b], n_ghost[b], ft) for b in range(B)])
    return dict(img8=img8, lab=lab, d=d, tgc=tgc, a_snr=a_snr, rho=rho, n_ghost=n_ghost, warp_taps=wt, feather_taps=ft,
                k1_img=k1_img, k1_lab=k1_lab, cm=cm, energy=energy, out_img=out, out_u8=snr_reverb.quantize_u8(out))


if __name__ == "__main__":
    out = Path(__file__).with_name("golden_v2.npz")
    np.savez_compressed(out, **build())
    print("wrote", out, out.stat().st_size, "bytes")
