"""Oracle for the SNR/attenuation change (A.4) and the reverberation artifact (A.5).

Test infrastructure, self-written from SURVEY.md Appendix A.4-A.5 (reference mount =
README.md:1-2 only; parity unpinned).  fp32 throughout.
"""
from __future__ import annotations

import numpy as np

from .deform import bone_surface

f32 = np.float32


def snr_apply(image, cm, a_snr):
    """A.4 with E = I:  clip(I + (a-1) * C * I, 0, 1)."""
    image = np.asarray(image, dtype=f32)
    if cm is None:
        return image.copy()
    cm = np.asarray(cm, dtype=f32)
    k = f32(f32(a_snr) - f32(1.0))
    out = (image + ((k * cm).astype(f32) * image).astype(f32)).astype(f32)
    return np.clip(out, f32(0), f32(1)).astype(f32)


def feather_mask(label, taps):
    """A.5 feathered mask: separable Gaussian of 1[L != 0], zero outside the image.

    Row pass (along x, increasing tap index) then column pass (along y, increasing tap index).
    """
    taps = np.asarray(taps, dtype=f32)
    R = (len(taps) - 1) // 2
    m = (np.asarray(label) != 0).astype(f32)
    H, W = m.shape
    pad = np.zeros((H + 2 * R, W + 2 * R), f32)
    pad[R:R + H, R:R + W] = m
    rows = np.zeros((H + 2 * R, W), f32)
    for i in range(2 * R + 1):
        rows = (rows + (taps[i] * pad[:, i:i + W]).astype(f32)).astype(f32)
    out = np.zeros((H, W), f32)
    for j in range(2 * R + 1):
        out = (out + (taps[j] * rows[j:j + H, :]).astype(f32)).astype(f32)
    return out


def reverb(image, label, rho, n_ghost, taps, y_top=None):
    """A.5:  I_out = clip(I + sum_{k=1..n} rho^k * Mf[y-k*D, x] * I[y-k*D, x], 0, 1).

    D = first bone row of the image (y_top); no bone or D < 1 or n < 1 or rho == 0 -> identity.
    """
    image = np.asarray(image, dtype=f32)
    H, W = image.shape
    if y_top is None:
        _, y_top, has = bone_surface(np.asarray(label))
    else:
        has = y_top >= 0
    n = int(n_ghost)
    if (not has) or y_top < 1 or n < 1:
        return image.copy()
    g = (feather_mask(label, taps) * image).astype(f32)
    acc = image.copy()
    coef = f32(1.0)
    for k in range(1, n + 1):
        coef = f32(coef * f32(rho))
        sh = k * int(y_top)
        if sh >= H:
            break
        acc[sh:, :] = (acc[sh:, :] + (coef * g[:H - sh, :]).astype(f32)).astype(f32)
    return np.clip(acc, f32(0), f32(1)).astype(f32)


def snr_reverb(image, cm, label, a_snr, rho, n_ghost, taps):
    """A.4 then A.5 on one image (the GPU fuses the two)."""
    return reverb(snr_apply(image, cm, a_snr), label, rho, n_ghost, taps)


def quantize_u8(image: np.ndarray) -> np.ndarray:
    """8-bit egress: floor(255 * v + 0.5), the product and the sum rounded to fp32 on their own."""
    v = np.asarray(image, dtype=f32)
    q = np.floor(((v * f32(255.0)).astype(f32) + f32(0.5)).astype(f32))
    return q.astype(np.uint8)


def snr_reverb_batch(images, cms, labels, a_snr, rho, n_ghost, taps):
    outs = []
    for b in range(images.shape[0]):
        cm = None if cms is None else cms[b]
        outs.append(snr_reverb(images[b], cm, labels[b], a_snr[b], rho[b], n_ghost[b], taps))
    return np.stack(outs)
