"""Oracle for the monogenic-signal local-energy map (A.4 general form
``I_out = clip(I + (a_snr - 1) * C * E, 0, 1)`` with E = normalised local energy instead of E = I).

Test infrastructure, self-written (the reference mount /root/reference/README.md:1-2 names the topic only;
parity unpinned).  Method [LIT: Felsberg & Sommer 2001 monogenic signal; Kovesi log-Gabor filters]:

    F      = FFT2(I)
    G(rho) = sum_s exp(-(ln(rho * lambda_s))^2 / (2 ln^2(sigma_f))),  lambda_s = lambda_min * mult^s,  G(0) = 0
    even   = Re IFFT2(G F)
    q      = IFFT2(G (v' - i u')/rho F)        (Riesz pair: odd1 = Re q, odd2 = Im q; both real for real I;
                                                u', v' = u, v with the Nyquist bins |.| = 1/2 set to 0)
    E      = sqrt(even^2 + odd1^2 + odd2^2);   E_hat = E / max E   (0 for a constant image)

u, v are the FFT frequencies in cycles/pixel along x (columns) and y (rows), rho = sqrt(u^2 + v^2).
Sizes that are not powers of two: the image is mirror-extended (numpy.pad mode="symmetric", at the bottom and
right) to the next power of two in that direction, transformed there, and E is cropped back before normalising.
Everything here is fp64 (numpy.fft); the CUDA path computes the same in fp32 (stated tolerance in the tests).
Pins that do not depend on any implementation: a constant image has E = 0; a plane wave cos(2 pi k0 x / W)
has E = G(k0 / W) everywhere (even and odd parts are in quadrature), so E_hat = 1.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32

DEFAULT_LAMBDA_MIN = 8.0
DEFAULT_MULT = 2.1
DEFAULT_SCALES = 3
DEFAULT_SIGMA_F = 0.55


def log_gabor_sum(H: int, W: int, lambda_min=DEFAULT_LAMBDA_MIN, mult=DEFAULT_MULT, scales=DEFAULT_SCALES,
                  sigma_f=DEFAULT_SIGMA_F):
    """Returns (G [H,W], u [1,W], v [H,1], rho [H,W]) in fp64."""
    u = np.fft.fftfreq(W)[None, :]
    v = np.fft.fftfreq(H)[:, None]
    rho = np.sqrt(u * u + v * v)
    safe = np.where(rho > 0, rho, 1.0)
    G = np.zeros((H, W))
    den = 2.0 * np.log(sigma_f) ** 2
    for s in range(int(scales)):
        lam = float(lambda_min) * float(mult) ** s
        G += np.exp(-(np.log(safe * lam) ** 2) / den)
    G[rho == 0] = 0.0
    return G, u, v, rho


def next_pow2(n: int) -> int:
    p = 1
    while p < n:
        p *= 2
    return p


def local_energy_raw(image, **kw) -> np.ndarray:
    """E (not normalised), fp64 [H,W]."""
    img = np.asarray(image, dtype=np.float64)
    H0, W0 = img.shape
    Hp, Wp = max(next_pow2(H0), 8), max(next_pow2(W0), 8)
    if (Hp, Wp) != (H0, W0):
        if Hp - H0 > H0 or Wp - W0 > W0:
            raise ValueError("image too small to mirror-extend (needs at least 4 pixels per side)")
        ext = np.pad(img, ((0, Hp - H0), (0, Wp - W0)), mode="symmetric")
        return local_energy_raw(ext, **kw)[:H0, :W0]
    H, W = img.shape
    G, u, v, rho = log_gabor_sum(H, W, **kw)
    F = np.fft.fft2(img)
    even = np.fft.ifft2(G * F).real
    safe = np.where(rho > 0, rho, 1.0)
    # odd multipliers vanish at the Nyquist bins (|u| = 1/2 has no +/- partner on an even-length DFT):
    # keeps q's two parts real for real images and makes the map commute with flips
    uo = np.where(np.abs(u) == 0.5, 0.0, u)
    vo = np.where(np.abs(v) == 0.5, 0.0, v)
    riesz = (vo - 1j * uo) / safe
    q = np.fft.ifft2(G * riesz * F)
    return np.sqrt(even * even + q.real * q.real + q.imag * q.imag)


def local_energy(image, **kw) -> np.ndarray:
    """E_hat = E / max E in [0,1], fp64 [H,W]; all zeros for an image without band-pass content."""
    E = local_energy_raw(image, **kw)
    m = E.max()
    return E / m if m > 0 else np.zeros_like(E)


def local_energy_batch(images, **kw) -> np.ndarray:
    return np.stack([local_energy(im, **kw) for im in images])


def snr_apply_signal(image, cm, signal, a_snr) -> np.ndarray:
    """A.4 general form: clip(I + ((a-1) * C) * E, 0, 1), fp32, the products rounded one by one."""
    image = np.asarray(image, dtype=f32)
    if cm is None:
        return image.copy()
    cm = np.asarray(cm, dtype=f32)
    sig = np.asarray(signal, dtype=f32)
    k = f32(f32(a_snr) - f32(1.0))
    out = (image + ((k * cm).astype(f32) * sig).astype(f32)).astype(f32)
    return np.clip(out, f32(0), f32(1)).astype(f32)
