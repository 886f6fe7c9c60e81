"""Oracle for probe-pressure deformation + TGC + remap.

Test infrastructure, self-written (reference = /root/reference/README.md:2 names the topic
only; parity unpinned).  The coordinate path is written as individually rounded float32
operations (no fused multiply-add) because label parity is bit-exact: a 1-ulp difference
in ``y_src`` near a .5 tie flips a nearest-neighbour pick.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def gaussian_taps(sigma: float, radius: int | None = None) -> np.ndarray:
    """A.1 step 3 taps: exp(-k^2 / 2 sigma^2), normalised in fp64, rounded to fp32."""
    if radius is None:
        radius = int(np.ceil(3.0 * float(sigma)))
    radius = max(int(radius), 0)
    k = np.arange(-radius, radius + 1, dtype=np.float64)
    if sigma < 1e-6:                      # degenerate sigma: identity tap
        t = (k == 0).astype(np.float64)
    else:
        t = np.exp(-(k * k) / (2.0 * float(sigma) ** 2))
    t /= t.sum()
    return t.astype(f32)


def bone_surface(label: np.ndarray):
    """A.1 steps 1-2.  label [H,W] (bone = non-zero).

    Returns (s int32[W], y_top int, has_bone bool).  s(x) = first bone row of column x;
    columns without bone take the value of the nearest bone column (ties -> lower x);
    an image without bone gets s = H-1 everywhere and y_top = -1.
    """
    H, W = label.shape
    bone = label != 0
    col_has = bone.any(axis=0)
    first = bone.argmax(axis=0).astype(np.int32)
    if not col_has.any():
        return np.full(W, H - 1, dtype=np.int32), -1, False
    cols = np.flatnonzero(col_has)
    x = np.arange(W)
    idx = np.searchsorted(cols, x)              # first bone column >= x
    has_r = idx < len(cols)
    has_l = idx > 0
    right = cols[np.minimum(idx, len(cols) - 1)]
    left = cols[np.maximum(idx - 1, 0)]
    big = np.iinfo(np.int64).max
    dist_r = np.where(has_r, right - x, big)    # 0 when x itself is a bone column
    dist_l = np.where(has_l, x - left, big)     # >= 1 always (left < x)
    pick = np.where(dist_l <= dist_r, left, right)
    s = first[pick].astype(np.int32)
    y_top = int(first[cols].min())
    return s, y_top, True


def smooth_profile(s: np.ndarray, taps: np.ndarray) -> np.ndarray:
    """A.1 step 3: s_hat(x) = sum_k t_k * s(clamp(x+k)), fp32, increasing k, mul then add."""
    W = s.shape[0]
    R = (len(taps) - 1) // 2
    sf = s.astype(f32)
    x = np.arange(W)
    acc = np.zeros(W, dtype=f32)
    for k in range(-R, R + 1):
        idx = np.clip(x + k, 0, W - 1)
        prod = (f32(taps[k + R]) * sf[idx]).astype(f32)
        acc = (acc + prod).astype(f32)
    return acc


def clamp_displacement(d: float, s_hat: np.ndarray) -> np.float32:
    """A.1 step 4: |d| <= 0.5 * min_x s_hat(x) (fp32)."""
    lim = f32(f32(0.5) * f32(s_hat.min()))
    return f32(min(max(f32(d), f32(-lim)), lim))


def source_rows(H: int, s_hat: np.ndarray, d_eff: np.float32) -> np.ndarray:
    """A.1 step 5 backward map; returns y_src [H,W] fp32.

    den = s_hat - d;  t = y/den if y < den else 1;  y_src = y + d*t  -- each op rounded
    to fp32 on its own.  ``y < den`` replaces min(y/den, 1) so den <= 0 is well defined.
    """
    y = np.arange(H, dtype=f32)[:, None]
    den = (s_hat[None, :].astype(f32) - f32(d_eff)).astype(f32)
    den_b = np.broadcast_to(den, (H, s_hat.shape[0]))
    y_b = np.broadcast_to(y, den_b.shape)
    below = y_b < den_b
    safe = np.where(below, den_b, f32(1.0)).astype(f32)
    t = np.where(below, (y_b / safe).astype(f32), f32(1.0)).astype(f32)
    prod = (f32(d_eff) * t).astype(f32)
    return (y_b + prod).astype(f32)


def tgc_gain(H: int, gains: np.ndarray) -> np.ndarray:
    """A.2 TGC: K control gains at depths k(H-1)/(K-1), linear in output depth y. fp32."""
    gains = np.asarray(gains, dtype=f32)
    K = gains.shape[0]
    if K == 1 or H == 1:
        return np.full(H, gains[0], dtype=f32)
    scale = f32(float(K - 1) / float(H - 1))
    u = (np.arange(H, dtype=f32) * scale).astype(f32)
    k0 = np.minimum(np.floor(u).astype(np.int64), K - 2)
    fr = (u - k0.astype(f32)).astype(f32)
    g0 = gains[k0]
    g1 = gains[k0 + 1]
    return (g0 + fr * (g1 - g0)).astype(f32)


def remap(image: np.ndarray, label: np.ndarray, y_src: np.ndarray, gain_y: np.ndarray):
    """A.2: axial 2-tap bilinear on the image (zeros outside), TGC, clip; nearest on label."""
    H, W = image.shape
    x = np.broadcast_to(np.arange(W)[None, :], (H, W))
    y0f = np.floor(y_src).astype(f32)
    fr = (y_src - y0f).astype(f32)
    i0 = y0f.astype(np.int64)
    i1 = i0 + 1
    in0 = (i0 >= 0) & (i0 <= H - 1)
    in1 = (i1 >= 0) & (i1 <= H - 1)
    v0 = np.where(in0, image[np.clip(i0, 0, H - 1), x], f32(0)).astype(f32)
    v1 = np.where(in1, image[np.clip(i1, 0, H - 1), x], f32(0)).astype(f32)
    v = ((f32(1.0) - fr) * v0 + fr * v1).astype(f32)
    out = np.clip((gain_y[:, None] * v).astype(f32), f32(0), f32(1)).astype(f32)

    yn = np.floor((y_src + f32(0.5)).astype(f32)).astype(np.int64)
    inn = (yn >= 0) & (yn <= H - 1)
    lab = np.where(inn, label[np.clip(yn, 0, H - 1), x], 0).astype(label.dtype)
    return out, lab


def ingest_u8(image_u8: np.ndarray) -> np.ndarray:
    """8-bit B-mode ingest: v / 255, one correctly rounded fp32 division."""
    return (np.asarray(image_u8, dtype=np.uint8).astype(f32) / f32(255.0)).astype(f32)


def fused_tgc_deform(image, label, d, gains, taps):
    """One image: A.1 + A.2.  Returns (image_out f32, label_out u8, info dict).
    A uint8 image is ingested with ``ingest_u8`` first."""
    if np.asarray(image).dtype == np.uint8:
        image = ingest_u8(image)
    image = np.asarray(image, dtype=f32)
    label = np.asarray(label)
    H, W = image.shape
    s, y_top, has_bone = bone_surface(label)
    s_hat = smooth_profile(s, np.asarray(taps, dtype=f32))
    d_eff = clamp_displacement(d, s_hat)
    y_src = source_rows(H, s_hat, d_eff)
    g = tgc_gain(H, gains)
    out, lab = remap(image, label, y_src, g)
    return out, lab, dict(s=s, s_hat=s_hat, d_eff=d_eff, y_top=y_top, has_bone=has_bone,
                          y_src=y_src, gain=g)


def fused_tgc_deform_batch(images, labels, d, gains, taps):
    """Batch wrapper: images [B,H,W] f32, labels [B,H,W] u8, d [B], gains [B,K]."""
    outs, labs = [], []
    for b in range(images.shape[0]):
        o, l, _ = fused_tgc_deform(images[b], labels[b], d[b], gains[b], taps)
        outs.append(o)
        labs.append(l)
    return np.stack(outs), np.stack(labs)
