"""Seeded synthetic B-mode speckle + bone arc + acoustic shadow.

Test infrastructure.  NumPy twin of the on-device generator in
``ultrasoundaugmentation_b200/synth.py`` (device and NumPy RNG streams do not bit-match, so
parity tests always generate here and upload).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def bone_geometry(H: int, W: int):
    """Arc surf(x) = 0.45H + 0.08H cos((x-W/2)/(0.25W)) for |x-W/2| < 0.3W; thickness 0.02H+2."""
    x = np.arange(W, dtype=np.float64)
    surf = 0.45 * H + 0.08 * H * np.cos((x - W / 2.0) / (0.25 * W))
    inside = np.abs(x - W / 2.0) < 0.3 * W
    thick = 0.02 * H + 2.0
    return surf, inside, thick


def make_bmode(H: int, W: int, seed: int = 1234, bone: bool = True):
    """Returns (image f32 [H,W] in [0,1], label u8 [H,W] with 1 on bone)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n1 = rng.standard_normal((H, W))
    n2 = rng.standard_normal((H, W))
    env = np.hypot(n1, n2) / np.sqrt(2.0)
    y = np.arange(H, dtype=np.float64)[:, None]
    echo = 0.35 * np.exp(-1.5 * y / H) * np.ones((1, W))
    label = np.zeros((H, W), np.uint8)
    if bone:
        surf, inside, thick = bone_geometry(H, W)
        on = inside[None, :] & (y >= surf[None, :]) & (y < surf[None, :] + thick)
        below = inside[None, :] & (y >= surf[None, :] + thick)
        echo = np.where(on, 1.0, echo)
        echo = np.where(below, 0.02, echo)
        label[on] = 1
    sig = env * echo
    img = np.log1p(255.0 * sig / sig.max()) / np.log(256.0)
    return img.astype(f32), label


def make_batch(B: int, H: int, W: int, seed0: int = 1234, first_index: int = 0):
    """Seed = seed0 + global sample index, so shards reproduce the same images."""
    ims, labs = [], []
    for b in range(B):
        im, lb = make_bmode(H, W, seed0 + first_index + b)
        ims.append(im)
        labs.append(lb)
    return np.stack(ims), np.stack(labs)
