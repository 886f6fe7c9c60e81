"""On-device synthetic B-mode speckle + bone arc + acoustic shadow.

Benchmark data only (torch device RNG; H2D stays out of the timed region).  Parity tests use
the NumPy twin in oracle/synth.py instead, because device and NumPy RNG streams differ.
"""
from __future__ import annotations

import math

import torch


def make_batch(B: int, H: int, W: int, seed: int = 1234, device="cuda", chunk: int = 256):
    """Returns (image f32 [B,H,W] in [0,1], label u8 [B,H,W]) resident on `device`."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    y = torch.arange(H, device=device, dtype=torch.float32).view(1, H, 1)
    x = torch.arange(W, device=device, dtype=torch.float32).view(1, 1, W)
    surf = 0.45 * H + 0.08 * H * torch.cos((x - W / 2.0) / (0.25 * W))
    inside = (x - W / 2.0).abs() < 0.3 * W
    thick = 0.02 * H + 2.0
    on = inside & (y >= surf) & (y < surf + thick)
    below = inside & (y >= surf + thick)
    echo = 0.35 * torch.exp(-1.5 * y / H).expand(1, H, W).clone()
    echo[on] = 1.0
    echo[below] = 0.02
    img = torch.empty((B, H, W), dtype=torch.float32, device=device)
    for lo in range(0, B, chunk):
        n = min(chunk, B - lo)
        n1 = torch.randn((n, H, W), generator=gen, device=device)
        n2 = torch.randn((n, H, W), generator=gen, device=device)
        sig = torch.hypot(n1, n2) * (echo / math.sqrt(2.0))
        mx = sig.amax(dim=(1, 2), keepdim=True)
        img[lo:lo + n] = torch.log1p(255.0 * sig / mx) / math.log(256.0)
    label = on.to(torch.uint8).expand(B, H, W).contiguous()
    return img.clamp_(0.0, 1.0), label
