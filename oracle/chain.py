"""Oracle for the full chain: deformation(+TGC) -> confidence map -> SNR -> reverb.

Test infrastructure (SURVEY.md Appendix A "Chain order"; reference mount = README.md:1-2
only; parity unpinned).
"""
from __future__ import annotations

import numpy as np

from . import confmap, deform, snr_reverb

f32 = np.float32


def full_chain(image, label, d, gains, a_snr, rho, n_ghost, warp_taps, feather_taps,
               alpha=2.0, beta=90.0, gamma=0.05, return_cm=False, sink="bottom_row"):
    """One image through A.1-A.5.  Returns (image_out f32, label_out u8[, cm f64]).
    sink: "bottom_row", "bottom_and_sides" or "label" (the deformed label's bone pixels are further sinks)."""
    img1, lab1, _ = deform.fused_tgc_deform(image, label, d, gains, warp_taps)
    if sink == "label":
        cm = confmap.confidence_map(img1, alpha, beta, gamma, sink="mask", mask=lab1)
    else:
        cm = confmap.confidence_map(img1, alpha, beta, gamma, sink=sink)
    out = snr_reverb.snr_reverb(img1, cm.astype(f32), lab1, a_snr, rho, n_ghost, feather_taps)
    if return_cm:
        return out, lab1, cm
    return out, lab1


def full_chain_batch(images, labels, params, warp_taps, feather_taps,
                     alpha=2.0, beta=90.0, gamma=0.05, sink="bottom_row"):
    outs, labs = [], []
    for b in range(images.shape[0]):
        o, l = full_chain(images[b], labels[b], params["d"][b], params["tgc"][b],
                          params["a_snr"][b], params["rho"][b], params["n_ghost"][b],
                          warp_taps, feather_taps, alpha, beta, gamma, sink=sink)
        outs.append(o)
        labs.append(l)
    return np.stack(outs), np.stack(labs)
