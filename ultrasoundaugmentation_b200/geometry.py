"""Probe geometry of the curvilinear (convex-array) case.

The chain's physics is formulated along scanlines.  For a linear probe the scanlines are the image columns; for a
convex probe they are rays through a virtual apex above the image, so the Cartesian display image is resampled
onto the polar grid of rays x samples (``ops.cart_to_polar``), the chain runs there unchanged (rows = samples along
the ray, columns = rays), and the result is resampled back (``ops.polar_to_cart``).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class ConvexGeometry:
    """Virtual apex (apex_x, apex_y) in Cartesian pixel coordinates (apex_y <= 0: above the image), rays
    theta_j = -theta_max + j * dtheta measured from the +y axis, samples r_i = r_min + i * dr [pixels]."""
    apex_x: float
    apex_y: float
    theta_max: float
    r_min: float
    r_max: float
    n_samples: int
    n_rays: int

    def __post_init__(self):
        if not (0.0 < self.theta_max < np.pi / 2):
            raise ValueError("theta_max must lie in (0, pi/2)")
        if not (self.r_max > self.r_min >= 0.0):
            raise ValueError("need 0 <= r_min < r_max")
        if self.n_samples < 2 or self.n_rays < 2:
            raise ValueError("need at least 2 samples and 2 rays")

    @classmethod
    def default(cls, H: int, W: int, n_samples: int | None = None, n_rays: int | None = None,
                theta_max: float = 0.6, apex_above: float = 0.35) -> "ConvexGeometry":
        """A fan that covers most of an H x W display image: apex ``apex_above * H`` pixels above the top centre."""
        ay = -apex_above * H
        return cls(apex_x=(W - 1) / 2.0, apex_y=ay, theta_max=theta_max, r_min=-ay, r_max=float(H - 1 - ay),
                   n_samples=n_samples or H, n_rays=n_rays or W)

    def steps(self):
        dr = (self.r_max - self.r_min) / (self.n_samples - 1)
        dth = 2.0 * self.theta_max / (self.n_rays - 1)
        return dr, dth

    def tables(self):
        """(sin, cos) of the ray angles, fp64 -> fp32: passed to the kernel as data, like the filter taps."""
        _, dth = self.steps()
        th = -self.theta_max + dth * np.arange(self.n_rays, dtype=np.float64)
        return np.sin(th).astype(np.float32), np.cos(th).astype(np.float32)

    def scalars(self):
        """fp32 scalars of the C ABI: apex_x, apex_y, theta_max, r_min, dr, 1/dr, 1/dtheta."""
        dr, dth = self.steps()
        f = np.float32
        return (f(self.apex_x), f(self.apex_y), f(self.theta_max), f(self.r_min), f(dr), f(1.0 / dr), f(1.0 / dth))
