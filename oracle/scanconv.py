"""Oracle for curvilinear (convex-probe) geometry: scan conversion between the
Cartesian display image and the polar grid of rays x samples on which the chain (A.1-A.5) runs unchanged --
rows of the polar image are radial samples (depth along the ray), columns are rays (scanlines).

Test infrastructure, self-written (the reference mount /root/reference/README.md:1-2 names the topic only;
parity unpinned).  Geometry: virtual apex (apex_x, apex_y) in Cartesian pixel coordinates (apex_y <= 0: above the
image), rays theta_j = -theta_max + j * dtheta measured from the +y axis, samples r_i = r_min + i * dr.

Label parity is bit-exact under nearest sampling, so every COORDINATE is a fixed sequence of individually rounded
fp32 operations (no FMA), identical here and in csrc/scanconv.cu:
  * sin/cos of the ray angles are host tables (fp64 -> fp32) passed to both sides as data;
  * the arctangent of polar <- Cartesian is the polynomial of Abramowitz & Stegun 4.4.49 (|error| <= 2e-8 rad
    before rounding), evaluated by Horner's rule in fp32 -- part of the specification, not an approximation of it.
Image values are 4-tap bilinear (zeros outside), tolerance 1e-5 relative as for the other image transforms.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

f32 = np.float32

# Abramowitz & Stegun 4.4.49: atan(z) = z * (1 + sum_k a_2k z^2k), 0 <= z <= 1
ATAN_COEF = tuple(f32(c) for c in (-0.3333314528, 0.1999355085, -0.1420889944, 0.1065626393,
                                   -0.0752896400, 0.0429096138, -0.0161657367, 0.0028662257))
HALF_PI = f32(np.pi / 2)


@dataclass(frozen=True)
class ConvexGeometry:
    apex_x: float
    apex_y: float
    theta_max: float      # half opening angle [rad], < pi/2
    r_min: float          # radius of the first sample [px from the apex]
    r_max: float          # radius of the last sample
    n_samples: int        # rows of the polar image
    n_rays: int           # columns of the polar image

    def steps(self):
        dr = (self.r_max - self.r_min) / max(self.n_samples - 1, 1)
        dth = 2.0 * self.theta_max / max(self.n_rays - 1, 1)
        return dr, dth

    def tables(self):
        """(sin_j, cos_j) of the ray angles, fp64 -> fp32."""
        _, dth = self.steps()
        th = -self.theta_max + dth * np.arange(self.n_rays, dtype=np.float64)
        return np.sin(th).astype(f32), np.cos(th).astype(f32)

    def scalars(self):
        """fp32 scalars both sides use: apex_x, apex_y, theta_max, r_min, dr, 1/dr, 1/dtheta."""
        dr, dth = self.steps()
        return (f32(self.apex_x), f32(self.apex_y), f32(self.theta_max), f32(self.r_min), f32(dr),
                f32(1.0 / dr) if dr > 0 else f32(0), f32(1.0 / dth) if dth > 0 else f32(0))


def spec_atan(z: np.ndarray) -> np.ndarray:
    """atan(z) for 0 <= z <= 1: Horner in z^2, every multiply and add rounded to fp32 on its own."""
    z = np.asarray(z, dtype=f32)
    z2 = (z * z).astype(f32)
    p = np.full(z.shape, ATAN_COEF[-1], dtype=f32)
    for c in ATAN_COEF[-2::-1]:
        p = ((p * z2).astype(f32) + c).astype(f32)
    p = ((p * z2).astype(f32) + f32(1.0)).astype(f32)
    return (z * p).astype(f32)


def spec_atan2_pos(dx: np.ndarray, dy: np.ndarray) -> np.ndarray:
    """theta = atan2(dx, dy) for dy > 0 (angle from the +y axis, sign of dx), built on spec_atan."""
    dx = np.asarray(dx, dtype=f32)
    dy = np.asarray(dy, dtype=f32)
    ax = np.abs(dx)
    small = ax <= dy                                            # t = ax/dy <= 1
    num = np.where(small, ax, dy).astype(f32)
    den = np.where(small, dy, ax).astype(f32)
    den = np.where(den > 0, den, f32(1.0)).astype(f32)
    a = spec_atan((num / den).astype(f32))
    a = np.where(small, a, (HALF_PI - a).astype(f32)).astype(f32)
    return np.where(dx < 0, -a, a).astype(f32)


def _bilinear(img: np.ndarray, ys: np.ndarray, xs: np.ndarray) -> np.ndarray:
    """4-tap bilinear gather at fp32 coordinates (ys, xs); samples outside the image read 0."""
    H, W = img.shape
    y0f = np.floor(ys).astype(f32); x0f = np.floor(xs).astype(f32)
    fy = (ys - y0f).astype(f32); fx = (xs - x0f).astype(f32)
    y0 = y0f.astype(np.int64); x0 = x0f.astype(np.int64)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy <= H - 1) & (xx >= 0) & (xx <= W - 1)
        return np.where(ok, img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], f32(0)).astype(f32)

    top = ((f32(1) - fx) * tap(y0, x0) + fx * tap(y0, x0 + 1)).astype(f32)
    bot = ((f32(1) - fx) * tap(y0 + 1, x0) + fx * tap(y0 + 1, x0 + 1)).astype(f32)
    return ((f32(1) - fy) * top + fy * bot).astype(f32)


def _nearest(lab: np.ndarray, ys: np.ndarray, xs: np.ndarray) -> np.ndarray:
    H, W = lab.shape
    yn = np.floor((ys + f32(0.5)).astype(f32)).astype(np.int64)
    xn = np.floor((xs + f32(0.5)).astype(f32)).astype(np.int64)
    ok = (yn >= 0) & (yn <= H - 1) & (xn >= 0) & (xn <= W - 1)
    return np.where(ok, lab[np.clip(yn, 0, H - 1), np.clip(xn, 0, W - 1)], 0).astype(lab.dtype)


def polar_coords_of_grid(g: ConvexGeometry):
    """Cartesian (y, x) of every polar grid node, fp32 [n_samples, n_rays]."""
    ax, ay, _, rmin, dr, _, _ = g.scalars()
    s, c = g.tables()
    i = np.arange(g.n_samples, dtype=f32)[:, None]
    r = (rmin + (i * dr).astype(f32)).astype(f32)
    x = (ax + (r * s[None, :]).astype(f32)).astype(f32)
    y = (ay + (r * c[None, :]).astype(f32)).astype(f32)
    return y, x


def cart_to_polar(image, label, g: ConvexGeometry):
    """Cartesian [H,W] -> polar [n_samples, n_rays]: image bilinear, label nearest."""
    image = np.asarray(image, dtype=f32)
    y, x = polar_coords_of_grid(g)
    return _bilinear(image, y, x), _nearest(np.asarray(label), y, x)


def grid_coords_of_cart(g: ConvexGeometry, H: int, W: int):
    """Polar grid coordinates (i*, j*) of every Cartesian pixel, fp32 [H,W], and the mask of pixels with dy > 0."""
    ax, ay, thmax, rmin, _, inv_dr, inv_dth = g.scalars()
    dx = (np.arange(W, dtype=f32)[None, :] - ax).astype(f32)
    dy = (np.arange(H, dtype=f32)[:, None] - ay).astype(f32)
    dxb = np.broadcast_to(dx, (H, W)); dyb = np.broadcast_to(dy, (H, W))
    below = dyb > 0
    r = np.sqrt(((dxb * dxb).astype(f32) + (dyb * dyb).astype(f32)).astype(f32)).astype(f32)
    th = spec_atan2_pos(dxb, np.where(below, dyb, f32(1.0)).astype(f32))
    i_s = ((r - rmin).astype(f32) * inv_dr).astype(f32)
    j_s = ((th + thmax).astype(f32) * inv_dth).astype(f32)
    return i_s, j_s, below


def polar_to_cart(image_p, label_p, g: ConvexGeometry, H: int, W: int):
    """polar [n_samples, n_rays] -> Cartesian [H,W]; pixels outside the sector are 0 (image and label)."""
    image_p = np.asarray(image_p, dtype=f32)
    i_s, j_s, below = grid_coords_of_cart(g, H, W)
    img = np.where(below, _bilinear(image_p, i_s, j_s), f32(0)).astype(f32)
    lab = np.where(below, _nearest(np.asarray(label_p), i_s, j_s), 0).astype(np.asarray(label_p).dtype)
    return img, lab


def default_geometry(H: int, W: int, n_samples: int | None = None, n_rays: int | None = None,
                     theta_max: float = 0.6, apex_above: float = 0.35) -> ConvexGeometry:
    """A fan that covers most of an H x W display image: apex apex_above * H pixels above the top centre."""
    ay = -apex_above * H
    return ConvexGeometry(apex_x=(W - 1) / 2.0, apex_y=ay, theta_max=theta_max, r_min=-ay,
                          r_max=float(H - 1 - ay), n_samples=n_samples or H, n_rays=n_rays or W)
