"""Curvilinear (convex-probe) geometry: CUDA scan conversion vs the oracle.

Labels bit-exact in BOTH directions (nearest sampling; every coordinate is a fixed sequence of individually rounded
fp32 operations, incl. the spec'd arctangent polynomial); image within 1e-5 relative (4-tap bilinear).
"""
import numpy as np
import pytest
import torch

from oracle import chain as ochain
from oracle import deform as odeform
from oracle import scanconv as osc
from tests import util

pytestmark = pytest.mark.gpu
RTOL = 1e-8


def pkg_geom(g):
    import ultrasoundaugmentation_b200 as us
    return us.ConvexGeometry(g.apex_x, g.apex_y, g.theta_max, g.r_min, g.r_max, g.n_samples, g.n_rays)


GEOMS = [
    (2, 128, 160, dict()),                                               # polar grid = image size
    (2, 97, 131, dict(n_samples=120, n_rays=77)),                        # odd sizes, different polar grid
    (1, 256, 256, dict(theta_max=0.35, n_samples=300, n_rays=190)),      # narrow fan, oversampled radially
    (1, 64, 200, dict(theta_max=1.2, apex_above=0.1)),                   # wide fan: most rays leave the image
]


@pytest.mark.parametrize("B,H,W,kw", GEOMS)
def test_scan_conversion_parity_both_directions(cuda_device, B, H, W, kw):
    from ultrasoundaugmentation_b200 import ops
    img, lab = util.synth_batch(B, H, W)
    g = osc.default_geometry(H, W, **kw)
    gp = pkg_geom(g)
    assert all(np.array_equal(a, b) for a, b in zip(g.tables(), gp.tables()))
    assert all(a == b for a, b in zip(g.scalars(), gp.scalars()))
    pi, pl = ops.cart_to_polar(util.cuda(img, cuda_device), util.cuda(lab, cuda_device), gp)
    assert tuple(pi.shape) == (B, g.n_samples, g.n_rays)
    for b in range(B):
        ri, rl = osc.cart_to_polar(img[b], lab[b], g)
        assert np.array_equal(pl[b].cpu().numpy(), rl)
        assert util.rel_err(pi[b].cpu().numpy(), ri) <= util.REL_TOL
    # back, from the ORACLE's polar image so that the second direction is tested on identical inputs
    rp = np.stack([osc.cart_to_polar(img[b], lab[b], g)[0] for b in range(B)])
    rpl = np.stack([osc.cart_to_polar(img[b], lab[b], g)[1] for b in range(B)])
    ci, cl = ops.polar_to_cart(util.cuda(rp, cuda_device), util.cuda(rpl, cuda_device), gp, H, W)
    for b in range(B):
        ri, rl = osc.polar_to_cart(rp[b], rpl[b], g, H, W)
        assert np.array_equal(cl[b].cpu().numpy(), rl)
        assert util.rel_err(ci[b].cpu().numpy(), ri) <= util.REL_TOL


def test_scan_conversion_label_codes_and_determinism(cuda_device):
    """Multi-valued labels travel unchanged (nearest picks, never blends); repeated runs are bitwise identical."""
    from ultrasoundaugmentation_b200 import ops
    H, W = 120, 140
    rng = np.random.default_rng(5)
    lab = (rng.integers(0, 4, (2, H // 8, W // 10)).repeat(8, 1).repeat(10, 2)[:, :H, :W] * 60).astype(np.uint8)
    img = rng.random((2, H, W), dtype=np.float32)
    g = osc.default_geometry(H, W)
    gp = pkg_geom(g)
    a = ops.cart_to_polar(util.cuda(img, cuda_device), util.cuda(lab, cuda_device), gp)
    b = ops.cart_to_polar(util.cuda(img, cuda_device), util.cuda(lab, cuda_device), gp)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert set(np.unique(a[1].cpu().numpy())) <= {0, 60, 120, 180}
    assert np.array_equal(a[1][0].cpu().numpy(), osc.cart_to_polar(img[0], lab[0], g)[1])


def test_chain_on_convex_geometry(cuda_device):
    """Full chain on the polar grid: Cartesian -> polar -> A.1-A.5 along the rays -> Cartesian, against the oracle."""
    import ultrasoundaugmentation_b200 as us
    B, H, W = 2, 128, 128
    img, lab = util.synth_batch(B, H, W)
    g = osc.default_geometry(H, W, theta_max=0.4, n_samples=112, n_rays=96)
    chain = us.PhysicsChain.full(seed=5, rtol=RTOL, geometry=pkg_geom(g))
    params = chain.get_params(B, image_size=chain._scanline_grid(H, W))
    gi, gl = chain(util.cuda(img, cuda_device), util.cuda(lab, cuda_device), params)
    torch.cuda.synchronize()
    assert tuple(gi.shape) == (B, H, W)
    p = {k: v.numpy() for k, v in params.items()}
    wt, ft = odeform.gaussian_taps(0.02 * g.n_rays), odeform.gaussian_taps(2.0)
    for b in range(B):
        pi, pl = osc.cart_to_polar(img[b], lab[b], g)
        oi, ol = ochain.full_chain(pi, pl, p["d"][b], p["tgc"][b], p["a_snr"][b], p["rho"][b], p["n_ghost"][b], wt, ft)
        ri, rl = osc.polar_to_cart(oi, ol, g, H, W)
        assert np.array_equal(gl[b].cpu().numpy(), rl)
        err = np.abs(gi[b].cpu().numpy().astype(np.float64) - ri)
        tol = 0.4 * util.tol_cm(RTOL)
        assert (err <= tol + 2 * util.REL_TOL * np.maximum(np.abs(ri), util.REL_DEN_FLOOR)).all(), err.max()


def test_geometry_rejects_u8_and_bad_parameters(cuda_device):
    import ultrasoundaugmentation_b200 as us
    with pytest.raises(ValueError):
        us.ConvexGeometry(0, -10, 2.0, 10, 100, 64, 64)                  # theta_max >= pi/2
    with pytest.raises(ValueError):
        us.ConvexGeometry(0, -10, 0.5, 100, 10, 64, 64)                  # r_max <= r_min
    g = us.ConvexGeometry.default(64, 64)
    with pytest.raises(ValueError):
        us.PhysicsChain([us.TimeGainCompensation()], output_dtype=torch.uint8, geometry=g)
    chain = us.PhysicsChain([us.TimeGainCompensation()], geometry=g)
    with pytest.raises(TypeError):
        chain(torch.zeros(1, 64, 64, dtype=torch.uint8, device=cuda_device), torch.zeros(1, 64, 64, dtype=torch.uint8, device=cuda_device))
