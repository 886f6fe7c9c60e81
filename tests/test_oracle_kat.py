"""Known-answer tests that pin the SELF-WRITTEN oracle: every expected value here is
closed-form or comes from an independent library, not from the oracle itself."""
import numpy as np
import pytest
from scipy.ndimage import map_coordinates

from oracle import confmap as ocm, deform as odeform, snr_reverb as osr, synth as osynth
from tests import util

f32 = np.float32


def test_identity_warp_is_bit_exact():
    img, lab = osynth.make_bmode(64, 48)
    o, l, info = odeform.fused_tgc_deform(img, lab, 0.0, np.ones(8, f32), util.warp_taps(48))
    assert np.array_equal(o, img) and np.array_equal(l, lab)


def test_tgc_unit_gains_identity_and_constant_gain():
    img, lab = osynth.make_bmode(40, 32)
    o, _, _ = odeform.fused_tgc_deform(img, lab, 0.0, np.full(4, 0.5, f32), util.warp_taps(32))
    assert np.allclose(o, 0.5 * img, rtol=1e-6, atol=0)
    g = odeform.tgc_gain(33, np.array([1.0, 2.0, 4.0], f32))
    assert g[0] == 1.0 and g[16] == 2.0 and g[32] == 4.0 and np.isclose(g[8], 1.5) and np.isclose(g[24], 3.0)


def test_bilinear_matches_scipy_map_coordinates():
    img, lab = osynth.make_bmode(96, 64)
    o, _, info = odeform.fused_tgc_deform(img, lab, 7.3, np.ones(2, f32), util.warp_taps(64))
    ys = info["y_src"].astype(np.float64)
    xs = np.broadcast_to(np.arange(64)[None, :], ys.shape).astype(np.float64)
    # grid-constant = zero padding that takes part in the interpolation (A.2: "samples outside read as 0")
    ref = map_coordinates(img.astype(np.float64), [ys, xs], order=1, mode="grid-constant", cval=0.0)
    assert np.abs(ref - o).max() <= 2e-7


def test_horizontal_line_phantom_moves_as_the_forward_model_says():
    """A bright line at depth Y above the bone lands at Y - d*min(Y/s_hat, 1) (+-1 px)."""
    H, W, Y, d = 128, 64, 30, 8.0
    img = np.zeros((H, W), f32); img[Y] = 1.0
    lab = np.zeros((H, W), np.uint8); lab[80:84] = 1
    o, l, info = odeform.fused_tgc_deform(img, lab, d, np.ones(2, f32), util.warp_taps(W))
    assert info["d_eff"] == f32(d) and np.allclose(info["s_hat"], 80.0, atol=1e-3)
    expect = Y - d * min(Y / 80.0, 1.0)
    assert abs(np.argmax(o[:, W // 2]) - expect) <= 1.0
    assert np.flatnonzero(l[:, 3])[0] == 80 - int(d)            # bone moves rigidly with the full d


def test_bone_surface_hole_fill_and_no_bone():
    lab = np.zeros((10, 9), np.uint8)
    lab[4, 2] = 1; lab[7, 6] = 1
    s, y_top, has = odeform.bone_surface(lab)
    assert has and y_top == 4
    assert s.tolist() == [4, 4, 4, 4, 4, 7, 7, 7, 7]           # x=4 is equidistant: tie -> lower x (col 2)
    s, y_top, has = odeform.bone_surface(np.zeros((10, 9), np.uint8))
    assert (not has) and y_top == -1 and (s == 9).all()


def test_displacement_clamp():
    s_hat = np.full(16, 20.0, f32)
    assert odeform.clamp_displacement(100.0, s_hat) == f32(10.0)
    assert odeform.clamp_displacement(-100.0, s_hat) == f32(-10.0)
    assert odeform.clamp_displacement(3.0, s_hat) == f32(3.0)


@pytest.mark.parametrize("H,W", [(33, 17), (64, 64), (9, 3)])
def test_constant_image_confidence_map_is_linear_ramp(H, W):
    C = ocm.confidence_map(np.full((H, W), 0.37, f32))
    assert np.abs(C - (1.0 - np.arange(H) / (H - 1.0))[:, None]).max() < 1e-12


def test_constant_image_has_three_distinct_weights():
    wS, wE, wSE, wSW = ocm.edge_weights(np.full((8, 8), 0.5, f32), beta=90.0, gamma=0.05)
    assert np.allclose(np.unique(wS[:-1]), [1.000001], rtol=1e-6)
    assert np.allclose(np.unique(wE[:, :-1]), [np.exp(-4.5) + 1e-6], rtol=1e-5)
    assert np.allclose(np.unique(wSE[:-1, :-1]), [np.exp(-90 * np.sqrt(2) * 0.05) + 1e-6], rtol=1e-5)
    assert np.array_equal(wSE[:-1, :-1], wSW[:-1, 1:])


def test_confidence_map_boundaries_range_residual():
    img, _ = osynth.make_bmode(64, 48)
    C, info = ocm.confidence_map(img, return_info=True)
    assert (C[0] == 1).all() and (C[-1] == 0).all()
    assert C.min() >= -1e-12 and C.max() <= 1 + 1e-12
    assert info["relres"] < 1e-12
    A = info["A"]
    assert abs(A - A.T).max() < 1e-15 and (A.diagonal() > 0).all()


def test_pcg_reference_agrees_with_direct_solve():
    img, _ = osynth.make_bmode(48, 40)
    C, info = ocm.confidence_map(img, return_info=True)
    for pc in ("jacobi", "line"):
        x, it, _ = ocm.pcg_reference(info["A"], info["b"], pc, W=40, rtol=1e-10)
        assert np.abs(x.reshape(46, 40) - C[1:-1]).max() < 1e-5


def test_snr_and_reverb_identities():
    img, lab = osynth.make_bmode(64, 48)
    cm = np.random.default_rng(0).uniform(0, 1, img.shape).astype(f32)
    taps = odeform.gaussian_taps(2.0)
    assert np.array_equal(osr.snr_apply(img, cm, 1.0), img)
    assert np.array_equal(osr.reverb(img, np.zeros_like(lab), 0.5, 2, taps), img)     # empty mask
    assert np.array_equal(osr.reverb(img, lab, 0.0, 2, taps), np.clip(img, 0, 1))      # rho = 0
    assert np.array_equal(osr.reverb(img, lab, 0.5, 0, taps), img)                     # no ghosts
    out = osr.reverb(img, lab, 0.5, 1, taps)
    _, y_top, _ = odeform.bone_surface(lab)
    assert np.array_equal(out[:y_top], img[:y_top])          # nothing above the first ghost row
    assert (out >= img - 1e-7).all() and out.max() <= 1.0


def test_feather_mask_mass_and_support():
    lab = np.zeros((40, 40), np.uint8); lab[20, 20] = 1
    taps = odeform.gaussian_taps(2.0)
    m = osr.feather_mask(lab, taps)
    assert np.isclose(m.sum(), 1.0, atol=1e-5) and np.isclose(m[20, 20], taps[6] ** 2, rtol=1e-5)
    assert m[:13].sum() == 0 and m[:, 27:].sum() == 0


def test_synth_is_seeded_and_in_range():
    a, la = osynth.make_bmode(64, 64, 5); b, lb = osynth.make_bmode(64, 64, 5); c, _ = osynth.make_bmode(64, 64, 6)
    assert np.array_equal(a, b) and np.array_equal(la, lb) and not np.array_equal(a, c)
    assert a.min() >= 0 and a.max() == 1.0 and la.sum() > 0 and set(np.unique(la)) == {0, 1}


def test_u8_ingest_and_egress_round_trip():
    """8-bit pipeline KATs: v/255 is exact at the ends, egress inverts ingest for all 256 codes,
    and the quantiser rounds half up."""
    from oracle import snr_reverb as osr
    codes = np.arange(256, dtype=np.uint8)
    f = odeform.ingest_u8(codes)
    assert f.dtype == np.float32 and f[0] == 0.0 and f[255] == 1.0
    assert np.all(np.diff(f) > 0)
    assert np.array_equal(f, (codes.astype(np.float64) / 255.0).astype(np.float32))   # correctly rounded
    assert np.array_equal(osr.quantize_u8(f), codes)
    assert np.array_equal(osr.quantize_u8(np.array([0.0, 0.5 / 255, 0.49 / 255, 1.0], np.float32)), [0, 1, 0, 255])


def test_local_energy_known_answers():
    """Monogenic local energy, pins independent of any implementation:
    constant image -> 0; plane wave -> constant energy G(f0) (quadrature); flips commute; scaling is linear."""
    from oracle import local_energy as ole
    H, W = 64, 128
    assert np.all(ole.local_energy_raw(np.full((H, W), 0.37)) < 1e-12)
    assert np.all(ole.local_energy(np.full((H, W), 0.37)) == 0.0)
    for k0, axis in [(8, 1), (5, 0), (16, 1)]:
        n = W if axis == 1 else H
        t = np.cos(2 * np.pi * k0 * np.arange(n) / n + 0.3)
        img = np.tile(t[None, :], (H, 1)) if axis == 1 else np.tile(t[:, None], (1, W))
        G = ole.log_gabor_sum(H, W)[0]
        want = G[0, k0] if axis == 1 else G[k0, 0]
        E = ole.local_energy_raw(img)
        assert np.allclose(E, want, rtol=0, atol=1e-12), (k0, axis)
        assert np.allclose(ole.local_energy(img), 1.0, atol=1e-10)
    img, _ = osynth.make_bmode(H, W, 3)
    E = ole.local_energy_raw(img)
    assert np.allclose(ole.local_energy_raw(img[:, ::-1]), E[:, ::-1], atol=1e-12)
    assert np.allclose(ole.local_energy_raw(img[::-1]), E[::-1], atol=1e-12)
    assert np.allclose(ole.local_energy_raw(2.5 * img), 2.5 * E, atol=1e-12)
    Eh = ole.local_energy(img)
    assert Eh.min() >= 0.0 and Eh.max() == 1.0
    # A.4 general form reduces to the hot-path form for E = I
    cm = np.random.default_rng(0).random((H, W)).astype(np.float32)
    assert np.array_equal(ole.snr_apply_signal(img, cm, img, 1.3), osr.snr_apply(img, cm, 1.3))


def test_scan_conversion_known_answers():
    """Convex-probe scan conversion: the spec'd arctangent is accurate, the two coordinate
    maps invert each other, an affine image survives Cartesian -> polar exactly (bilinear is exact on affine
    functions) and the round trip reproduces a smooth image inside the sector."""
    from oracle import scanconv as osc
    z = np.linspace(0, 1, 20001, dtype=np.float32)
    assert np.abs(osc.spec_atan(z).astype(np.float64) - np.arctan(z.astype(np.float64))).max() <= 2e-7
    dx = np.array([-3.0, -1.0, 0.0, 0.5, 2.0, 7.0], np.float32); dy = np.array([1.0, 1.0, 2.0, 4.0, 2.0, 1.0], np.float32)
    assert np.abs(osc.spec_atan2_pos(dx, dy) - np.arctan2(dx.astype(np.float64), dy)).max() <= 3e-7
    H, W = 96, 128
    g = osc.default_geometry(H, W, theta_max=0.35)      # narrow enough that the fan stays inside the image sideways
    y, x = osc.polar_coords_of_grid(g)
    i_s, j_s, below = osc.grid_coords_of_cart(g, H, W)
    # the two maps invert each other: the (i*, j*) fields, interpolated at the Cartesian position of node (i, j), give (i, j)
    node_ok = (y >= 1) & (y <= H - 2) & (x >= 1) & (x <= W - 2)
    ii, jj = np.meshgrid(np.arange(g.n_samples), np.arange(g.n_rays), indexing="ij")
    assert node_ok.mean() > 0.3
    assert np.abs(osc._bilinear(i_s, y, x)[node_ok] - ii[node_ok]).max() < 2e-2
    assert np.abs(osc._bilinear(j_s, y, x)[node_ok] - jj[node_ok]).max() < 2e-2
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    affine = (0.1 + 0.002 * xx + 0.004 * yy).astype(np.float32)
    lab = ((yy - 50) ** 2 + (xx - 64) ** 2 < 15 ** 2).astype(np.uint8)
    pi, pl = osc.cart_to_polar(affine, lab, g)
    inside = (y >= 0) & (y <= H - 1) & (x >= 0) & (x <= W - 1)
    want = 0.1 + 0.002 * x.astype(np.float64) + 0.004 * y.astype(np.float64)
    assert np.abs(pi[inside] - want[inside]).max() <= 2e-6
    assert pl.dtype == np.uint8 and 0 < pl.sum() < pl.size
    smooth = (0.5 + 0.4 * np.sin(xx / 17.0) * np.cos(yy / 23.0)).astype(np.float32)
    ps, _ = osc.cart_to_polar(smooth, lab, g)
    back, back_l = osc.polar_to_cart(ps, pl, g, H, W)
    core = (i_s > 5) & (i_s < g.n_samples - 2) & (j_s > 1) & (j_s < g.n_rays - 2) & below
    assert core.mean() > 0.3
    assert np.abs(back[core] - smooth[core]).max() <= 5e-3
    assert (back_l[core] != lab[core]).mean() <= 0.02          # only pixels on the disc's boundary may flip
    assert np.all(back[~below] == 0)


# ---- sink modes of the confidence map --------------------------------------------------------
def _dense_confidence_map(image, sinks):
    """Independent restatement for tiny images: dense 8-neighbour Laplacian built edge by edge, Dirichlet rows removed."""
    H, W = image.shape
    wS, wE, wSE, wSW = ocm.edge_weights(image)
    L = np.zeros((H * W, H * W))
    def edge(i, j, w):
        L[i, i] += w; L[j, j] += w; L[i, j] -= w; L[j, i] -= w
    for y in range(H):
        for x in range(W):
            i = y * W + x
            if x + 1 < W: edge(i, i + 1, wE[y, x])
            if y + 1 < H:
                edge(i, i + W, wS[y, x])
                if x + 1 < W: edge(i, i + W + 1, wSE[y, x])
                if x > 0: edge(i, i + W - 1, wSW[y, x])
    fixed = np.zeros(H * W, bool); val = np.zeros(H * W)
    fixed[:W] = True; val[:W] = 1.0
    fixed[-W:] = True
    fixed |= sinks.ravel()
    U = ~fixed
    x = np.zeros(H * W); x[fixed] = val[fixed]
    x[U] = np.linalg.solve(L[np.ix_(U, U)], -L[np.ix_(U, fixed)] @ val[fixed])
    return x.reshape(H, W)


@pytest.mark.parametrize("mode", ["bottom_row", "bottom_and_sides", "mask"])
def test_sink_modes_match_a_dense_dirichlet_solve(mode):
    img, lab = osynth.make_bmode(14, 11)
    mask = np.zeros((14, 11), np.uint8); mask[5:8, 3:6] = 1; mask[0, 2] = 1; mask[13, 4] = 1; mask[9, 10] = 1
    sinks = ocm.sink_mask(14, 11, mode, mask if mode == "mask" else None)
    C = ocm.confidence_map(img, sink=mode, mask=mask if mode == "mask" else None)
    assert np.abs(C - _dense_confidence_map(img, sinks)).max() < 1e-9        # two direct solves of a system with cond ~ 1e6
    assert (C[sinks] == 0).all() and (C[0] == 1).all() and (C[-1] == 0).all()
    # maximum principle: further sinks can only lower the potential
    assert (C <= ocm.confidence_map(img) + 1e-13).all()


def test_sink_row_gives_a_ramp_above_and_zero_below():
    """Closed form: constant image, a full interior row r of sinks => C = 1 - y/r above it and 0 below."""
    H, W, r = 21, 9, 12
    mask = np.zeros((H, W), np.uint8); mask[r] = 1
    C = ocm.confidence_map(np.full((H, W), 0.4, f32), sink="mask", mask=mask)
    ref = np.clip(1.0 - np.arange(H) / r, 0.0, None)[:, None]
    assert np.abs(C - ref).max() < 1e-12


def test_sink_mask_without_sinks_is_the_plain_map():
    img, _ = osynth.make_bmode(24, 20)
    assert np.array_equal(ocm.confidence_map(img, sink="mask", mask=np.zeros((24, 20), np.uint8)), ocm.confidence_map(img))
    with pytest.raises(ValueError):
        ocm.confidence_map(img, sink="mask")
    with pytest.raises(ValueError):
        ocm.confidence_map(img, sink="everywhere")
