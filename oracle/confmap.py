"""Oracle for the random-walk ultrasound confidence map.

Test infrastructure, self-written from the published Karamalis-2012 B-mode definition as
restated in SURVEY.md A.3 (reference mount = README.md:1-2 only; parity unpinned).
Weights are formed in fp32 in the order A.3 gives; the linear system is solved directly in
fp64 with SuperLU (scipy.sparse.linalg.splu, scipy 1.18.1) -- the ground truth the GPU's
matrix-free PCG is compared with.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

f32 = np.float32
EPS32 = f32(2.220446e-16)
WFLOOR = f32(1e-6)
SQRT2 = f32(1.41421356237309515)


def depth_profile(H: int, alpha: float) -> np.ndarray:
    """A.3 step 2 depth factor 1 - exp(-alpha*y/(H-1)), evaluated in fp64, rounded to fp32."""
    y = np.arange(H, dtype=np.float64)
    return (1.0 - np.exp(-float(alpha) * y / float(max(H - 1, 1)))).astype(f32)


def attenuated(image: np.ndarray, alpha: float) -> np.ndarray:
    """A.3 steps 1-2: min/max normalise, weight by depth.  fp32."""
    img = np.asarray(image, dtype=f32)
    mn = f32(img.min())
    mx = f32(img.max())
    den = f32(f32(mx - mn) + EPS32)
    norm = ((img - mn).astype(f32) / den).astype(f32)
    return (norm * depth_profile(img.shape[0], alpha)[:, None]).astype(f32)


def edge_weights(image: np.ndarray, alpha=2.0, beta=90.0, gamma=0.05):
    """A.3 steps 3-5.  Returns four [H,W] fp32 planes (zero where the edge does not exist):

    wS[y,x]  edge (y,x)-(y+1,x)      penalty 0
    wE[y,x]  edge (y,x)-(y,x+1)      penalty gamma
    wSE[y,x] edge (y,x)-(y+1,x+1)    penalty sqrt(2)*gamma
    wSW[y,x] edge (y,x)-(y+1,x-1)    penalty sqrt(2)*gamma
    """
    a = attenuated(image, alpha)
    H, W = a.shape
    dS = np.abs(a[:-1, :] - a[1:, :]).astype(f32)
    dE = np.abs(a[:, :-1] - a[:, 1:]).astype(f32)
    dSE = np.abs(a[:-1, :-1] - a[1:, 1:]).astype(f32)
    dSW = np.abs(a[:-1, 1:] - a[1:, :-1]).astype(f32)
    parts = [p for p in (dS, dE, dSE, dSW) if p.size]
    mn = f32(min(p.min() for p in parts))
    mx = f32(max(p.max() for p in parts))
    den = f32(f32(mx - mn) + EPS32)
    b = f32(beta)
    pen_h = f32(gamma)
    pen_d = f32(SQRT2 * f32(gamma))

    def w(delta, pen):
        dn = ((delta - mn).astype(f32) / den).astype(f32)
        arg = (-(b * (dn + pen).astype(f32))).astype(f32)
        return (np.exp(arg).astype(f32) + WFLOOR).astype(f32)

    wS = np.zeros((H, W), f32)
    wE = np.zeros((H, W), f32)
    wSE = np.zeros((H, W), f32)
    wSW = np.zeros((H, W), f32)
    wS[:-1, :] = w(dS, f32(0))
    wE[:, :-1] = w(dE, pen_h)
    wSE[:-1, :-1] = w(dSE, pen_d)
    wSW[:-1, 1:] = w(dSW, pen_d)
    return wS, wE, wSE, wSW


def laplacian(wS, wE, wSE, wSW) -> sp.csr_matrix:
    """Full-grid graph Laplacian L = D - W (fp64) from the four edge planes."""
    H, W = wS.shape
    idx = np.arange(H * W).reshape(H, W)
    rows, cols, vals = [], [], []

    def add(i, j, w):
        i = i.ravel(); j = j.ravel(); w = w.ravel().astype(np.float64)
        rows.extend([i, j]); cols.extend([j, i]); vals.extend([w, w])

    add(idx[:-1, :], idx[1:, :], wS[:-1, :])
    add(idx[:, :-1], idx[:, 1:], wE[:, :-1])
    add(idx[:-1, :-1], idx[1:, 1:], wSE[:-1, :-1])
    add(idx[:-1, 1:], idx[1:, :-1], wSW[:-1, 1:])
    r = np.concatenate(rows); c = np.concatenate(cols); v = np.concatenate(vals)
    Wm = sp.coo_matrix((v, (r, c)), shape=(H * W, H * W)).tocsr()
    D = sp.diags(np.asarray(Wm.sum(axis=1)).ravel())
    return (D - Wm).tocsr()


def sink_mask(H: int, W: int, sink="bottom_row", mask=None) -> np.ndarray:
    """Boolean [H,W] plane of the EXTRA sink pixels (potential 0) among rows 1..H-2.

    The top row is always the source (potential 1) and the bottom row always a sink.
    ``bottom_row``: nothing else.  ``bottom_and_sides``: also the first and the last column.
    ``mask``: also every pixel where ``mask`` is non-zero.
    """
    m = np.zeros((H, W), bool)
    if sink == "bottom_and_sides":
        m[:, 0] = True
        m[:, -1] = True
    elif sink == "mask":
        if mask is None:
            raise ValueError("sink='mask' needs a mask plane")
        m |= np.asarray(mask) != 0
    elif sink != "bottom_row":
        raise ValueError(f"unknown sink mode {sink!r}")
    m[0] = False
    m[-1] = False
    return m


def system(wS, wE, wSE, wSW, sinks=None):
    """A.3 step 6: (L_UU, b) with top row = 1, bottom row = 0, unknown rows 1..H-2.

    ``sinks`` (bool [H,W], optional): further pixels held at potential 0.  They stay in the vector as decoupled
    identity rows (solution exactly 0), so the layout is independent of the mask.
    """
    H, W = wS.shape
    L = laplacian(wS, wE, wSE, wSW)
    lo, hi = W, (H - 1) * W
    A = L[lo:hi, lo:hi].tocsc()
    b = np.asarray(-(L[lo:hi, 0:W] @ np.ones(W))).ravel()
    if sinks is not None and sinks[1:-1].any():
        keep = (~sinks[1:-1].ravel()).astype(np.float64)
        Dk = sp.diags(keep)
        A = (Dk @ A @ Dk + sp.diags(1.0 - keep)).tocsc()      # couplings to a sink only count on the neighbour's diagonal
        b = b * keep
    return A, b


def confidence_map(image, alpha=2.0, beta=90.0, gamma=0.05, return_info=False, sink="bottom_row", mask=None):
    """A.3 complete: confidence map C [H,W] in fp64 (seed rows exactly 1 and 0)."""
    image = np.asarray(image, dtype=f32)
    H, W = image.shape
    if H < 3:
        raise ValueError("confidence map needs H >= 3")
    planes = edge_weights(image, alpha, beta, gamma)
    A, b = system(*planes, sinks=sink_mask(H, W, sink, mask))
    x = spla.splu(A).solve(b)
    C = np.empty((H, W), np.float64)
    C[0] = 1.0
    C[-1] = 0.0
    C[1:-1] = x.reshape(H - 2, W)
    if return_info:
        res = np.linalg.norm(b - A @ x) / max(np.linalg.norm(b), 1e-300)
        return C, dict(relres=res, A=A, b=b, planes=planes)
    return C


def confidence_map_batch(images, alpha=2.0, beta=90.0, gamma=0.05, sink="bottom_row", masks=None):
    return np.stack([confidence_map(im, alpha, beta, gamma, sink=sink, mask=None if masks is None else masks[i])
                     for i, im in enumerate(images)])


def pcg_reference(A, b, precond="line", W=None, rtol=1e-8, maxiter=20000):
    """Plain fp64 PCG used only to cross-check iteration counts of the GPU solver.

    precond: "jacobi" or "line" (column tridiagonal = vertical couplings + full diagonal).
    Returns (x, iters, relres_history).
    """
    n = A.shape[0]
    diag = A.diagonal()
    if precond == "jacobi":
        def M(r):
            return r / diag
    else:
        off = -A.diagonal(W)          # vertical coupling between (y,x) and (y+1,x)
        T = sp.diags([-off, diag, -off], [-W, 0, W], format="csc")
        lu = spla.splu(T)
        def M(r):
            return lu.solve(r)
    x = np.zeros(n)
    r = b.copy()
    z = M(r)
    p = z.copy()
    rz = r @ z
    bn = np.linalg.norm(b)
    hist = []
    for it in range(1, maxiter + 1):
        q = A @ p
        a = rz / (p @ q)
        x += a * p
        r -= a * q
        rr = np.linalg.norm(r) / bn
        hist.append(rr)
        if rr <= rtol:
            return x, it, hist
        z = M(r)
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, maxiter, hist
