"""Development aid (CPU, fp64): iteration counts of candidate two-level preconditioners for the
confidence-map system, before anything is written in CUDA.

    python tests/manual/precond_proto.py [H] [W] [deformed]

M^-1 = T^-1 + sum_k P_k (P_k^T A P_k)^-1 P_k^T   (additive), T = column-line tridiagonal part of A.
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import confmap as ocm, deform as odeform, synth as osynth


def make_system(H, W, seed=1234, deformed=False):
    img, lab = osynth.make_bmode(H, W, seed)
    if deformed:
        rng = np.random.default_rng(seed)
        gains = rng.uniform(0.6, 1.4, 8).astype(np.float32)
        img, lab, _ = odeform.fused_tgc_deform(img, lab, np.float32(0.1 * H), gains, odeform.gaussian_taps(0.02 * W))
    planes = ocm.edge_weights(img)
    A, b = ocm.system(*planes)
    return A.tocsr(), b, planes


def line_solver(A, W):
    diag = A.diagonal()
    off = A.diagonal(W)
    T = sp.diags([off, diag, off], [-W, 0, W], format="csc")
    lu = spla.splu(T, permc_spec="NATURAL")
    return lu.solve


def agg_prolongation(Hu, W, by, bx, y_off=0, x_off=0):
    """piecewise-constant prolongation from by x bx aggregates (offsets shift the aggregate grid)"""
    y = (np.arange(Hu) + y_off) // by
    x = (np.arange(W) + x_off) // bx
    nby, nbx = y.max() + 1, x.max() + 1
    idx = (y[:, None] * nbx + x[None, :]).ravel()
    n = Hu * W
    return sp.csr_matrix((np.ones(n), (np.arange(n), idx)), shape=(n, nby * nbx)), (nby, nbx)


def smooth_prolongation(A, P, omega=0.66, line=None, W=None):
    """one damped-Jacobi smoothing step on the tentative prolongation (smoothed aggregation)"""
    d = A.diagonal()
    return (P - omega * sp.diags(1.0 / d) @ (A @ P)).tocsr()


def coarse_solver(A, P):
    Ac = (P.T @ A @ P).tocsc()
    lu = spla.splu(Ac)
    return lambda r: P @ lu.solve(P.T @ r), Ac


def pcg(A, b, M, rtol=1e-8, maxiter=5000, x0=None):
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = b - A @ x
    z = M(r)
    p = z.copy()
    rz = r @ z
    bn = np.linalg.norm(b)
    for it in range(1, maxiter + 1):
        q = A @ p
        a = rz / (p @ q)
        x += a * p
        r -= a * q
        if np.linalg.norm(r) <= rtol * bn:
            return x, it
        z = M(r)
        rzn = r @ z
        p = z + (rzn / rz) * p
        rz = rzn
    return x, maxiter


def ramp_guess(H, W):
    y = np.arange(1, H - 1, dtype=np.float64)
    return np.repeat(1.0 - y / (H - 1), W)


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    W = int(sys.argv[2]) if len(sys.argv) > 2 else H
    deformed = len(sys.argv) > 3 and sys.argv[3] == "deformed"
    A, b, planes = make_system(H, W, deformed=deformed)
    Hu = H - 2
    x0 = ramp_guess(H, W)
    Tsolve = line_solver(A, W)
    t0 = time.time()
    xref = spla.splu(A.tocsc()).solve(b)
    print(f"{H}x{W} deformed={deformed} direct solve {time.time() - t0:.1f}s", flush=True)

    def run(name, M):
        t0 = time.time()
        x, it = pcg(A, b, M, x0=x0)
        print(f"{name:48s} iters {it:5d}  err {np.abs(x - xref).max():.2e}  ({time.time() - t0:.1f}s)", flush=True)
        return it

    cases = sys.argv[4].split(",") if len(sys.argv) > 4 else ["line", "4x32", "2x32", "1x32", "1x64", "2x16", "1x16"]
    for c in cases:
        if c == "line":
            run("line only", Tsolve)
            continue
        parts = c.split("+")
        solvers = []
        for p_ in parts:
            smooth = p_.endswith("s")
            p_ = p_.rstrip("s")
            by, bx = (int(v) for v in p_.split("x"))
            P, (nby, nbx) = agg_prolongation(Hu, W, by, bx)
            if smooth:
                P = smooth_prolongation(A, P)
            cs, Ac = coarse_solver(A, P)
            solvers.append(cs)
        run(f"line + {c} (nc={P.shape[1]})", lambda r: Tsolve(r) + sum(s(r) for s in solvers))


if __name__ == "__main__":
    main()
