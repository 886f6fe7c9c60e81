"""Development aid (CPU): the block-tridiagonal coarse solve of csrc/confmap_coarse.cuh in NumPy.

Ac (9-point operator on the nby x nbx aggregate grid) is block tridiagonal with nbx x nbx blocks:
    Delta_0 = D_0,  Delta_I = D_I - E_I W_{I-1} E_I^T,  W_I = Delta_I^-1          (fp64 setup, W rounded to fp32)
    forward   v_I = r_I - E_I (W_{I-1} v_{I-1});   backward  z_I = W_I (v_I - E_{I+1}^T z_{I+1})
Checks the fp32 solve against a direct solve and the PCG iteration count with the rounded operator.
"""
from __future__ import annotations

import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np
import scipy.sparse.linalg as spla

from tests.manual.precond_proto import agg_prolongation, line_solver, make_system, pcg, ramp_guess


def block_factor(Ac, nby, nbx):
    A = Ac.toarray() if nby * nbx <= 20000 else None
    Wl, El = [], []
    Wprev = None
    for I in range(nby):
        s = slice(I * nbx, (I + 1) * nbx)
        D = Ac[s, s].toarray()
        if I == 0:
            E = np.zeros((nbx, nbx))
            M = D
        else:
            sp_ = slice((I - 1) * nbx, I * nbx)
            E = Ac[s, sp_].toarray()
            M = D - E @ Wprev @ E.T
        Wprev = np.linalg.inv(M)
        Wprev = 0.5 * (Wprev + Wprev.T)
        Wl.append(Wprev.astype(np.float32))
        El.append(E.astype(np.float32))
    return Wl, El


def block_solve(Wl, El, r, nby, nbx):
    f32 = np.float32
    r = r.astype(f32).reshape(nby, nbx)
    v = np.zeros((nby, nbx), f32)
    z = np.zeros((nby, nbx), f32)
    w = np.zeros(nbx, f32)
    for I in range(nby):
        v[I] = r[I] - El[I] @ w
        w = Wl[I] @ v[I]
    zn = np.zeros(nbx, f32)
    for I in range(nby - 1, -1, -1):
        u = v[I] - (El[I + 1].T @ zn if I + 1 < nby else 0)
        zn = Wl[I] @ u.astype(f32)
        z[I] = zn
    return z.ravel().astype(np.float64)


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    by, bx = (int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "2x32").split("x"))
    A, b, _ = make_system(H, H)
    P, (nby, nbx) = agg_prolongation(H - 2, H, by, bx)
    Ac = (P.T @ A @ P).tocsr()
    lu = spla.splu(Ac.tocsc())
    Wl, El = block_factor(Ac, nby, nbx)
    mineig = min(np.linalg.eigvalsh(w.astype(np.float64)).min() for w in Wl)
    cond = max(np.linalg.cond(w.astype(np.float64)) for w in Wl)
    print(f"{H}^2 {by}x{bx}: nby={nby} nbx={nbx}  min eig of rounded W {mineig:.3e}  max cond {cond:.3e}")
    rng = np.random.default_rng(0)
    r = P.T @ (A @ rng.standard_normal(A.shape[0]))
    ze = lu.solve(r)
    zb = block_solve(Wl, El, r, nby, nbx)
    print("fp32 block solve vs direct: rel err", np.abs(zb - ze).max() / np.abs(ze).max())
    T = line_solver(A, H)
    x0 = ramp_guess(H, H)
    _, it_exact = pcg(A, b, lambda q: T(q) + P @ lu.solve(P.T @ q), x0=x0)
    _, it_block = pcg(A, b, lambda q: T(q) + P @ block_solve(Wl, El, P.T @ q, nby, nbx), x0=x0)
    print("PCG iterations: exact coarse", it_exact, " fp32 block coarse", it_block)


if __name__ == "__main__":
    main()
