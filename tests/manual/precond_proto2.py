"""Development aid (CPU, fp64): weighted additive combinations of the shipped preconditioner's three terms, and further
line levels, before anything is written in CUDA.

    python tests/manual/precond_proto2.py [H] [W] [deformed]

M^-1 = T^-1 + wc * P Ac^-1 P^T + sum_k w_k * P_k T_k^-1 P_k^T
T = column-line part of A; Ac = Galerkin operator of by x 32 aggregates; T_k = column-line part of the Galerkin operator of
1 x k aggregates (k = 4 is the shipped "level" term).
"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from tests.manual.precond_proto import make_system, line_solver, agg_prolongation, coarse_solver, pcg, ramp_guess

def level_line(A, Hu, W, k):
    P, (nby, nbx) = agg_prolongation(Hu, W, 1, k)
    Ak = (P.T @ A @ P).tocsr()
    Ts = line_solver(Ak, nbx)
    return lambda r: P @ Ts(P.T @ r)

def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    W = int(sys.argv[2]) if len(sys.argv) > 2 else H
    deformed = len(sys.argv) > 3 and sys.argv[3] == "deformed"
    A, b, _ = make_system(H, W, deformed=deformed)
    Hu = H - 2
    x0 = ramp_guess(H, W)
    T = line_solver(A, W)
    lv = {k: level_line(A, Hu, W, k) for k in (2, 4, 8, 16)}
    cs = {}
    for by in (2, 4):
        P, _ = agg_prolongation(Hu, W, by, 32)
        cs[by], _ = coarse_solver(A, P)
    def run(name, M):
        t0 = time.time(); x, it = pcg(A, b, M, x0=x0)
        print(f"{name:60s} iters {it:5d} ({time.time()-t0:.1f}s)", flush=True)
    for by in (4, 2):
        run(f"shipped: line + {by}x32 + level4", lambda r: T(r) + cs[by](r) + lv[4](r))
        for wc, w4 in ((0.7, 1.0), (1.5, 1.0), (2.0, 1.0), (1.0, 0.7), (1.0, 1.5), (1.0, 2.0), (2.0, 2.0), (1.5, 1.5)):
            run(f"  wc={wc} w4={w4}", lambda r: T(r) + wc * cs[by](r) + w4 * lv[4](r))
        run(f"  + level16", lambda r: T(r) + cs[by](r) + lv[4](r) + lv[16](r))
        run(f"  + level2 + level8", lambda r: T(r) + cs[by](r) + lv[4](r) + lv[2](r) + lv[8](r))
        run(f"  + level2 + level8 + level16", lambda r: T(r) + cs[by](r) + lv[4](r) + lv[2](r) + lv[8](r) + lv[16](r))
        run(f"  level2 instead of level4", lambda r: T(r) + cs[by](r) + lv[2](r))
        run(f"  level8 instead of level4", lambda r: T(r) + cs[by](r) + lv[8](r))

if __name__ == "__main__":
    main()
