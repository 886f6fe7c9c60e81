"""Shared helpers of the parity tests (oracle = checker, CUDA path = thing under test)."""
import numpy as np
import torch

from oracle import deform as odeform
from oracle import synth as osynth

REL_TOL = 1e-5          # north_star: image transforms within 1e-5 relative error
REL_DEN_FLOOR = 1e-3    # denominator max(|ref|, 1e-3)


def rel_err(got: np.ndarray, ref: np.ndarray) -> float:
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), REL_DEN_FLOOR)))


# Measured amplification K = max|C_gpu - C_oracle| / (true relative residual) of the shipped solver on B200
# (profiles/r2g_error_estimate_study.log: 7 shapes x {speckle, white noise, smooth} x rtol 1e-6..1e-9 x 2 pinned solver
# variants): speckle <= 2.2e4, white noise <= 6.0e4, smooth <= 7.7e3.  The bound below is that measurement with margin.
# It is NOT derived: DESIGN.md section 4 explains why neither a rigorous bound nor a solver-side estimate is usable
# (the error sits in near-null modes the residual barely sees).
K_CM = {"speckle": 5e4, "noise": 1e5}


def tol_cm(rtol: float, image_class: str = "speckle") -> float:
    """max|dC| <= K * rtol + 1e-6, K = 5e4 for speckle / chain images, 1e5 for white noise."""
    return K_CM[image_class] * rtol + 1e-6


def cm_rigorous_bound(image, cm_gpu):
    """||A^-1||_inf * ||b - A x||_inf for one image and the solver's result x = the fp32 map itself (small images only).
    A = the oracle's L_UU is an M-matrix: A^-1 >= 0 entrywise, so ||A^-1||_inf = max(A^-1 1).  Rigorous, but 300-6500x above
    the typical error, so only the fuzz test uses it, as the fallback of the measured tolerance."""
    import scipy.sparse.linalg as spla
    from oracle import confmap as ocm
    A, b = ocm.system(*ocm.edge_weights(np.asarray(image, np.float32)))
    lu = spla.splu(A.tocsc())
    kinf = float(lu.solve(np.ones_like(b)).max())
    x = np.asarray(cm_gpu, np.float64)[1:-1].ravel()
    rinf = float(np.abs(b - A @ x).max())
    return 1.01 * kinf * rinf + 1e-9       # x* - x = A^-1 (b - A x) exactly; 1 % for the fp64 arithmetic of the check itself


def synth_batch(B, H, W, seed0=1234):
    return osynth.make_batch(B, H, W, seed0)


def sample_params(B, H, K=8, seed=7):
    rng = np.random.default_rng(seed)
    return dict(
        d=rng.uniform(-0.05 * H, 0.15 * H, B).astype(np.float32),
        tgc=rng.uniform(0.6, 1.4, (B, K)).astype(np.float32),
        a_snr=rng.uniform(0.7, 1.4, B).astype(np.float32),
        rho=rng.uniform(0.3, 0.7, B).astype(np.float32),
        n_ghost=rng.integers(1, 3, B).astype(np.int32),
    )


def cuda(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def warp_taps(W, sigma_rel=0.02):
    return odeform.gaussian_taps(sigma_rel * W)
