// Coarse-grid correction of the confidence-map PCG (included by confmap_pcg.cu).
//
// Two-level ADDITIVE preconditioner   M^-1 = T^-1 + P Ac^-1 P^T
//   T   column-line tridiagonal part of A (the sweeps in confmap_fused.cuh)
//   P   piecewise-constant prolongation from aggregates of `by` rows x `bx` columns of unknowns
//   Ac  = P^T A P, the Laplacian of the aggregate graph (+ the couplings to the seed rows on the diagonal):
//         a 9-point operator on an nby x nbx grid, symmetric positive definite.
// The line preconditioner moves information one column per CG iteration; the coarse space carries the
// slow, blob-like error modes of these heterogeneous systems across the image at once.
//
// Ac is BLOCK TRIDIAGONAL with nbx x nbx blocks (one block row = one row of aggregates):
//     D_I (tridiagonal) on the diagonal,  E_I (tridiagonal) = coupling of aggregate row I to row I-1.
// Setup (once per image, fp64): block LDL^T
//     Delta_0 = D_0,   Delta_I = D_I - E_I W_{I-1} E_I^T,   W_I = Delta_I^-1   (dense, symmetrised, stored in fp32)
// Per iteration (one warp per image, confmap_fused.cuh: coarse_task):
//     forward    v_I = rc_I - E_I (W_{I-1} v_{I-1})
//     backward   zc_I = W_I (v_I - E_{I+1}^T zc_{I+1})              and   rc . zc = sum_I v_I . (W_I v_I)
// i.e. per aggregate ROW one dense nbx x nbx matrix-vector product spread over the 32 lanes plus a three-point
// stencil -- a dependent chain of ~130 cycles per row of aggregates, instead of one shuffle + FMA step (~57 cycles)
// per aggregate in a scalar banded triangular solve (round 1: 129 us for 4080 aggregates; now ~35 us for 8160).
// The operator applied is exactly (I+S)^-T W (I+S)^-1 with S_I = E_I W_{I-1}: symmetric for ANY stored E and
// symmetric W, positive definite as long as the rounded W_I are (their condition numbers are O(100)).
// Nothing is kept in shared memory across steps, so the number of aggregates is not limited by it.
#pragma once

namespace usaug {

struct CoarseGeom {
  int by;        // aggregate height in rows (power of two), 0 = coarse correction disabled
  int by_shift;  // log2(by)
  int bx_shift;  // log2(aggregate width): 4 ... 7 (nbx <= 32)
  int nbx, nby;  // coarse grid
  int nb;        // block size: nbx padded to 8, 16 or 32 = row pitch of rc / zc and of the W blocks
  int nc;        // nby * nb: padded coarse vector length per image
  int rec;       // floats per block-row record: nb*(nb+4) (W_I, rows padded like the shared-memory slot) + 4*nb (E_I: centre, left, right, spare)
  int tw;        // twisted block factorisation: the aggregate row where the two elimination fronts meet; -1 = plain (top-down)
};

// min_shift: log2 of the smallest aggregate height tried.  Thin aggregates converge faster (CPU prototype at
// 512x512: 8x32 -> 512, 4x32 -> 435, 2x32 -> 404 iterations; 8x8 -> 462 at the same count as 2x32).
static inline CoarseGeom coarse_geom(int H, int W, int min_shift, int min_bx_shift = 5) {
  CoarseGeom g{};
  const int hu = H - 2;
  g.bx_shift = min_bx_shift;
  while (((W + (1 << g.bx_shift) - 1) >> g.bx_shift) > 32) ++g.bx_shift;
  g.nbx = (W + (1 << g.bx_shift) - 1) >> g.bx_shift;
  if (g.bx_shift > 7 || hu < 8) return g;                   // by = 0: disabled (W > 4096 or a very short image)
  g.nb = g.nbx <= 8 ? 8 : g.nbx <= 16 ? 16 : 32;
  int sh = min_shift;
  while (((hu + (1 << sh) - 1) >> sh) > 1024) ++sh;         // at most 1024 aggregate rows
  g.by = 1 << sh; g.by_shift = sh;
  g.nby = (hu + g.by - 1) >> sh;
  g.nc = g.nby * g.nb;
  g.rec = g.nb * (g.nb + 4) + 4 * g.nb;
  g.tw = -1;
  return g;
}

#ifdef __CUDACC__

// ---------------------------------------------------------------------------------------------
// Ac in block form: Ab[b][I][k][nb] (fp64), k = 0 diagonal, 1 coupling to (I, J-1), 2 / 3 / 4 coupling to
// (I-1, J) / (I-1, J-1) / (I-1, J+1); zero for the padding columns J >= nbx.  One warp per aggregate,
// lane = pixel column inside the aggregate; butterfly sums => deterministic.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) coarse_assemble_kernel(int H, int W, CoarseGeom g,
                                                              const float* __restrict__ wS, const float* __restrict__ wE,
                                                              const float* __restrict__ wSE, const float* __restrict__ wSW,
                                                              double* __restrict__ Ab, const uint8_t* __restrict__ snk) {
  const int lane = threadIdx.x & 31;
  const int agg = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (agg >= g.nc) return;
  const int I = agg / g.nb, J = agg - I * g.nb;
  double* out = Ab + ((size_t)b * g.nby + I) * 5 * g.nb + J;
  if (J >= g.nbx) {
    if (lane < 5) out[lane * g.nb] = 0.0;
    return;
  }
  const size_t base = (size_t)b * H * W;
  const float* S = wS + base; const float* E = wE + base; const float* SE = wSE + base; const float* SW = wSW + base;
  const uint8_t* K = snk ? snk + base : nullptr;   // sink pixels: out of the aggregate; an edge to one counts like an edge to a seed row
  double sum[9];   // sum[(dI+1)*3 + (dJ+1)]: weight to the neighbouring aggregate; [4] collects the seed couplings
#pragma unroll
  for (int k = 0; k < 9; ++k) sum[k] = 0.0;
  const int bx = 1 << g.bx_shift;
  for (int xo = lane; xo < bx; xo += 32) {
    const int x = J * bx + xo;
    if (x >= W) continue;
    const int y0 = I * g.by + 1, y1 = min(y0 + g.by, H - 1);
    for (int y = y0; y < y1; ++y) {
      if (K && K[(size_t)y * W + x]) continue;
      // the 8 neighbours of (y, x): (dy, dx, weight)
      const float wgt[8] = {
          S[(size_t)(y - 1) * W + x],                                         // N
          S[(size_t)y * W + x],                                               // S
          x > 0 ? E[(size_t)y * W + x - 1] : 0.f,                             // W
          E[(size_t)y * W + x],                                               // E   (0 in the last column)
          x > 0 ? SE[(size_t)(y - 1) * W + x - 1] : 0.f,                      // NW
          x < W - 1 ? SW[(size_t)(y - 1) * W + x + 1] : 0.f,                  // NE
          SW[(size_t)y * W + x],                                              // SW  (0 in the first column)
          SE[(size_t)y * W + x]};                                             // SE  (0 in the last column)
      const int dy[8] = {-1, 1, 0, 0, -1, -1, 1, 1};
      const int dx[8] = {0, 0, -1, 1, -1, 1, -1, 1};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int yy = y + dy[k], xx = x + dx[k];
        if (xx < 0 || xx >= W) continue;
        if (yy == 0 || yy == H - 1 || (K && K[(size_t)yy * W + xx])) { sum[4] += (double)wgt[k]; continue; }   // seed row / sink: diagonal only
        const int dI = ((yy - 1) >> g.by_shift) - I, dJ = (xx >> g.bx_shift) - J;
        if (dI == 0 && dJ == 0) continue;                                     // inside the aggregate
        sum[(dI + 1) * 3 + (dJ + 1)] += (double)wgt[k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) sum[k] = warp_sum(sum[k]);
  if (lane == 0) {
    double diag = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) diag += sum[k];
    out[0] = (diag > 0.0) ? diag : 1.0;                                        // an aggregate of sinks only: decoupled identity row
    out[1 * g.nb] = (J > 0) ? -sum[1 * 3 + 0] : 0.0;                          // (I, J-1)
    out[2 * g.nb] = (I > 0) ? -sum[0 * 3 + 1] : 0.0;                          // (I-1, J)
    out[3 * g.nb] = (I > 0 && J > 0) ? -sum[0 * 3 + 0] : 0.0;                 // (I-1, J-1)
    out[4 * g.nb] = (I > 0 && J < g.nbx - 1) ? -sum[0 * 3 + 2] : 0.0;         // (I-1, J+1)
  }
}

// ---------------------------------------------------------------------------------------------
// Block LDL^T of Ac, one 128-thread block per image, fp64 in shared memory.
//   plain   (g.tw < 0):   Delta_I = D_I - E_I W_{I-1} E_I^T                         for I = 0 .. nby-1
//   twisted (g.tw = M):   the same for I < M;   Delta_I = D_I - E_{I+1}^T W_{I+1} E_{I+1}   for I > M (from the bottom);
//                         Delta_M = D_M - E_M W_{M-1} E_M^T - E_{M+1}^T W_{M+1} E_{M+1}   where the two fronts meet
//   W_I = Delta_I^-1 by in-place Gauss-Jordan (SPD, no pivoting).
// Record I of Wt[b] = { W_I [NB][NB + 4] (symmetrised, fp32, zero in the padding; the row pitch of the solver's shared-memory
// slots, so that a record is staged by straight 16-byte copies), eC[NB], eL[NB], eR[NB], 0[NB] } with
// E_I = the coupling of aggregate row I to row I-1 (eC: same column, eL: column J-1, eR: column J+1).
// ---------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(128) coarse_factor_kernel(CoarseGeom g, const double* __restrict__ Ab,
                                                            float* __restrict__ Wt) {
  constexpr int P = NB + 1;                      // odd pitch: column accesses are conflict-free
  __shared__ double WpT[NB * P], WpB[NB * P], M[NB * P], X[NB * P];
  __shared__ double ab[5 * NB], en[3 * NB], prow[NB], pcol[NB];
  const int tid = threadIdx.x, b = blockIdx.x;
  const int nbx = g.nbx, nby = g.nby;
  const int mid = (g.tw >= 0) ? g.tw : nby;      // rows below `mid` are eliminated from the top
  const double* A = Ab + (size_t)b * nby * 5 * NB;
  float* out = Wt + (size_t)b * nby * g.rec;
  // order of elimination: 0 .. mid-1 from the top, nby-1 .. mid+1 from the bottom, then the meeting row
  for (int step = 0; step < nby; ++step) {
    const int nt = mid < nby ? mid : nby;
    const int I = (step < nt) ? step : (step < nby - 1 ? nby - 1 - (step - nt) : mid);
    const bool from_top = I > 0 && I <= mid;                                         // subtract E_I W_{I-1} E_I^T
    const bool from_bot = I >= mid && I < nby - 1;                                   // subtract E_{I+1}^T W_{I+1} E_{I+1}
    for (int i = tid; i < 5 * NB; i += 128) ab[i] = A[(size_t)I * 5 * NB + i];
    if (from_bot) for (int i = tid; i < 3 * NB; i += 128) en[i] = A[(size_t)(I + 1) * 5 * NB + 2 * NB + i];
    __syncthreads();
    const double *dC = ab, *dL = ab + NB, *eC = ab + 2 * NB, *eL = ab + 3 * NB, *eR = ab + 4 * NB;
    const double *nC = en, *nL = en + NB, *nR = en + 2 * NB;                          // E_{I+1}
    for (int i = tid; i < NB * NB; i += 128) {   // M = D_I  (identity in the padding)
      const int j = i / NB, k = i - j * NB;
      double v = 0.0;
      if (j >= nbx || k >= nbx) v = (j == k) ? 1.0 : 0.0;
      else if (j == k) v = dC[j];
      else if (k == j - 1) v = dL[j];
      else if (k == j + 1) v = dL[k];
      M[j * P + k] = v;
    }
    if (from_top) {                              // X = E_I W_{I-1};  M -= X E_I^T
      for (int i = tid; i < NB * NB; i += 128) {
        const int j = i / NB, k = i - j * NB;
        double v = eC[j] * WpT[j * P + k];
        if (j > 0) v += eL[j] * WpT[(j - 1) * P + k];
        if (j < NB - 1) v += eR[j] * WpT[(j + 1) * P + k];
        X[j * P + k] = v;
      }
      __syncthreads();
      for (int i = tid; i < NB * NB; i += 128) {
        const int j = i / NB, k = i - j * NB;
        if (j >= nbx || k >= nbx) continue;
        double t = X[j * P + k] * eC[k];
        if (k > 0) t += X[j * P + k - 1] * eL[k];
        if (k < NB - 1) t += X[j * P + k + 1] * eR[k];
        M[j * P + k] -= t;
      }
    }
    __syncthreads();
    if (from_bot) {                              // X = E_{I+1}^T W_{I+1};  M -= X E_{I+1}
      for (int i = tid; i < NB * NB; i += 128) {
        const int j = i / NB, k = i - j * NB;
        double v = nC[j] * WpB[j * P + k];
        if (j < NB - 1) v += nL[j + 1] * WpB[(j + 1) * P + k];
        if (j > 0) v += nR[j - 1] * WpB[(j - 1) * P + k];
        X[j * P + k] = v;
      }
      __syncthreads();
      for (int i = tid; i < NB * NB; i += 128) {
        const int j = i / NB, k = i - j * NB;
        if (j >= nbx || k >= nbx) continue;
        double t = X[j * P + k] * nC[k];
        if (k < NB - 1) t += X[j * P + k + 1] * nL[k + 1];
        if (k > 0) t += X[j * P + k - 1] * nR[k - 1];
        M[j * P + k] -= t;
      }
    }
    __syncthreads();
    for (int p = 0; p < nbx; ++p) {              // in-place Gauss-Jordan inversion
      if (tid < NB) { prow[tid] = M[p * P + tid]; pcol[tid] = M[tid * P + p]; }
      __syncthreads();
      const double piv = 1.0 / fmax(prow[p], 1e-300);
      for (int i = tid; i < NB * NB; i += 128) {
        const int j = i / NB, k = i - j * NB;
        if (j >= nbx || k >= nbx) continue;
        double v;
        if (j == p) v = (k == p) ? piv : prow[k] * piv;
        else if (k == p) v = -pcol[j] * piv;
        else v = M[j * P + k] - pcol[j] * prow[k] * piv;
        M[j * P + k] = v;
      }
      __syncthreads();
    }
    float* rec = out + (size_t)I * g.rec;
    double* Wp = (I < mid) ? WpT : WpB;          // the next Schur complement of this front uses the ROUNDED inverse
    for (int i = tid; i < NB * NB; i += 128) {
      const int j = i / NB, k = i - j * NB;
      const bool in = j < nbx && k < nbx;
      const double w = in ? 0.5 * (M[j * P + k] + M[k * P + j]) : 0.0;
      rec[j * (NB + 4) + k] = (float)w;
      if (k < 4) rec[j * (NB + 4) + NB + k] = 0.f;       // row padding
      Wp[j * P + k] = (double)(float)w;
    }
    for (int i = tid; i < 4 * NB; i += 128) rec[NB * (NB + 4) + i] = (i < 3 * NB) ? (float)ab[2 * NB + i] : 0.f;
    __syncthreads();
  }
}

#endif  // __CUDACC__

}  // namespace usaug
