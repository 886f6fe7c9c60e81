// Coarse-grid correction of the confidence-map PCG (included by confmap_pcg.cu).
//
// Two-level ADDITIVE preconditioner   M^-1 = T^-1 + P Ac^-1 P^T
//   T   column-line tridiagonal part of A (the sweeps in confmap_pcg.cu)
//   P   piecewise-constant prolongation from aggregates of `by` rows x 32 columns of unknowns
//   Ac  = P^T A P, the Laplacian of the aggregate graph (+ the couplings to the seed rows on the diagonal):
//         a 9-point operator on an nby x nbx grid, symmetric positive definite, band width nbx + 1.
// The line preconditioner moves information one column per CG iteration; the coarse space carries the
// slow, blob-like error modes of these heterogeneous systems across the image at once.  Measured on the
// CPU prototype at 512x512 (beta = 90): 1460 -> ~390 iterations with 4 x 32 aggregates.
//
// Setup (once per image):   coarse_assemble_kernel -> coarse_factor_kernel (banded Cholesky, fp64 -> fp32)
// Per iteration:            DOWN restricts the new residual (rc = P^T r), the single-warp COARSE task solves
//                           Ac zc = rc with two triangular sweeps in shared memory, UP adds P zc to z.
// M^-1 stays symmetric positive definite for ANY stored factor L: the operator applied is exactly (L L^T)^-1.
#pragma once

namespace usaug {

// aggregate width: 32 pixel columns, or 64 for images wider than 960 (keeps the band <= 31)
constexpr int kCoarseMaxN = 4096;     // coarse unknowns per image (the solve keeps the vector in shared memory)
constexpr int kCoarseMaxBand = 31;    // nbx + 1: the register-window solver holds one element per lane
constexpr int kCoarseTileRows = 32;   // factor rows staged per cp.async tile = one 32-step block of the sweeps

struct CoarseGeom {
  int by;        // aggregate height in rows (power of two), 0 = coarse correction disabled
  int by_shift;  // log2(by)
  int bx_shift;  // log2(aggregate width): 5 or 6
  int nbx, nby;  // coarse grid
  int nc;        // nbx * nby
  int kd;        // band width nbx + 1
  int kd1;       // row pitch of the factor arrays: 32 floats (kd <= 31), zero padded
};

// min_shift: log2 of the smallest aggregate height tried.  Thin aggregates converge faster (CPU prototype at
// 512x512: 8x32 -> 512, 4x32 -> 402, 2x32 -> 351 iterations; 4x16 -> 393, 8x8 -> 431 at the same count) but the
// serial COARSE solve grows with nc: 2 rows pay off only when the batch is bandwidth-bound (cm_layout decides).
static inline CoarseGeom coarse_geom(int H, int W, int min_shift) {
  CoarseGeom g{};
  const int hu = H - 2;
  g.bx_shift = 5;
  g.nbx = (W + 31) >> 5;
  if (g.nbx + 1 > kCoarseMaxBand) { g.bx_shift = 6; g.nbx = (W + 63) >> 6; }
  g.kd = g.nbx + 1;
  if (g.kd > kCoarseMaxBand || hu < 8) return g;            // by = 0: disabled
  for (int sh = min_shift; sh <= 8; ++sh) {
    const int by = 1 << sh;
    const int nby = (hu + by - 1) / by;
    if (nby * g.nbx <= kCoarseMaxN) {
      g.by = by; g.by_shift = sh; g.nby = nby; g.nc = nby * g.nbx;
      g.kd1 = 32;                          // rows padded with zeros to one entry per lane
      return g;
    }
  }
  return g;
}

#ifdef __CUDACC__

// ---------------------------------------------------------------------------------------------
// Ac in lower band form: Ab[i][d] = Ac[i, i-d], d = 0..kd, i = I*nbx + J.  One warp per aggregate,
// lane = pixel column inside the aggregate; butterfly sums => deterministic.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) coarse_assemble_kernel(int H, int W, CoarseGeom g,
                                                              const float* __restrict__ wS, const float* __restrict__ wE,
                                                              const float* __restrict__ wSE, const float* __restrict__ wSW,
                                                              double* __restrict__ Ab) {
  const int lane = threadIdx.x & 31;
  const int agg = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (agg >= g.nc) return;
  const int I = agg / g.nbx, J = agg - I * g.nbx;
  const size_t base = (size_t)b * H * W;
  const float* S = wS + base; const float* E = wE + base; const float* SE = wSE + base; const float* SW = wSW + base;
  double sum[9];   // sum[(dI+1)*3 + (dJ+1)]: weight to the neighbouring aggregate; [4] collects the seed couplings
#pragma unroll
  for (int k = 0; k < 9; ++k) sum[k] = 0.0;
  const int bx = 1 << g.bx_shift;
  for (int xo = lane; xo < bx; xo += 32) {
    const int x = J * bx + xo;
    if (x >= W) continue;
    const int y0 = I * g.by + 1, y1 = min(y0 + g.by, H - 1);
    for (int y = y0; y < y1; ++y) {
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
        if (yy == 0 || yy == H - 1) { sum[4] += (double)wgt[k]; continue; }   // seed row: diagonal only
        const int dI = ((yy - 1) >> g.by_shift) - I, dJ = (xx >> g.bx_shift) - J;
        if (dI == 0 && dJ == 0) continue;                                     // inside the aggregate
        sum[(dI + 1) * 3 + (dJ + 1)] += (double)wgt[k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) sum[k] = warp_sum(sum[k]);
  double diag = 0.0;
#pragma unroll
  for (int k = 0; k < 9; ++k) diag += sum[k];
  double* row = Ab + ((size_t)b * g.nc + agg) * (g.kd + 1);
  for (int d = lane; d <= g.kd; d += 32) {
    double v = 0.0;
    if (d == 0) v = diag;
    // lower neighbours (index i - d): several can share one offset when nbx <= 2, so accumulate
    if (J > 0 && d == 1) v -= sum[1 * 3 + 0];                                 // (I, J-1)
    if (I > 0) {
      if (J < g.nbx - 1 && d == g.nbx - 1) v -= sum[0 * 3 + 2];               // (I-1, J+1)
      if (d == g.nbx) v -= sum[0 * 3 + 1];                                    // (I-1, J)
      if (J > 0 && d == g.nbx + 1) v -= sum[0 * 3 + 0];                       // (I-1, J-1)
    }
    row[d] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Banded Cholesky Ac = L L^T, right-looking, one warp per image, fp64 in a sliding shared-memory window.
//   Lr[i][0] = 1/L[i,i],  Lr[i][d] = L[i, i-d]              (row i of L:    backward sweep in saxpy form)
//   Ur[i][0] = 1/L[i,i],  Ur[i][d] = L[i+d, i] / L[i,i]     (column i of L, pre-scaled: forward sweep)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) coarse_factor_kernel(CoarseGeom g, const double* __restrict__ Ab,
                                                           float* __restrict__ Lr, float* __restrict__ Ur) {
  extern __shared__ double win[];          // [(kd+1)][(kd+1)]: win[r % (kd+1)][d] = current A[j+r, j+r-d]
  const int lane = threadIdx.x, b = blockIdx.x;
  const int n = g.nc, kd = g.kd, m = kd + 1;
  // (r1, r2) of the t-th trailing-update pair, r1-major (pairs with r1 <= rmax form a prefix), unranked once
  unsigned short* pair_tab = reinterpret_cast<unsigned short*>(win + m * m);   // [kd (kd + 1) / 2]
  for (int t = lane; t < kd * (kd + 1) / 2; t += 32) {
    int r1 = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
    while ((r1 + 1) * (r1 + 2) / 2 <= t) ++r1;
    while (r1 * (r1 + 1) / 2 > t) --r1;
    const int r2 = t - r1 * (r1 + 1) / 2 + 1;
    pair_tab[t] = (unsigned short)(((r1 + 1) << 8) | r2);
  }
  const double* A = Ab + (size_t)b * n * m;
  float* L_ = Lr + (size_t)b * n * g.kd1;
  float* U_ = Ur + (size_t)b * n * g.kd1;
  for (int i = lane; i < n * g.kd1; i += 32) { L_[i] = 0.f; U_[i] = 0.f; }
  for (int r = 0; r < m && r < n; ++r)
    for (int d = lane; d < m; d += 32) win[r * m + d] = A[(size_t)r * m + d];
  __syncwarp();
  int slot = 0;                                    // j % m, kept incrementally
  for (int j = 0; j < n; ++j) {
    // column j of L lives on the diagonal-offset line of the window: L[j+r, j] = win[(j+r) % m][r] / L[j,j]
    const double ajj = win[slot * m + 0];
    const double ljj = sqrt(fmax(ajj, 1e-300));
    const double inv = 1.0 / ljj;
    const int rmax = min(kd, n - 1 - j);
    // scale the column
    for (int r = 1 + lane; r <= rmax; r += 32) { int rs = slot + r; if (rs >= m) rs -= m; win[rs * m + r] *= inv; }
    __syncwarp();
    // trailing update: A[j+r1, j+r2] -= L[j+r1, j] * L[j+r2, j]   for 1 <= r2 <= r1 <= rmax
    const int npair = rmax * (rmax + 1) / 2;
    for (int t = lane; t < npair; t += 32) {
      const int pr = pair_tab[t], r1 = pr >> 8, r2 = pr & 255;
      int s1 = slot + r1; if (s1 >= m) s1 -= m;
      int s2 = slot + r2; if (s2 >= m) s2 -= m;
      const double l1 = win[s1 * m + r1], l2 = win[s2 * m + r2];
      win[s1 * m + (r1 - r2)] -= l1 * l2;
    }
    __syncwarp();
    // write out column j (as row j of Ur) and scatter it into the rows of Lr
    if (lane == 0) { U_[(size_t)j * g.kd1] = (float)inv; L_[(size_t)j * g.kd1] = (float)inv; }
    for (int r = 1 + lane; r <= rmax; r += 32) {
      int rs = slot + r; if (rs >= m) rs -= m;
      const float v = (float)win[rs * m + r];
      U_[(size_t)j * g.kd1 + r] = v * (float)inv;
      L_[(size_t)(j + r) * g.kd1 + r] = v;
    }
    __syncwarp();
    // slide the window: row j leaves, row j + m enters in its slot
    if (j + m < n)
      for (int d = lane; d < m; d += 32) win[slot * m + d] = A[(size_t)(j + m) * m + d];
    __syncwarp();
    if (++slot == m) slot = 0;
  }
}

#endif  // __CUDACC__

}  // namespace usaug
