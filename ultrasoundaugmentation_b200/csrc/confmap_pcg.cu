// K2: ultrasound confidence map -- random-walk potentials on the
// 8-neighbour image graph, solved per image with a matrix-free preconditioned conjugate gradient.
//
//   setup   : min/max normalise, depth attenuation, edge differences, weights
//             w = exp(-beta*(dn + pen)) + 1e-6  (4 planes: S, E, SE, SW), column-line LDL^T factors.
//   solve   : ONE persistent cooperative kernel for the whole batch.  Every image carries its own
//             state machine (RESID -> PRECOND -> INNER* -> FOLD -> RESID ... -> FINAL -> DONE):
//             an inner PCG in the vector type T (fp32 by default) wrapped in fp64 iterative
//             refinement (x64 += e; r = b - A x64 evaluated in fp64), because a pure-fp32 CG
//             stagnates at a true relative residual of ~1e-5 on these badly scaled systems.
//             A step of the kernel is   slot 1 (stencil tasks)  | grid sync |
//                                       slot 2 (column sweeps)  | grid sync.
//             Dot products: fp64 per-lane accumulation -> warp butterfly -> one slot per task ->
//             the LAST arriving task of an image sums the slots in index order (deterministic;
//             no floating-point atomics) and advances that image's state.
//   precond : column-line tridiagonal (vertical couplings + full diagonal), applied as
//             z = L^-T D^-1 L^-1 r with one down and one up sweep per column, thread per column.
#include <cooperative_groups.h>
#include <stdlib.h>
#include <cuda/std/type_traits>

#include <type_traits>

#include "kernels.cuh"
#include "confmap_coarse.cuh"

namespace cg = cooperative_groups;

namespace usaug {

constexpr int kChunks = 32;        // per-image partial reductions in the setup kernels
constexpr float kEps = 2.220446e-16f;
constexpr float kFloor = 1e-6f;
constexpr float kSqrt2 = 1.41421356237309515f;
constexpr int kStripCols = 30;     // productive columns per warp in stencil tasks (2 halo lanes)
constexpr int kRowBlock = 64;      // rows per stencil task

enum Mode : int { M_RESID = 0, M_PRECOND = 1, M_INNER = 2, M_FOLD = 3, M_FINAL = 4, M_DONE = 5 };

struct ImgState {
  int mode, iters, first, outer;
  double rz, alpha, beta, bb, rr_true, rr_prev, rr_target, rr_rec;
  unsigned cnt1, cnt2;
  // fused kernel: (first interval in which the mode may run) << 32 | mode.  Phases have different
  // task counts, so a mode published by the last arriver must not be picked up by tasks of the
  // same image that are still to come in the SAME interval.
  unsigned long long fword;
  int pcur, poor;     // pcur: which p plane holds the current search direction; poor: consecutive weak refinement rounds
  double inner2;      // squared residual reduction asked of the next inner solve (adapts to the fp32 drift floor)
  double alpha_prev;  // fused kernel: step length of the iteration whose e-update is still pending
  int epend, fresh;   // epend: 1 = e += alpha_prev * p_prev is deferred to the next DOWN / FOLD; fresh: COARSE follows a PRECOND
  double rz_line;     // line part of r.z of the current iteration (the COARSE task adds rc.zc)
};

// ---------------------------------------------------------------------------------------------
// setup kernels
// ---------------------------------------------------------------------------------------------
__global__ void cm_depth_kernel(float* __restrict__ depth, int H, float alpha) {
  const int y = blockIdx.x * blockDim.x + threadIdx.x;
  if (y < H) depth[y] = (float)(1.0 - exp(-(double)alpha * (double)y / (double)max(H - 1, 1)));
}

__device__ __forceinline__ void block_minmax(float& mn, float& mx) {
  __shared__ float s_mn[8], s_mx[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  mn = warp_min(mn); mx = warp_max(mx);
  __syncthreads();
  if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
  __syncthreads();
  mn = s_mn[0]; mx = s_mx[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); }
}

__global__ void __launch_bounds__(256) cm_minmax_kernel(const float* __restrict__ img, int N,
                                                        float* __restrict__ mm) {
  const int b = blockIdx.y, c = blockIdx.x;
  const float* I = img + (size_t)b * N;
  const int lo = (int)((long long)N * c / kChunks), hi = (int)((long long)N * (c + 1) / kChunks);
  float mn = 3.0e38f, mx = -3.0e38f;
  for (int i = lo + threadIdx.x; i < hi; i += 256) { const float v = I[i]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
  block_minmax(mn, mx);
  if (threadIdx.x == 0) { mm[((size_t)b * kChunks + c) * 2] = mn; mm[((size_t)b * kChunks + c) * 2 + 1] = mx; }
}

__device__ __forceinline__ void read_mm(const float* __restrict__ mm, int b, float& mn, float& den) {
  float lo = 3.0e38f, hi = -3.0e38f;
#pragma unroll 8
  for (int c = 0; c < kChunks; ++c) {
    lo = fminf(lo, mm[((size_t)b * kChunks + c) * 2]);
    hi = fmaxf(hi, mm[((size_t)b * kChunks + c) * 2 + 1]);
  }
  mn = lo;
  den = __fadd_rn(__fsub_rn(hi, lo), kEps);
}

// attenuated, normalised intensity a = ((I - mn) / den) * depth[y]   (A.3 steps 1-2)
__device__ __forceinline__ float atten(const float* __restrict__ I, const float* __restrict__ depth,
                                       int y, int x, int W, float mn, float den) {
  return __fmul_rn(__fdiv_rn(__fsub_rn(I[(size_t)y * W + x], mn), den), depth[y]);
}

__global__ void __launch_bounds__(256) cm_delta_kernel(const float* __restrict__ img, int H, int W,
                                                       const float* __restrict__ depth,
                                                       const float* __restrict__ mm1,
                                                       float* __restrict__ mm2) {
  const int b = blockIdx.y, c = blockIdx.x, N = H * W;
  const float* I = img + (size_t)b * N;
  float mn1, den1;
  read_mm(mm1, b, mn1, den1);
  const int lo = (int)((long long)N * c / kChunks), hi = (int)((long long)N * (c + 1) / kChunks);
  float mn = 3.0e38f, mx = -3.0e38f;
  for (int i = lo + threadIdx.x; i < hi; i += 256) {
    const int y = i / W, x = i - y * W;
    const float a0 = atten(I, depth, y, x, W, mn1, den1);
    if (x < W - 1) { const float d = fabsf(__fsub_rn(a0, atten(I, depth, y, x + 1, W, mn1, den1))); mn = fminf(mn, d); mx = fmaxf(mx, d); }
    if (y < H - 1) {
      { const float d = fabsf(__fsub_rn(a0, atten(I, depth, y + 1, x, W, mn1, den1))); mn = fminf(mn, d); mx = fmaxf(mx, d); }
      if (x < W - 1) { const float d = fabsf(__fsub_rn(a0, atten(I, depth, y + 1, x + 1, W, mn1, den1))); mn = fminf(mn, d); mx = fmaxf(mx, d); }
      if (x > 0) { const float d = fabsf(__fsub_rn(a0, atten(I, depth, y + 1, x - 1, W, mn1, den1))); mn = fminf(mn, d); mx = fmaxf(mx, d); }
    }
  }
  block_minmax(mn, mx);
  if (threadIdx.x == 0) { mm2[((size_t)b * kChunks + c) * 2] = mn; mm2[((size_t)b * kChunks + c) * 2 + 1] = mx; }
}

// Edge weight w = exp(-beta * ((|a0 - a1| - mn) / den + pen)) + 1e-6  (A.3 steps 3-5), evaluated as
//   ex2(k1 * |a0 - a1| + k0) + 1e-6,   k1 = -beta*log2(e)/den,   k0 = beta*log2(e)*(mn/den - pen)
// with per-image constants (k1, k0 per edge type) formed in fp64.  ONE function for every kernel (weight
// planes, line factors, fp64 residual, and the fused UP phase that recomputes the weights from the `a`
// plane instead of reading four planes), so the matrix is bit-identical wherever it is applied.
// Relative deviation from the oracle's exp(): <= ~5e-6 where w > 1e-6 matters (|arg| <= 20).
__device__ __forceinline__ float edge_w_fast(float a0, float a1, float k1, float k0) {
  float e;
  const float arg = fmaf(fabsf(__fsub_rn(a0, a1)), k1, k0);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(arg));
  return __fadd_rn(e, kFloor);
}

// wk[b] = {k1, k0_vertical, k0_horizontal, k0_diagonal}
__global__ void cm_wconst_kernel(int B, const float* __restrict__ mm2, float beta, float gamma,
                                 float4* __restrict__ wk) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float mn2, den2;
  read_mm(mm2, b, mn2, den2);
  const double L2E = 1.4426950408889634074, bd = (double)beta, dn = (double)den2, m = (double)mn2 / dn;
  const double pen_h = (double)gamma, pen_d = (double)__fmul_rn(kSqrt2, gamma);
  wk[b] = make_float4((float)(-bd * L2E / dn), (float)(bd * L2E * m), (float)(bd * L2E * (m - pen_h)),
                      (float)(bd * L2E * (m - pen_d)));
}

__global__ void __launch_bounds__(256) cm_weights_kernel(
    const float* __restrict__ img, int H, int W, const float* __restrict__ depth,
    const float* __restrict__ mm1, const float4* __restrict__ wk, float* __restrict__ aplane,
    float* __restrict__ wS, float* __restrict__ wE, float* __restrict__ wSE, float* __restrict__ wSW) {
  const int b = blockIdx.y, N = H * W;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= N) return;
  const float* I = img + (size_t)b * N;
  float mn1, den1;
  read_mm(mm1, b, mn1, den1);
  const float4 k = wk[b];
  const int y = i / W, x = i - y * W;
  const float a0 = atten(I, depth, y, x, W, mn1, den1);
  float s = 0.f, e = 0.f, se = 0.f, sw = 0.f;
  if (x < W - 1) e = edge_w_fast(a0, atten(I, depth, y, x + 1, W, mn1, den1), k.x, k.z);
  if (y < H - 1) {
    s = edge_w_fast(a0, atten(I, depth, y + 1, x, W, mn1, den1), k.x, k.y);
    if (x < W - 1) se = edge_w_fast(a0, atten(I, depth, y + 1, x + 1, W, mn1, den1), k.x, k.w);
    if (x > 0) sw = edge_w_fast(a0, atten(I, depth, y + 1, x - 1, W, mn1, den1), k.x, k.w);
  }
  const size_t o = (size_t)b * N + i;
  aplane[o] = a0;
  wS[o] = s; wE[o] = e; wSE[o] = se; wSW[o] = sw;
}

__device__ __forceinline__ uint32_t bf16_bits(float v) {      // round to nearest even, finite inputs
  const uint32_t u = __float_as_uint(v);
  return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// Column-line factorisation T = L D L^T (T = vertical couplings + full diagonal of L_UU):
//   piv[y] = diag[y] - wS[y-1]^2 / piv[y-1],  ip = 1/piv,  c[y] = wS[y] * ip[y].
// Jacobi: ip = 1/diag, c = 0.  Also writes the initial guess into x64 (seed rows 1 / 0) and the
// per-block partial of ||b||^2.
__global__ void __launch_bounds__(128) cm_factor_kernel(
    int H, int W, const float* __restrict__ wS, const float* __restrict__ wE,
    const float* __restrict__ wSE, const float* __restrict__ wSW, int precond, int zero_guess,
    float* __restrict__ ip, float* __restrict__ c, uint32_t* __restrict__ ipc, double* __restrict__ x64,
    double* __restrict__ bbpart) {
  __shared__ double s_bb[4];
  const int b = blockIdx.y;
  const int x = blockIdx.x * 128 + threadIdx.x;
  const size_t base = (size_t)b * H * W;
  double bb = 0.0;
  if (x < W) {
    const float* S = wS + base; const float* E = wE + base;
    const float* SE = wSE + base; const float* SW = wSW + base;
    {
      double bv = (double)S[x];
      if (x > 0) bv += (double)SE[x - 1];
      if (x < W - 1) bv += (double)SW[x + 1];
      bb = bv * bv;
    }
    x64[base + x] = 1.0;
    x64[base + (size_t)(H - 1) * W + x] = 0.0;
    ip[base + x] = 0.f; c[base + x] = 0.f; ipc[base + x] = 0u;
    ip[base + (size_t)(H - 1) * W + x] = 0.f; c[base + (size_t)(H - 1) * W + x] = 0.f;
    ipc[base + (size_t)(H - 1) * W + x] = 0u;
    double ip_prev = 0.0;   // row 0 is a seed: not part of T
    float s_prev = S[x];
    for (int y = 1; y <= H - 2; ++y) {
      const size_t o = (size_t)y * W + x;
      const float s = S[o];
      double dg = (double)s + (double)s_prev + (double)E[o] + (double)SE[o] + (double)SW[o];
      if (x > 0) dg += (double)E[o - 1] + (double)SE[o - W - 1];
      if (x < W - 1) dg += (double)SW[o - W + 1];
      double piv = dg;
      if (precond == USAUG_PRECOND_LINE && y > 1) piv -= (double)s_prev * (double)s_prev * ip_prev;
      const double ipv = 1.0 / piv;
      const float ipf = (float)ipv;
      const float cf = (precond == USAUG_PRECOND_LINE && y < H - 2) ? (float)((double)s * ipv) : 0.0f;
      ip[base + o] = ipf;
      c[base + o] = cf;
      // fused kernel: both factors as bf16 in one word (low = ip, high = c).  M = L D L^T stays SPD for
      // any positive ip and any c, and the iteration count is unchanged.
      ipc[base + o] = bf16_bits(ipf) | (bf16_bits(cf) << 16);
      x64[base + o] = zero_guess ? 0.0 : 1.0 - (double)y / (double)(H - 1);
      // keep the recurrence consistent with the ROUNDED factors the sweeps will use
      ip_prev = (double)(float)ipv;
      s_prev = s;
    }
  }
  bb = warp_sum(bb);
  if ((threadIdx.x & 31) == 0) s_bb[threadIdx.x >> 5] = bb;
  __syncthreads();
  if (threadIdx.x == 0) bbpart[(size_t)b * gridDim.x + blockIdx.x] = s_bb[0] + s_bb[1] + s_bb[2] + s_bb[3];
}

__global__ void cm_init_state_kernel(int B, int nblk, const double* __restrict__ bbpart,
                                     ImgState* __restrict__ st, unsigned* __restrict__ done) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) *done = 0u;
  if (b >= B) return;
  double bb = 0.0;
  for (int i = 0; i < nblk; ++i) bb += bbpart[(size_t)b * nblk + i];
  ImgState s;
  s.mode = M_RESID; s.iters = 0; s.first = 1; s.outer = 0;
  s.rz = 0; s.alpha = 0; s.beta = 0; s.bb = bb; s.rr_true = 0; s.rr_prev = 1e300; s.rr_target = 0; s.rr_rec = 0;
  s.cnt1 = 0; s.cnt2 = 0; s.fword = 0ull; s.pcur = 0; s.poor = 0; s.inner2 = 0.0; s.alpha_prev = 0.0; s.epend = 0; s.fresh = 0; s.rz_line = 0.0;
  st[b] = s;
}

// Decision after a true-residual evaluation (both PCG kernels).  Returns true when the image is finished.
// Refinement rounds adapt: a round that truly gained less than it was asked for (fp32 drift floor) makes
// the next round ask for less; only two consecutive rounds with < 2x gain count as stagnation.
__device__ __forceinline__ bool after_true_residual(ImgState* st, double s0, double rtol2, double inner_red2, int maxiter) {
  const double bb = __ldcg(&st->bb), prev = __ldcg(&st->rr_prev);
  const int it = __ldcg(&st->iters), outer = __ldcg(&st->outer);
  st->rr_true = s0;
  if (s0 <= rtol2 * bb || it >= maxiter) return true;
  double ask = (outer == 0) ? inner_red2 : __ldcg(&st->inner2);
  int poor = 0;
  if (outer > 0 && inner_red2 > 0.0) {
    const double achieved = s0 / prev;                      // true squared reduction of the last round
    poor = (achieved > 0.25) ? __ldcg(&st->poor) + 1 : 0;
    if (poor >= 2) return true;
    ask = fmin(1e-2, fmax(inner_red2, 10.0 * achieved));     // ask ~3x less (in norm) than was achieved
  } else if (outer > 0 && s0 > 0.25 * prev) {
    return true;                                            // fp64 vectors: no refinement to adapt
  }
  st->poor = poor;
  st->inner2 = ask;
  st->rr_prev = s0;
  st->rr_target = fmax(0.49 * rtol2 * bb, ask * s0);
  return false;
}

// ---------------------------------------------------------------------------------------------
// persistent PCG kernel
// ---------------------------------------------------------------------------------------------
template <typename T>
struct PcgArgs {
  int B, H, W, maxiter;
  double rtol2;      // rtol^2
  double inner_red2; // squared residual reduction asked of one inner solve (0 => run to rtol)
  const float *wS, *wE, *wSE, *wSW, *ip, *c;
  const uint32_t* ipc;   // fused kernel: bf16 ip (low half) | bf16 c (high half)
  const float* aplane;   // attenuated normalised intensity a (A.3 step 2); fused UP recomputes weights from it
  const float4* wk;      // per image {k1, k0_vertical, k0_horizontal, k0_diagonal} of edge_w_fast
  double* x64;
  T *e, *r, *p, *q, *z;   // q doubles as h (forward-eliminated residual) inside slot 2
  T *p2;                  // fused kernel: second search-direction plane (ping-pong); z holds h there
  ImgState* st;
  double* part;      // [B][pmax][2]
  unsigned* done;
  float* cm;
  int* iters_out;
  double* relres_out;
  int nstrip1, nrb, nstrip2, pmax, max_steps;
  unsigned long long *qslots, *qhead, *qtail;   // fused kernel: ready queue of (image, strip) tasks
  unsigned qcap;
  int wpb;           // fused kernel: worker warps per block actually used (<= AW)
  CoarseGeom cg;     // coarse-grid correction geometry (cg.by == 0: disabled)
  float *rc, *zc;    // [B][nc] restricted residual / coarse correction
  const float *Lr, *Ur;   // [B][nc][kd1] banded Cholesky factor of Ac, row- and column-oriented
  int vec4, nstrip4; // vectorised stencil (fp32 vectors, W % 4 == 0): 128-column strips
  int restart;       // experimental (USAUG_CM_CONTINUE): keep p across refinement rounds, replace r every 2 decades
  int debug;         // USAUG_CM_DEBUG_TIMING: block 0 prints per-slot wall times of a few steps
};

template <typename T>
__device__ __forceinline__ T p_new(T z, T p, T beta, bool first) {
  return first ? z : (T)fma(beta, p, z);
}

// arrival of one task of image b; returns true in all lanes of the last task to arrive
__device__ __forceinline__ bool arrive_last(unsigned* cnt, unsigned ntasks, int lane) {
  unsigned t = 0;
  __syncwarp();
  if (lane == 0) { __threadfence(); t = atomicAdd(cnt, 1u); }
  t = __shfl_sync(0xffffffffu, t, 0);
  const bool last = (t == ntasks - 1);
  if (last) __threadfence();
  return last;
}

__device__ __forceinline__ void reduce_parts(const double* part, int n, int lane, double& s0, double& s1) {
  double a0 = 0.0, a1 = 0.0;
  for (int i = lane; i < n; i += 32) { a0 += __ldcg(part + 2 * i); a1 += __ldcg(part + 2 * i + 1); }
  s0 = warp_sum(a0); s1 = warp_sum(a1);
}

// ---- slot 1: q = A p_new (INNER) or r = b - A x64 (RESID) on a 30-column x kRowBlock-row tile ---
// Warp-marching stencil: lane = column (lanes 0 / 31 are halo columns), rows march down through
// registers; horizontal neighbours come from warp shuffles, so every plane is read once with fully
// coalesced 128-byte rows.  Loads run kQ rows ahead of the arithmetic in a rolling register queue
// (each slot is refilled right after it is consumed) to keep ~kQ*6 lines per warp in flight.
template <typename A, typename T>
struct StRow { A v; T z, p; float s, e, se, sw; };

template <typename T, bool RESID>
__device__ __forceinline__ double stencil_task(const PcgArgs<T>& a, int b, int xbase, int ncols, int rbk, int lane,
                                               T beta, bool first) {
  // productive columns xbase .. xbase+ncols-1 (ncols <= 30) on lanes 1..ncols; lanes 0 and ncols+1 are halo
  using A = typename std::conditional<RESID, double, T>::type;
  constexpr int kQ = RESID ? 4 : 8;
  const int H = a.H, W = a.W;
  const size_t base = (size_t)b * H * W;
  const int x = xbase + lane - 1;
  const bool valid = (x >= 0) && (x < W) && lane <= ncols + 1;
  const bool own = valid && lane >= 1 && lane <= ncols;
  const int ya = 1 + rbk * kRowBlock, yb = min(ya + kRowBlock, H - 1);
  const float* __restrict__ S = a.wS + base; const float* __restrict__ E = a.wE + base;
  const float* __restrict__ SE = a.wSE + base; const float* __restrict__ SW = a.wSW + base;
  const double* X = a.x64 + base;
  const T* Z = a.z + base; const T* P = a.p + base;
  T* __restrict__ Q = a.q + base; T* __restrict__ Rv = a.r + base;

  // loads only -- no arithmetic on the loaded values here, or the in-order warp would wait for
  // the loads it has just issued and the queue would collapse to a depth of one row
  auto ldrow = [&](int yy) -> StRow<A, T> {
    StRow<A, T> r; r.v = (A)0; r.z = r.p = (T)0; r.s = r.e = r.se = r.sw = 0.f;
    if (valid && yy <= yb) {
      const size_t o = (size_t)yy * W + x;
      if (RESID) r.v = (A)__ldcg(X + o);      // x64 may have been folded by another SM: read at L2
      else if (yy != 0 && yy != H - 1) { r.z = Z[o]; if (!first) r.p = P[o]; }
      r.s = S[o]; r.e = E[o]; r.se = SE[o]; r.sw = SW[o];
    }
    return r;
  };
  auto value = [&](const StRow<A, T>& r) -> A {
    if (RESID) return r.v;
    return (A)p_new<T>(r.z, r.p, beta, first);      // seed rows and invalid lanes hold z = p = 0
  };
  const unsigned full = 0xffffffffu;
  StRow<A, T> m = ldrow(ya - 1), c = ldrow(ya);
  StRow<A, T> qu[kQ];
#pragma unroll
  for (int u = 0; u < kQ; ++u) qu[u] = ldrow(ya + 1 + u);
  A Pm = value(m), Pc = value(c);
  A PmL = __shfl_up_sync(full, Pm, 1), PmR = __shfl_down_sync(full, Pm, 1);
  A PcL = __shfl_up_sync(full, Pc, 1), PcR = __shfl_down_sync(full, Pc, 1);
  float wSEmL = __shfl_up_sync(full, m.se, 1), wSWmR = __shfl_down_sync(full, m.sw, 1);
  float mS = m.s;
  double acc = 0.0;
  for (int y0 = ya; y0 < yb; y0 += kQ) {
#pragma unroll
    for (int u = 0; u < kQ; ++u) {
      const int y = y0 + u;
      if (y < yb) {
        const StRow<A, T> n = qu[u];
        qu[u] = ldrow(y + 1 + kQ);                    // refill first: the new loads go out before we block on n
        const A Pn = value(n);
        const A PnL = __shfl_up_sync(full, Pn, 1), PnR = __shfl_down_sync(full, Pn, 1);
        const float eL = __shfl_up_sync(full, c.e, 1);
        A q = (A)c.s * (Pc - Pn);
        q += (A)mS * (Pc - Pm);
        q += (A)c.e * (Pc - PcR);
        q += (A)eL * (Pc - PcL);
        q += (A)c.se * (Pc - PnR);
        q += (A)c.sw * (Pc - PnL);
        q += (A)wSEmL * (Pc - PmL);
        q += (A)wSWmR * (Pc - PmR);
        if (own) {
          const size_t o = (size_t)y * W + x;
          if (RESID) { const double rr = -(double)q; Rv[o] = (T)rr; acc += rr * rr; }
          else { Q[o] = (T)q; acc += (double)Pc * (double)q; }
        }
        wSEmL = __shfl_up_sync(full, c.se, 1);
        wSWmR = __shfl_down_sync(full, c.sw, 1);
        mS = c.s;
        Pm = Pc; PmL = PcL; PmR = PcR;
        Pc = Pn; PcL = PnL; PcR = PnR;
        c = n;
      }
    }
  }
  return warp_sum(acc);
}

// ---- slot 1, vectorised (fp32 vectors, W % 4 == 0): lane = 4 adjacent columns, warp = 128-column strip.
// Global -> shared staging with cp.async (LDGSTS): every lane copies ITS OWN 16-byte pieces of the next
// kRing rows into a private ring in shared memory (zero-fill predication instead of branches), so no
// barrier is ever needed -- cp.async.wait_group orders a lane against its own copies only -- and the
// bytes in flight are bounded by shared memory (~21 KB per warp), not by registers.
// Horizontal neighbours inside the lane are registers, across lanes two shuffles per row, across strips
// one 4-byte halo copy on lane 0 / lane 31.
constexpr int kStrip4 = 128;
constexpr int kRing = 6;

struct RingRow {            // one staged image row of a 128-column strip (3584 bytes)
  float4 z[32], p[32], s[32], e[32], se[32], sw[32];
  float hz[32], hp[32], hw1[32], hw2[32];   // lane 0: (z, p, wE, wSE) at column x0-1; lane 31: (z, p, wSW, -) at x0+128
};
constexpr size_t kRingBytesPerWarp = sizeof(RingRow) * kRing;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int n = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem, bool pred) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int n = pred ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sa), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ double stencil_task_v4(const PcgArgs<float>& a, RingRow* ring, int b, int strip, int rbk,
                                                  int lane, float beta, bool first) {
  const int H = a.H, W = a.W;
  const size_t base = (size_t)b * H * W;
  const int x0 = strip * kStrip4, x = x0 + 4 * lane;
  const bool valid = x < W;
  const int hx = (lane == 0) ? x0 - 1 : x0 + kStrip4;
  const bool hvalid = (lane == 0 && x0 > 0) || (lane == 31 && x0 + kStrip4 < W);
  const int ya = 1 + rbk * kRowBlock, yb = min(ya + kRowBlock, H - 1);
  const float* __restrict__ S = a.wS + base; const float* __restrict__ E = a.wE + base;
  const float* __restrict__ SE = a.wSE + base; const float* __restrict__ SW = a.wSW + base;
  const float* Z = a.z + base; const float* P = a.p + base;
  float* __restrict__ Q = a.q + base;
  const float* __restrict__ HW1 = (lane == 0) ? E : SW;
  const int xs = valid ? x : 0, hxs = hvalid ? hx : 0;      // keep predicated-off addresses in bounds

  // stage row yy into ring slot (yy - (ya-1)) % kRing; always commits a group so the count stays in step
  auto issue = [&](int yy) {
    RingRow* r = ring + ((yy - (ya - 1)) % kRing);
    const bool in = yy <= yb;
    const bool interior = in && (yy != 0) && (yy != H - 1);
    const size_t o = (size_t)(in ? yy : 0) * W;
    cp_async16(&r->z[lane], Z + o + xs, interior && valid);
    cp_async16(&r->p[lane], P + o + xs, interior && valid && !first);
    cp_async16(&r->s[lane], S + o + xs, in && valid);
    cp_async16(&r->e[lane], E + o + xs, in && valid);
    cp_async16(&r->se[lane], SE + o + xs, in && valid);
    cp_async16(&r->sw[lane], SW + o + xs, in && valid);
    cp_async4(&r->hz[lane], Z + o + hxs, interior && hvalid);
    cp_async4(&r->hp[lane], P + o + hxs, interior && hvalid && !first);
    cp_async4(&r->hw1[lane], HW1 + o + hxs, in && hvalid);
    cp_async4(&r->hw2[lane], SE + o + hxs, in && hvalid && lane == 0);
    cp_async_commit();
  };
  const unsigned full = 0xffffffffu;
  struct WRow { float4 s, e, se, sw; float hw1, hw2; };
  // v[0] = left neighbour column, v[1..4] = own columns, v[5] = right neighbour column
  auto fetch = [&](int yy, float (&v)[6], WRow& w) {
    const RingRow* r = ring + ((yy - (ya - 1)) % kRing);
    const float4 z = r->z[lane], p = r->p[lane];
    w.s = r->s[lane]; w.e = r->e[lane]; w.se = r->se[lane]; w.sw = r->sw[lane];
    w.hw1 = r->hw1[lane]; w.hw2 = r->hw2[lane];
    v[1] = p_new<float>(z.x, p.x, beta, first); v[2] = p_new<float>(z.y, p.y, beta, first);
    v[3] = p_new<float>(z.z, p.z, beta, first); v[4] = p_new<float>(z.w, p.w, beta, first);
    const float h = p_new<float>(r->hz[lane], r->hp[lane], beta, first);
    const float l = __shfl_up_sync(full, v[4], 1), rt = __shfl_down_sync(full, v[1], 1);
    v[0] = (lane == 0) ? h : l;
    v[5] = (lane == 31) ? h : rt;
  };

#pragma unroll
  for (int u = 0; u < kRing; ++u) issue(ya - 1 + u);
  float pm[6], pc[6], pn[6];
  WRow m, c, n;
  cp_async_wait<kRing - 2>();            // rows ya-1 and ya have landed
  fetch(ya - 1, pm, m);
  fetch(ya, pc, c);
  issue(ya - 1 + kRing);
  issue(ya + kRing);
  double acc = 0.0;
#pragma unroll 2
  for (int y = ya; y < yb; ++y) {
    cp_async_wait<kRing - 1>();          // row y+1 has landed (kRing-1 younger groups may still fly)
    fetch(y + 1, pn, n);
    // edge weights shared with the neighbouring lanes / strips
    const float eL_s = __shfl_up_sync(full, c.e.w, 1);
    const float semL_s = __shfl_up_sync(full, m.se.w, 1);
    const float swmR_s = __shfl_down_sync(full, m.sw.x, 1);
    const float eL = (lane == 0) ? c.hw1 : eL_s;          // wE[y, x-1]
    const float semL = (lane == 0) ? m.hw2 : semL_s;      // wSE[y-1, x-1]
    const float swmR = (lane == 31) ? m.hw1 : swmR_s;     // wSW[y-1, x+4]
    const float S_[4] = {c.s.x, c.s.y, c.s.z, c.s.w};
    const float Sm[4] = {m.s.x, m.s.y, m.s.z, m.s.w};
    const float E5[5] = {eL, c.e.x, c.e.y, c.e.z, c.e.w};
    const float SE_[4] = {c.se.x, c.se.y, c.se.z, c.se.w};
    const float SW_[4] = {c.sw.x, c.sw.y, c.sw.z, c.sw.w};
    const float SEm[4] = {semL, m.se.x, m.se.y, m.se.z};
    const float SWm[4] = {m.sw.y, m.sw.z, m.sw.w, swmR};
    float q[4], dot = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float pcj = pc[j + 1];
      float t = S_[j] * (pcj - pn[j + 1]);
      t = fmaf(Sm[j], pcj - pm[j + 1], t);
      t = fmaf(E5[j + 1], pcj - pc[j + 2], t);
      t = fmaf(E5[j], pcj - pc[j], t);
      t = fmaf(SE_[j], pcj - pn[j + 2], t);
      t = fmaf(SW_[j], pcj - pn[j], t);
      t = fmaf(SEm[j], pcj - pm[j], t);
      t = fmaf(SWm[j], pcj - pm[j + 2], t);
      q[j] = t;
      dot = fmaf(pcj, t, dot);
    }
    if (valid) {
      *reinterpret_cast<float4*>(Q + (size_t)y * W + x) = make_float4(q[0], q[1], q[2], q[3]);
      acc += (double)dot;
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) { pm[j] = pc[j]; pc[j] = pn[j]; }
    m = c; c = n;
    issue(y + 1 + kRing);                // refill the slot row y+1 was read from (its values are consumed)
  }
  cp_async_wait<0>();                    // drain before the ring is reused by the next task
  return warp_sum(acc);
}

// ---- slot 2 bodies: one lane = one column, rows 1..H-2 ------------------------------------------
// INNER: p = z + beta p; e += alpha p; r -= alpha q; h = L^-1 r (stored over q); then z = L^-T D^-1 h.
// The recurrences are sequential in y, so parallelism is one lane per column; bandwidth comes from
// rolling register prefetch queues (kD rows x 7 planes down, kUp rows x 3 planes up per lane).
template <typename T>
__device__ __forceinline__ void sweep_inner(const PcgArgs<T>& a, int b, int x, bool valid, T alpha, T beta,
                                            bool first, double& rr_out, double& rz_out) {
  constexpr int kD = sizeof(T) == 4 ? 8 : 4;
  constexpr int kUp = sizeof(T) == 4 ? 16 : 8;
  const int H = a.H, W = a.W;
  const size_t base = (size_t)b * H * W + x;
  const float* __restrict__ IP = a.ip + base; const float* __restrict__ C = a.c + base;
  T* __restrict__ Z = a.z + base; T* __restrict__ P = a.p + base; T* __restrict__ E = a.e + base;
  T* __restrict__ R = a.r + base; T* __restrict__ Q = a.q + base;
  double rr = 0.0, rz = 0.0;
  if (valid) {
    {
      T zv[kD], pv[kD], ev[kD], rv[kD], qv[kD]; float iv[kD], cv[kD];
      auto ld = [&](int u, int y) {
        if (y <= H - 2) {
          const size_t o = (size_t)y * W;
          zv[u] = Z[o]; pv[u] = first ? (T)0 : P[o]; ev[u] = E[o]; rv[u] = R[o]; qv[u] = Q[o];
          iv[u] = IP[o]; cv[u] = C[o];
        }
      };
#pragma unroll
      for (int u = 0; u < kD; ++u) ld(u, 1 + u);
      T hprev = (T)0; float cprev = 0.f;
      for (int y0 = 1; y0 <= H - 2; y0 += kD) {
#pragma unroll
        for (int u = 0; u < kD; ++u) {
          const int y = y0 + u;
          if (y <= H - 2) {
            const size_t o = (size_t)y * W;
            const T pn = p_new<T>(zv[u], pv[u], beta, first);
            P[o] = pn;
            E[o] = (T)fma(alpha, pn, ev[u]);
            const T rn = (T)fma(-alpha, qv[u], rv[u]);
            R[o] = rn;
            rr += (double)rn * (double)rn;
            const T h = (T)fma((T)cprev, hprev, rn);
            Q[o] = h;
            rz += (double)iv[u] * ((double)h * (double)h);
            hprev = h; cprev = cv[u];
            ld(u, y + kD);
          }
        }
      }
    }
    {
      T hv[kUp]; float iv[kUp], cv[kUp];
      auto ld = [&](int u, int y) {
        if (y >= 1) { const size_t o = (size_t)y * W; hv[u] = Q[o]; iv[u] = IP[o]; cv[u] = C[o]; }
      };
#pragma unroll
      for (int u = 0; u < kUp; ++u) ld(u, H - 2 - u);
      T znext = (T)0;
      for (int y0 = H - 2; y0 >= 1; y0 -= kUp) {
#pragma unroll
        for (int u = 0; u < kUp; ++u) {
          const int y = y0 - u;
          if (y >= 1) {
            const size_t o = (size_t)y * W;
            const T zz = (T)fma((T)cv[u], znext, (T)iv[u] * hv[u]);
            Z[o] = zz; znext = zz;
            ld(u, y - kUp);
          }
        }
      }
    }
  }
  rr_out = warp_sum(rr); rz_out = warp_sum(rz);
}

// PRECOND: z = M^-1 r for a fresh residual, e = 0
template <typename T>
__device__ __forceinline__ double sweep_precond(const PcgArgs<T>& a, int b, int x, bool valid) {
  const int H = a.H, W = a.W;
  const size_t base = (size_t)b * H * W + x;
  const float* __restrict__ IP = a.ip + base; const float* __restrict__ C = a.c + base;
  T* __restrict__ Z = a.z + base; T* __restrict__ E = a.e + base;
  const T* __restrict__ R = a.r + base; T* __restrict__ Q = a.q + base;
  double rz = 0.0;
  if (valid) {
    T hprev = (T)0; float cprev = 0.f;
#pragma unroll 4
    for (int y = 1; y <= H - 2; ++y) {
      const size_t o = (size_t)y * W;
      const T h = (T)fma((T)cprev, hprev, R[o]);
      Q[o] = h; E[o] = (T)0;
      rz += (double)IP[o] * ((double)h * (double)h);
      hprev = h; cprev = C[o];
    }
    T znext = (T)0;
#pragma unroll 4
    for (int y = H - 2; y >= 1; --y) {
      const size_t o = (size_t)y * W;
      const T zz = (T)fma((T)C[o], znext, (T)IP[o] * Q[o]);
      Z[o] = zz; znext = zz;
    }
  }
  return warp_sum(rz);
}

template <typename T>
__global__ void __launch_bounds__(256) pcg_persistent_kernel(const PcgArgs<T> a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  RingRow* ring = reinterpret_cast<RingRow*>(smem_raw) + (size_t)(threadIdx.x >> 5) * kRing;
  const int gw = blockIdx.x * wpb + (threadIdx.x >> 5);
  const int nw = gridDim.x * wpb;
  const int B = a.B, H = a.H, W = a.W;
  const int per_rs = a.nstrip1 * a.nrb;                       // RESID: scalar 30-column strips
  const int per_in = a.vec4 ? a.nstrip4 * a.nrb : per_rs;     // INNER: 128-column strips when vectorised
  const int tot1 = (per_in > per_rs ? per_in : per_rs) * B, tot2 = a.nstrip2 * B;

  auto now = []() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
  for (int step = 0; step < a.max_steps; ++step) {
    unsigned long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    const bool dbg = a.debug && blockIdx.x == 0 && threadIdx.x == 0 && step >= 10 && step < 14;
    if (dbg) t0 = now();
    // ------------------------------ slot 1: stencil tasks --------------------------------------
    for (int t = gw; t < tot1; t += nw) {
      const int b = t % B, ti = t / B;
      ImgState* st = a.st + b;
      const int mode = __ldcg(&st->mode);
      if (mode != M_INNER && mode != M_RESID) continue;
      const int per1 = (mode == M_INNER) ? per_in : per_rs;
      if (ti >= per1) continue;
      const int nstr = (mode == M_INNER && a.vec4) ? a.nstrip4 : a.nstrip1;
      const int strip = ti % nstr, rbk = ti / nstr;
      double s;
      if (mode == M_INNER) {
        const bool first = __ldcg(&st->first) != 0;
        const T beta = (T)__ldcg(&st->beta);
        if constexpr (std::is_same<T, float>::value) {
          if (a.vec4) s = stencil_task_v4(a, ring, b, strip, rbk, lane, beta, first);
          else s = stencil_task<T, false>(a, b, strip * kStripCols, kStripCols, rbk, lane, beta, first);
        } else {
          s = stencil_task<T, false>(a, b, strip * kStripCols, kStripCols, rbk, lane, beta, first);
        }
      } else {
        s = stencil_task<T, true>(a, b, strip * kStripCols, kStripCols, rbk, lane, (T)0, true);
      }
      double* part = a.part + (size_t)b * a.pmax * 2;
      if (lane == 0) { part[2 * ti] = s; part[2 * ti + 1] = 0.0; }
      if (arrive_last(&st->cnt1, (unsigned)per1, lane)) {
        double s0, s1;
        reduce_parts(part, per1, lane, s0, s1);
        if (lane == 0) {
          st->cnt1 = 0;
          if (mode == M_INNER) {
            if (s0 > 0.0) st->alpha = __ldcg(&st->rz) / s0;
            else { st->alpha = 0.0; st->rr_target = 1e300; }   // breakdown: leave the inner solve
          } else {
            st->mode = after_true_residual(st, s0, a.rtol2, a.inner_red2, a.maxiter) ? M_FINAL : M_PRECOND;
          }
        }
      }
    }
    if (dbg) t1 = now();
    grid.sync();
    if (dbg) t2 = now();
    // ------------------------------ slot 2: column sweeps --------------------------------------
    for (int t = gw; t < tot2; t += nw) {
      const int b = t % B, strip = t / B;
      ImgState* st = a.st + b;
      const int mode = __ldcg(&st->mode);
      if (mode == M_DONE || mode == M_RESID) continue;
      const int x = strip * 32 + lane;
      const bool valid = x < W;
      const size_t base = (size_t)b * H * W + x;
      double s0 = 0.0, s1 = 0.0;
      if (mode == M_INNER) {
        const bool first = __ldcg(&st->first) != 0;
        sweep_inner<T>(a, b, x, valid, (T)__ldcg(&st->alpha), (T)__ldcg(&st->beta), first, s0, s1);
      } else if (mode == M_PRECOND) {
        s1 = sweep_precond<T>(a, b, x, valid);
      } else if (mode == M_FOLD) {
        if (valid) {
#pragma unroll 4
          for (int y = 1; y <= H - 2; ++y) { const size_t o = base + (size_t)y * W; a.x64[o] += (double)a.e[o]; }
        }
      } else {  // M_FINAL
        if (valid) {
#pragma unroll 4
          for (int y = 0; y < H; ++y) { const size_t o = base + (size_t)y * W; a.cm[o] = (float)a.x64[o]; }
        }
      }
      double* part = a.part + (size_t)b * a.pmax * 2;
      if (lane == 0) { part[2 * strip] = s0; part[2 * strip + 1] = s1; }
      if (arrive_last(&st->cnt2, (unsigned)a.nstrip2, lane)) {
        reduce_parts(part, a.nstrip2, lane, s0, s1);
        if (lane == 0) {
          st->cnt2 = 0;
          if (mode == M_INNER) {
            const double rz_old = __ldcg(&st->rz);
            const int it = __ldcg(&st->iters) + 1;
            st->iters = it; st->rr_rec = s0; st->rz = s1; st->first = 0;
            st->beta = (rz_old > 0.0) ? s1 / rz_old : 0.0;
            if (s0 <= __ldcg(&st->rr_target) || it >= a.maxiter || !(s1 > 0.0)) st->mode = M_FOLD;
          } else if (mode == M_PRECOND) {
            st->rz = s1; st->first = 1; st->beta = 0.0;
            st->mode = (s1 > 0.0) ? M_INNER : M_FINAL;
          } else if (mode == M_FOLD) {
            st->mode = M_RESID; st->outer = __ldcg(&st->outer) + 1;
          } else {
            st->mode = M_DONE;
            const double bb = __ldcg(&st->bb);
            if (a.iters_out) a.iters_out[b] = __ldcg(&st->iters);
            if (a.relres_out) a.relres_out[b] = (bb > 0.0) ? sqrt(__ldcg(&st->rr_true) / bb) : 0.0;
            __threadfence();
            atomicAdd(a.done, 1u);
          }
        }
      }
    }
    if (dbg) t3 = now();
    grid.sync();
    if (dbg) printf("[pcg dbg] step %d: slot1 %llu ns (block 0), sync wait %llu ns, slot2 %llu ns, sync wait %llu ns\n",
                    step, t1 - t0, t2 - t1, t3 - t2, now() - t3);
    if (__ldcg(a.done) >= (unsigned)B) break;
  }
}

// =============================================================================================
// Fused fp32 kernel (W % 4 == 0): the production path.
//   One grid-sync interval = one PHASE per image (images may be in different phases):
//     F_UP    z = L^-T D^-1 h on the fly (never stored), p = z + beta p, q = A p, partial p.q
//             -- the up sweep of iteration k fused with the stencil of iteration k+1
//     F_DOWN  e += alpha p; r -= alpha q; h = L^-1 r (stored over q); partial r.r and r.z = sum ip h^2
//   19 plane touches per unknown and iteration (UP: read h, ip, c, p, 4 weights, write p, q;
//   DOWN: read p, e, r, q, c, ip, write e, r, h) instead of 22, and no z plane.
//   Task = (image, 128-column strip), lane = 4 adjacent columns, rows staged through a private
//   cp.async ring in shared memory (see stencil_task_v4); the column recurrences are 4-way
//   independent per lane.  Strip halo columns: lane 0 / lane 31 carry one extra scalar recurrence;
//   their six scalars per row are fetched by ONE cp.async.4 instruction (lanes 0-5 left, 16-21 right).
// =============================================================================================
enum FMode : int { F_RESID = 0, F_PRECOND = 1, F_UP = 2, F_DOWN = 3, F_FOLD = 4, F_FINAL = 5, F_DONE = 6, F_COARSE = 7 };

struct RingU {   // one staged row of a 128-column strip for F_UP (2560 bytes)
  float4 h[32], p[32], a[32];
  uint4 ipc[32];
  // 16-byte chunks that contain the strip's halo columns (cp.async.cg = L2, never a stale L1 line):
  // [0..3] columns x0-4..x0-1 of h, ipc, p, a (.w is the left neighbour column);
  // [16..19] columns x0+128..x0+131 of the same planes (.x is the right neighbour column)
  float4 halo[32];
};
struct RingD {   // one staged row for F_DOWN / F_PRECOND (3072 bytes)
  float4 p[32], pp[32], e[32], r[32], q[32];
  uint4 ipc[32];
};
constexpr int kRingRowsPerCta = 80;                                   // 80 x 2560 B = 200 KB per CTA
constexpr size_t kFusedSmem = sizeof(RingU) * kRingRowsPerCta;

__device__ __forceinline__ float hsel(const float4& v, int lane) { return lane == 31 ? v.x : v.w; }

struct RingUP {   // one staged row of a 128-column strip for F_UP, weight planes read from memory (4096 bytes)
  float4 h[32], p[32], s[32], e[32], se[32], sw[32];
  uint4 ipc[32];
  // 16-byte chunks that contain the strip's halo columns (cp.async.cg = L2, never a stale L1 line):
  // [0..4] columns x0-4..x0-1 of h, ipc, p, wE, wSE (.w is the left neighbour column);
  // [16..19] columns x0+128..x0+131 of h, ipc, p, wSW (.x is the right neighbour column)
  float4 halo[32];
};

// Prolongation P zc inside F_UP: the coarse value of this lane's aggregate column (and of the strip's halo
// column on lanes 0 / 31), fetched one aggregate row ahead of use while the sweep marches upward.
struct CoarseZ {
  const float* zc;      // this image's coarse correction, nullptr = disabled
  int sh, nbx, jown, jhalo, icur;
  float cur, curh, nxt, nxth;
  bool own, halo;
  __device__ __forceinline__ void init(const PcgArgs<float>& a, int b, int x, int x0, int lane, bool valid, int W) {
    zc = a.cg.by ? a.zc + (size_t)b * a.cg.nc : nullptr;
    sh = a.cg.by_shift; nbx = a.cg.nbx; icur = -1;
    own = valid; jown = valid ? (x >> a.cg.bx_shift) : 0;
    const int hx = (lane == 0) ? x0 - 1 : x0 + kStrip4;
    halo = (lane == 0 && x0 > 0) || (lane == 31 && x0 + kStrip4 < W);
    jhalo = halo ? (hx >> a.cg.bx_shift) : 0;
    cur = curh = nxt = nxth = 0.f;
  }
  // returns the additive terms for row y (0 outside the unknown rows)
  __device__ __forceinline__ void at_row(int y, int H, float& add, float& addh) {
    add = 0.f; addh = 0.f;
    if (zc == nullptr || y < 1 || y > H - 2) return;
    const int I = (y - 1) >> sh;
    if (I != icur) {
      if (icur < 0) {
        cur = own ? __ldcg(zc + I * nbx + jown) : 0.f;
        curh = halo ? __ldcg(zc + I * nbx + jhalo) : 0.f;
      } else {
        cur = nxt; curh = nxth;
      }
      icur = I;
      if (I > 0) {
        nxt = own ? __ldcg(zc + (I - 1) * nbx + jown) : 0.f;
        nxth = halo ? __ldcg(zc + (I - 1) * nbx + jhalo) : 0.f;
      }
    }
    add = cur; addh = curh;
  }
};

// Edge weights of one image row for the 4 columns of a lane (plus the three that belong to its
// neighbours' columns), recomputed from the `a` plane.
struct WRow {
  float s[4], e[4], se[4], sw[4];
  float eL;    // wE [row, x-1]
  float seL;   // wSE[row, x-1]
  float swR;   // wSW[row, x+4]
};

template <int R>
__device__ __forceinline__ double up_task(const PcgArgs<float>& a, unsigned char* ringmem, int b, int strip, int lane,
                                          float beta, bool first, int pcur) {
  RingU* ring = reinterpret_cast<RingU*>(ringmem);
  const int H = a.H, W = a.W;
  const size_t base = (size_t)b * H * W;
  const int x0 = strip * kStrip4, x = x0 + 4 * lane;
  const bool valid = x < W;
  const bool hasL = valid && x > 0, hasR = valid && (x + 4 < W);
  const int xs = valid ? x : 0;
  const uint32_t* __restrict__ IPC = a.ipc + base;
  const float* __restrict__ Ap = a.aplane + base;
  const float4 wk = a.wk[b];
  // reads: h (z plane), p_old (pin); writes: p_new (pout), q.  Nothing is updated in place, because the
  // halo columns of this strip belong to a neighbouring strip whose warp runs asynchronously.
  const float* __restrict__ Hb = a.z + base;
  const float* __restrict__ Pin = (pcur ? a.p2 : a.p) + base;
  float* __restrict__ P = (pcur ? a.p : a.p2) + base;
  float* __restrict__ Q = a.q + base;
  // halo fetch role of this lane: k = which plane, left (lanes 0-3) or right (lanes 16-19) chunk
  const int k = lane & 15;
  const bool isL = lane < 16;
  const int hx = isL ? x0 - 4 : x0 + kStrip4;
  const bool hcol = isL ? (x0 > 0) : (x0 + kStrip4 < W);
  const bool hrole = hcol && k <= 3;
  const float* hsrc = (k == 0) ? Hb : (k == 1) ? reinterpret_cast<const float*>(IPC) : (k == 2) ? Pin : Ap;
  const int hxs = hrole ? hx : 0;

  auto issue = [&](int y) {                       // stage row y (rows are visited bottom -> top, H-1 .. 0)
    RingU* r = ring + ((H - 1 - y) % R);
    const bool in = y >= 0;
    const bool interior = y >= 1 && y <= H - 2;
    const size_t o = (size_t)(in ? y : 0) * W;
    cp_async16(&r->h[lane], Hb + o + xs, interior && valid);
    cp_async16(&r->ipc[lane], IPC + o + xs, interior && valid);
    cp_async16(&r->p[lane], Pin + o + xs, interior && valid && !first);
    cp_async16(&r->a[lane], Ap + o + xs, in && valid);
    cp_async16(&r->halo[lane], hsrc + o + hxs, in && hrole && (k == 3 || interior) && !(k == 2 && first));
    cp_async_commit();
  };
  const unsigned full = 0xffffffffu;
  float zn[4] = {0.f, 0.f, 0.f, 0.f}, zhn = 0.f;            // z of the row below (own columns, halo column)
  float pl[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, pc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, pu[6];
  float ab[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, au[6];     // a of the row below / of this row, columns x-1..x+4
  WRow wc, wu;
#pragma unroll
  for (int j = 0; j < 4; ++j) wc.s[j] = wc.e[j] = wc.se[j] = wc.sw[j] = 0.f;
  wc.eL = wc.seL = wc.swR = 0.f;
  const int hoff = (lane == 31) ? 16 : 0;
  CoarseZ cz;
  cz.init(a, b, x, x0, lane, valid, W);

#pragma unroll
  for (int u = 0; u < R; ++u) issue(H - 1 - u);
  double acc = 0.0;
#pragma unroll 2
  for (int y = H - 1; y >= 0; --y) {
    cp_async_wait<R - 1>();
    __syncwarp();          // the halo chunks were copied by OTHER lanes: their wait_group must have passed too
    {
      const RingU* r = ring + ((H - 1 - y) % R);
      const float4 h = r->h[lane], p = r->p[lane], av = r->a[lane];
      const uint4 f = r->ipc[lane];
      // lane 0 uses the left chunks' .w, lane 31 the right chunks' .x; other lanes read but never use them
      const float hh = hsel(r->halo[hoff + 0], lane);
      const uint32_t hf = __float_as_uint(hsel(r->halo[hoff + 1], lane));
      const float hp = hsel(r->halo[hoff + 2], lane);
      const float ha = hsel(r->halo[hoff + 3], lane);
      float z[4];     // line part of z (the recurrence runs on it alone); the coarse part is added on top
      z[0] = fmaf(bf16_hi(f.x), zn[0], bf16_lo(f.x) * h.x); z[1] = fmaf(bf16_hi(f.y), zn[1], bf16_lo(f.y) * h.y);
      z[2] = fmaf(bf16_hi(f.z), zn[2], bf16_lo(f.z) * h.z); z[3] = fmaf(bf16_hi(f.w), zn[3], bf16_lo(f.w) * h.w);
      const float zh = fmaf(bf16_hi(hf), zhn, bf16_lo(hf) * hh);
      float zadd, zaddh;
      cz.at_row(y, H, zadd, zaddh);
      pu[1] = fmaf(beta, p.x, z[0] + zadd); pu[2] = fmaf(beta, p.y, z[1] + zadd);
      pu[3] = fmaf(beta, p.z, z[2] + zadd); pu[4] = fmaf(beta, p.w, z[3] + zadd);
      const float ph = fmaf(beta, hp, zh + zaddh);
      const float l = __shfl_up_sync(full, pu[4], 1), rt = __shfl_down_sync(full, pu[1], 1);
      pu[0] = (lane == 0) ? ph : l;
      pu[5] = (lane == 31) ? ph : rt;
      zn[0] = z[0]; zn[1] = z[1]; zn[2] = z[2]; zn[3] = z[3]; zhn = zh;
      if (valid && y >= 1 && y <= H - 2)
        *reinterpret_cast<float4*>(P + (size_t)y * W + x) = make_float4(pu[1], pu[2], pu[3], pu[4]);
      au[1] = av.x; au[2] = av.y; au[3] = av.z; au[4] = av.w;
      const float al = __shfl_up_sync(full, av.w, 1), ar = __shfl_down_sync(full, av.x, 1);
      au[0] = (lane == 0) ? ha : al;
      au[5] = (lane == 31) ? ha : ar;
    }
    __syncwarp();          // every lane has read this slot's halo before any lane refills it (issue below)
    // weights of row y (edges inside the row and down to row y+1), masked where the edge does not exist
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      wu.s[j] = valid ? edge_w_fast(au[j + 1], ab[j + 1], wk.x, wk.y) : 0.f;
      wu.e[j] = valid ? edge_w_fast(au[j + 1], au[j + 2], wk.x, wk.z) : 0.f;
      wu.se[j] = valid ? edge_w_fast(au[j + 1], ab[j + 2], wk.x, wk.w) : 0.f;
      wu.sw[j] = valid ? edge_w_fast(au[j + 1], ab[j], wk.x, wk.w) : 0.f;
    }
    wu.eL = hasL ? edge_w_fast(au[0], au[1], wk.x, wk.z) : 0.f;
    wu.seL = hasL ? edge_w_fast(au[0], ab[1], wk.x, wk.w) : 0.f;
    wu.swR = hasR ? edge_w_fast(au[5], ab[4], wk.x, wk.w) : 0.f;
    if (!hasR) { wu.e[3] = 0.f; wu.se[3] = 0.f; }
    if (!hasL) wu.sw[0] = 0.f;

    const int yc = y + 1;                          // row whose q is now computable (u = y, c = y+1, l = y+2)
    if (yc <= H - 2) {
      const float E5[5] = {wc.eL, wc.e[0], wc.e[1], wc.e[2], wc.e[3]};
      const float SEu[4] = {wu.seL, wu.se[0], wu.se[1], wu.se[2]};
      const float SWu[4] = {wu.sw[1], wu.sw[2], wu.sw[3], wu.swR};
      float q[4], dot = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pcj = pc[j + 1];
        float t = wc.s[j] * (pcj - pl[j + 1]);
        t = fmaf(wu.s[j], pcj - pu[j + 1], t);
        t = fmaf(E5[j + 1], pcj - pc[j + 2], t);
        t = fmaf(E5[j], pcj - pc[j], t);
        t = fmaf(wc.se[j], pcj - pl[j + 2], t);
        t = fmaf(wc.sw[j], pcj - pl[j], t);
        t = fmaf(SEu[j], pcj - pu[j], t);
        t = fmaf(SWu[j], pcj - pu[j + 2], t);
        q[j] = t;
        dot = fmaf(pcj, t, dot);
      }
      if (valid) {
        *reinterpret_cast<float4*>(Q + (size_t)yc * W + x) = make_float4(q[0], q[1], q[2], q[3]);
        acc += (double)dot;
      }
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) { pl[j] = pc[j]; pc[j] = pu[j]; ab[j] = au[j]; }
    wc = wu;
    issue(y - R);
  }
  cp_async_wait<0>();
  return warp_sum(acc);
}

// Variant of F_UP that READS the four weight planes (9 plane touches instead of 6, but ~25 % fewer
// instructions per row): used when there are too few tasks to be bandwidth-bound, where the single-warp
// instruction latency of a strip, not HBM, sets the pace.
template <int R>
__device__ __forceinline__ double up_task_planes(const PcgArgs<float>& a, unsigned char* ringmem, int b, int strip, int lane,
                                          float beta, bool first, int pcur) {
  RingUP* ring = reinterpret_cast<RingUP*>(ringmem);
  const int H = a.H, W = a.W;
  const size_t base = (size_t)b * H * W;
  const int x0 = strip * kStrip4, x = x0 + 4 * lane;
  const bool valid = x < W;
  const int xs = valid ? x : 0;
  const float* __restrict__ S = a.wS + base; const float* __restrict__ E = a.wE + base;
  const float* __restrict__ SE = a.wSE + base; const float* __restrict__ SW = a.wSW + base;
  const uint32_t* __restrict__ IPC = a.ipc + base;
  // reads: h (z plane), p_old (pin); writes: p_new (pout), q.  Nothing is updated in place, because the
  // halo columns of this strip belong to a neighbouring strip whose warp runs asynchronously.
  const float* __restrict__ Hb = a.z + base;
  const float* __restrict__ Pin = (pcur ? a.p2 : a.p) + base;
  float* __restrict__ P = (pcur ? a.p : a.p2) + base;
  float* __restrict__ Q = a.q + base;
  // halo fetch role of this lane: k = which plane, left (lanes 0-4) or right (lanes 16-19) chunk
  const int k = lane & 15;
  const bool isL = lane < 16;
  const int hx = isL ? x0 - 4 : x0 + kStrip4;
  const bool hcol = isL ? (x0 > 0) : (x0 + kStrip4 < W);
  const bool hrole = hcol && (k <= 3 || (k == 4 && isL));
  const float* hsrc = (k == 0) ? Hb : (k == 1) ? reinterpret_cast<const float*>(IPC) : (k == 2) ? Pin
                      : (k == 3) ? (isL ? E : SW) : SE;
  const int hxs = hrole ? hx : 0;

  auto issue = [&](int y) {                       // stage row y (rows are visited bottom -> top)
    RingUP* r = ring + ((H - 2 - y) % R);
    const bool in = y >= 0;
    const bool interior = y >= 1;
    const size_t o = (size_t)(in ? y : 0) * W;
    cp_async16(&r->h[lane], Hb + o + xs, interior && valid);
    cp_async16(&r->ipc[lane], IPC + o + xs, interior && valid);
    cp_async16(&r->p[lane], Pin + o + xs, interior && valid && !first);
    cp_async16(&r->s[lane], S + o + xs, in && valid);
    cp_async16(&r->e[lane], E + o + xs, interior && valid);
    cp_async16(&r->se[lane], SE + o + xs, in && valid);
    cp_async16(&r->sw[lane], SW + o + xs, in && valid);
    cp_async16(&r->halo[lane], hsrc + o + hxs, in && hrole && (k >= 3 || interior) && !(k == 2 && first));
    cp_async_commit();
  };
  const unsigned full = 0xffffffffu;
  struct WRowP { float4 s, e, se, sw; float hw1, hw2; };
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float zn[4] = {0.f, 0.f, 0.f, 0.f}, zhn = 0.f;            // z of the row below (own columns, halo column)
  float pl[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, pc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, pu[6];
  WRowP wc, wu;
  wc.s = wc.e = wc.se = wc.sw = zero4; wc.hw1 = wc.hw2 = 0.f;
  const int hoff = (lane == 31) ? 16 : 0;
  CoarseZ cz;
  cz.init(a, b, x, x0, lane, valid, W);

#pragma unroll
  for (int u = 0; u < R; ++u) issue(H - 2 - u);
  double acc = 0.0;
#pragma unroll 2
  for (int y = H - 2; y >= 0; --y) {
    cp_async_wait<R - 1>();
    __syncwarp();          // the halo chunks were copied by OTHER lanes: their wait_group must have passed too
    {
      const RingUP* r = ring + ((H - 2 - y) % R);
      const float4 h = r->h[lane], p = r->p[lane];
      const uint4 f = r->ipc[lane];
      wu.s = r->s[lane]; wu.e = r->e[lane]; wu.se = r->se[lane]; wu.sw = r->sw[lane];
      // lane 0 uses the left chunks' .w, lane 31 the right chunks' .x; other lanes read but never use them
      const float hh = hsel(r->halo[hoff + 0], lane);
      const uint32_t hf = __float_as_uint(hsel(r->halo[hoff + 1], lane));
      const float hp = hsel(r->halo[hoff + 2], lane);
      wu.hw1 = hsel(r->halo[hoff + 3], lane); wu.hw2 = hsel(r->halo[hoff + 4], lane);
      float z[4];     // line part of z (the recurrence runs on it alone); the coarse part is added on top
      z[0] = fmaf(bf16_hi(f.x), zn[0], bf16_lo(f.x) * h.x); z[1] = fmaf(bf16_hi(f.y), zn[1], bf16_lo(f.y) * h.y);
      z[2] = fmaf(bf16_hi(f.z), zn[2], bf16_lo(f.z) * h.z); z[3] = fmaf(bf16_hi(f.w), zn[3], bf16_lo(f.w) * h.w);
      const float zh = fmaf(bf16_hi(hf), zhn, bf16_lo(hf) * hh);
      float zadd, zaddh;
      cz.at_row(y, H, zadd, zaddh);
      pu[1] = fmaf(beta, p.x, z[0] + zadd); pu[2] = fmaf(beta, p.y, z[1] + zadd);
      pu[3] = fmaf(beta, p.z, z[2] + zadd); pu[4] = fmaf(beta, p.w, z[3] + zadd);
      const float ph = fmaf(beta, hp, zh + zaddh);
      const float l = __shfl_up_sync(full, pu[4], 1), rt = __shfl_down_sync(full, pu[1], 1);
      pu[0] = (lane == 0) ? ph : l;
      pu[5] = (lane == 31) ? ph : rt;
      zn[0] = z[0]; zn[1] = z[1]; zn[2] = z[2]; zn[3] = z[3]; zhn = zh;
      if (valid && y >= 1) *reinterpret_cast<float4*>(P + (size_t)y * W + x) = make_float4(pu[1], pu[2], pu[3], pu[4]);
    }
    __syncwarp();          // every lane has read this slot's halo before any lane refills it (issue below)
    const int yc = y + 1;                          // row whose q is now computable (u = y, c = y+1, l = y+2)
    if (yc <= H - 2) {
      const float eL_s = __shfl_up_sync(full, wc.e.w, 1);
      const float seL_s = __shfl_up_sync(full, wu.se.w, 1);
      const float swR_s = __shfl_down_sync(full, wu.sw.x, 1);
      const float eL = (lane == 0) ? wc.hw1 : eL_s;           // wE[c, x-1]
      const float seL = (lane == 0) ? wu.hw2 : seL_s;         // wSE[u, x-1]
      const float swR = (lane == 31) ? wu.hw1 : swR_s;        // wSW[u, x+4]
      const float Sc[4] = {wc.s.x, wc.s.y, wc.s.z, wc.s.w};
      const float Su[4] = {wu.s.x, wu.s.y, wu.s.z, wu.s.w};
      const float E5[5] = {eL, wc.e.x, wc.e.y, wc.e.z, wc.e.w};
      const float SEc[4] = {wc.se.x, wc.se.y, wc.se.z, wc.se.w};
      const float SWc[4] = {wc.sw.x, wc.sw.y, wc.sw.z, wc.sw.w};
      const float SEu[4] = {seL, wu.se.x, wu.se.y, wu.se.z};
      const float SWu[4] = {wu.sw.y, wu.sw.z, wu.sw.w, swR};
      float q[4], dot = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pcj = pc[j + 1];
        float t = Sc[j] * (pcj - pl[j + 1]);
        t = fmaf(Su[j], pcj - pu[j + 1], t);
        t = fmaf(E5[j + 1], pcj - pc[j + 2], t);
        t = fmaf(E5[j], pcj - pc[j], t);
        t = fmaf(SEc[j], pcj - pl[j + 2], t);
        t = fmaf(SWc[j], pcj - pl[j], t);
        t = fmaf(SEu[j], pcj - pu[j], t);
        t = fmaf(SWu[j], pcj - pu[j + 2], t);
        q[j] = t;
        dot = fmaf(pcj, t, dot);
      }
      if (valid) {
        *reinterpret_cast<float4*>(Q + (size_t)yc * W + x) = make_float4(q[0], q[1], q[2], q[3]);
        acc += (double)dot;
      }
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) { pl[j] = pc[j]; pc[j] = pu[j]; }
    wc = wu;
    issue(y - R);
  }
  cp_async_wait<0>();
  return warp_sum(acc);
}

// F_DOWN / F_PRECOND.  KIND: 0 = PRECOND (fresh residual: h = L^-1 r, e = 0, nothing else updated),
//   1 = DOWN that defers the e-update (r -= alpha q; h):              reads r, q, (ip,c)      writes r, h
//   2 = DOWN that applies two e-updates (e += alpha_prev p_prev + alpha p; r -= alpha q; h):
//                                                                     reads p, p_prev, e, r, q, (ip,c)  writes e, r, h
// Alternating 1 / 2 halves the traffic of the e-update (the previous search direction is still intact
// in the other p plane), 13 instead of 14 plane touches per iteration on average.
template <int RD, int KIND>
__device__ __forceinline__ void down_task(const PcgArgs<float>& a, unsigned char* ringmem, int b, int strip, int lane,
                                          float alpha, float alpha_prev, int pcur, double& rr_out, double& rz_out) {
  RingD* ring = reinterpret_cast<RingD*>(ringmem);
  const int H = a.H, W = a.W;
  const size_t base = (size_t)b * H * W;
  const int x = strip * kStrip4 + 4 * lane;
  const bool valid = x < W;
  const int xs = valid ? x : 0;
  const uint32_t* __restrict__ IPC = a.ipc + base;
  const float* __restrict__ P = (pcur ? a.p2 : a.p) + base;        // p of this iteration (written by the last F_UP)
  const float* __restrict__ PP = (pcur ? a.p : a.p2) + base;       // p of the previous iteration
  float* __restrict__ Ev = a.e + base; float* __restrict__ Rv = a.r + base;
  const float* __restrict__ Q = a.q + base;
  float* __restrict__ Hb = a.z + base;          // h = L^-1 r goes to the z plane (read by the next F_UP)
  auto issue = [&](int y) {
    RingD* r = ring + ((y - 1) % RD);
    const bool in = (y <= H - 2) && valid;
    const size_t o = (size_t)(y <= H - 2 ? y : 0) * W + xs;
    if (KIND == 2) {
      cp_async16(&r->p[lane], P + o, in);
      cp_async16(&r->pp[lane], PP + o, in);
      cp_async16(&r->e[lane], Ev + o, in);
    }
    if (KIND != 0) cp_async16(&r->q[lane], Q + o, in);
    cp_async16(&r->r[lane], Rv + o, in);
    cp_async16(&r->ipc[lane], IPC + o, in);
    cp_async_commit();
  };
#pragma unroll
  for (int u = 0; u < RD; ++u) issue(1 + u);
  float hp[4] = {0.f, 0.f, 0.f, 0.f}, cp[4] = {0.f, 0.f, 0.f, 0.f};
  float racc = 0.f;
  double rr = 0.0, rz = 0.0;
#pragma unroll 2
  for (int y = 1; y <= H - 2; ++y) {
    cp_async_wait<RD - 1>();
    const RingD* r = ring + ((y - 1) % RD);
    const float4 rv = r->r[lane];
    const uint4 f = r->ipc[lane];
    const float4 cv = make_float4(bf16_hi(f.x), bf16_hi(f.y), bf16_hi(f.z), bf16_hi(f.w));
    const float4 iv = make_float4(bf16_lo(f.x), bf16_lo(f.y), bf16_lo(f.z), bf16_lo(f.w));
    float rn[4] = {rv.x, rv.y, rv.z, rv.w};
    if (KIND != 0) {
      const float4 qv = r->q[lane];
      rn[0] = fmaf(-alpha, qv.x, rv.x); rn[1] = fmaf(-alpha, qv.y, rv.y);
      rn[2] = fmaf(-alpha, qv.z, rv.z); rn[3] = fmaf(-alpha, qv.w, rv.w);
      if (valid) *reinterpret_cast<float4*>(Rv + (size_t)y * W + x) = make_float4(rn[0], rn[1], rn[2], rn[3]);
      if (KIND == 2) {
        const float4 pv = r->p[lane], pq = r->pp[lane], ev = r->e[lane];
        const float4 en = make_float4(fmaf(alpha, pv.x, fmaf(alpha_prev, pq.x, ev.x)), fmaf(alpha, pv.y, fmaf(alpha_prev, pq.y, ev.y)),
                                      fmaf(alpha, pv.z, fmaf(alpha_prev, pq.z, ev.z)), fmaf(alpha, pv.w, fmaf(alpha_prev, pq.w, ev.w)));
        if (valid) *reinterpret_cast<float4*>(Ev + (size_t)y * W + x) = en;
      }
    } else if (valid) {
      *reinterpret_cast<float4*>(Ev + (size_t)y * W + x) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = fmaf(cp[j], hp[j], rn[j]);
    if (valid) *reinterpret_cast<float4*>(Hb + (size_t)y * W + x) = make_float4(h[0], h[1], h[2], h[3]);
    const float srr = fmaf(rn[0], rn[0], fmaf(rn[1], rn[1], fmaf(rn[2], rn[2], rn[3] * rn[3])));
    const float srz = fmaf(iv.x * h[0], h[0], fmaf(iv.y * h[1], h[1], fmaf(iv.z * h[2], h[2], iv.w * h[3] * h[3])));
    rr += (double)srr; rz += (double)srz;
    if (a.cg.by) {                               // restriction rc = P^T r: sum of the new residual over each aggregate
      racc += (rn[0] + rn[1]) + (rn[2] + rn[3]);
      if ((y & (a.cg.by - 1)) == 0 || y == H - 2) {          // (y-1) is the last row of its aggregate row
        float v = racc;                                      // 8 (or 16) lanes share one aggregate column
        v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
        if (a.cg.bx_shift == 6) v += __shfl_xor_sync(0xffffffffu, v, 8);
        if ((lane & ((1 << (a.cg.bx_shift - 2)) - 1)) == 0 && valid)
          a.rc[(size_t)b * a.cg.nc + (size_t)((y - 1) >> a.cg.by_shift) * a.cg.nbx + (x >> a.cg.bx_shift)] = v;
        racc = 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) hp[j] = h[j];
    cp[0] = cv.x; cp[1] = cv.y; cp[2] = cv.z; cp[3] = cv.w;
    issue(y + RD);
  }
  cp_async_wait<0>();
  rr_out = warp_sum(rr); rz_out = warp_sum(rz);
}

// F_COARSE: one warp solves Ac zc = rc for one image and returns rc . zc.
// Two triangular sweeps in saxpy form with a ROTATING REGISTER WINDOW: lane l holds the vector element
// j == l (mod 32) of the 32-wide window that starts at the current row, so a step is
//   broadcast the head element (one shuffle)  ->  every lane: w -= F[i][d] * head  (one FMA)
// and the dependent chain per row is shuffle + FMA (~30 cycles) instead of a shared-memory round trip.
// Needs band width kd <= 31.  The factor rows are pre-scaled (F[i][d] = L[.]/L[i,i], F[i][0] = 1/L[i,i])
// and staged in shared memory by cp.async tiles; the vector lives in shared memory only as input/output.
__device__ __forceinline__ double coarse_task(const PcgArgs<float>& a, unsigned char* ringmem, int b, int lane) {
  const CoarseGeom g = a.cg;
  const int n = g.nc, kd1 = g.kd1;
  float* ysb = reinterpret_cast<float*>(ringmem);                // 32 zeros | vector [n] | zeros up to kCoarseMaxN + 64
  float* ys = ysb + 32;
  float* tiles = ysb + kCoarseMaxN + 96;                         // 2 x [32][32]
  const float* rc = a.rc + (size_t)b * n;
  float* zc = a.zc + (size_t)b * n;
  const float* U = a.Ur + (size_t)b * n * kd1;
  const float* L = a.Lr + (size_t)b * n * kd1;
  const unsigned full = 0xffffffffu;
  const int ntile = (n + kCoarseTileRows - 1) / kCoarseTileRows;
  auto issue_tile = [&](const float* F, int t, float* buf) {
    if (t >= 0 && t < ntile) {
      const int r0 = t * kCoarseTileRows, rows = min(kCoarseTileRows, n - r0);
      const int chunks = rows * kd1 / 4;                         // rows past n are zero-filled (src-size 0)
      for (int c = lane; c < kCoarseTileRows * 32 / 4; c += 32)
        cp_async16(buf + 4 * c, F + (size_t)r0 * kd1 + 4 * (c < chunks ? c : 0), c < chunks);
    }
    cp_async_commit();
  };
  ysb[lane] = 0.f;
  for (int i = lane; i < n + 64; i += 32) ys[i] = (i < n) ? __ldcg(rc + i) : 0.f;
  __syncwarp();
  // Both sweeps advance in blocks of 32 rows (= one factor tile).  Inside a block lane l admits exactly one
  // element (prefetched into `nx` before the block) and retires exactly one (kept in `out`, stored after the
  // block), so the inner step has no shared-memory store and its dependent chain is shuffle -> FMA -> select.
  // ---- forward: L y = rc, column-oriented (row i of Ur = scaled column i of L) -----------------------
  issue_tile(U, 0, tiles);
  float w = ys[lane];                                            // window [0, 32): lane l holds element l
  for (int t = 0; t < ntile; ++t) {
    issue_tile(U, t + 1, tiles + ((t + 1) & 1) * kCoarseTileRows * 32);
    const int i0 = t * 32;
    const float nx = ys[i0 + 32 + lane];                         // element this lane admits during the block
    float out = 0.f;
    cp_async_wait<1>();
    __syncwarp();
    const float* T = tiles + (t & 1) * kCoarseTileRows * 32;
#pragma unroll 8
    for (int hl = 0; hl < 32; ++hl) {                            // row i = i0 + hl; its element sits on lane hl
      const float* row = T + hl * 32;
      const int d = (lane - hl) & 31;                            // this lane holds element i + d
      const float f = row[d];                                    // zero padded beyond the band
      const float r0 = row[0];
      const float head = __shfl_sync(full, w, hl);               // element i, fully updated
      const float wn = fmaf(-f, head, w);
      w = (d == 0) ? nx : wn;                                    // retire element i, admit element i + 32
      out = (d == 0) ? head * r0 : out;                          // y_i
    }
    if (i0 + lane < n) ys[i0 + lane] = out;
    __syncwarp();
  }
  cp_async_wait<0>();
  __syncwarp();
  // ---- backward: L^T x = y, row-oriented (row i of Lr) ----------------------------------------------------
  issue_tile(L, ntile - 1, tiles);
  {
    const int top = ntile * 32 - 1;                              // window (top-32, top]: lane l holds element j == l (mod 32)
    const int j = top - ((top - lane) & 31);
    w = ys[j];                                                   // ys is zero padded past n
  }
  for (int k = 0; k < ntile; ++k) {
    const int t = ntile - 1 - k;
    issue_tile(L, t - 1, tiles + ((k + 1) & 1) * kCoarseTileRows * 32);
    const int i0 = t * 32;
    const float nx = ysb[i0 + lane];                             // = ys[i0 - 32 + lane]: element admitted during the block
    float out = 0.f;
    cp_async_wait<1>();
    __syncwarp();
    const float* T = tiles + (k & 1) * kCoarseTileRows * 32;
#pragma unroll 8
    for (int hl = 31; hl >= 0; --hl) {                           // row i = i0 + hl
      const float* row = T + hl * 32;
      const int d = (hl - lane) & 31;                            // this lane holds element i - d
      const float f = row[d];
      const float r0 = row[0];
      const float head = __shfl_sync(full, w, hl);               // element i with all later rows eliminated
      const float xi = head * r0;                                // x_i = head / L[i,i]   (0 for padding rows: r0 = 0)
      const float wn = fmaf(-f, xi, w);
      w = (d == 0) ? nx : wn;                                    // retire element i, admit element i - 32
      out = (d == 0) ? xi : out;
    }
    if (i0 + lane < n) ys[i0 + lane] = out;
    __syncwarp();
  }
  cp_async_wait<0>();
  __syncwarp();
  double acc = 0.0;
  for (int i = lane; i < n; i += 32) {
    const float v = ys[i];
    zc[i] = v;
    acc += (double)v * (double)__ldcg(rc + i);
  }
  __syncwarp();
  return warp_sum(acc);
}

// ---- ready queue (multi-producer / multi-consumer ticket ring in global memory) -------------------
// Only the strips of ONE image have to meet between phases, so the fused kernel has no grid-wide
// barrier: the last-arriving strip of an image publishes the image's next phase as nstrip4 new
// tasks, and any idle warp pops the next task.  Work-conserving: no SM waits for a slower SM.
constexpr unsigned kPoison = 0xffffffffu;

// pushes n entries: payload0, payload0+stride, ...  (one lane; one reservation, one release fence)
__device__ __forceinline__ void q_push_n(const PcgArgs<float>& a, unsigned payload0, unsigned stride, unsigned n) {
  const unsigned long long k0 = atomicAdd(a.qtail, (unsigned long long)n);
  __threadfence();                                          // release: state written before the stamps
  for (unsigned i = 0; i < n; ++i) {
    const unsigned long long k = k0 + i;
    volatile unsigned long long* slot = a.qslots + (k % a.qcap);
    *slot = ((k + 1ull) << 32) | (unsigned long long)(payload0 + i * stride);
  }
}

__device__ __forceinline__ unsigned q_pop(const PcgArgs<float>& a, int lane) {        // whole warp
  // lane 0 takes the ticket; then ALL lanes poll the slot together (a warp with 31 lanes parked at a
  // convergence barrier while one lane spins was measured to slow the other warps of the SM ~4x)
  unsigned long long k = 0;
  if (lane == 0) k = atomicAdd(a.qhead, 1ull);
  k = __shfl_sync(0xffffffffu, k, 0);
  volatile unsigned long long* slot = a.qslots + (k % a.qcap);
  const unsigned stamp = (unsigned)(k + 1ull);
  unsigned long long v;
  unsigned ns = 32;
  while ((unsigned)((v = *slot) >> 32) != stamp) { __nanosleep(ns); if (ns < 512) ns *= 2; }
  __threadfence();                                          // acquire
  return (unsigned)(v & 0xffffffffull);
}

// slots 0..T-1 hold the first phase (RESID) of every image, image index fastest
__global__ void cm_init_queue_kernel(unsigned long long* __restrict__ slots, unsigned cap, int B, int nstrip4,
                                     unsigned long long* __restrict__ head, unsigned long long* __restrict__ tail) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned T = (unsigned)(B * nstrip4);
  if (i == 0) { *head = 0ull; *tail = (unsigned long long)T; }
  if (i >= cap) return;
  unsigned long long v = 0ull;
  if (i < T) {
    const unsigned b = i % (unsigned)B, ti = i / (unsigned)B;
    v = ((unsigned long long)(i + 1u) << 32) | (unsigned long long)(b * (unsigned)nstrip4 + ti);
  }
  slots[i] = v;
}

// fp64 residual r = b - A x64 over one 128-column strip (rare: once per refinement round)
__device__ __forceinline__ double resid_strip(const PcgArgs<float>& a, int b, int strip, int lane) {
  const int x0 = strip * kStrip4, xe = min(x0 + kStrip4, a.W);
  double s = 0.0;
  for (int xb = x0; xb < xe; xb += kStripCols)
    for (int rbk = 0; rbk < a.nrb; ++rbk)
      s += stencil_task<float, true>(a, b, xb, min(kStripCols, xe - xb), rbk, lane, 0.f, true);
  return s;
}

// AW = active warps per CTA (8 or 4): fewer active warps own deeper rings.
// RECOMP = F_UP recomputes the edge weights from the `a` plane (bandwidth-bound regime) or reads them.
template <int AW, bool RECOMP>
__global__ void __launch_bounds__(256, 1) pcg_fused_kernel(const PcgArgs<float> a) {
  constexpr size_t kWarpRing = kFusedSmem / AW;                   // ring bytes per active warp
  constexpr int R = (int)(kWarpRing / (RECOMP ? sizeof(RingU) : sizeof(RingUP)));   // F_UP ring depth in rows
  constexpr int RD = (int)(kWarpRing / sizeof(RingD));                              // F_DOWN ring depth
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= AW || warp >= a.wpb) return;                        // no block-wide barrier anywhere below
  if (a.debug == 2 && (blockIdx.x != 0 || warp != 0)) return;     // debug: a single worker runs every task
  unsigned char* ringmem = smem_raw + (size_t)warp * kWarpRing;
  const int B = a.B, H = a.H, W = a.W;
  const int per = a.nstrip4;
  const unsigned nworkers = (a.debug == 2) ? 1u : gridDim.x * (unsigned)a.wpb;

  auto now = []() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
  unsigned long long t_wait = 0, t_mode[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t_arr = 0;
  unsigned n_mode[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (;;) {
    const unsigned long long tq0 = a.debug ? now() : 0ull;
    const unsigned payload = q_pop(a, lane);
    const unsigned long long tq1 = a.debug ? now() : 0ull;
    t_wait += tq1 - tq0;
    if (payload == kPoison) break;
    const int b = (int)(payload / (unsigned)per), ti = (int)(payload % (unsigned)per);
    ImgState* st = a.st + b;
    const int mode = (int)(__ldcg(&st->fword) & 0xffffffffull);
    double s0 = 0.0, s1 = 0.0;
    if (mode == F_UP) {
      if constexpr (RECOMP)
        s0 = up_task<R>(a, ringmem, b, ti, lane, (float)__ldcg(&st->beta), __ldcg(&st->first) != 0, __ldcg(&st->pcur));
      else
        s0 = up_task_planes<R>(a, ringmem, b, ti, lane, (float)__ldcg(&st->beta), __ldcg(&st->first) != 0, __ldcg(&st->pcur));
    } else if (mode == F_DOWN) {
      if (__ldcg(&st->epend))
        down_task<RD, 2>(a, ringmem, b, ti, lane, (float)__ldcg(&st->alpha), (float)__ldcg(&st->alpha_prev), __ldcg(&st->pcur), s0, s1);
      else
        down_task<RD, 1>(a, ringmem, b, ti, lane, (float)__ldcg(&st->alpha), 0.f, __ldcg(&st->pcur), s0, s1);
    } else if (mode == F_PRECOND) {
      down_task<RD, 0>(a, ringmem, b, ti, lane, 0.f, 0.f, 0, s0, s1);
    } else if (mode == F_RESID) {
      s0 = resid_strip(a, b, ti, lane);
    } else if (mode == F_COARSE) {
      s0 = coarse_task(a, ringmem, b, lane);
    } else {
      const int x = ti * kStrip4 + 4 * lane;
      if (x < W) {
        const size_t base = (size_t)b * H * W + x;
        if (mode == F_FOLD) {
          // x64 += e (+ the e-update of the last iteration if it is still pending)
          const float ap = __ldcg(&st->epend) ? (float)__ldcg(&st->alpha_prev) : 0.f;
          const float* Pl = (__ldcg(&st->pcur) ? a.p2 : a.p);
#pragma unroll 4
          for (int y = 1; y <= H - 2; ++y) {
            const size_t o = base + (size_t)y * W;
            float4 e = __ldcg(reinterpret_cast<const float4*>(a.e + o));
            if (ap != 0.f) {
              const float4 pv = __ldcg(reinterpret_cast<const float4*>(Pl + o));
              e.x = fmaf(ap, pv.x, e.x); e.y = fmaf(ap, pv.y, e.y); e.z = fmaf(ap, pv.z, e.z); e.w = fmaf(ap, pv.w, e.w);
            }
            double2* xp = reinterpret_cast<double2*>(a.x64 + o);
            double2 v0 = __ldcg(xp), v1 = __ldcg(xp + 1);
            v0.x += (double)e.x; v0.y += (double)e.y; v1.x += (double)e.z; v1.y += (double)e.w;
            xp[0] = v0; xp[1] = v1;
          }
        } else {   // F_FINAL
#pragma unroll 4
          for (int y = 0; y < H; ++y) {
            const size_t o = base + (size_t)y * W;
            const double2* xp = reinterpret_cast<const double2*>(a.x64 + o);
            const double2 v0 = __ldcg(xp), v1 = __ldcg(xp + 1);
            *reinterpret_cast<float4*>(a.cm + o) = make_float4((float)v0.x, (float)v0.y, (float)v1.x, (float)v1.y);
          }
        }
      }
    }
    const unsigned long long tq2 = a.debug ? now() : 0ull;
    if (a.debug) { t_mode[mode] += tq2 - tq1; n_mode[mode]++; }
    // ---- arrival: every lane publishes its own stores, then lane 0 takes a ticket -------------------
    double* part = a.part + (size_t)b * a.pmax * 2;
    if (lane == 0) { part[2 * ti] = s0; part[2 * ti + 1] = s1; }
    __threadfence();
    __syncwarp();
    unsigned tk = 0;
    if (lane == 0) tk = atomicAdd(&st->cnt1, 1u);
    tk = __shfl_sync(0xffffffffu, tk, 0);
    if (a.debug) t_arr += now() - tq2;
    const int ntask = (mode == F_COARSE) ? 1 : per;               // F_COARSE is a single task per image
    if (tk != (unsigned)ntask - 1) continue;
    // ---- last strip of this image and phase: reduce, advance the state, publish the next phase ------
    __threadfence();
    reduce_parts(part, ntask, lane, s0, s1);
    if (lane != 0) continue;
    st->cnt1 = 0;
    int next = F_DONE;
    if (mode == F_UP) {
      if (s0 > 0.0) st->alpha = __ldcg(&st->rz) / s0;
      else { st->alpha = 0.0; st->rr_target = 1e300; }          // breakdown: leave the inner solve
      st->pcur = __ldcg(&st->pcur) ^ 1;                           // F_UP wrote the other p plane
      next = F_DOWN;
    } else if (mode == F_DOWN) {
      const double rz_old = __ldcg(&st->rz);
      const int it = __ldcg(&st->iters) + 1;
      if (__ldcg(&st->epend)) st->epend = 0;                       // this DOWN applied both pending updates
      else { st->epend = 1; st->alpha_prev = __ldcg(&st->alpha); }  // this DOWN deferred its own update
      st->iters = it; st->rr_rec = s0;
      if (s0 <= __ldcg(&st->rr_target) || it >= a.maxiter || !(s1 > 0.0)) {
        next = F_FOLD;
      } else if (a.cg.by) {                                        // r.z is completed by the coarse solve
        st->rz_line = s1; st->fresh = 0;
        next = F_COARSE;
      } else {
        st->rz = s1; st->first = 0;
        st->beta = (rz_old > 0.0) ? s1 / rz_old : 0.0;
        next = F_UP;
      }
    } else if (mode == F_PRECOND) {
      // First round: start CG (p = z).  USAUG_CM_CONTINUE (experimental, rejected: more iterations): keep p.
      const double rz_old = __ldcg(&st->rz);
      const bool cont = __ldcg(&st->outer) > 0 && rz_old > 0.0 && a.restart;
      st->epend = 0;                                               // F_FOLD consumed any pending update, e = 0 now
      if (a.cg.by) {
        st->rz_line = s1; st->fresh = 1;
        next = (s1 > 0.0) ? F_COARSE : F_FINAL;
      } else {
        st->rz = s1; st->first = cont ? 0 : 1; st->beta = cont ? s1 / rz_old : 0.0;
        next = (s1 > 0.0) ? F_UP : F_FINAL;
      }
    } else if (mode == F_COARSE) {
      // z = T^-1 r + P Ac^-1 P^T r  =>  r.z = sum ip h^2 (line part, from DOWN) + rc.zc (this task)
      const double rz_new = __ldcg(&st->rz_line) + s0, rz_old = __ldcg(&st->rz);
      if (__ldcg(&st->fresh)) { st->rz = rz_new; st->first = 1; st->beta = 0.0; next = F_UP; }
      else if (rz_new > 0.0 && rz_old > 0.0) { st->beta = rz_new / rz_old; st->rz = rz_new; st->first = 0; next = F_UP; }
      else next = F_FOLD;
    } else if (mode == F_RESID) {
      next = after_true_residual(st, s0, a.rtol2, a.inner_red2, a.maxiter) ? F_FINAL : F_PRECOND;
    } else if (mode == F_FOLD) {
      next = F_RESID; st->outer = __ldcg(&st->outer) + 1;
    } else {   // F_FINAL
      const double bb = __ldcg(&st->bb);
      if (a.iters_out) a.iters_out[b] = __ldcg(&st->iters);
      if (a.relres_out) a.relres_out[b] = (bb > 0.0) ? sqrt(__ldcg(&st->rr_true) / bb) : 0.0;
      next = F_DONE;
    }
    st->fword = (unsigned long long)next;
    if (next == F_DONE) {
      __threadfence();
      if (atomicAdd(a.done, 1u) == (unsigned)B - 1u) q_push_n(a, kPoison, 0u, nworkers);   // release every worker
    } else {
      q_push_n(a, (unsigned)(b * per), 1u, (next == F_COARSE) ? 1u : (unsigned)per);
    }
  }
  if ((a.debug & 1) && lane == 0 && warp == 0 && (blockIdx.x % 74) == 0)
    printf("[pcg queue dbg] block %d: wait %.1f us | UP %u x %.1f us | DOWN %u x %.1f us | PRECOND %u x %.1f | RESID %u x %.1f | COARSE %u x %.1f | arrive %.1f us total\n",
           blockIdx.x, t_wait * 1e-3, n_mode[F_UP], n_mode[F_UP] ? t_mode[F_UP] * 1e-3 / n_mode[F_UP] : 0.0, n_mode[F_DOWN],
           n_mode[F_DOWN] ? t_mode[F_DOWN] * 1e-3 / n_mode[F_DOWN] : 0.0, n_mode[F_PRECOND],
           n_mode[F_PRECOND] ? t_mode[F_PRECOND] * 1e-3 / n_mode[F_PRECOND] : 0.0, n_mode[F_RESID],
           n_mode[F_RESID] ? t_mode[F_RESID] * 1e-3 / n_mode[F_RESID] : 0.0, n_mode[F_COARSE],
           n_mode[F_COARSE] ? t_mode[F_COARSE] * 1e-3 / n_mode[F_COARSE] : 0.0, t_arr * 1e-3);
}

struct CmLayout {
  float *depth, *mm1, *mm2, *wS, *wE, *wSE, *wSW, *ip, *c, *aplane;
  uint32_t* ipc;
  float4* wk;
  double *x64, *bbpart, *part;
  void *e, *r, *p, *q, *z, *p2;
  ImgState* st;
  unsigned* done;
  unsigned long long *qslots, *qhead, *qtail;
  unsigned qcap;
  CoarseGeom cg;
  float *rc, *zc, *Lr, *Ur;
  double* Ab;
  int nblk, nstrip1, nrb, nstrip2, pmax;
  size_t bytes;
};

static CmLayout cm_layout(void* ws, int B, int H, int W, int flags) {
  CmLayout L;
  const size_t N = (size_t)B * H * W;
  const size_t tsz = (flags & USAUG_CM_FP64_VECTORS) ? 8 : 4;
  L.nblk = (W + 127) / 128;
  L.nstrip1 = (W + kStripCols - 1) / kStripCols;
  L.nrb = (H - 2 + kRowBlock - 1) / kRowBlock;
  L.nstrip2 = (W + 31) / 32;
  L.pmax = L.nstrip1 * L.nrb > L.nstrip2 ? L.nstrip1 * L.nrb : L.nstrip2;
  Carver cv(ws);
  L.depth = cv.take<float>(H);
  L.mm1 = cv.take<float>((size_t)B * kChunks * 2);
  L.mm2 = cv.take<float>((size_t)B * kChunks * 2);
  L.wS = cv.take<float>(N); L.wE = cv.take<float>(N); L.wSE = cv.take<float>(N); L.wSW = cv.take<float>(N);
  L.ip = cv.take<float>(N); L.c = cv.take<float>(N); L.ipc = cv.take<uint32_t>(N);
  L.aplane = cv.take<float>(N); L.wk = cv.take<float4>(B);
  L.x64 = cv.take<double>(N);
  L.e = cv.take<char>(N * tsz); L.r = cv.take<char>(N * tsz); L.p = cv.take<char>(N * tsz);
  L.q = cv.take<char>(N * tsz); L.z = cv.take<char>(N * tsz); L.p2 = cv.take<char>(N * tsz);
  L.bbpart = cv.take<double>((size_t)B * L.nblk);
  L.part = cv.take<double>((size_t)B * L.pmax * 2);
  L.st = cv.take<ImgState>(B);
  L.done = cv.take<unsigned>(1);
  // ready queue: every image has at most nstrip4 ready tasks at a time; plus one poison per worker warp
  L.qcap = (unsigned)(2 * ((W + kStrip4 - 1) / kStrip4) * B + 8192);
  L.qslots = cv.take<unsigned long long>(L.qcap);
  L.qhead = cv.take<unsigned long long>(1); L.qtail = cv.take<unsigned long long>(1);
  // coarse-grid correction (fused fp32 kernel only)
  L.cg = ((flags & (USAUG_CM_FP64_VECTORS | USAUG_CM_NO_COARSE)) || W % 4 != 0) ? CoarseGeom{} : coarse_geom(H, W);
  const size_t nc = (size_t)L.cg.nc;
  L.rc = cv.take<float>((size_t)B * nc); L.zc = cv.take<float>((size_t)B * nc);
  L.Lr = cv.take<float>((size_t)B * nc * L.cg.kd1); L.Ur = cv.take<float>((size_t)B * nc * L.cg.kd1);
  L.Ab = cv.take<double>((size_t)B * nc * (L.cg.kd + 1));
  L.bytes = cv.off;
  return L;
}

// Optional measurement hook: a (start, stop) cudaEvent_t pair recorded on the launching stream
// immediately around the persistent kernel (bench.py's roofline uses it).  Thread-local.
static thread_local cudaEvent_t g_ev_start = nullptr, g_ev_stop = nullptr;
static thread_local int g_debug = 0;

static thread_local int g_restart = 0;
static thread_local int g_legacy = 0;   // USAUG_CM_LEGACY_KERNEL: force the two-slot kernel (A/B testing)

template <int AW, bool RECOMP>
static int launch_fused_aw(PcgArgs<float>& a, int blocks_cap_sms, cudaStream_t stream) {
  int per_sm = 0;
  USAUG_CUDA(cudaFuncSetAttribute(pcg_fused_kernel<AW, RECOMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem));
  USAUG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_fused_kernel<AW, RECOMP>, 256, kFusedSmem));
  if (per_sm < 1) { set_error("pcg_fused_kernel cannot be resident"); return USAUG_E_CUDA; }
  const int tasks = a.nstrip4 * a.B;
  int blocks = tasks;
  if (blocks > blocks_cap_sms) blocks = blocks_cap_sms;
  if (blocks < 1) blocks = 1;
  // no more workers than tasks that can be ready at once: an idle worker only polls the queue
  a.wpb = (tasks + blocks - 1) / blocks;
  if (a.wpb > AW) a.wpb = AW;
  if (getenv("USAUG_DEBUG_WPB")) a.wpb = atoi(getenv("USAUG_DEBUG_WPB"));
  cm_init_queue_kernel<<<(a.qcap + 255) / 256, 256, 0, stream>>>(a.qslots, a.qcap, a.B, a.nstrip4, a.qhead, a.qtail);
  USAUG_CHECK_LAUNCH("cm_init_queue_kernel");
  void* params[] = {(void*)&a};
  if (g_ev_start) USAUG_CUDA(cudaEventRecord(g_ev_start, stream));
  // cooperative launch only for its co-residency guarantee: workers spin on the ready queue
  USAUG_CUDA(cudaLaunchCooperativeKernel((const void*)pcg_fused_kernel<AW, RECOMP>, dim3(blocks), dim3(256), params,
                                         kFusedSmem, stream));
  if (g_ev_stop) USAUG_CUDA(cudaEventRecord(g_ev_stop, stream));
  return USAUG_OK;
}

static int launch_fused(PcgArgs<float>& a, const CmLayout& L, int B, cudaStream_t stream) {
  (void)L;
  int dev = 0, sms = 0;
  USAUG_CUDA(cudaGetDevice(&dev));
  USAUG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // every inner iteration is two phases; a refinement round adds four
  a.max_steps = 6 * a.maxiter + 64;   // worst case: a refinement round (5 phases) after every iteration
  const int tasks = a.nstrip4 * B;
  // few tasks: 4 active warps per CTA with rings twice as deep keep enough bytes in flight
  if (const char* v = getenv("USAUG_DEBUG_VARIANT")) {          // development: force a kernel instance
    if (v[0] == 'r') return launch_fused_aw<8, true>(a, sms, stream);
    if (v[0] == 'p') return launch_fused_aw<4, false>(a, sms, stream);
  }
  if (tasks >= 6 * sms) return launch_fused_aw<8, true>(a, sms, stream);
  return launch_fused_aw<4, false>(a, sms, stream);
}

template <typename T>
static int launch_pcg(const CmLayout& L, float* cm, int B, int H, int W, double rtol, int maxiter,
                      int* iters, double* relres, cudaStream_t stream) {
  PcgArgs<T> a;
  a.B = B; a.H = H; a.W = W; a.maxiter = maxiter;
  a.rtol2 = rtol * rtol;
  const double inner_red = std::is_same<T, float>::value ? (g_restart ? 1e-2 : 1e-4) : 0.0;
  a.inner_red2 = inner_red * inner_red;
  a.wS = L.wS; a.wE = L.wE; a.wSE = L.wSE; a.wSW = L.wSW; a.ip = L.ip; a.c = L.c; a.ipc = L.ipc; a.aplane = L.aplane; a.wk = L.wk;
  a.x64 = L.x64;
  a.e = (T*)L.e; a.r = (T*)L.r; a.p = (T*)L.p; a.q = (T*)L.q; a.z = (T*)L.z; a.p2 = (T*)L.p2;
  a.st = L.st; a.part = L.part; a.done = L.done;
  a.qslots = L.qslots; a.qhead = L.qhead; a.qtail = L.qtail; a.qcap = L.qcap;
  a.cg = L.cg; a.rc = L.rc; a.zc = L.zc; a.Lr = L.Lr; a.Ur = L.Ur; a.cm = cm; a.iters_out = iters; a.relres_out = relres;
  a.nstrip1 = L.nstrip1; a.nrb = L.nrb; a.nstrip2 = L.nstrip2; a.pmax = L.pmax;
  a.vec4 = (std::is_same<T, float>::value && W % 4 == 0) ? 1 : 0;
  a.nstrip4 = (W + kStrip4 - 1) / kStrip4;
  // every inner iteration is one step; each refinement round adds <= 3 steps and needs >= 1 iteration
  a.max_steps = 4 * maxiter + 16;
  a.debug = g_debug;
  a.restart = g_restart;

  if constexpr (std::is_same<T, float>::value) {
    if (a.vec4 && !g_legacy) return launch_fused(a, L, B, stream);
  }
  int dev = 0, sms = 0, per_sm = 0;
  const size_t smem = a.vec4 ? kRingBytesPerWarp * 8 : 0;
  USAUG_CUDA(cudaGetDevice(&dev));
  USAUG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (smem > 48 * 1024)
    USAUG_CUDA(cudaFuncSetAttribute(pcg_persistent_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  USAUG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_persistent_kernel<T>, 256, smem));
  if (per_sm < 1) { set_error("pcg_persistent_kernel cannot be resident"); return USAUG_E_CUDA; }
  // enough warps for every column task, never more blocks than can be co-resident
  const int tasks1 = (a.vec4 ? a.nstrip4 : L.nstrip1) * L.nrb * B;
  const int want_warps = L.nstrip2 * B > tasks1 ? L.nstrip2 * B : tasks1;
  int blocks = (want_warps + 7) / 8;
  const int cap = sms * (per_sm > 4 ? 4 : per_sm);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  void* params[] = {(void*)&a};
  if (g_ev_start) USAUG_CUDA(cudaEventRecord(g_ev_start, stream));
  USAUG_CUDA(cudaLaunchCooperativeKernel((const void*)pcg_persistent_kernel<T>, dim3(blocks), dim3(256),
                                         params, smem, stream));
  if (g_ev_stop) USAUG_CUDA(cudaEventRecord(g_ev_stop, stream));
  return USAUG_OK;
}

}  // namespace usaug

using namespace usaug;

extern "C" void usaug_confmap_set_timing_events(void* start_event, void* stop_event) {
  g_ev_start = static_cast<cudaEvent_t>(start_event);
  g_ev_stop = static_cast<cudaEvent_t>(stop_event);
}

extern "C" size_t usaug_confmap_workspace_bytes(int B, int H, int W, int precond, int flags) {
  (void)precond;
  if (B <= 0 || H < 3 || W <= 0) return 0;
  return cm_layout(nullptr, B, H, W, flags).bytes;
}

extern "C" int usaug_confmap_pcg(const float* img, float* cm, int B, int H, int W, float alpha,
                                 float beta, float gamma, double rtol, int maxiter, int precond,
                                 int flags, int* iters, double* relres, void* ws, size_t ws_bytes,
                                 void* stream_) {
  if (!img || !cm || !ws) { set_error("usaug_confmap_pcg: null pointer"); return USAUG_E_INVALID_ARG; }
  if (B <= 0 || H < 3 || W <= 0 || (long long)H * W > (1ll << 30)) {
    set_error("usaug_confmap_pcg: bad shape B=%d H=%d W=%d (need H>=3)", B, H, W); return USAUG_E_INVALID_ARG;
  }
  if (!(rtol > 0.0) || maxiter < 0 || (precond != USAUG_PRECOND_JACOBI && precond != USAUG_PRECOND_LINE)) {
    set_error("usaug_confmap_pcg: bad solver options rtol=%g maxiter=%d precond=%d", rtol, maxiter, precond);
    return USAUG_E_INVALID_ARG;
  }
  const size_t need = usaug_confmap_workspace_bytes(B, H, W, precond, flags);
  if (ws_bytes < need) { set_error("usaug_confmap_pcg: workspace %zu < %zu", ws_bytes, need); return USAUG_E_WORKSPACE_TOO_SMALL; }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CmLayout L = cm_layout(ws, B, H, W, flags);
  const int N = H * W;
  g_debug = (flags & USAUG_CM_DEBUG_TIMING) ? 1 : 0;
  if (getenv("USAUG_DEBUG_SINGLE")) g_debug = 2;   // development: one worker warp runs every task
  g_legacy = (flags & USAUG_CM_LEGACY_KERNEL) ? 1 : 0;
  g_restart = (flags & USAUG_CM_CONTINUE) ? 1 : 0;

  cm_depth_kernel<<<(H + 127) / 128, 128, 0, stream>>>(L.depth, H, alpha);
  USAUG_CHECK_LAUNCH("cm_depth_kernel");
  cm_minmax_kernel<<<dim3(kChunks, B), 256, 0, stream>>>(img, N, L.mm1);
  USAUG_CHECK_LAUNCH("cm_minmax_kernel");
  cm_delta_kernel<<<dim3(kChunks, B), 256, 0, stream>>>(img, H, W, L.depth, L.mm1, L.mm2);
  USAUG_CHECK_LAUNCH("cm_delta_kernel");
  cm_wconst_kernel<<<(B + 127) / 128, 128, 0, stream>>>(B, L.mm2, beta, gamma, L.wk);
  USAUG_CHECK_LAUNCH("cm_wconst_kernel");
  cm_weights_kernel<<<dim3((N + 255) / 256, B), 256, 0, stream>>>(img, H, W, L.depth, L.mm1, L.wk, L.aplane,
                                                                 L.wS, L.wE, L.wSE, L.wSW);
  USAUG_CHECK_LAUNCH("cm_weights_kernel");
  cm_factor_kernel<<<dim3(L.nblk, B), 128, 0, stream>>>(H, W, L.wS, L.wE, L.wSE, L.wSW, precond,
                                                        (flags & USAUG_CM_ZERO_GUESS) ? 1 : 0, L.ip, L.c,
                                                        L.ipc, L.x64, L.bbpart);
  USAUG_CHECK_LAUNCH("cm_factor_kernel");
  if (L.cg.by && !(flags & USAUG_CM_LEGACY_KERNEL)) {
    coarse_assemble_kernel<<<dim3((L.cg.nc + 7) / 8, B), 256, 0, stream>>>(H, W, L.cg, L.wS, L.wE, L.wSE, L.wSW, L.Ab);
    USAUG_CHECK_LAUNCH("coarse_assemble_kernel");
    const size_t fsm = (size_t)(L.cg.kd + 1) * (L.cg.kd + 1) * sizeof(double);
    coarse_factor_kernel<<<B, 32, fsm, stream>>>(L.cg, L.Ab, L.Lr, L.Ur);
    USAUG_CHECK_LAUNCH("coarse_factor_kernel");
  }
  cm_init_state_kernel<<<(B + 127) / 128, 128, 0, stream>>>(B, L.nblk, L.bbpart, L.st, L.done);
  USAUG_CHECK_LAUNCH("cm_init_state_kernel");
  if (flags & USAUG_CM_FP64_VECTORS)
    return launch_pcg<double>(L, cm, B, H, W, rtol, maxiter, iters, relres, stream);
  return launch_pcg<float>(L, cm, B, H, W, rtol, maxiter, iters, relres, stream);
}
