// Shared types and small device helpers of the confidence-map solver (included by confmap_pcg.cu).
#pragma once
#include <cooperative_groups.h>

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
  // fused kernel: the phase (FMode) the image's queued tasks belong to; written by the last arriver of the
  // previous phase before it publishes the tasks
  unsigned long long fword;
  int pcur, poor;     // pcur: which p plane holds the current search direction; poor: consecutive weak refinement rounds
  double inner2;      // squared residual reduction asked of the next inner solve (adapts to the fp32 drift floor)
  double alpha_prev;  // fused kernel: step length of the iteration whose e-update is still pending
  int epend, fresh;   // epend: 1 = e += alpha_prev * p_prev is deferred to the next DOWN / FOLD; fresh: COARSE follows a PRECOND
  double rz_line;     // line part of r.z of the current iteration (the COARSE task adds rc.zc)
  int prio, counted;  // fused kernel: ready-queue class of this image's tasks (1 = served first); counted in qsum/qcnt
  float rem;          // predicted remaining iterations (the contribution of this image to qsum)
  unsigned zprog;     // fused kernel: outward steps of the running coarse solve whose rows of zc are final (F_UP waits on it)
};

constexpr int kStrip4 = 128;

__device__ __forceinline__ uint32_t bf16_bits(float v) {      // round to nearest even, finite inputs
  const uint32_t u = __float_as_uint(v);
  return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

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

// ---- kernel arguments --------------------------------------------------------------------------
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
  unsigned* done;    // [0] images finished, [1] set when the idle watchdog gave up
  float* cm;
  int* iters_out;
  double* relres_out;
  int nstrip1, nrb, nstrip2, pmax, max_steps;
  // fused kernel: two ready rings of (image, strip) tasks, [0] = normal, [1] = served first
  unsigned long long *qslots, *qhead, *qtail;   // qslots[2][qcap], qhead[2], qtail[2]
  unsigned qcap;
  int* qavail;       // [2] entries written and not yet claimed
  double* qsum;      // sum over the running images of their predicted remaining iterations
  unsigned* qcnt;    // number of images in qsum
  int wpb;           // fused kernel: worker warps per block actually used (<= AW)
  unsigned long long watchdog_ns;   // fused kernel: idle workers give up when no task was claimed by anybody for this long
  CoarseGeom cg;     // coarse-grid correction geometry (cg.by == 0: disabled)
  float *rc, *zc;    // [B][nc] restricted residual / coarse correction
  const float* Wt;   // [B][nby][rec] block LDL^T of Ac: per aggregate row the dense inverse W_I and the couplings E_I
  int vec4, nstrip4; // vectorised stencil (fp32 vectors, W % 4 == 0): 128-column strips
  // fused kernel, twisted line factorisation: tw_m = 0 (off) or the row m where the two elimination fronts meet;
  // per = tasks per image and phase = nstrip4 * (tw_m ? 2 : 1)
  int tw_m, per;
  const float *wS2, *wSE2, *wSW2;   // mirrored weight planes (edges towards the row above), read by the lower half
  // level preconditioner (fused kernel): column-line solve on the aggregates of 1 row x 4 columns, added to z.
  // h4 = nullptr: disabled.  Planes [B][H][w4p], w4p = W/4 padded to a multiple of 4.
  float* h4;               // forward-eliminated level residual (written by DOWN, read by UP)
  const uint32_t* ipc4;    // level line factors, bf16 ip | bf16 c << 16 like ipc
  int w4p;
  // extra sink pixels (potential 0) among rows 1..H-2: [B][H][W] bytes, nullptr = bottom row only.  A sink keeps its
  // place in every vector as a decoupled row: x = p = q = r = 0 there, line factors ip = c = 0 (the sweeps test ip == 0).
  const uint8_t* snk;
  // fused kernel, latency-bound launches: the coarse solve of an iteration is queued WITH the F_DOWN / F_PRECOND tasks and
  // runs its inward phase behind them (spec = 1; the words of rc are then self-validating, see kRcSentinel)
  int spec;
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

}  // namespace usaug
