// Generic confidence-map PCG kernel: grid-synchronous, two slots per iteration (stencil | column sweeps),
// line preconditioner only, vector type T = float (with fp64 refinement) or double.  It serves fp64 vectors
// and widths that are not a multiple of 4, and is the A/B reference of the fused fp32 kernel
// (USAUG_CM_LEGACY_KERNEL).  Its scalar stencil also evaluates the fp64 residual for the fused kernel.
#pragma once
#include "confmap_types.cuh"

namespace usaug {

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
          USAUG_DEV_ASSERT(y >= 1 && y <= H - 2 && x >= 0 && x < W);
          if (a.snk != nullptr && a.snk[base + o] != 0) q = (A)0;      // sink pixel: decoupled row, r = q = 0
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
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int gw = blockIdx.x * wpb + (threadIdx.x >> 5);
  const int nw = gridDim.x * wpb;
  const int B = a.B, H = a.H, W = a.W;
  const int per1 = a.nstrip1 * a.nrb;                         // stencil tasks per image: 30-column strips x row blocks
  const int tot1 = per1 * B, tot2 = a.nstrip2 * B;

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
      const int strip = ti % a.nstrip1, rbk = ti / a.nstrip1;
      double s;
      if (mode == M_INNER) {
        const bool first = __ldcg(&st->first) != 0;
        const T beta = (T)__ldcg(&st->beta);
        s = stencil_task<T, false>(a, b, strip * kStripCols, kStripCols, rbk, lane, beta, first);
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

}  // namespace usaug
