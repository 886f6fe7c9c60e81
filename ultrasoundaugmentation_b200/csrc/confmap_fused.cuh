// Fused fp32 confidence-map PCG kernel (production path).  Included by confmap_pcg.cu.
#pragma once
#include "confmap_generic.cuh"

// rows consumed per cp.async wait in the F_UP / F_DOWN sweeps (instruction-level parallelism across rows; 3 and 4 were measured
// slower: fewer rows in flight), and the shared memory of a CTA in staged F_UP rows of 2304 B (the worker warps split it into
// their private rings; 88 and 96 rows were not faster)
constexpr int kRowsPerWaitUp = 2, kRowsPerWaitDown = 2;

namespace usaug {

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

struct RingU {   // one staged row of a 128-column strip for F_UP (2304 bytes)
  uint2 hq[32];    // h as bf16: 4 columns per lane (copied 16 bytes at a time by the even lanes)
  float4 p[32], a[32];
  uint4 ipc[32];
  // 16-byte chunks that contain the strip's halo columns (cp.async.cg = L2, never a stale L1 line):
  // [0..3] columns x0-4..x0-1 of h, ipc, p, a (.w is the left neighbour column);
  // [16..19] columns x0+128..x0+131 of the same planes (.x is the right neighbour column)
  float4 halo[32];
};
struct RingD {   // one staged row for F_DOWN / F_PRECOND (3200 bytes)
  float4 p[32], pp[32], e[32], r[32], q[32];
  uint4 ipc[32];
  uint32_t ipc4[32];   // level preconditioner: line factor of each lane's four-column aggregate
};
constexpr int kRingRowsPerCta = 80;                                   // x 2304 B = 180 KB per CTA
constexpr size_t kFusedSmem = sizeof(RingU) * kRingRowsPerCta;

__device__ __forceinline__ float hsel(const float4& v, int lane) { return lane == 31 ? v.x : v.w; }

// The h plane of the fused kernel (h = L^-1 r, written by F_DOWN, read by the next F_UP) is stored as bf16 in the z
// allocation, row pitch = W rounded up to 8 elements (16-byte copies): 4 of the ~60 bytes an unknown costs per iteration.
// h only feeds the preconditioner z = L^-T D^-1 h; F_DOWN forms r.z = sum ip h h~ with the rounded h~ that F_UP will see,
// so the pair (z, r.z) stays consistent, and the recurrences themselves run in fp32.
__device__ __forceinline__ int h_pitch(int W) { return (W + 7) & ~7; }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {      // round to nearest even, one instruction
  uint32_t w;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(hi), "f"(lo));
  return w;
}
__device__ __forceinline__ float hq_halo(const float4& v, int lane) {      // column x0-1 of a left chunk / x0+128 of a right one
  return lane == 31 ? bf16_lo(__float_as_uint(v.x)) : bf16_hi(__float_as_uint(v.w));
}

struct RingUP {   // one staged row of a 128-column strip for F_UP, weight planes read from memory (3840 bytes)
  uint2 hq[32];    // h as bf16: 4 columns per lane (copied 16 bytes at a time by the even lanes)
  float4 p[32], s[32], e[32], se[32], sw[32];
  uint4 ipc[32];
  // 16-byte chunks that contain the strip's halo columns (cp.async.cg = L2, never a stale L1 line):
  // [0..4] columns x0-4..x0-1 of h, ipc, p, wE, wSE (.w is the left neighbour column);
  // [16..19] columns x0+128..x0+131 of h, ipc, p, wSW (.x is the right neighbour column)
  float4 halo[32];
};

// Row range and addressing of one F_UP / F_DOWN task.  Untwisted (a.tw_m == 0): one task per strip sweeps all
// unknown rows.  Twisted line factorisation (a.tw_m = m > 0): T = N D N^T is eliminated from BOTH ends towards
// row m, so DOWN runs two independent half sweeps (rows 1..m downwards, rows H-2..m+1 upwards) and UP runs
// two independent half sweeps from row m outwards: twice the tasks, half the dependent chain, the same T^-1.
// The lower half works on the vertically mirrored view yv = H-1-y, in which it looks exactly like the upper
// half: the factor plane stores, per row, the coupling TOWARDS row m, and the edge weights are either
// recomputed from the `a` plane (symmetric in the two pixels) or read from mirrored weight planes.
struct HalfView {
  bool flip;            // lower half: mirrored view
  int yown;             // own rows of the task in the view: 1 .. yown
  int ys;               // F_UP starts here (yown + 1 = the first row of the other half, or the seed row H-1)
  ptrdiff_t rs;         // row stride in the view (+-W)
  size_t base;          // element offset of view row 0
  __device__ __forceinline__ HalfView(const PcgArgs<float>& a, int b, int half) {
    const int H = a.H, W = a.W, m = a.tw_m;
    flip = (half != 0);
    yown = m ? (flip ? H - 2 - m : m) : H - 2;
    ys = m ? yown + 1 : H - 1;
    rs = flip ? -(ptrdiff_t)W : (ptrdiff_t)W;
    base = (size_t)b * H * W + (flip ? (size_t)(H - 1) * W : 0);
  }
};

// What the otherwise idle lanes of an F_UP row's halo-copy instruction fetch (16-byte chunks into Ring*::halo):
//   halo[5..7]    three chunks of this image's coarse correction zc around the strip's aggregate columns
//   halo[8..15]   h4: the level preconditioner's forward-eliminated residual of the strip's 32 four-column aggregates
//   halo[20..27]  ipc4: their line factors
//   halo[28..31]  the chunks that hold the aggregates left / right of the strip: h4 left, ipc4 left, h4 right, ipc4 right
// Everything is zero-filled where it does not exist (seed rows, image border, feature disabled), so the consumer adds
// the terms unconditionally.  (Round 1 fetched zc with plain loads one aggregate row ahead; ncu showed the sweep
// waiting on exactly those loads at the next warp barrier: 19 % of the F_UP stall samples of a latency-bound launch.)
constexpr unsigned kCoarseDone = 0x7fffffffu;   // ImgState::zprog: the coarse solve of this iteration is complete

// Speculative coarse solve: the restriction rc is consumed while the image's F_DOWN tasks are still writing it.  Every word of
// rc is self-validating: the solve leaves a sentinel (a NaN no arithmetic produces) in each word it has consumed, the next
// phase's sweeps overwrite it with plain stores (no fence, no counter in the sweep), and a staged row that still holds a
// sentinel was simply staged too early -- its words are then polled in global memory.
constexpr unsigned kRcSentinel = 0x7fffdeadu;
__device__ __forceinline__ bool rc_pending(float v) { return __float_as_uint(v) == kRcSentinel; }
// this lane's word of a block row, polled until it is written -- or until every F_DOWN task has arrived (then whatever is
// there is final: a solve that has broken down must not hang the device)
__device__ __noinline__ float rc_poll(const float* word, const volatile unsigned* arrived, unsigned per) {
  unsigned spins = 0;
  for (;;) {
    const float v = __ldcg(word);
    if (!rc_pending(v)) return v;
    if (*arrived >= per) return __ldcg(word);
    __nanosleep(100);
    if (++spins > (1u << 24)) __trap();
  }
}

// slow path of UpAux::wait_zc: until more than `k` outward steps of the coarse solve are published (whole warp)
__device__ __noinline__ unsigned zc_spin(volatile unsigned* prog, unsigned k) {
  unsigned v, spins = 0;
  for (;;) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(const_cast<const unsigned*>(prog)) : "memory");
    if (v > k) return v;
    __nanosleep(64);
    if (++spins > (1u << 24)) __trap();                    // (~10 s: the producer is a running warp, this cannot happen)
  }
}

struct UpAux {
  const float* base;     // this lane's source: view row 0 of a level plane / row 0 of zc, chunk offset included
  ptrdiff_t stride;      // elements per view row (0 for the coarse lanes, whose row comes from the aggregate index)
  bool on, isz, flip;
  int sh, nb, H;
  int zown, zhal;        // float index inside halo[5..7] of the lane's own / halo aggregate column
  int hoff4;             // 28 (lane 0: left neighbour aggregate) or 30 (lane 31: right one)
  float z4n, z4hn;       // level recurrence: z4 of the previous row (own aggregate, halo aggregate)
  // the coarse solve may still be producing zc (its outward phase runs behind the publication of this phase's tasks):
  // zM = the meeting aggregate row (-1: zc is complete before any task starts), zknown = outward steps known to be done
  volatile unsigned* zprog;
  int zM;
  unsigned zknown;
  // Whole warp, before view row yv is staged: its aggregate row of zc must be final.  Aggregate row I is produced by outward step
  // |I - M| - 1 (the meeting row itself before the tasks were published); the sweeps move away from the meeting row like the solve
  // does, about half as fast, so after a short wait at the start this is a register compare per row and a reload every few dozen rows.
  // Called ONCE per group of staged rows, outside the row bodies (a spin loop inside `issue` split the basic blocks of the
  // row pair and cost F_UP 30 %): yv = the row of the group that lies furthest from the meeting row.
  __device__ __forceinline__ int zc_step(int yv) const {     // outward step of the coarse solve that writes view row yv's aggregate row
    const int yc = min(max(yv, 1), H - 2);
    const int y = flip ? H - 1 - yc : yc;
    const int I = (y - 1) >> sh;
    return (I > zM ? I - zM : zM - I) - 1;
  }
  // view rows ya .. yb are about to be staged: the step is a convex function of the row, so its maximum is at one of the ends
  // (twisted sweeps only ever move away from the meeting row; the plain sweep approaches it first)
  __device__ __forceinline__ void wait_zc(int ya, int yb) {
    if (zM < 0) return;
    const int k = max(zc_step(ya), zc_step(yb));
    if (k >= (int)zknown) zknown = zc_spin(zprog, (unsigned)k);
  }
  __device__ __forceinline__ void init(const PcgArgs<float>& a, const HalfView& hv, int b, int x, int x0, int lane, int W) {
    flip = hv.flip; H = a.H; sh = a.cg.by_shift; nb = a.cg.nb;
    zprog = &a.st[b].zprog; zM = a.cg.by ? a.cg.tw : -1; zknown = 0u;
    on = false; isz = false; base = a.z; stride = 0;
    z4n = z4hn = 0.f;
    hoff4 = (lane == 31) ? 30 : 28;
    const int cb = ((x0 >> a.cg.bx_shift) >> 2) - 1;                 // first of the three zc chunks
    const int hx = (lane == 0) ? x0 - 1 : x0 + kStrip4;
    zown = min(max((x >> a.cg.bx_shift) - 4 * cb, 0), 11);
    zhal = min(max((hx >> a.cg.bx_shift) - 4 * cb, 0), 11);
    if (lane >= 5 && lane <= 7) {
      const int c = cb + lane - 5;
      if (a.cg.by && c >= 0 && 4 * c < nb) { on = isz = true; base = a.zc + (size_t)b * a.cg.nc + 4 * c; }
    }
    if (a.h4 != nullptr) {
      const int W4 = a.w4p, X0 = x0 >> 2;
      const size_t base4 = (size_t)b * H * W4 + (flip ? (size_t)(H - 1) * W4 : 0);
      const ptrdiff_t rs4 = flip ? -(ptrdiff_t)W4 : (ptrdiff_t)W4;
      const float* F4 = reinterpret_cast<const float*>(a.ipc4);
      int e = -1; const float* pl = nullptr;
      if (lane >= 8 && lane < 16) { e = X0 + 4 * (lane - 8); pl = a.h4; }
      else if (lane >= 20 && lane < 28) { e = X0 + 4 * (lane - 20); pl = F4; }
      else if (lane >= 28 && lane < 30 && x0 > 0) { e = X0 - 4; pl = (lane == 28) ? a.h4 : F4; }
      else if (lane >= 30 && x0 + kStrip4 < W) { e = X0 + 32; pl = (lane == 30) ? a.h4 : F4; }
      if (pl != nullptr && e >= 0 && e < W4) { on = true; base = pl + base4 + e; stride = rs4; }
    }
  }
  // Source of this lane's halo-slot chunk, row by row (the sweeps stage rows y0, y0 - 1, ...): `cur` runs down the lane's
  // plane for every role that is linear in the view row (strip halo columns and level planes; plain = source of a lane
  // without an auxiliary role); the coarse lanes take aggregate row ((y - 1) >> sh) of zc.  Branch-free: as a two-way
  // branch on the lane's role this was ~30 instructions per row executed by every warp of an F_UP task.
  const float* cur;
  ptrdiff_t cstep;
  __device__ __forceinline__ void start(const float* plain, ptrdiff_t plain_step, int y0) {
    const float* b0 = on ? base : plain;
    cstep = on ? stride : plain_step;
    cur = isz ? base : b0 + (ptrdiff_t)y0 * cstep;
    if (isz) cstep = 0;
  }
  __device__ __forceinline__ const float* next(int yv, bool interior) {
    const int y = flip ? H - 1 - yv : yv;
    const int zo = interior ? ((y - 1) >> sh) * nb : 0;          // (only read by the coarse lanes)
    const float* r = isz ? base + zo : cur;
    cur -= cstep;
    return r;
  }
  // additive terms of view row y for the lane's own columns and for its halo column: P zc + level correction
  __device__ __forceinline__ void at_row(const float4* halo, int lane, float& add, float& addh) {
    const float* f = reinterpret_cast<const float*>(halo);
    const float h4 = f[32 + lane];
    const uint32_t f4 = __float_as_uint(f[80 + lane]);
    const float h4h = f[4 * hoff4 + (lane == 31 ? 0 : 3)];
    const uint32_t f4h = __float_as_uint(f[4 * hoff4 + 4 + (lane == 31 ? 0 : 3)]);
    z4n = fmaf(bf16_hi(f4), z4n, bf16_lo(f4) * h4);
    z4hn = fmaf(bf16_hi(f4h), z4hn, bf16_lo(f4h) * h4h);
    add = f[20 + zown] + z4n;
    addh = f[20 + zhal] + z4hn;
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

// Twisted factorisation, upper half: its F_UP sweep starts one row beyond row m (the first row of the lower half,
// which the stencil of row m needs), and that row's recurrence runs the other way: z[m+1] = c[m+1] z[m] + ip[m+1] h[m+1].
// z[m] = ip[m] h[m] has no recurrence term, so it is formed directly here and seeds the sweep.
__device__ __forceinline__ void twist_seed(const PcgArgs<float>& a, const HalfView& hv, const uint16_t* HQ, const uint32_t* IPC,
                                           int b, int x, int x0, int lane, bool valid, float (&zn)[4], float& zhn, UpAux& ax) {
  if (a.tw_m == 0 || hv.flip) return;
  const size_t o = (size_t)a.tw_m * a.W, oh = (size_t)a.tw_m * h_pitch(a.W);
  if (a.h4 != nullptr) {                                          // the level preconditioner's recurrences, same seed
    const size_t o4 = ((size_t)b * a.H + a.tw_m) * a.w4p;
    const int X = x >> 2, Xh = ((lane == 0) ? x0 - 1 : x0 + kStrip4) >> 2;
    if (valid) ax.z4n = bf16_lo(__ldcg(a.ipc4 + o4 + X)) * __ldcg(a.h4 + o4 + X);
    if ((lane == 0 && x0 > 0) || (lane == 31 && x0 + kStrip4 < a.W)) ax.z4hn = bf16_lo(__ldcg(a.ipc4 + o4 + Xh)) * __ldcg(a.h4 + o4 + Xh);
  }
  if (valid) {
    const uint2 hw = __ldcg(reinterpret_cast<const uint2*>(HQ + oh + x));
    const uint4 f = __ldcg(reinterpret_cast<const uint4*>(IPC + o + x));
    zn[0] = bf16_lo(f.x) * bf16_lo(hw.x); zn[1] = bf16_lo(f.y) * bf16_hi(hw.x);
    zn[2] = bf16_lo(f.z) * bf16_lo(hw.y); zn[3] = bf16_lo(f.w) * bf16_hi(hw.y);
  }
  const int hx = (lane == 0) ? x0 - 1 : x0 + kStrip4;
  if ((lane == 0 && x0 > 0) || (lane == 31 && x0 + kStrip4 < a.W)) zhn = bf16_lo(__ldcg(IPC + o + hx)) * __uint_as_float((uint32_t)__ldcg(HQ + oh + hx) << 16);
}

template <int R, bool SINK>
__device__ __forceinline__ double up_task(const PcgArgs<float>& a, unsigned char* ringmem, int b, int strip, int half, int lane,
                                          float beta, bool first, int pcur) {
  RingU* ring = reinterpret_cast<RingU*>(ringmem);
  const int H = a.H, W = a.W;
  const HalfView hv(a, b, half);
  const size_t base = hv.base;
  const ptrdiff_t rs = hv.rs;
  const int ys = hv.ys, yown = hv.yown;
  const int x0 = strip * kStrip4, x = x0 + 4 * lane;
  const bool valid = x < W;
  const bool hasL = valid && x > 0, hasR = valid && (x + 4 < W);
  const int xs = valid ? x : 0;
  const uint32_t* __restrict__ IPC = a.ipc + base;
  const float* __restrict__ Ap = a.aplane + base;
  const float4 wk = a.wk[b];
  // reads: h (z plane), p_old (pin); writes: p_new (pout), q.  Nothing is updated in place, because the
  // halo columns of this strip belong to a neighbouring strip whose warp runs asynchronously.
  const int hp = h_pitch(W);                                       // bf16 h plane: its own row pitch and stride
  const ptrdiff_t rsh = hv.flip ? -(ptrdiff_t)hp : (ptrdiff_t)hp;
  const uint16_t* __restrict__ HQ = reinterpret_cast<const uint16_t*>(a.z) + (size_t)b * H * hp + (hv.flip ? (size_t)(H - 1) * hp : 0);
  const float* __restrict__ Pin = (pcur ? a.p2 : a.p) + base;
  float* __restrict__ P = (pcur ? a.p : a.p2) + base;
  float* __restrict__ Q = a.q + base;
  // halo fetch role of this lane: k = which plane, left (lanes 0-3) or right (lanes 16-19) chunk
  const int k = lane & 15;
  const bool isL = lane < 16;
  const int hx = isL ? x0 - 4 : x0 + kStrip4;
  const bool hcol = isL ? (x0 > 0) : (x0 + kStrip4 < W);
  const bool hrole = hcol && k <= 3;
  // (k == 0: the h halo chunk holds 8 bf16 columns, x0-8 .. x0-1 or x0+128 .. x0+135, and steps by the bf16 pitch)
  const float* hsrc = (k == 0) ? reinterpret_cast<const float*>(HQ + (hrole ? (isL ? x0 - 8 : x0 + kStrip4) : 0))
                      : ((k == 1) ? reinterpret_cast<const float*>(IPC) : (k == 2) ? Pin : Ap) + (hrole ? hx : 0);
  const ptrdiff_t hstep = (k == 0) ? rsh / 2 : rs;
  UpAux ax;
  ax.init(a, hv, b, x, x0, lane, W);
  const bool h_in_only = hrole && k == 3;         // the a plane is also needed from the seed rows
  const bool h_interior = (hrole && k != 3 && !(k == 2 && first)) || ax.on;

  int slot_i = 0;                                 // ring slot of the next row to stage
  ptrdiff_t o_i = (ptrdiff_t)ys * rs;             // element offset of that row (running: no 64-bit multiply per row)
  ptrdiff_t oh_i = (ptrdiff_t)ys * rsh;           // ... and in the bf16 h plane
  auto issue = [&](int y) {                       // stage view row y (rows are visited ys .. 0)
    RingU* r = ring + slot_i;
    slot_i = (slot_i + 1 == R) ? 0 : slot_i + 1;
    const bool in = y >= 0;
    const bool interior = y >= 1 && y <= H - 2;
    const ptrdiff_t o = in ? o_i : (ptrdiff_t)0;
    if (!(lane & 1)) cp_async16(&r->hq[lane], HQ + (in ? oh_i : (ptrdiff_t)0) + xs, interior && valid);   // 8 columns: this lane's and the next one's
    oh_i -= rsh;
    cp_async16(&r->ipc[lane], IPC + o + xs, interior && valid);
    cp_async16(&r->p[lane], Pin + o + xs, interior && valid && !first);
    cp_async16(&r->a[lane], Ap + o + xs, in && valid);
    cp_async16(&r->halo[lane], ax.next(y, interior), (in && h_in_only) || (interior && h_interior));
    cp_async_commit();
    o_i -= rs;
  };
  const unsigned full = 0xffffffffu;
  float zn[4] = {0.f, 0.f, 0.f, 0.f}, zhn = 0.f;            // z of the row below (own columns, halo column)
  float pl[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, pc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float ab[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};            // a of the row below, columns x-1..x+4
  WRow wc;
#pragma unroll
  for (int j = 0; j < 4; ++j) wc.s[j] = wc.e[j] = wc.se[j] = wc.sw[j] = 0.f;
  wc.eL = wc.seL = wc.swR = 0.f;
  const int hoff = (lane == 31) ? 16 : 0;

  ax.start(hsrc, hstep, ys);
  ax.wait_zc(ys, ys - (R - 1));
#pragma unroll
  for (int u = 0; u < R; ++u) issue(ys - u);
  twist_seed(a, hv, HQ, IPC, b, x, x0, lane, valid, zn, zhn, ax);
  double acc = 0.0;
  unsigned mc = 15u;                              // sink mask of row c: bit j set = column j is an unknown
  int slot_c = 0;                                 // ring slot of the row being consumed
  ptrdiff_t o_c = (ptrdiff_t)ys * rs + x;         // offset of the row being consumed
  // G rows are consumed per wait: only the one-FMA recurrences (z, z4) and the p / a / weight hand-down tie a row to
  // the previous one, so the loads, shuffles, ex2 and stores of neighbouring rows overlap in the single warp of a task.
  constexpr int G = kRowsPerWaitUp;
  static_assert(R > G, "UP ring too shallow");
  float pu[G][6], au[G][6];
  unsigned mu[G];
  // part A of a row: everything that reads the staged slot -- z recurrence, p = z + beta p, the row's a values
  auto rowA = [&](int y, int g) {
    const RingU* r = ring + slot_c;
    slot_c = (slot_c + 1 == R) ? 0 : slot_c + 1;
    const uint2 hw = r->hq[lane];
    const float4 h = make_float4(bf16_lo(hw.x), bf16_hi(hw.x), bf16_lo(hw.y), bf16_hi(hw.y));
    const float4 p = r->p[lane], av = r->a[lane];
    const uint4 f = r->ipc[lane];
    // lane 0 uses the left chunks' .w, lane 31 the right chunks' .x; other lanes read but never use them
    const float hh = hq_halo(r->halo[hoff + 0], lane);
    const uint32_t hf = __float_as_uint(hsel(r->halo[hoff + 1], lane));
    const float hp = hsel(r->halo[hoff + 2], lane);
    const float ha = hsel(r->halo[hoff + 3], lane);
    float z[4];     // line part of z (the recurrence runs on it alone); the coarse part is added on top
    z[0] = fmaf(bf16_hi(f.x), zn[0], bf16_lo(f.x) * h.x); z[1] = fmaf(bf16_hi(f.y), zn[1], bf16_lo(f.y) * h.y);
    z[2] = fmaf(bf16_hi(f.z), zn[2], bf16_lo(f.z) * h.z); z[3] = fmaf(bf16_hi(f.w), zn[3], bf16_lo(f.w) * h.w);
    const float zh = fmaf(bf16_hi(hf), zhn, bf16_lo(hf) * hh);
    float zadd, zaddh;
    ax.at_row(r->halo, lane, zadd, zaddh);
    pu[g][1] = fmaf(beta, p.x, z[0] + zadd); pu[g][2] = fmaf(beta, p.y, z[1] + zadd);
    pu[g][3] = fmaf(beta, p.z, z[2] + zadd); pu[g][4] = fmaf(beta, p.w, z[3] + zadd);
    float ph = fmaf(beta, hp, zh + zaddh);
    mu[g] = 15u;
    if (SINK) {                                     // sink pixels carry ip = 0: the search direction stays 0 there
      mu[g] = ((f.x & 0xffffu) ? 1u : 0u) | ((f.y & 0xffffu) ? 2u : 0u) | ((f.z & 0xffffu) ? 4u : 0u) | ((f.w & 0xffffu) ? 8u : 0u);
      if (!(mu[g] & 1u)) pu[g][1] = 0.f;
      if (!(mu[g] & 2u)) pu[g][2] = 0.f;
      if (!(mu[g] & 4u)) pu[g][3] = 0.f;
      if (!(mu[g] & 8u)) pu[g][4] = 0.f;
      if (!(hf & 0xffffu)) ph = 0.f;
    }
    const float l = __shfl_up_sync(full, pu[g][4], 1), rt = __shfl_down_sync(full, pu[g][1], 1);
    pu[g][0] = (lane == 0) ? ph : l;
    pu[g][5] = (lane == 31) ? ph : rt;
    zn[0] = z[0]; zn[1] = z[1]; zn[2] = z[2]; zn[3] = z[3]; zhn = zh;
    USAUG_DEV_ASSERT(slot_c >= 0 && slot_c < R);
    if (valid && y >= 1 && y <= yown) {
      USAUG_DEV_ASSERT(o_c == (ptrdiff_t)y * rs + x && y <= H - 2 && x + 3 < W);
      *reinterpret_cast<float4*>(P + o_c) = make_float4(pu[g][1], pu[g][2], pu[g][3], pu[g][4]);
    }
    au[g][1] = av.x; au[g][2] = av.y; au[g][3] = av.z; au[g][4] = av.w;
    const float al = __shfl_up_sync(full, av.w, 1), ar = __shfl_down_sync(full, av.x, 1);
    au[g][0] = (lane == 0) ? ha : al;
    au[g][5] = (lane == 31) ? ha : ar;
    o_c -= rs;
  };
  // part B: weights of row y (edges inside the row and down to row y+1), then q = A p of row y+1
  auto rowB = [&](int y, int g) {
    WRow wu;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      wu.s[j] = valid ? edge_w_fast(au[g][j + 1], ab[j + 1], wk.x, wk.y) : 0.f;
      wu.e[j] = valid ? edge_w_fast(au[g][j + 1], au[g][j + 2], wk.x, wk.z) : 0.f;
      wu.se[j] = valid ? edge_w_fast(au[g][j + 1], ab[j + 2], wk.x, wk.w) : 0.f;
      wu.sw[j] = valid ? edge_w_fast(au[g][j + 1], ab[j], wk.x, wk.w) : 0.f;
    }
    wu.eL = hasL ? edge_w_fast(au[g][0], au[g][1], wk.x, wk.z) : 0.f;
    wu.seL = hasL ? edge_w_fast(au[g][0], ab[1], wk.x, wk.w) : 0.f;
    wu.swR = hasR ? edge_w_fast(au[g][5], ab[4], wk.x, wk.w) : 0.f;
    if (!hasR) { wu.e[3] = 0.f; wu.se[3] = 0.f; }
    if (!hasL) wu.sw[0] = 0.f;

    const int yc = y + 1;                          // row whose q is now computable (u = y, c = y+1, l = y+2)
    {
      const float E5[5] = {wc.eL, wc.e[0], wc.e[1], wc.e[2], wc.e[3]};
      const float SEu[4] = {wu.seL, wu.se[0], wu.se[1], wu.se[2]};
      const float SWu[4] = {wu.sw[1], wu.sw[2], wu.sw[3], wu.swR};
      float q[4], dot = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pcj = pc[j + 1];
        float t = wc.s[j] * (pcj - pl[j + 1]);
        t = fmaf(wu.s[j], pcj - pu[g][j + 1], t);
        t = fmaf(E5[j + 1], pcj - pc[j + 2], t);
        t = fmaf(E5[j], pcj - pc[j], t);
        t = fmaf(wc.se[j], pcj - pl[j + 2], t);
        t = fmaf(wc.sw[j], pcj - pl[j], t);
        t = fmaf(SEu[j], pcj - pu[g][j], t);
        t = fmaf(SWu[j], pcj - pu[g][j + 2], t);
        if (SINK && !((mc >> j) & 1u)) t = 0.f;                 // ... and so does A p (decoupled row)
        q[j] = t;
        dot = fmaf(pcj, t, dot);
      }
      const bool stq = valid && yc >= 1 && yc <= yown;
      USAUG_DEV_ASSERT(!stq || (o_c + (ptrdiff_t)(G + 1 - g) * rs == (ptrdiff_t)yc * rs + x && yc <= H - 2));
      if (stq) *reinterpret_cast<float4*>(Q + o_c + (ptrdiff_t)(G + 1 - g) * rs) = make_float4(q[0], q[1], q[2], q[3]);
      acc += (double)(stq ? dot : 0.f);          // (predicated store, unconditional add: no branch between the rows of a pair)
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) { pl[j] = pc[j]; pc[j] = pu[g][j]; ab[j] = au[g][j]; }
    wc = wu;
    mc = mu[g];
  };
  for (int y = ys; y >= 0; y -= G) {
    cp_async_wait<R - G>();
    __syncwarp();          // the halo chunks were copied by OTHER lanes: their wait_group must have passed too
#pragma unroll
    for (int g = 0; g < G; ++g) rowA(y - g, g);
    __syncwarp();          // every lane has read these slots' halo chunks before any lane refills them (issue below)
#pragma unroll
    for (int g = 0; g < G; ++g) rowB(y - g, g);
    ax.wait_zc(y - R, y - (G - 1) - R);
#pragma unroll
    for (int g = 0; g < G; ++g) issue(y - g - R);
  }
  cp_async_wait<0>();
  return warp_sum(acc);
}

// Variant of F_UP that READS the four weight planes (9 plane touches instead of 6, but ~25 % fewer
// instructions per row): used when there are too few tasks to be bandwidth-bound, where the single-warp
// instruction latency of a strip, not HBM, sets the pace.
template <int R, bool SINK>
__device__ __forceinline__ double up_task_planes(const PcgArgs<float>& a, unsigned char* ringmem, int b, int strip, int half, int lane,
                                          float beta, bool first, int pcur) {
  RingUP* ring = reinterpret_cast<RingUP*>(ringmem);
  const int H = a.H, W = a.W;
  const HalfView hv(a, b, half);
  const size_t base = hv.base;
  const ptrdiff_t rs = hv.rs;
  const int ys = hv.ys, yown = hv.yown;
  const int x0 = strip * kStrip4, x = x0 + 4 * lane;
  const bool valid = x < W;
  const int xs = valid ? x : 0;
  // the mirrored view reads the mirrored weight planes (edges towards the row ABOVE in the image)
  const float* __restrict__ S = (hv.flip ? a.wS2 : a.wS) + base; const float* __restrict__ E = a.wE + base;
  const float* __restrict__ SE = (hv.flip ? a.wSE2 : a.wSE) + base; const float* __restrict__ SW = (hv.flip ? a.wSW2 : a.wSW) + base;
  const uint32_t* __restrict__ IPC = a.ipc + base;
  // reads: h (z plane), p_old (pin); writes: p_new (pout), q.  Nothing is updated in place, because the
  // halo columns of this strip belong to a neighbouring strip whose warp runs asynchronously.
  const int hp = h_pitch(W);                                       // bf16 h plane: its own row pitch and stride
  const ptrdiff_t rsh = hv.flip ? -(ptrdiff_t)hp : (ptrdiff_t)hp;
  const uint16_t* __restrict__ HQ = reinterpret_cast<const uint16_t*>(a.z) + (size_t)b * H * hp + (hv.flip ? (size_t)(H - 1) * hp : 0);
  const float* __restrict__ Pin = (pcur ? a.p2 : a.p) + base;
  float* __restrict__ P = (pcur ? a.p : a.p2) + base;
  float* __restrict__ Q = a.q + base;
  // halo fetch role of this lane: k = which plane, left (lanes 0-4) or right (lanes 16-19) chunk
  const int k = lane & 15;
  const bool isL = lane < 16;
  const int hx = isL ? x0 - 4 : x0 + kStrip4;
  const bool hcol = isL ? (x0 > 0) : (x0 + kStrip4 < W);
  const bool hrole = hcol && (k <= 3 || (k == 4 && isL));
  // (k == 0: the h halo chunk holds 8 bf16 columns, x0-8 .. x0-1 or x0+128 .. x0+135, and steps by the bf16 pitch)
  const float* hsrc = (k == 0) ? reinterpret_cast<const float*>(HQ + (hrole ? (isL ? x0 - 8 : x0 + kStrip4) : 0))
                      : ((k == 1) ? reinterpret_cast<const float*>(IPC) : (k == 2) ? Pin : (k == 3) ? (isL ? E : SW) : SE) + (hrole ? hx : 0);
  const ptrdiff_t hstep = (k == 0) ? rsh / 2 : rs;
  UpAux ax;
  ax.init(a, hv, b, x, x0, lane, W);
  const bool h_in_only = hrole && k >= 3;         // weights are also needed from the seed row 0
  const bool h_interior = (hrole && k < 3 && !(k == 2 && first)) || ax.on;

  const int y1 = ys < H - 2 ? ys : H - 2;         // first row of the sweep (the seed row H-1 carries nothing here)
  int slot_i = 0;                                 // ring slot of the next row to stage
  ptrdiff_t o_i = (ptrdiff_t)y1 * rs;             // element offset of that row (running: no 64-bit multiply per row)
  ptrdiff_t oh_i = (ptrdiff_t)y1 * rsh;           // ... and in the bf16 h plane
  auto issue = [&](int y) {                       // stage view row y (rows are visited y1 .. 0)
    RingUP* r = ring + slot_i;
    slot_i = (slot_i + 1 == R) ? 0 : slot_i + 1;
    const bool in = y >= 0;
    const bool interior = y >= 1;
    const ptrdiff_t o = in ? o_i : (ptrdiff_t)0;
    if (!(lane & 1)) cp_async16(&r->hq[lane], HQ + (in ? oh_i : (ptrdiff_t)0) + xs, interior && valid);   // 8 columns: this lane's and the next one's
    oh_i -= rsh;
    cp_async16(&r->ipc[lane], IPC + o + xs, interior && valid);
    cp_async16(&r->p[lane], Pin + o + xs, interior && valid && !first);
    cp_async16(&r->s[lane], S + o + xs, in && valid);
    cp_async16(&r->e[lane], E + o + xs, interior && valid);
    cp_async16(&r->se[lane], SE + o + xs, in && valid);
    cp_async16(&r->sw[lane], SW + o + xs, in && valid);
    cp_async16(&r->halo[lane], ax.next(y, interior), (in && h_in_only) || (interior && h_interior));
    cp_async_commit();
    o_i -= rs;
  };
  const unsigned full = 0xffffffffu;
  struct WRowP { float4 s, e, se, sw; float hw1, hw2; };
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float zn[4] = {0.f, 0.f, 0.f, 0.f}, zhn = 0.f;            // z of the row below (own columns, halo column)
  float pl[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, pc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  WRowP wc;
  wc.s = wc.e = wc.se = wc.sw = zero4; wc.hw1 = wc.hw2 = 0.f;
  const int hoff = (lane == 31) ? 16 : 0;

  ax.start(hsrc, hstep, y1);
  ax.wait_zc(y1, y1 - (R - 1));
#pragma unroll
  for (int u = 0; u < R; ++u) issue(y1 - u);
  twist_seed(a, hv, HQ, IPC, b, x, x0, lane, valid, zn, zhn, ax);
  double acc = 0.0;
  unsigned mc = 15u;                              // sink mask of row c: bit j set = column j is an unknown
  int slot_c = 0;                                 // ring slot of the row being consumed
  ptrdiff_t o_c = (ptrdiff_t)y1 * rs + x;         // offset of the row being consumed
  // G rows per wait, as in up_task: the rows' loads, shuffles and stores overlap in the single warp of a task
  constexpr int G = kRowsPerWaitUp;
  static_assert(R > G, "UP ring too shallow");
  float pu[G][6];
  WRowP wu[G];
  unsigned mu[G];
  auto rowA = [&](int y, int g) {
    const RingUP* r = ring + slot_c;
    slot_c = (slot_c + 1 == R) ? 0 : slot_c + 1;
    const uint2 hw = r->hq[lane];
    const float4 h = make_float4(bf16_lo(hw.x), bf16_hi(hw.x), bf16_lo(hw.y), bf16_hi(hw.y));
    const float4 p = r->p[lane];
    const uint4 f = r->ipc[lane];
    wu[g].s = r->s[lane]; wu[g].e = r->e[lane]; wu[g].se = r->se[lane]; wu[g].sw = r->sw[lane];
    // lane 0 uses the left chunks' .w, lane 31 the right chunks' .x; other lanes read but never use them
    const float hh = hq_halo(r->halo[hoff + 0], lane);
    const uint32_t hf = __float_as_uint(hsel(r->halo[hoff + 1], lane));
    const float hp = hsel(r->halo[hoff + 2], lane);
    wu[g].hw1 = hsel(r->halo[hoff + 3], lane); wu[g].hw2 = hsel(r->halo[hoff + 4], lane);
    float z[4];     // line part of z (the recurrence runs on it alone); the coarse part is added on top
    z[0] = fmaf(bf16_hi(f.x), zn[0], bf16_lo(f.x) * h.x); z[1] = fmaf(bf16_hi(f.y), zn[1], bf16_lo(f.y) * h.y);
    z[2] = fmaf(bf16_hi(f.z), zn[2], bf16_lo(f.z) * h.z); z[3] = fmaf(bf16_hi(f.w), zn[3], bf16_lo(f.w) * h.w);
    const float zh = fmaf(bf16_hi(hf), zhn, bf16_lo(hf) * hh);
    float zadd, zaddh;
    ax.at_row(r->halo, lane, zadd, zaddh);
    pu[g][1] = fmaf(beta, p.x, z[0] + zadd); pu[g][2] = fmaf(beta, p.y, z[1] + zadd);
    pu[g][3] = fmaf(beta, p.z, z[2] + zadd); pu[g][4] = fmaf(beta, p.w, z[3] + zadd);
    float ph = fmaf(beta, hp, zh + zaddh);
    mu[g] = 15u;
    if (SINK) {                                     // sink pixels carry ip = 0: the search direction stays 0 there
      mu[g] = ((f.x & 0xffffu) ? 1u : 0u) | ((f.y & 0xffffu) ? 2u : 0u) | ((f.z & 0xffffu) ? 4u : 0u) | ((f.w & 0xffffu) ? 8u : 0u);
      if (!(mu[g] & 1u)) pu[g][1] = 0.f;
      if (!(mu[g] & 2u)) pu[g][2] = 0.f;
      if (!(mu[g] & 4u)) pu[g][3] = 0.f;
      if (!(mu[g] & 8u)) pu[g][4] = 0.f;
      if (!(hf & 0xffffu)) ph = 0.f;
    }
    const float l = __shfl_up_sync(full, pu[g][4], 1), rt = __shfl_down_sync(full, pu[g][1], 1);
    pu[g][0] = (lane == 0) ? ph : l;
    pu[g][5] = (lane == 31) ? ph : rt;
    zn[0] = z[0]; zn[1] = z[1]; zn[2] = z[2]; zn[3] = z[3]; zhn = zh;
    USAUG_DEV_ASSERT(slot_c >= 0 && slot_c < R);
    if (valid && y >= 1 && y <= yown) {
      USAUG_DEV_ASSERT(o_c == (ptrdiff_t)y * rs + x && y <= H - 2 && x + 3 < W);
      *reinterpret_cast<float4*>(P + o_c) = make_float4(pu[g][1], pu[g][2], pu[g][3], pu[g][4]);
    }
    o_c -= rs;
  };
  auto rowB = [&](int y, int g) {
    const int yc = y + 1;                          // row whose q is now computable (u = y, c = y+1, l = y+2)
    {
      const float eL_s = __shfl_up_sync(full, wc.e.w, 1);
      const float seL_s = __shfl_up_sync(full, wu[g].se.w, 1);
      const float swR_s = __shfl_down_sync(full, wu[g].sw.x, 1);
      const float eL = (lane == 0) ? wc.hw1 : eL_s;           // wE[c, x-1]
      const float seL = (lane == 0) ? wu[g].hw2 : seL_s;      // wSE[u, x-1]
      const float swR = (lane == 31) ? wu[g].hw1 : swR_s;     // wSW[u, x+4]
      const float Sc[4] = {wc.s.x, wc.s.y, wc.s.z, wc.s.w};
      const float Su[4] = {wu[g].s.x, wu[g].s.y, wu[g].s.z, wu[g].s.w};
      const float E5[5] = {eL, wc.e.x, wc.e.y, wc.e.z, wc.e.w};
      const float SEc[4] = {wc.se.x, wc.se.y, wc.se.z, wc.se.w};
      const float SWc[4] = {wc.sw.x, wc.sw.y, wc.sw.z, wc.sw.w};
      const float SEu[4] = {seL, wu[g].se.x, wu[g].se.y, wu[g].se.z};
      const float SWu[4] = {wu[g].sw.y, wu[g].sw.z, wu[g].sw.w, swR};
      float q[4], dot = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pcj = pc[j + 1];
        float t = Sc[j] * (pcj - pl[j + 1]);
        t = fmaf(Su[j], pcj - pu[g][j + 1], t);
        t = fmaf(E5[j + 1], pcj - pc[j + 2], t);
        t = fmaf(E5[j], pcj - pc[j], t);
        t = fmaf(SEc[j], pcj - pl[j + 2], t);
        t = fmaf(SWc[j], pcj - pl[j], t);
        t = fmaf(SEu[j], pcj - pu[g][j], t);
        t = fmaf(SWu[j], pcj - pu[g][j + 2], t);
        if (SINK && !((mc >> j) & 1u)) t = 0.f;                 // ... and so does A p (decoupled row)
        q[j] = t;
        dot = fmaf(pcj, t, dot);
      }
      const bool stq = valid && yc >= 1 && yc <= yown;
      USAUG_DEV_ASSERT(!stq || (o_c + (ptrdiff_t)(G + 1 - g) * rs == (ptrdiff_t)yc * rs + x && yc <= H - 2));
      if (stq) *reinterpret_cast<float4*>(Q + o_c + (ptrdiff_t)(G + 1 - g) * rs) = make_float4(q[0], q[1], q[2], q[3]);
      acc += (double)(stq ? dot : 0.f);          // (predicated store, unconditional add: no branch between the rows of a pair)
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) { pl[j] = pc[j]; pc[j] = pu[g][j]; }
    wc = wu[g];
    mc = mu[g];
  };
  for (int y = y1; y >= 0; y -= G) {
    cp_async_wait<R - G>();
    __syncwarp();          // the halo chunks were copied by OTHER lanes: their wait_group must have passed too
#pragma unroll
    for (int g = 0; g < G; ++g) rowA(y - g, g);
    __syncwarp();          // every lane has read these slots' halo chunks before any lane refills them (issue below)
#pragma unroll
    for (int g = 0; g < G; ++g) rowB(y - g, g);
    ax.wait_zc(y - R, y - (G - 1) - R);
#pragma unroll
    for (int g = 0; g < G; ++g) issue(y - g - R);
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
__device__ __forceinline__ void down_task(const PcgArgs<float>& a, unsigned char* ringmem, int b, int strip, int half, int lane,
                                          float alpha, float alpha_prev, int pcur, double& rr_out, double& rz_out) {
  RingD* ring = reinterpret_cast<RingD*>(ringmem);
  const int H = a.H, W = a.W;
  const HalfView hv(a, b, half);
  const size_t base = hv.base;
  const ptrdiff_t rs = hv.rs;
  // rows 1 .. ylast of the view.  Twisted: the upper half also takes row m, but only its part of h[m] (the last
  // arriver of the phase adds the lower half's term and the row's share of r.z, see twist_fix)
  const int ylast = hv.yown;
  const int ymid = (a.tw_m && !hv.flip) ? a.tw_m : -1;
  const int x = strip * kStrip4 + 4 * lane;
  const bool valid = x < W;
  const int xs = valid ? x : 0;
  const uint32_t* __restrict__ IPC = a.ipc + base;
  const float* __restrict__ P = (pcur ? a.p2 : a.p) + base;        // p of this iteration (written by the last F_UP)
  const float* __restrict__ PP = (pcur ? a.p : a.p2) + base;       // p of the previous iteration
  float* __restrict__ Ev = a.e + base; float* __restrict__ Rv = a.r + base;
  const float* __restrict__ Q = a.q + base;
  // h = L^-1 r goes to the z allocation as bf16 (read by the next F_UP; see h_pitch)
  const int hpit = h_pitch(W);
  const ptrdiff_t rsh = hv.flip ? -(ptrdiff_t)hpit : (ptrdiff_t)hpit;
  uint16_t* __restrict__ HQ = reinterpret_cast<uint16_t*>(a.z) + (size_t)b * H * hpit + (hv.flip ? (size_t)(H - 1) * hpit : 0);
  // level preconditioner (four-column aggregates = one lane): h4 = L4^-1 (P4^T r), its own plane of pitch w4p
  const bool lv = a.h4 != nullptr;
  const int W4 = a.w4p;
  const size_t base4 = (size_t)b * H * W4 + (hv.flip ? (size_t)(H - 1) * W4 : 0);
  const ptrdiff_t rs4 = hv.flip ? -(ptrdiff_t)W4 : (ptrdiff_t)W4;
  // every lane stages its own level factor: 4-byte cp.async goes through L1, which is safe for a plane that is
  // constant during the solve (and needs no cross-lane hand-off, unlike a 16-byte chunk copied by a quarter of the lanes)
  const uint32_t* __restrict__ F4 = lv ? a.ipc4 + base4 + (valid ? (x >> 2) : 0) : a.ipc;
  float* __restrict__ H4 = lv ? a.h4 + base4 + (x >> 2) : a.z;
  float h4p = 0.f, c4p = 0.f;
  int slot_i = 0, slot_c = 0;
  ptrdiff_t o_i = rs + xs, o4_i = rs4;             // element offsets of the next row to stage (running: no 64-bit multiply per row)
  auto issue = [&](int y) {
    RingD* r = ring + slot_i;
    slot_i = (slot_i + 1 == RD) ? 0 : slot_i + 1;
    const bool in = (y <= ylast) && valid;
    const ptrdiff_t o = (y <= ylast) ? o_i : (ptrdiff_t)xs;
    if (lv) cp_async4(&r->ipc4[lane], F4 + ((y <= ylast) ? o4_i : (ptrdiff_t)0), in);
    if (KIND == 2) {
      cp_async16(&r->p[lane], P + o, in);
      cp_async16(&r->pp[lane], PP + o, in);
      cp_async16(&r->e[lane], Ev + o, in);
    }
    if (KIND != 0) cp_async16(&r->q[lane], Q + o, in);
    cp_async16(&r->r[lane], Rv + o, in);
    cp_async16(&r->ipc[lane], IPC + o, in);
    cp_async_commit();
    o_i += rs; o4_i += rs4;
  };
#pragma unroll
  for (int u = 0; u < RD; ++u) issue(1 + u);
  float hp[4] = {0.f, 0.f, 0.f, 0.f}, cp[4] = {0.f, 0.f, 0.f, 0.f};
  float racc = 0.f;
  double rr = 0.0, rz = 0.0;
  ptrdiff_t o_c = rs + x, o4_c = rs4, oh_c = rsh + x;   // offsets of the row being consumed
  // restriction: rows until the first flush, destination of that flush, and whether this lane stores
  //   upper half / untwisted: image rows 1, 2, ...: an aggregate row ends where (ya & (by - 1)) == 0 (and at row H - 2)
  //   lower half (flipped):   image rows H-2, H-3, ...: an aggregate row ends where ((ya - 1) & (by - 1)) == 0
  const int by = a.cg.by ? a.cg.by : 1;
  int agc = hv.flip ? (((H - 2) - 1) & (by - 1)) + 1 : by;
  const int ya0 = hv.flip ? H - 2 : 1;
  const bool rc_st = a.cg.by != 0 && valid && (lane & ((1 << (a.cg.bx_shift - 2)) - 1)) == 0;
  float* rcp = a.rc + (a.cg.by ? (size_t)b * a.cg.nc + (size_t)((ya0 - 1) >> a.cg.by_shift) * a.cg.nb + (size_t)(x >> a.cg.bx_shift) : 0);
  const ptrdiff_t rc_step = a.cg.by ? (hv.flip ? -(ptrdiff_t)a.cg.nb : (ptrdiff_t)a.cg.nb) : 0;

  // One row.  Only the one-FMA recurrences (h, h4) tie a row to its predecessor; G rows are consumed per wait so that the
  // loads, updates and stores of neighbouring rows overlap in the instruction stream of the (single) warp of a task.
  auto row = [&](int y) {
    const RingD* r = ring + slot_c;
    slot_c = (slot_c + 1 == RD) ? 0 : slot_c + 1;
    const bool live = valid && y <= ylast;         // rows past the end are staged as zeros and store nothing
    USAUG_DEV_ASSERT(!live || (o_c == (ptrdiff_t)y * rs + x && y >= 1 && y <= H - 2 && x + 3 < W && o4_c == (ptrdiff_t)y * rs4));
    const float4 rv = r->r[lane];
    const uint4 f = r->ipc[lane];
    const float4 cv = make_float4(bf16_hi(f.x), bf16_hi(f.y), bf16_hi(f.z), bf16_hi(f.w));
    const float4 iv = make_float4(bf16_lo(f.x), bf16_lo(f.y), bf16_lo(f.z), bf16_lo(f.w));
    float rn[4] = {rv.x, rv.y, rv.z, rv.w};
    if (KIND != 0) {
      const float4 qv = r->q[lane];
      rn[0] = fmaf(-alpha, qv.x, rv.x); rn[1] = fmaf(-alpha, qv.y, rv.y);
      rn[2] = fmaf(-alpha, qv.z, rv.z); rn[3] = fmaf(-alpha, qv.w, rv.w);
      if (live) *reinterpret_cast<float4*>(Rv + o_c) = make_float4(rn[0], rn[1], rn[2], rn[3]);
      if (KIND == 2) {
        const float4 pv = r->p[lane], pq = r->pp[lane], ev = r->e[lane];
        const float4 en = make_float4(fmaf(alpha, pv.x, fmaf(alpha_prev, pq.x, ev.x)), fmaf(alpha, pv.y, fmaf(alpha_prev, pq.y, ev.y)),
                                      fmaf(alpha, pv.z, fmaf(alpha_prev, pq.z, ev.z)), fmaf(alpha, pv.w, fmaf(alpha_prev, pq.w, ev.w)));
        if (live) *reinterpret_cast<float4*>(Ev + o_c) = en;
      }
    } else if (live) {
      *reinterpret_cast<float4*>(Ev + o_c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = fmaf(cp[j], hp[j], rn[j]);
    const uint32_t hw0 = pack_bf16x2(h[0], h[1]), hw1 = pack_bf16x2(h[2], h[3]);
    if (live) *reinterpret_cast<uint2*>(HQ + oh_c) = make_uint2(hw0, hw1);
    const float srr = fmaf(rn[0], rn[0], fmaf(rn[1], rn[1], fmaf(rn[2], rn[2], rn[3] * rn[3])));
    // r.z = sum ip h h~ with the rounded h~ that F_UP will read: consistent with the z it forms
    float srz = fmaf(iv.x * h[0], bf16_lo(hw0), fmaf(iv.y * h[1], bf16_hi(hw0), fmaf(iv.z * h[2], bf16_lo(hw1), iv.w * h[3] * bf16_hi(hw1))));
    const float r4 = (rn[0] + rn[1]) + (rn[2] + rn[3]);             // P4^T r: the lane's four columns
    if (lv) {
      const uint32_t f4 = r->ipc4[lane];
      const float h4 = fmaf(c4p, h4p, r4);
      if (live) H4[o4_c] = h4;
      srz = fmaf(bf16_lo(f4) * h4, h4, srz);
      h4p = h4; c4p = bf16_hi(f4);
    }
    rr += (double)srr;
    rz += (double)((y != ymid) ? srz : 0.f);
    // restriction rc = P^T r: sum of the new residual over each aggregate, flushed at the last row of an aggregate row in
    // the order of travel (a countdown and a running destination: the test and the address were ~20 instructions per row)
    racc += r4;
    agc -= (y <= ylast) ? 1 : 0;                   // (the zero-filled row past the end of an odd sweep does not count)
    if (agc == 0) {
      float v = racc;                                        // 4 ... 32 lanes share one aggregate column
      v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2);
      const float t4 = __shfl_xor_sync(0xffffffffu, v, 4); v += (a.cg.bx_shift >= 5) ? t4 : 0.f;
      const float t8 = __shfl_xor_sync(0xffffffffu, v, 8); v += (a.cg.bx_shift >= 6) ? t8 : 0.f;
      const float t16 = __shfl_xor_sync(0xffffffffu, v, 16); v += (a.cg.bx_shift == 7) ? t16 : 0.f;
      USAUG_DEV_ASSERT(!rc_st || (rcp >= a.rc + (size_t)b * a.cg.nc && rcp < a.rc + (size_t)(b + 1) * a.cg.nc));
      if (rc_st) *rcp = v;
      rcp += rc_step;
      racc = 0.f;
      agc = a.cg.by;

    }
#pragma unroll
    for (int j = 0; j < 4; ++j) hp[j] = h[j];
    cp[0] = cv.x; cp[1] = cv.y; cp[2] = cv.z; cp[3] = cv.w;
    o_c += rs; o4_c += rs4; oh_c += rsh;
  };
  constexpr int G = kRowsPerWaitDown;              // rows per wait
  static_assert(RD > G, "DOWN ring too shallow");
  for (int y = 1; y <= ylast; y += G) {
    cp_async_wait<RD - G>();
#pragma unroll
    for (int g = 0; g < G; ++g) row(y + g);
#pragma unroll
    for (int g = 0; g < G; ++g) issue(y + g + RD);
  }
  if (agc != by) {                                   // a partial aggregate row at the end of the sweep (image row H - 2, or row m)
    float v = racc;
    v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2);
    const float t4 = __shfl_xor_sync(0xffffffffu, v, 4); v += (a.cg.bx_shift >= 5) ? t4 : 0.f;
    const float t8 = __shfl_xor_sync(0xffffffffu, v, 8); v += (a.cg.bx_shift >= 6) ? t8 : 0.f;
    const float t16 = __shfl_xor_sync(0xffffffffu, v, 16); v += (a.cg.bx_shift == 7) ? t16 : 0.f;
    if (rc_st) *rcp = v;
  }
  cp_async_wait<0>();
  rr_out = warp_sum(rr); rz_out = warp_sum(rz);
}

// F_COARSE: one warp solves Ac zc = rc for one image and returns rc . zc (block LDL^T, see confmap_coarse.cuh).
//   forward    v_I = rc_I - E_I w_{I-1},  w_I = W_I v_I          (v_I parked in the zc array)
//   backward   zc_I = W_I (v_I - E_{I+1}^T zc_{I+1})
// Lane l owns element row = l % NB of the current block row (32 / NB lanes share a row and split the columns of
// W_I between them); the vector is broadcast through a double-buffered shared-memory line, W_I is read from a
// cp.async ring of block-row records with 128-bit, conflict-free loads (row pitch NB + 4).  One __syncwarp per step.
template <int NB>
struct CoarseSlot {
  float W[NB * (NB + 4)];
  float E[4 * NB];     // couplings of this block row to the one above: centre, left (J-1), right (J+1), spare
  float vec[NB];       // forward: rc_I; backward: v_I
};

// (__noinline__ with narrow arguments: the solves get their own register allocation and instruction schedule, and the
//  sweeps' code no longer changes when this code does -- F_UP was measured 17 % slower after an unrelated edit here)
template <int NB, int RING_BYTES>
__device__ __noinline__ double coarse_task_nb(const CoarseGeom g, const float* __restrict__ rec, const float* __restrict__ rc,
                                              float* __restrict__ zc, unsigned char* ringmem, int lane) {
  using Slot = CoarseSlot<NB>;
  constexpr int Dfit = (RING_BYTES - 2 * NB * (int)sizeof(float)) / (int)sizeof(Slot);
  constexpr int D = Dfit > 16 ? 16 : Dfit;                       // ring depth in block rows
  static_assert(D >= 3, "coarse ring too shallow");
  constexpr int LPR = 32 / NB;                                   // lanes per row
  constexpr int CP = NB / LPR;                                   // columns of W per lane
  const int nby = g.nby;
  Slot* ring = reinterpret_cast<Slot*>(ringmem);
  float* vs = reinterpret_cast<float*>(ringmem + (size_t)D * sizeof(Slot));      // [2][NB]
  const int row = lane & (NB - 1), c0 = (lane / NB) * CP;
  const bool owner = lane < NB;
  const unsigned full = 0xffffffffu;
  constexpr int RCH = (NB * (NB + 4) + 4 * NB) / 128;            // 16-byte chunks of a record per lane: 1, 3 or 10
  static_assert((NB * (NB + 4) + 4 * NB) % 128 == 0, "a record is a whole number of 512-byte warp copies");
  const bool v_on = lane < NB / 4;                               // the vector: NB / 4 chunks

  auto issue = [&](int k, int I, const float* vec) {             // stage block row I into slot k % D
    if (I >= 0 && I < nby) {
      Slot* s = ring + (k % D);
      const float* src = rec + (size_t)I * g.rec + 4 * lane;      // a record is the image of Slot::W and Slot::E
      float* dst = s->W + 4 * lane;
#pragma unroll
      for (int j = 0; j < RCH; ++j) cp_async16(dst + 128 * j, src + 128 * j, true);
      if (v_on) cp_async16(&s->vec[4 * lane], vec + (size_t)I * NB + 4 * lane, true);
    }
    cp_async_commit();
  };
  auto matvec = [&](const Slot* s, const float* v) {             // (W_I v)[row], complete in every lane
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if constexpr (CP >= 4) {
#pragma unroll
      for (int q = 0; q < CP / 4; ++q) {
        const float4 wv = *reinterpret_cast<const float4*>(&s->W[row * (NB + 4) + c0 + 4 * q]);
        const float4 vv = *reinterpret_cast<const float4*>(&v[c0 + 4 * q]);
        acc[0] = fmaf(wv.x, vv.x, acc[0]); acc[1] = fmaf(wv.y, vv.y, acc[1]);
        acc[2] = fmaf(wv.z, vv.z, acc[2]); acc[3] = fmaf(wv.w, vv.w, acc[3]);
      }
    } else {
      const float2 wv = *reinterpret_cast<const float2*>(&s->W[row * (NB + 4) + c0]);
      const float2 vv = *reinterpret_cast<const float2*>(&v[c0]);
      acc[0] = wv.x * vv.x; acc[1] = wv.y * vv.y;
    }
    float r = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#pragma unroll
    for (int o = NB; o < 32; o <<= 1) r += __shfl_xor_sync(full, r, o);
    return r;
  };

  // ---- forward ----------------------------------------------------------------------------------
  for (int u = 0; u < D - 1; ++u) issue(u, u, rc);
  cp_async_wait<D - 2>();
  __syncwarp();
  float w = 0.f;                                                 // (W_{I-1} v_{I-1})[row]
  float pvw = 0.f;                                               // v . w of the previous step (added one step late:
  double dot = 0.0;                                              //  the fp64 add then hides behind the shuffles)
  for (int I = 0; I < nby; ++I) {
    const Slot* s = ring + (I % D);
    const float wls = __shfl_up_sync(full, w, 1), wrs = __shfl_down_sync(full, w, 1);
    const float e0 = s->E[row], e1 = s->E[NB + row], e2 = s->E[2 * NB + row], r0 = s->vec[row];
    dot += (double)pvw;
    const float wl = (row == 0) ? 0.f : wls, wr = (row == NB - 1) ? 0.f : wrs;
    const float v = r0 - fmaf(e0, w, fmaf(e1, wl, e2 * wr));
    float* vb = vs + (I & 1) * NB;
    if (owner) vb[row] = v;
    cp_async_wait<D - 3>();                                      // this lane's pieces of block row I + 1 have landed
    __syncwarp();                                                // publishes vb and block row I + 1; everybody is done with row I - 1
    w = matvec(s, vb);
    if (owner) zc[(size_t)I * NB + row] = v;
    issue(I + D - 1, I + D - 1, rc);                             // refills the slot of block row I - 1
    pvw = owner ? v * w : 0.f;
  }
  dot += (double)pvw;
  cp_async_wait<0>();
  __threadfence();                                               // v (in zc) is re-read through cp.async below
  __syncwarp();
  // ---- backward ---------------------------------------------------------------------------------
  for (int u = 0; u < D - 1; ++u) issue(u, nby - 1 - u, zc);
  cp_async_wait<D - 2>();
  __syncwarp();
  float z = 0.f, eC = 0.f, eL = 0.f, eR = 0.f;                   // zc_{I+1}[row] and E_{I+1}[row]
  for (int k = 0; k < nby; ++k) {
    const int I = nby - 1 - k;
    const Slot* s = ring + (k % D);
    // (E_{I+1}^T z)[j] = eC[j] z[j] + eL[j+1] z[j+1] + eR[j-1] z[j-1]
    const float tb = eL * z, tc = eR * z;
    const float tbs = __shfl_down_sync(full, tb, 1), tcs = __shfl_up_sync(full, tc, 1);
    const float v0 = s->vec[row];
    const float u = v0 - (fmaf(eC, z, (row == NB - 1) ? 0.f : tbs) + ((row == 0) ? 0.f : tcs));
    float* vb = vs + (k & 1) * NB;
    if (owner) vb[row] = u;
    eC = s->E[row]; eL = s->E[NB + row]; eR = s->E[2 * NB + row];
    cp_async_wait<D - 3>();
    __syncwarp();
    z = matvec(s, vb);
    if (owner) zc[(size_t)I * NB + row] = z;
    issue(k + D - 1, I - (D - 1), zc);
  }
  cp_async_wait<0>();
  __syncwarp();
  return warp_sum(dot);
}

// Twisted form of the same solve (a.cg.tw = M >= 0): the block factorisation runs from both ends towards aggregate row M,
// so ONE warp advances two independent recurrences per step -- rows I (from the top) and nby-1-I (from the bottom) --
// whose shuffle -> shared memory -> matrix-vector chains overlap in its instruction stream: half the dependent steps.
//   inward     top     v_I = rc_I - E_I w_{I-1},              w_I = W_I v_I      I = 0 .. M-1
//              bottom  v_I = rc_I - E_{I+1}^T w_{I+1},        w_I = W_I v_I      I = nby-1 .. M+1
//   meeting    v_M = rc_M - E_M w_{M-1} - E_{M+1}^T w_{M+1},  zc_M = W_M v_M
//   outward    top     zc_I = W_I (v_I - E_{I+1}^T zc_{I+1})                     I = M-1 .. 0
//              bottom  zc_I = W_I (v_I - E_I zc_{I-1})                           I = M+1 .. nby-1
//   rc . zc = sum_I v_I . (W_I v_I)
// The solve runs in two calls: phase 0 = inward + meeting row (returns rc . zc, which completes r . z and therefore beta, and
// leaves the meeting row's z and couplings in `mid`), phase 1 = outward.  Between them the caller publishes the image's F_UP
// tasks: F_UP consumes the aggregate rows of zc from the meeting row outwards -- the order in which phase 1 produces them -- so
// the sweeps start behind the outward front instead of after it.  Phase 1 publishes its progress after 8 steps and then every 16
// (`prog` = outward steps completed, after a fence); up_task waits on it before it stages a row (UpAux::wait_zc).
struct CoarseMid { float zM, mC, mL, mR; };

template <int NB, int RING_BYTES>
__device__ __noinline__ double coarse_task_tw_nb(const CoarseGeom g, const float* __restrict__ rec, float* rc,
                                                 float* __restrict__ zc, unsigned char* ringmem, int lane, int phase,
                                                 CoarseMid* mid, volatile unsigned* prog, bool spec, int per,
                                                 const volatile unsigned* arrived) {
  using Slot = CoarseSlot<NB>;
  constexpr int Dfit = (RING_BYTES - 4 * NB * (int)sizeof(float)) / (int)sizeof(Slot) / 2;
  constexpr int D = Dfit > 8 ? 8 : Dfit;                         // ring depth per front, in block rows
  static_assert(D >= 3, "coarse ring too shallow for two fronts");
  constexpr int LPR = 32 / NB;                                   // lanes per row
  constexpr int CP = NB / LPR;                                   // columns of W per lane
  const int nby = g.nby, M = g.tw, nB = nby - 1 - M;             // M rows above the meeting row, nB below (M - 1 or M)
  Slot* ringT = reinterpret_cast<Slot*>(ringmem);
  Slot* ringB = ringT + D;
  float* vsT = reinterpret_cast<float*>(ringmem + (size_t)2 * D * sizeof(Slot));   // [2][NB]
  float* vsB = vsT + 2 * NB;
  const int row = lane & (NB - 1), c0 = (lane / NB) * CP;
  const bool owner = lane < NB;
  const unsigned full = 0xffffffffu;
  constexpr int RCH = (NB * (NB + 4) + 4 * NB) / 128;            // 16-byte chunks of a record per lane: 1, 3 or 10
  static_assert((NB * (NB + 4) + 4 * NB) % 128 == 0, "a record is a whole number of 512-byte warp copies");
  const bool v_on = lane < NB / 4;                               // the vector: NB / 4 chunks
  const size_t rstep = (size_t)g.rec;

  // block row I into slot k % D of a ring (no commit).  A record is the image of Slot::W and Slot::E: straight copies.
  auto stage = [&](Slot* ring, int k, int I, const float* vec) {
    if (I < 0 || I >= nby) return;
    Slot* s = ring + (k % D);
    const float* src = rec + (size_t)I * rstep + 4 * lane;
    float* dst = s->W + 4 * lane;
#pragma unroll
    for (int j = 0; j < RCH; ++j) cp_async16(dst + 128 * j, src + 128 * j, true);
    if (v_on) cp_async16(&s->vec[4 * lane], vec + (size_t)I * NB + 4 * lane, true);
  };
  auto matvec = [&](const Slot* s, const float* v) {             // (W_I v)[row], complete in every lane
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if constexpr (CP >= 4) {
#pragma unroll
      for (int q = 0; q < CP / 4; ++q) {
        const float4 wv = *reinterpret_cast<const float4*>(&s->W[row * (NB + 4) + c0 + 4 * q]);
        const float4 vv = *reinterpret_cast<const float4*>(&v[c0 + 4 * q]);
        acc[0] = fmaf(wv.x, vv.x, acc[0]); acc[1] = fmaf(wv.y, vv.y, acc[1]);
        acc[2] = fmaf(wv.z, vv.z, acc[2]); acc[3] = fmaf(wv.w, vv.w, acc[3]);
      }
    } else {
      const float2 wv = *reinterpret_cast<const float2*>(&s->W[row * (NB + 4) + c0]);
      const float2 vv = *reinterpret_cast<const float2*>(&v[c0]);
      acc[0] = wv.x * vv.x; acc[1] = wv.y * vv.y;
    }
    float r = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#pragma unroll
    for (int o = NB; o < 32; o <<= 1) r += __shfl_xor_sync(full, r, o);
    return r;
  };
  // (E n)[row] with E from a staged record; (E^T n)[row] with E held in registers (the record of the previous step)
  auto e_direct = [&](const Slot* s, float n) {
    const float nl = __shfl_up_sync(full, n, 1), nr = __shfl_down_sync(full, n, 1);
    return fmaf(s->E[row], n, fmaf(s->E[NB + row], (row == 0) ? 0.f : nl, s->E[2 * NB + row] * ((row == NB - 1) ? 0.f : nr)));
  };
  auto e_transp = [&](float eC, float eL, float eR, float n) {
    const float tb = eL * n, tc = eR * n;
    const float tbs = __shfl_down_sync(full, tb, 1), tcs = __shfl_up_sync(full, tc, 1);
    return fmaf(eC, n, (row == NB - 1) ? 0.f : tbs) + ((row == 0) ? 0.f : tcs);
  };

  if (phase == 0) {
  // ---- inward: the top front consumes block rows 0 .. M (round k = row k), the bottom front nby-1 .. M+1 -------------
  // Speculative mode: rc is still being written by the image's F_DOWN tasks, from the outside in -- the order in which the
  // fronts consume it; see kRcSentinel.  The meeting row is staged after ALL tasks have arrived.
  const int mlast = spec ? M - 1 : M;                            // last block row the top ring stages inside the pipeline
  const bool mine = owner && row < g.nbx;                        // this lane owns a real (not padding) word of every block row
  for (int u = 0; u < D - 1; ++u) { stage(ringT, u, (u <= mlast) ? u : -1, rc); stage(ringB, u, (u < nB) ? nby - 1 - u : -1, rc); cp_async_commit(); }
  cp_async_wait<D - 2>();
  __syncwarp();
  float wT = 0.f, wB = 0.f;                                      // (W v)[row] of the previous row of each front
  float bC = 0.f, bL = 0.f, bR = 0.f;                            // E of the previous row of the bottom front
  float pvw = 0.f;
  double dot = 0.0;
  // both fronts in ONE basic block (no branch between them, so that their chains interleave); the top front has at
  // most one row more than the bottom one: that step runs the bottom front on a zero-filled slot and discards it
  for (int k = 0; k < M; ++k) {
    const bool hasB = k < nB;
    const Slot* sT = ringT + (k % D);
    const Slot* sB = ringB + (k % D);
    float rT = sT->vec[row], rB = sB->vec[row];
    if (spec) {
      const bool late = mine && (rc_pending(rT) || (hasB && rc_pending(rB)));
      if (__any_sync(full, late)) {                              // staged before the sweeps had written it: poll the words themselves
        if (mine) {
          rT = rc_poll(rc + (size_t)k * NB + row, arrived, (unsigned)per);
          if (hasB) rB = rc_poll(rc + (size_t)(nby - 1 - k) * NB + row, arrived, (unsigned)per);
        }
        rT = __shfl_sync(full, rT, row); rB = __shfl_sync(full, rB, row);     // (the lanes that share a row)
      }
      if (mine) {                                                // consumed: leave the sentinel for the next phase
        rc[(size_t)k * NB + row] = __uint_as_float(kRcSentinel);
        if (hasB) rc[(size_t)(nby - 1 - k) * NB + row] = __uint_as_float(kRcSentinel);
      }
    }
    const float vT = rT - e_direct(sT, wT);
    const float vB = rB - e_transp(bC, bL, bR, wB);
    dot += (double)pvw;
    float* vbT = vsT + (k & 1) * NB;
    float* vbB = vsB + (k & 1) * NB;
    if (owner) { vbT[row] = vT; vbB[row] = vB; }
    { const float tC = sB->E[row], tL = sB->E[NB + row], tR = sB->E[2 * NB + row]; bC = hasB ? tC : bC; bL = hasB ? tL : bL; bR = hasB ? tR : bR; }
    cp_async_wait<D - 3>();                                      // this lane's pieces of round k + 1 have landed
    __syncwarp();                                                // publishes the vectors and round k + 1; everybody is done with round k - 1
    wT = matvec(sT, vbT);
    const float wBn = matvec(sB, vbB);
    USAUG_DEV_ASSERT(k < nby && nby - 1 - k >= 0 && (k % D) < D);
    if (owner) { zc[(size_t)k * NB + row] = vT; if (hasB) zc[(size_t)(nby - 1 - k) * NB + row] = vB; }
    { const int kn = k + D - 1; stage(ringT, kn, (kn <= mlast) ? kn : -1, rc); stage(ringB, kn, (kn < nB) ? nby - 1 - kn : -1, rc); cp_async_commit(); }
    pvw = owner ? fmaf(vT, wT, hasB ? vB * wBn : 0.f) : 0.f;
    if (hasB) wB = wBn;
  }
  dot += (double)pvw;
  if (spec) {                                                    // every F_DOWN task of the image has arrived: rc is complete
    unsigned spins = 0;
    while (*arrived < (unsigned)per) { __nanosleep(100); if (++spins > (1u << 24)) __trap(); }
    __threadfence();
    cp_async_wait<0>();
    __syncwarp();
    stage(ringT, M, M, rc);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
  }
  // ---- the meeting row ------------------------------------------------------------------------------------------------
  float zM, mC, mL, mR;
  {
    const Slot* sT = ringT + (M % D);
    float v = sT->vec[row] - e_direct(sT, wT);
    if (nB > 0) v -= e_transp(bC, bL, bR, wB);
    float* vb = vsT + (M & 1) * NB;
    if (owner) vb[row] = v;
    mC = sT->E[row]; mL = sT->E[NB + row]; mR = sT->E[2 * NB + row];
    __syncwarp();
    zM = matvec(sT, vb);
    if (owner) { zc[(size_t)M * NB + row] = zM; dot += (double)(v * zM); }
    if (spec && mine) rc[(size_t)M * NB + row] = __uint_as_float(kRcSentinel);
  }
  cp_async_wait<0>();
  __threadfence();                                               // the parked v (in zc) are re-read through cp.async in phase 1;
  __syncwarp();                                                  // zc of the meeting row is final and visible
  mid->zM = zM; mid->mC = mC; mid->mL = mL; mid->mR = mR;
  return warp_sum(dot);
  }
  // ---- outward: top front rows M-1 .. 0, bottom front rows M+1 .. nby-1 -----------------------------------------------
  const float zM = mid->zM;
  float mC = mid->mC, mL = mid->mL, mR = mid->mR;
  for (int u = 0; u < D - 1; ++u) { stage(ringT, u, M - 1 - u, zc); stage(ringB, u, (u < nB) ? M + 1 + u : -1, zc); cp_async_commit(); }
  cp_async_wait<D - 2>();
  __syncwarp();
  float zT = zM, zB = zM;
  for (int k = 0; k < M; ++k) {
    const bool hasB = k < nB;
    const Slot* sT = ringT + (k % D);
    const Slot* sB = ringB + (k % D);
    const float uT = sT->vec[row] - e_transp(mC, mL, mR, zT);
    const float uB = sB->vec[row] - e_direct(sB, zB);
    float* vbT = vsT + (k & 1) * NB;
    float* vbB = vsB + (k & 1) * NB;
    if (owner) { vbT[row] = uT; vbB[row] = uB; }
    mC = sT->E[row]; mL = sT->E[NB + row]; mR = sT->E[2 * NB + row];
    cp_async_wait<D - 3>();
    __syncwarp();
    zT = matvec(sT, vbT);
    zB = matvec(sB, vbB);
    if (owner) { zc[(size_t)(M - 1 - k) * NB + row] = zT; if (hasB) zc[(size_t)(M + 1 + k) * NB + row] = zB; }
    { const int kn = k + D - 1; stage(ringT, kn, M - 1 - kn, zc); stage(ringB, kn, (kn < nB) ? M + 1 + kn : -1, zc); cp_async_commit(); }
    if ((k & 15) == 15 || k == 7) {                              // the first k + 1 aggregate rows of both fronts are final
      __threadfence();
      __syncwarp();
      if (lane == 0) *prog = (unsigned)(k + 1);
    }
  }
  cp_async_wait<0>();
  __threadfence();
  __syncwarp();
  if (lane == 0) *prog = kCoarseDone;
  return 0.0;
}

// coarse_begin: everything of the solve that r . z needs (the whole solve in the plain form; inward phase + meeting row in the
// twisted form) -- returns rc . zc.  coarse_finish: the rest (outward phase), after the caller has published the F_UP tasks.
template <int RING_BYTES>
__device__ __forceinline__ double coarse_begin(const PcgArgs<float>& a, unsigned char* ringmem, int b, int lane, CoarseMid* mid, bool& pending,
                                               bool overlap, bool speculative = false) {
  const CoarseGeom g = a.cg;
  const float* rec = a.Wt + (size_t)b * g.nby * g.rec;
  float* rc = a.rc + (size_t)b * g.nc;
  float* zc = a.zc + (size_t)b * g.nc;
  volatile unsigned* prog = &a.st[b].zprog;
  // speculative: rc is still being produced by the image's F_DOWN tasks (arr = their arrival counter)
  const volatile unsigned* arr = &a.st[b].cnt1;
  const int per = a.per;
  pending = false;
  if (g.tw >= 0) {
    // (32 x 32 blocks need 5 KB per staged block row: two fronts fit only the deeper rings of the 4-warp kernel;
    //  the host leaves cg.tw = -1 otherwise)
    if (g.nb == 16 || g.nb == 8 || RING_BYTES >= 6 * (int)sizeof(CoarseSlot<32>) + 512) {
      if (lane == 0) *prog = 0u;                 // (the F_UP tasks that will read it are published after phase 0)
      double d = 0.0;
      if (g.nb == 16) d = coarse_task_tw_nb<16, RING_BYTES>(g, rec, rc, zc, ringmem, lane, 0, mid, prog, speculative, per, arr);
      else if (g.nb == 8) d = coarse_task_tw_nb<8, RING_BYTES>(g, rec, rc, zc, ringmem, lane, 0, mid, prog, speculative, per, arr);
      else if constexpr (RING_BYTES >= 6 * (int)sizeof(CoarseSlot<32>) + 512) d = coarse_task_tw_nb<32, RING_BYTES>(g, rec, rc, zc, ringmem, lane, 0, mid, prog, speculative, per, arr);
      // Bandwidth-bound launches finish the solve before anything is published: a sweep that waits for zc would hold a worker
      // warp other tasks could use (512 images of 512 x 512: 990 -> 917 img/s with the overlap).
      if (overlap) { pending = true; return d; }
      if (g.nb == 16) coarse_task_tw_nb<16, RING_BYTES>(g, rec, rc, zc, ringmem, lane, 1, mid, prog, false, per, arr);
      else if (g.nb == 8) coarse_task_tw_nb<8, RING_BYTES>(g, rec, rc, zc, ringmem, lane, 1, mid, prog, false, per, arr);
      else if constexpr (RING_BYTES >= 6 * (int)sizeof(CoarseSlot<32>) + 512) coarse_task_tw_nb<32, RING_BYTES>(g, rec, rc, zc, ringmem, lane, 1, mid, prog, false, per, arr);
      return d;
    }
  }
  double d;
  if (g.nb == 16) d = coarse_task_nb<16, RING_BYTES>(g, rec, rc, zc, ringmem, lane);
  else if (g.nb == 32) d = coarse_task_nb<32, RING_BYTES>(g, rec, rc, zc, ringmem, lane);
  else d = coarse_task_nb<8, RING_BYTES>(g, rec, rc, zc, ringmem, lane);
  if (lane == 0) *prog = kCoarseDone;            // complete before anything is published (the caller fences)
  return d;
}

template <int RING_BYTES>
__device__ __forceinline__ void coarse_finish(const PcgArgs<float>& a, unsigned char* ringmem, int b, int lane, CoarseMid* mid) {
  const CoarseGeom g = a.cg;
  const float* rec = a.Wt + (size_t)b * g.nby * g.rec;
  float* rc = a.rc + (size_t)b * g.nc;
  float* zc = a.zc + (size_t)b * g.nc;
  volatile unsigned* prog = &a.st[b].zprog;
  if (g.nb == 16) coarse_task_tw_nb<16, RING_BYTES>(g, rec, rc, zc, ringmem, lane, 1, mid, prog, false, a.per, prog);
  else if (g.nb == 8) coarse_task_tw_nb<8, RING_BYTES>(g, rec, rc, zc, ringmem, lane, 1, mid, prog, false, a.per, prog);
  else if constexpr (RING_BYTES >= 6 * (int)sizeof(CoarseSlot<32>) + 512) coarse_task_tw_nb<32, RING_BYTES>(g, rec, rc, zc, ringmem, lane, 1, mid, prog, false, a.per, prog);
}

// ---- ready queue (two multi-producer / multi-consumer rings in global memory) ----------------------
// Only the strips of ONE image have to meet between phases, so the fused kernel has no grid-wide
// barrier: the last-arriving strip of an image publishes the image's next phase as nstrip4 new
// tasks, and any idle warp pops the next task.  Work-conserving: no SM waits for a slower SM.
//
// Ring 1 is served before ring 0.  An image publishes into ring 1 while its predicted remaining
// iteration count is above the mean over the running images (update_priority), so slowly converging
// images advance at chain-latency speed while the memory system is kept busy by the rest, and all images
// finish at about the same time instead of leaving a latency-bound tail of stragglers.
// Capacity invariant: a ring entry is rewritten only after tail has advanced by qcap.  Unclaimed entries are bounded by
// the tasks that can be ready at once (<= per * B, every image has one phase in flight), claimed-but-unread tickets by
// the number of worker warps (<= 8 * 148); qcap = 2 * per * B + 1024 >= per * B + workers for every launch the
// host makes (launch_fused_aw never starts more worker warps than tasks), so a reader is never lapped.
// Slot k of a ring holds ((k + 1) << 32 | payload).  qavail[ring] counts entries completely written and not
// yet claimed: a consumer decrements it and, only if it was positive, takes ticket k = head++ -- fetch-adds
// throughout, so claims pipeline in L2 (a CAS on the head serialises the pops at one per round trip:
// measured 1 us per task, 3.5x slower than this at 512 images).
constexpr unsigned kNoTask = 0xffffffffu;
constexpr unsigned kSpecTask = 0x80000000u;   // payload flag: the speculative coarse solve of image (payload & ~flag), see pcg_fused_kernel
constexpr unsigned kIdleSleepMaxNs = 2048;
constexpr unsigned long long kWatchdogNs = 4000000000ull;   // default: idle workers give up when NOBODY has claimed a task for 4 s

// pushes n entries: payload0, payload0+stride, ...  (one lane)
// (extra != kNoTask: one more entry with that payload, behind the n others)
__device__ __forceinline__ void q_push_n(const PcgArgs<float>& a, int ring, unsigned payload0, unsigned stride, unsigned n,
                                         unsigned extra = kNoTask) {
  const unsigned nx = n + (extra != kNoTask ? 1u : 0u);
  const unsigned long long k0 = atomicAdd(a.qtail + ring, (unsigned long long)nx);
  __threadfence();                                          // release: state written before the stamps
  unsigned long long* slots = a.qslots + (size_t)ring * a.qcap;
  for (unsigned i = 0; i < n; ++i) {
    const unsigned long long k = k0 + i;
    // the ring never laps an unread entry (signed: another producer's count may already have let consumers claim tickets
    // beyond this one, they then wait for this entry's stamp)
    USAUG_DEV_ASSERT((long long)(k - __ldcg(a.qhead + ring)) < (long long)a.qcap);
    volatile unsigned long long* slot = slots + (k % a.qcap);
    *slot = ((k + 1ull) << 32) | (unsigned long long)(payload0 + i * stride);
  }
  if (extra != kNoTask) {
    const unsigned long long k = k0 + n;
    volatile unsigned long long* slot = slots + (k % a.qcap);
    *slot = ((k + 1ull) << 32) | (unsigned long long)extra;
  }
  __threadfence();
  atomicAdd(a.qavail + ring, (int)nx);
}

// Bail-out of a stalled scheduler (one lane): every image that is not DONE reports iters = -1, relres = NaN; its
// confidence map stays at the neutral 0 the C entry point initialised it with.  Then every worker is released.
__device__ __forceinline__ void q_give_up(const PcgArgs<float>& a) {
  if (atomicExch(a.done + 1, 1u) == 0u) {
    for (int b = 0; b < a.B; ++b)
      if ((int)(__ldcg(&a.st[b].fword) & 0xffffffffull) != F_DONE) {
        if (a.iters_out) a.iters_out[b] = -1;
        if (a.relres_out) a.relres_out[b] = __longlong_as_double(0x7ff8000000000000ll);
      }
    __threadfence();
  }
  atomicMax(a.done, (unsigned)a.B);
}

// Whole warp.  Returns the payload of the next task, or kNoTask when every image is finished.
// ALL lanes execute the polling loads (a warp with 31 lanes parked at a convergence barrier while one lane
// spins was measured to slow the other warps of the SM ~4x); lane 0's values are broadcast and lane 0 claims.
// sleep_max: longest back-off of an idle worker between polls (latency-bound launches poll eagerly: every hand-off of an
// image's phase to the next sits on its critical path, and their idle workers have nothing else to do)
__device__ __forceinline__ unsigned q_pop(const PcgArgs<float>& a, int lane, unsigned sleep_max) {
  unsigned ns = 32, polls = 0;
  unsigned long long t_idle = 0, last_claimed = 0;
  for (;;) {
    int av0, av1;
    asm volatile("ld.volatile.global.v2.s32 {%0, %1}, [%2];" : "=r"(av0), "=r"(av1) : "l"(a.qavail) : "memory");
    int ring = (av1 > 0) ? 1 : (av0 > 0) ? 0 : -1;
    ring = __shfl_sync(0xffffffffu, ring, 0);
    if (ring >= 0) {
      unsigned long long k = ~0ull;
      if (lane == 0) {
        if (atomicAdd(a.qavail + ring, -1) > 0) k = atomicAdd(a.qhead + ring, 1ull);
        else atomicAdd(a.qavail + ring, 1);                  // lost the race for the last entries: give it back
      }
      k = __shfl_sync(0xffffffffu, k, 0);
      if (k == ~0ull) continue;
      // the entry is reserved for this warp; its producer may still be writing it (another producer's count came first)
      volatile unsigned long long* slot = a.qslots + (size_t)ring * a.qcap + (k % a.qcap);
      const unsigned stamp = (unsigned)(k + 1ull);
      unsigned long long v;
      unsigned spins = 0;
      while ((unsigned)((v = *slot) >> 32) != stamp) {
        __nanosleep(32);
        if (++spins > (1u << 26)) { if (lane == 0) q_give_up(a); return kNoTask; }
      }
      __threadfence();                                       // acquire
      return (unsigned)(v & 0xffffffffull);
    }
    if (__ldcg(a.done) >= (unsigned)a.B) return kNoTask;
    __nanosleep(ns);
    if (ns < sleep_max) { ns *= 2; continue; }
    if ((++polls & 63u) != 0u) continue;                     // (the watchdog's own loads would halve the polling rate)
    // watchdog: a scheduling bug must not hang the device.  It watches GLOBAL progress (tasks claimed by anybody),
    // not this worker's own luck: with one slowly converging image left, a worker can lose the race for its few
    // tasks for a long time while the solve is perfectly healthy.
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned long long claimed = *(volatile unsigned long long*)(a.qhead) + *(volatile unsigned long long*)(a.qhead + 1);
    if (t_idle == 0 || claimed != last_claimed) { t_idle = t; last_claimed = claimed; }
    else if (t - t_idle > a.watchdog_ns) {
      if (lane == 0) q_give_up(a);
      __syncwarp();
      return kNoTask;
    }
  }
}

// ring 0 starts with the first phase (RESID) of every image, image index fastest; ring 1 starts empty
__global__ void cm_init_queue_kernel(unsigned long long* __restrict__ slots, unsigned cap, int B, int nstrip4,
                                     unsigned long long* __restrict__ head, unsigned long long* __restrict__ tail,
                                     int* __restrict__ avail, double* __restrict__ qsum, unsigned* __restrict__ qcnt) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned T = (unsigned)(B * nstrip4);
  if (i == 0) {
    head[0] = 0ull; head[1] = 0ull; tail[0] = (unsigned long long)T; tail[1] = 0ull;
    avail[0] = (int)T; avail[1] = 0; *qsum = 0.0; *qcnt = 0u;
  }
  if (i >= 2u * cap) return;
  unsigned long long v = 0ull;
  if (i < T) {
    const unsigned b = i % (unsigned)B, ti = i / (unsigned)B;
    v = ((unsigned long long)(i + 1u) << 32) | (unsigned long long)(b * (unsigned)nstrip4 + ti);
  }
  slots[i] = v;
}

// One lane, every 8th iteration of an image: predict the iterations it still needs from the fraction of the
// log-residual distance to rtol covered so far, keep the running sum over images current, pick the ring.
__device__ __forceinline__ void update_priority(const PcgArgs<float>& a, ImgState* st, int it, double rr) {
  const double bb = __ldcg(&st->bb);
  if (!(bb > 0.0) || !(rr > 0.0)) return;
  const float covered = __log2f((float)(rr / bb)), whole = __log2f((float)a.rtol2);   // both negative
  const float phi = fminf(fmaxf(covered / whole, 0.02f), 1.f);
  const float rem = (float)it * (1.f - phi) / phi;
  float old = 0.f;
  if (__ldcg(&st->counted)) old = __ldcg(&st->rem);
  else { st->counted = 1; atomicAdd(a.qcnt, 1u); }
  st->rem = rem;
  const double sum = atomicAdd(a.qsum, (double)rem - (double)old) + ((double)rem - (double)old);
  const unsigned cnt = __ldcg(a.qcnt);
  st->prio = ((double)rem * (double)(cnt ? cnt : 1u) > sum) ? 1 : 0;
}
__device__ __forceinline__ void leave_priority(const PcgArgs<float>& a, ImgState* st) {
  if (!__ldcg(&st->counted)) return;
  st->counted = 0;
  atomicAdd(a.qsum, -(double)__ldcg(&st->rem));
  atomicSub(a.qcnt, 1u);
}

// fp64 residual r = b - A x64 over one 128-column strip (rare: once per refinement round); with the twisted
// factorisation the two half tasks of a strip split the row blocks
__device__ __forceinline__ double resid_strip(const PcgArgs<float>& a, int b, int strip, int half, int lane) {
  const int x0 = strip * kStrip4, xe = min(x0 + kStrip4, a.W);
  const int r0 = (a.tw_m && half) ? a.nrb / 2 : 0, r1 = (a.tw_m && !half) ? a.nrb / 2 : a.nrb;
  double s = 0.0;
  for (int xb = x0; xb < xe; xb += kStripCols)
    for (int rbk = r0; rbk < r1; ++rbk)
      s += stencil_task<float, true>(a, b, xb, min(kStripCols, xe - xb), rbk, lane, 0.f, true);
  return s;
}

// Twisted factorisation: after F_DOWN / F_PRECOND the h plane holds, in row m, only the upper half's part
// r[m] + c[m-1] h[m-1].  The last arriver of the phase (one warp) adds the lower half's term c[m+1] h[m+1], stores the
// complete row and returns the row's share of r.z = sum ip h^2.
__device__ __forceinline__ double twist_fix(const PcgArgs<float>& a, int b, int lane) {
  const int W = a.W, m = a.tw_m, hp = h_pitch(W);
  const size_t o = (size_t)b * a.H * W + (size_t)m * W;
  uint16_t* Hm = reinterpret_cast<uint16_t*>(a.z) + ((size_t)b * a.H + m) * hp;       // bf16 h plane, rows m and m + 1
  const uint16_t* Hn = Hm + hp;
  const uint32_t* Fm = a.ipc + o;
  const uint32_t* Fn = a.ipc + o + W;
  double acc = 0.0;
  // four chunks per round with all their loads issued first: the warp that runs this is the last arriver of the phase,
  // i.e. it sits on the image's critical path, and a round costs one L2 round trip instead of four
  for (int xb = 4 * lane; xb < W; xb += 512) {
    uint2 ht[4], hb[4]; uint4 fm[4], fn[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int x = xb + 128 * u;
      if (x < W) {
        ht[u] = __ldcg(reinterpret_cast<const uint2*>(Hm + x)); hb[u] = __ldcg(reinterpret_cast<const uint2*>(Hn + x));
        fm[u] = __ldcg(reinterpret_cast<const uint4*>(Fm + x)); fn[u] = __ldcg(reinterpret_cast<const uint4*>(Fn + x));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int x = xb + 128 * u;
      if (x < W) {
        const float h0 = fmaf(bf16_hi(fn[u].x), bf16_lo(hb[u].x), bf16_lo(ht[u].x)), h1 = fmaf(bf16_hi(fn[u].y), bf16_hi(hb[u].x), bf16_hi(ht[u].x));
        const float h2 = fmaf(bf16_hi(fn[u].z), bf16_lo(hb[u].y), bf16_lo(ht[u].y)), h3 = fmaf(bf16_hi(fn[u].w), bf16_hi(hb[u].y), bf16_hi(ht[u].y));
        const uint32_t w0 = pack_bf16x2(h0, h1), w1 = pack_bf16x2(h2, h3);
        *reinterpret_cast<uint2*>(Hm + x) = make_uint2(w0, w1);
        acc += (double)fmaf(bf16_lo(fm[u].x) * h0, bf16_lo(w0), fmaf(bf16_lo(fm[u].y) * h1, bf16_hi(w0),
                       fmaf(bf16_lo(fm[u].z) * h2, bf16_lo(w1), bf16_lo(fm[u].w) * h3 * bf16_hi(w1))));
      }
    }
  }
  if (a.h4 != nullptr) {                                            // the same for the level preconditioner's plane
    const size_t o4 = ((size_t)b * a.H + m) * a.w4p;
    const int W4 = W >> 2;
    for (int Xb = lane; Xb < W4; Xb += 128) {
      float hn[4], hm[4]; uint32_t gn[4], gm[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int X = Xb + 32 * u;
        if (X < W4) {
          gn[u] = __ldcg(a.ipc4 + o4 + a.w4p + X); hn[u] = __ldcg(a.h4 + o4 + a.w4p + X);
          hm[u] = __ldcg(a.h4 + o4 + X); gm[u] = __ldcg(a.ipc4 + o4 + X);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int X = Xb + 32 * u;
        if (X < W4) {
          const float h = fmaf(bf16_hi(gn[u]), hn[u], hm[u]);
          a.h4[o4 + X] = h;
          acc += (double)(bf16_lo(gm[u]) * h * h);
        }
      }
    }
  }
  return warp_sum(acc);
}

// AW = active warps per CTA (8 or 4): fewer active warps own deeper rings.
// RECOMP = F_UP recomputes the edge weights from the `a` plane (bandwidth-bound regime) or reads them.
// SINK = the launch has extra sink pixels (a.snk): F_UP masks p and A p there (a separate instantiation, so that the
// masks cost the default path nothing -- as a run-time branch they lengthened an F_UP task by 8 %).
template <int AW, bool RECOMP, bool SINK>
__global__ void __launch_bounds__(256, 1) pcg_fused_kernel(const PcgArgs<float> a) {
  constexpr size_t kWarpRing = (kFusedSmem / AW) & ~(size_t)127;  // ring bytes per active warp
  constexpr int R = (int)(kWarpRing / (RECOMP ? sizeof(RingU) : sizeof(RingUP)));   // F_UP ring depth in rows
  constexpr int RD = (int)(kWarpRing / sizeof(RingD));                              // F_DOWN ring depth
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= AW || warp >= a.wpb) return;                        // no block-wide barrier anywhere below
  unsigned char* ringmem = smem_raw + (size_t)warp * kWarpRing;
  const int H = a.H, W = a.W;
  const int per = a.per;                                          // tasks per image and phase: strips x halves

  auto now = []() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
  unsigned long long t_wait = 0, t_mode[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t_arr = 0;
  unsigned n_mode[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (;;) {
    const unsigned long long tq0 = a.debug ? now() : 0ull;
    const unsigned payload = q_pop(a, lane, RECOMP ? kIdleSleepMaxNs : 128u);
    const unsigned long long tq1 = a.debug ? now() : 0ull;
    t_wait += tq1 - tq0;
    if (payload == kNoTask) break;
    // A speculative coarse task (a.spec) travels with the F_DOWN / F_PRECOND tasks of its image.  It runs the inward phase of the
    // coarse solve BEHIND those sweeps (they flush the restriction rc from the outside in, the order the two fronts consume
    // it), waits for all of them to arrive, and then arrives itself as the (per + 1)-th and, by construction, LAST task of the
    // phase: it reduces their partial sums, advances the image's state and publishes F_UP like any last arriver -- with
    // rc . zc already in hand, so the ~45 us of the inward phase are off the image's critical path.
    const bool spec_task = (payload & kSpecTask) != 0u;
    const int b = spec_task ? (int)(payload & ~kSpecTask) : (int)(payload / (unsigned)per);
    const int ti = spec_task ? 0 : (int)(payload % (unsigned)per);
    USAUG_DEV_ASSERT(b >= 0 && b < a.B && 2 * ti + 1 < 2 * a.pmax);
    const int strip = ti % a.nstrip4, half = ti / a.nstrip4;      // half = 1: lower half of a twisted factorisation
    ImgState* st = a.st + b;
    const int mode = (int)(__ldcg(&st->fword) & 0xffffffffull);
    double s0 = 0.0, s1 = 0.0;
    CoarseMid cmid;                       // a coarse solve this warp has begun in this round and must finish after publishing
    bool coarse_pending = false;
    double spec_d = 0.0;                  // rc . zc of a speculative coarse solve
    if (spec_task) {
      USAUG_DEV_ASSERT(mode == F_DOWN || mode == F_PRECOND);
      spec_d = coarse_begin<(int)kWarpRing>(a, ringmem, b, lane, &cmid, coarse_pending, true, true);
    } else if (mode == F_UP) {
      if constexpr (RECOMP)
        s0 = up_task<R, SINK>(a, ringmem, b, strip, half, lane, (float)__ldcg(&st->beta), __ldcg(&st->first) != 0, __ldcg(&st->pcur));
      else
        s0 = up_task_planes<R, SINK>(a, ringmem, b, strip, half, lane, (float)__ldcg(&st->beta), __ldcg(&st->first) != 0, __ldcg(&st->pcur));
    } else if (mode == F_DOWN) {
      if (__ldcg(&st->epend))
        down_task<RD, 2>(a, ringmem, b, strip, half, lane, (float)__ldcg(&st->alpha), (float)__ldcg(&st->alpha_prev), __ldcg(&st->pcur), s0, s1);
      else
        down_task<RD, 1>(a, ringmem, b, strip, half, lane, (float)__ldcg(&st->alpha), 0.f, __ldcg(&st->pcur), s0, s1);
    } else if (mode == F_PRECOND) {
      down_task<RD, 0>(a, ringmem, b, strip, half, lane, 0.f, 0.f, 0, s0, s1);
    } else if (mode == F_RESID) {
      s0 = resid_strip(a, b, strip, half, lane);
    } else if (mode == F_COARSE) {
      s0 = coarse_begin<(int)kWarpRing>(a, ringmem, b, lane, &cmid, coarse_pending, !RECOMP);
    } else {
      const int x = strip * kStrip4 + 4 * lane;
      // rows of this task: all of them, or one side of row m
      const int ylo = (a.tw_m && half) ? a.tw_m + 1 : 0, yhi = (a.tw_m && !half) ? a.tw_m : H - 1;
      if (x < W) {
        const size_t base = (size_t)b * H * W + x;
        if (mode == F_FOLD) {
          // x64 += e (+ the e-update of the last iteration if it is still pending)
          const float ap = __ldcg(&st->epend) ? (float)__ldcg(&st->alpha_prev) : 0.f;
          const float* Pl = (__ldcg(&st->pcur) ? a.p2 : a.p);
#pragma unroll 4
          for (int y = max(ylo, 1); y <= min(yhi, H - 2); ++y) {
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
          for (int y = ylo; y <= yhi; ++y) {
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
    if (lane == 0 && !spec_task) { part[2 * ti] = s0; part[2 * ti + 1] = s1; }
    __threadfence();
    __syncwarp();
    unsigned tk = 0;
    if (lane == 0) tk = atomicAdd(&st->cnt1, 1u);
    tk = __shfl_sync(0xffffffffu, tk, 0);
    if (a.debug) t_arr += now() - tq2;
    // F_COARSE is a single task per image; a phase with a speculative coarse task has per + 1 participants, and that task
    // arrives last (it has waited for the others)
    const bool spec_phase = a.spec && (mode == F_DOWN || mode == F_PRECOND);
    const int ntask = (mode == F_COARSE) ? 1 : per + (spec_phase ? 1 : 0);
    USAUG_DEV_ASSERT(!spec_phase || (tk == (unsigned)per) == spec_task);
    if (tk != (unsigned)ntask - 1) continue;
    // ---- last strip of this image and phase: reduce, advance the state, publish the next phase ------
    __threadfence();
    reduce_parts(part, (mode == F_COARSE) ? 1 : per, lane, s0, s1);

    if (a.tw_m && (mode == F_DOWN || mode == F_PRECOND)) {
      s1 += twist_fix(a, b, lane);
      __threadfence();                                            // every lane publishes its part of row m
      __syncwarp();
    }
    int next = F_DONE;
    // the F_COARSE decision, also used after the inline coarse solve below
    auto after_coarse = [&](double rc_zc) {
      // z = T^-1 r + P Ac^-1 P^T r  =>  r.z = sum ip h^2 (line part, from DOWN) + rc.zc (the coarse solve)
      const double rz_new = __ldcg(&st->rz_line) + rc_zc, rz_old = __ldcg(&st->rz);
      if (__ldcg(&st->fresh)) { st->rz = rz_new; st->first = 1; st->beta = 0.0; return (int)F_UP; }
      if (rz_new > 0.0 && rz_old > 0.0) { st->beta = rz_new / rz_old; st->rz = rz_new; st->first = 0; return (int)F_UP; }
      return (int)F_FOLD;
    };
    if (lane == 0) {
    st->cnt1 = 0;
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
        if ((it & 7) == 0 && it >= 32) update_priority(a, st, it, s0);
      } else {
        st->rz = s1; st->first = 0;
        st->beta = (rz_old > 0.0) ? s1 / rz_old : 0.0;
        next = F_UP;
      }
    } else if (mode == F_PRECOND) {
      // Every refinement round restarts CG (p = z).  Keeping p across rounds ("residual replacement") was
      // measured to need 2300-3700 instead of 1600 iterations here and was dropped.
      st->epend = 0;                                               // F_FOLD consumed any pending update, e = 0 now
      if (a.cg.by) {
        st->rz_line = s1; st->fresh = 1;
        next = (s1 > 0.0) ? F_COARSE : F_FINAL;
      } else {
        st->rz = s1; st->first = 1; st->beta = 0.0;
        next = (s1 > 0.0) ? F_UP : F_FINAL;
      }
    } else if (mode == F_COARSE) {
      next = after_coarse(s0);
    } else if (mode == F_RESID) {
      next = after_true_residual(st, s0, a.rtol2, a.inner_red2, a.maxiter) ? F_FINAL : F_PRECOND;
    } else if (mode == F_FOLD) {
      next = F_RESID; st->outer = __ldcg(&st->outer) + 1;
    } else {   // F_FINAL
      const double bb = __ldcg(&st->bb);
      if (a.iters_out) a.iters_out[b] = __ldcg(&st->iters);
      if (a.relres_out) a.relres_out[b] = (bb > 0.0) ? sqrt(__ldcg(&st->rr_true) / bb) : 0.0;
      leave_priority(a, st);
      next = F_DONE;
    }
    }
    next = __shfl_sync(0xffffffffu, next, 0);
    if (!RECOMP && next == F_COARSE && spec_task) {
      // the inward phase has already run behind the sweeps
      __threadfence();
      __syncwarp();
      if (lane == 0) next = after_coarse(spec_d);
      next = __shfl_sync(0xffffffffu, next, 0);
    } else if (!RECOMP && next == F_COARSE) {
      // Latency-bound launches (few tasks, the !RECOMP kernel): the coarse solve is a single-warp task, so the warp
      // that completed the phase runs it right here instead of handing it to another worker through the queue -- one
      // hand-off less on the UP -> DOWN -> COARSE chain (+2 % at B <= 64).  Bandwidth-bound launches keep it in
      // the queue: there it is the two service classes that decide who runs next (inline: -11 % at B = 512).
      const unsigned long long tc0 = a.debug ? now() : 0ull;
      const double rc_zc = coarse_begin<(int)kWarpRing>(a, ringmem, b, lane, &cmid, coarse_pending, !RECOMP);
      __threadfence();                                          // every lane publishes its part of zc
      __syncwarp();
      if (a.debug) { t_mode[F_COARSE] += now() - tc0; n_mode[F_COARSE]++; }
      if (lane == 0) next = after_coarse(rc_zc);
      next = __shfl_sync(0xffffffffu, next, 0);
    }
    if (lane == 0) {
      st->fword = (unsigned long long)next;
      if (next == F_DONE) {
        __threadfence();
        atomicAdd(a.done, 1u);                                     // idle workers leave when done == B
      } else {
        q_push_n(a, __ldcg(&st->prio), (unsigned)(b * per), 1u, (next == F_COARSE) ? 1u : (unsigned)per,
                 (a.spec && (next == F_DOWN || next == F_PRECOND)) ? (kSpecTask | (unsigned)b) : kNoTask);
      }
    }
    __syncwarp();
    // a coarse solve begun above: its F_UP tasks are published, now produce the rest of zc behind which they sweep
    // (next != F_UP: the inner solve has ended -- FOLD does not read zc)
    if (coarse_pending && next == F_UP) {
      const unsigned long long tc0 = a.debug ? now() : 0ull;
      coarse_finish<(int)kWarpRing>(a, ringmem, b, lane, &cmid);
      if (a.debug) t_mode[F_COARSE] += now() - tc0;
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

}  // namespace usaug
