// K2: ultrasound confidence map -- random-walk potentials on the
// 8-neighbour image graph, solved per image with a matrix-free preconditioned conjugate gradient.
// This translation unit holds the workspace layout, the kernel launches and the C ABI; the device code is in
//   confmap_types.cuh    per-image state, kernel arguments, small device helpers
//   confmap_setup.cuh    normalisation, attenuation, edge weights, column-line factors, initial state
//   confmap_coarse.cuh   aggregation coarse grid: Galerkin operator, banded Cholesky
//   confmap_fused.cuh    production kernel (fp32 vectors + fp64 refinement, W % 4 == 0): ready-queue scheduled
//                        phases UP / DOWN / COARSE, cp.async rings, two-level preconditioner
//   confmap_generic.cuh  grid-synchronous kernel for fp64 vectors and any width (also the A/B reference)
#include "confmap_setup.cuh"
#include "confmap_fused.cuh"

namespace cg = cooperative_groups;

namespace usaug {

struct CmLayout {
  float *depth, *mm1, *mm2, *wS, *wE, *wSE, *wSW, *ip, *c, *aplane;
  float *wS2, *wSE2, *wSW2;   // mirrored weight planes (twisted factorisation + weight-reading F_UP only)
  uint8_t* snk;               // extra sink pixels (USAUG_CM_SINKS), else nullptr
  float* h4; uint32_t* ipc4; int w4p;   // level preconditioner planes (fused kernel, line preconditioner)
  bool fused, recomp;         // fused fp32 kernel; its F_UP recomputes the edge weights
  int tw_m, per;              // twisted line factorisation: meeting row (0 = off); tasks per image and phase
  uint32_t* ipc;
  float4* wk;
  double *x64, *bbpart, *part;
  bool spec;                  // the coarse solve runs behind the F_DOWN sweeps (latency-bound launches)
  void *e, *r, *p, *q, *z, *p2;
  ImgState* st;
  unsigned* done;
  unsigned long long *qslots, *qhead, *qtail;
  unsigned qcap;
  int* qavail;
  double* qsum;
  unsigned* qcnt;
  CoarseGeom cg;
  float *rc, *zc, *Wt;
  double* Ab;
  int nblk, nstrip1, nrb, nstrip2, pmax;
  size_t bytes;
  Carver cv{nullptr};         // (its red zones: bounds-checked debug build)
};

static CmLayout cm_layout(void* ws, int B, int H, int W, int flags) {
  CmLayout L;
  const size_t N = (size_t)B * H * W;
  const size_t tsz = (flags & USAUG_CM_FP64_VECTORS) ? 8 : 4;
  // ---- solver variant: a function of the launch's size only (flags can pin every choice) ------------------
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  const int nstrip4 = (W + kStrip4 - 1) / kStrip4;
  const long long tasks = (long long)B * nstrip4;               // 128-column strips in flight
  L.fused = !(flags & (USAUG_CM_FP64_VECTORS | USAUG_CM_LEGACY_KERNEL)) && W % 4 == 0;
  // few tasks: the single-warp instruction latency of a strip sets the pace, so F_UP reads the weight planes
  // (fewer instructions per row); many tasks: HBM sets the pace, so it recomputes them (fewer bytes)
  L.recomp = (flags & USAUG_CM_FORCE_RECOMPUTE) ? true : (flags & USAUG_CM_FORCE_PLANES) ? false : tasks >= 6ll * sms;
  // Aggregates of 2 rows instead of 4: ~16 % fewer iterations for a coarse solve twice as long.  Since the solve takes
  // 40 us per 128 aggregate rows (round 2) that pays at every batch size while the image has <= 256 aggregate rows of 2
  // (profiles/r3g_solver_variants_by_batch.log: +7 % at 64 images of 512 x 512, +12 % at 128 ... 384); taller images keep 4-row
  // aggregates unless the launch is deeply bandwidth-bound, where the solves hide behind other images' sweeps.
  const bool many = tasks >= 12ll * sms;
  const bool fine = (flags & USAUG_CM_COARSE_FINE) ? true : (flags & USAUG_CM_COARSE_STANDARD) ? false : (many || H - 2 <= 512);
  int min_shift = fine ? 1 : 2, min_bx_shift = 5;
  // Latency-bound launches of wide images: 64-column aggregates keep the blocks of the coarse operator at 16 x 16 -- the
  // single-warp solve is on every image's critical path there, and a 32 x 32 block row costs 2.6x a 16 x 16 one
  // (32 images of 1024 x 1024: solve 219 -> 84 us, 444 -> 462 iterations, 138 -> 160 img/s).  Bandwidth-bound launches keep
  // the finer grid: their solves overlap with other images' sweeps, and iterations are what they pay for.
  if (!L.recomp && W > 512 && W <= 1024) min_bx_shift = 6;
  if (const char* e = getenv("USAUG_CM_COARSE_SHIFT")) min_shift = atoi(e);          // development knobs
  if (const char* e = getenv("USAUG_CM_COARSE_BXSHIFT")) min_bx_shift = atoi(e);
  L.cg = (!L.fused || (flags & USAUG_CM_NO_COARSE)) ? CoarseGeom{} : coarse_geom(H, W, min_shift, min_bx_shift);
  // Twisted block factorisation of the coarse operator (two fronts per solve, advanced by one warp): everywhere except
  // for 32 x 32 blocks in the 8-warp kernel, whose rings are too shallow for two fronts
  if (L.cg.by && !(L.recomp && L.cg.nb == 32) && !getenv("USAUG_CM_COARSE_PLAIN")) L.cg.tw = L.cg.nby / 2;
  // Twisted line factorisation (two half sweeps per strip): halves the dependent UP -> DOWN chain of an image, which
  // is what bounds launches that cannot fill the GPU -- and with twice the tasks of half the length the ready queue also
  // balances bandwidth-bound launches better (512 images of 512 x 512: 848 -> 916 img/s), so it is always on.
  L.tw_m = 0;
  const bool twist = (flags & USAUG_CM_TWIST_OFF) ? false : true;
  if (L.fused && twist && H - 2 >= 32) {
    const int by = L.cg.by ? L.cg.by : 1;
    const int m = ((H - 1) / 2) & ~(by - 1);                    // on an aggregate-row boundary: rows 1..m | m+1..H-2
    if (m >= 2 && m <= H - 4) L.tw_m = m;
  }
  L.per = nstrip4 * (L.tw_m ? 2 : 1);
  L.nblk = (W + 127) / 128;
  L.nstrip1 = (W + kStripCols - 1) / kStripCols;
  L.nrb = (H - 2 + kRowBlock - 1) / kRowBlock;
  L.nstrip2 = (W + 31) / 32;
  L.pmax = L.nstrip1 * L.nrb > L.nstrip2 ? L.nstrip1 * L.nrb : L.nstrip2;
  if (L.pmax < 2 * L.nblk) L.pmax = 2 * L.nblk;                 // fused kernel: one partial per (strip, half) task
  Carver cv(ws);
  L.depth = cv.take<float>(H);
  L.mm1 = cv.take<float>((size_t)B * kChunks * 2);
  L.mm2 = cv.take<float>((size_t)B * kChunks * 2);
  L.wS = cv.take<float>(N); L.wE = cv.take<float>(N); L.wSE = cv.take<float>(N); L.wSW = cv.take<float>(N);
  // fp32 line factors: read by the generic kernel only (the fused kernel reads the packed bf16 plane ipc)
  L.ip = cv.take<float>(L.fused ? 0 : N); L.c = cv.take<float>(L.fused ? 0 : N); L.ipc = cv.take<uint32_t>(N);
  if (L.fused) { L.ip = nullptr; L.c = nullptr; }
  L.aplane = cv.take<float>(N); L.wk = cv.take<float4>(B);
  const size_t nmir = (L.tw_m && !L.recomp) ? N : 0;
  L.wS2 = cv.take<float>(nmir); L.wSE2 = cv.take<float>(nmir); L.wSW2 = cv.take<float>(nmir);
  if (!nmir) L.wS2 = L.wSE2 = L.wSW2 = nullptr;
  L.w4p = ((W >> 2) + 3) & ~3;
  const size_t n4 = (L.fused && !(flags & USAUG_CM_NO_LEVEL)) ? (size_t)B * H * L.w4p : 0;
  L.h4 = cv.take<float>(n4); L.ipc4 = cv.take<uint32_t>(n4);
  if (!n4) { L.h4 = nullptr; L.ipc4 = nullptr; }
  L.snk = cv.take<uint8_t>((flags & USAUG_CM_SINKS) ? N : 0);
  if (!(flags & USAUG_CM_SINKS)) L.snk = nullptr;
  L.x64 = cv.take<double>(N);
  L.e = cv.take<char>(N * tsz); L.r = cv.take<char>(N * tsz); L.p = cv.take<char>(N * tsz);
  L.q = cv.take<char>(N * tsz); L.z = cv.take<char>(N * tsz); L.p2 = cv.take<char>(N * tsz);
  L.bbpart = cv.take<double>((size_t)B * L.nblk);
  L.part = cv.take<double>((size_t)B * L.pmax * 2);
  // the coarse solve runs behind the F_DOWN sweeps (its inward phase consumes the restriction in the order they flush it) in
  // latency-bound launches with both twisted factorisations; see confmap_fused.cuh
  L.spec = L.fused && !L.recomp && L.tw_m > 0 && L.cg.by && L.cg.tw >= 0 && !getenv("USAUG_CM_NO_SPEC");
  L.st = cv.take<ImgState>(B);
  L.done = cv.take<unsigned>(2);
  // two ready rings; every image has at most `per` ready tasks at a time, in one ring or the other
  L.qcap = (unsigned)(2 * L.per * B + 1024);
  L.qslots = cv.take<unsigned long long>(2 * (size_t)L.qcap);
  L.qhead = cv.take<unsigned long long>(2); L.qtail = cv.take<unsigned long long>(2);   // separate 256-byte lines
  L.qavail = cv.take<int>(2);
  L.qsum = cv.take<double>(1); L.qcnt = cv.take<unsigned>(1);
  // coarse-grid correction (fused fp32 kernel only)
  const size_t nc = (size_t)L.cg.nc;
  L.rc = cv.take<float>((size_t)B * nc); L.zc = cv.take<float>((size_t)B * nc);
  L.Wt = cv.take<float>((size_t)B * L.cg.nby * L.cg.rec);
  L.Ab = cv.take<double>((size_t)B * nc * 5);
  L.bytes = cv.off;
  L.cv = cv;
  return L;
}

// Optional measurement hook: a (start, stop) cudaEvent_t pair recorded on the launching stream
// immediately around the persistent kernel (bench.py's roofline uses it).  Thread-local.
static thread_local cudaEvent_t g_ev_start = nullptr, g_ev_stop = nullptr;
static thread_local int g_debug = 0;

static thread_local int g_legacy = 0;   // USAUG_CM_LEGACY_KERNEL: force the two-slot kernel (A/B testing)

template <int AW, bool RECOMP, bool SINK>
static int launch_fused_aw(PcgArgs<float>& a, int blocks_cap_sms, cudaStream_t stream) {
  int per_sm = 0;
  USAUG_CUDA(cudaFuncSetAttribute(pcg_fused_kernel<AW, RECOMP, SINK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem));
  USAUG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_fused_kernel<AW, RECOMP, SINK>, 256, kFusedSmem));
  if (per_sm < 1) { set_error("pcg_fused_kernel cannot be resident"); return USAUG_E_CUDA; }
  const int tasks = (a.per + (a.spec ? 1 : 0)) * a.B;           // (+ one speculative coarse task per image)
  int blocks = tasks;
  if (blocks > blocks_cap_sms) blocks = blocks_cap_sms;
  if (blocks < 1) blocks = 1;
  // no more workers than tasks that can be ready at once: an idle worker only polls the queue
  a.wpb = (tasks + blocks - 1) / blocks;
  if (a.wpb > AW) a.wpb = AW;
  cm_init_queue_kernel<<<(2 * a.qcap + 255) / 256, 256, 0, stream>>>(a.qslots, a.qcap, a.B, a.per, a.qhead, a.qtail,
                                                                        a.qavail, a.qsum, a.qcnt);
  USAUG_CHECK_LAUNCH("cm_init_queue_kernel");
  void* params[] = {(void*)&a};
  if (g_ev_start) USAUG_CUDA(cudaEventRecord(g_ev_start, stream));
  // cooperative launch only for its co-residency guarantee: workers spin on the ready queue
  USAUG_CUDA(cudaLaunchCooperativeKernel((const void*)pcg_fused_kernel<AW, RECOMP, SINK>, dim3(blocks), dim3(256), params,
                                         kFusedSmem, stream));
  if (g_ev_stop) USAUG_CUDA(cudaEventRecord(g_ev_stop, stream));
  return USAUG_OK;
}

static int launch_fused(PcgArgs<float>& a, const CmLayout& L, int B, cudaStream_t stream) {
  (void)B;
  int dev = 0, sms = 0;
  USAUG_CUDA(cudaGetDevice(&dev));
  USAUG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // every inner iteration is two phases; a refinement round adds four
  a.max_steps = 6 * a.maxiter + 64;   // worst case: a refinement round (5 phases) after every iteration
  // few tasks (weight-reading F_UP): 4 active warps per CTA with rings twice as deep keep enough bytes in flight
  if (a.snk != nullptr) {
    if (L.recomp) return launch_fused_aw<8, true, true>(a, sms, stream);
    return launch_fused_aw<4, false, true>(a, sms, stream);
  }
  if (L.recomp) return launch_fused_aw<8, true, false>(a, sms, stream);
  return launch_fused_aw<4, false, false>(a, sms, stream);
}

template <typename T>
static int launch_pcg(const CmLayout& L, float* cm, int B, int H, int W, double rtol, int maxiter,
                      int* iters, double* relres, cudaStream_t stream) {
  PcgArgs<T> a;
  a.B = B; a.H = H; a.W = W; a.maxiter = maxiter;
  a.rtol2 = rtol * rtol;
  double inner_red = std::is_same<T, float>::value ? 1e-4 : 0.0;   // first refinement round; later rounds adapt
  if (const char* e = getenv("USAUG_CM_INNER_RED")) { if (inner_red > 0.0) inner_red = atof(e); }   // development knob
  a.inner_red2 = inner_red * inner_red;
  a.wS = L.wS; a.wE = L.wE; a.wSE = L.wSE; a.wSW = L.wSW; a.ip = L.ip; a.c = L.c; a.ipc = L.ipc; a.aplane = L.aplane; a.wk = L.wk;
  a.x64 = L.x64;
  a.e = (T*)L.e; a.r = (T*)L.r; a.p = (T*)L.p; a.q = (T*)L.q; a.z = (T*)L.z; a.p2 = (T*)L.p2;
  a.st = L.st; a.part = L.part; a.done = L.done;
  a.qslots = L.qslots; a.qhead = L.qhead; a.qtail = L.qtail; a.qcap = L.qcap; a.qavail = L.qavail; a.qsum = L.qsum; a.qcnt = L.qcnt;
  a.cg = L.cg; a.rc = L.rc; a.zc = L.zc; a.Wt = L.Wt; a.cm = cm; a.iters_out = iters; a.relres_out = relres;
  a.nstrip1 = L.nstrip1; a.nrb = L.nrb; a.nstrip2 = L.nstrip2; a.pmax = L.pmax;
  a.vec4 = (std::is_same<T, float>::value && W % 4 == 0) ? 1 : 0;
  a.nstrip4 = (W + kStrip4 - 1) / kStrip4;
  a.tw_m = L.tw_m; a.per = L.per; a.wS2 = L.wS2; a.wSE2 = L.wSE2; a.wSW2 = L.wSW2;
  a.h4 = L.h4; a.ipc4 = L.ipc4; a.w4p = L.w4p;
  a.snk = L.snk;
  a.spec = L.spec ? 1 : 0;
  // every inner iteration is one step; each refinement round adds <= 3 steps and needs >= 1 iteration
  a.max_steps = 4 * maxiter + 16;
  a.debug = g_debug;
  a.watchdog_ns = kWatchdogNs;
  if (const char* w = getenv("USAUG_CM_WATCHDOG_MS")) a.watchdog_ns = (unsigned long long)atoll(w) * 1000000ull;   // test hook

  if constexpr (std::is_same<T, float>::value) {
    if (L.fused) return launch_fused(a, L, B, stream);
  }
  // generic grid-synchronous kernel (fp64 vectors, W % 4 != 0, or USAUG_CM_LEGACY_KERNEL)
  int dev = 0, sms = 0, per_sm = 0;
  USAUG_CUDA(cudaGetDevice(&dev));
  USAUG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  USAUG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_persistent_kernel<T>, 256, 0));
  if (per_sm < 1) { set_error("pcg_persistent_kernel cannot be resident"); return USAUG_E_CUDA; }
  // enough warps for every column task, never more blocks than can be co-resident
  const int tasks1 = L.nstrip1 * L.nrb * B;
  const int want_warps = L.nstrip2 * B > tasks1 ? L.nstrip2 * B : tasks1;
  int blocks = (want_warps + 7) / 8;
  const int cap = sms * (per_sm > 4 ? 4 : per_sm);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  void* params[] = {(void*)&a};
  if (g_ev_start) USAUG_CUDA(cudaEventRecord(g_ev_start, stream));
  USAUG_CUDA(cudaLaunchCooperativeKernel((const void*)pcg_persistent_kernel<T>, dim3(blocks), dim3(256),
                                         params, 0, stream));
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
  return usaug_confmap_pcg_sinks(img, cm, B, H, W, alpha, beta, gamma, rtol, maxiter, precond, flags & ~USAUG_CM_SINKS,
                                 USAUG_SINK_BOTTOM_ROW, nullptr, iters, relres, ws, ws_bytes, stream_);
}

extern "C" int usaug_confmap_pcg_sinks(const float* img, float* cm, int B, int H, int W, float alpha,
                                       float beta, float gamma, double rtol, int maxiter, int precond,
                                       int flags, int sink_mode, const uint8_t* sink_mask, int* iters,
                                       double* relres, void* ws, size_t ws_bytes, void* stream_) {
  if (!img || !cm || !ws) { set_error("usaug_confmap_pcg: null pointer"); return USAUG_E_INVALID_ARG; }
  if (sink_mode != USAUG_SINK_BOTTOM_ROW && sink_mode != USAUG_SINK_BOTTOM_AND_SIDES && sink_mode != USAUG_SINK_MASK) {
    set_error("usaug_confmap_pcg_sinks: unknown sink mode %d", sink_mode); return USAUG_E_INVALID_ARG;
  }
  if (sink_mode == USAUG_SINK_MASK && !sink_mask) {
    set_error("usaug_confmap_pcg_sinks: USAUG_SINK_MASK needs a mask plane"); return USAUG_E_INVALID_ARG;
  }
  flags = (sink_mode == USAUG_SINK_BOTTOM_ROW) ? (flags & ~USAUG_CM_SINKS) : (flags | USAUG_CM_SINKS);
  if (B <= 0 || H < 3 || W <= 0 || (long long)H * W > (1ll << 30)) {
    set_error("usaug_confmap_pcg: bad shape B=%d H=%d W=%d (need H>=3)", B, H, W); return USAUG_E_INVALID_ARG;
  }
  if (!(rtol > 0.0) || maxiter < 0 || (precond != USAUG_PRECOND_JACOBI && precond != USAUG_PRECOND_LINE)) {
    set_error("usaug_confmap_pcg: bad solver options rtol=%g maxiter=%d precond=%d", rtol, maxiter, precond);
    return USAUG_E_INVALID_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(img) & 15u) || (reinterpret_cast<uintptr_t>(cm) & 15u) || (reinterpret_cast<uintptr_t>(ws) & 255u)) {
    set_error("usaug_confmap_pcg: img and cm must be 16-byte aligned, ws 256-byte aligned"); return USAUG_E_INVALID_ARG;
  }
  const size_t need = usaug_confmap_workspace_bytes(B, H, W, precond, flags);
  if (ws_bytes < need) { set_error("usaug_confmap_pcg: workspace %zu < %zu", ws_bytes, need); return USAUG_E_WORKSPACE_TOO_SMALL; }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CmLayout L = cm_layout(ws, B, H, W, flags);
  const int N = H * W;
  g_debug = (flags & USAUG_CM_DEBUG_TIMING) ? 1 : 0;
  g_legacy = (flags & USAUG_CM_LEGACY_KERNEL) ? 1 : 0;

  if (int zrc = L.cv.arm(stream)) return zrc;
  // neutral result for images the solver could not finish (scheduler watchdog): C = 0 switches the SNR term off
  USAUG_CUDA(cudaMemsetAsync(cm, 0, (size_t)B * N * sizeof(float), stream));
  cm_depth_kernel<<<(H + 127) / 128, 128, 0, stream>>>(L.depth, H, alpha);
  USAUG_CHECK_LAUNCH("cm_depth_kernel");
  cm_minmax_kernel<<<dim3(kChunks, B), 256, 0, stream>>>(img, N, L.mm1);
  USAUG_CHECK_LAUNCH("cm_minmax_kernel");
  cm_delta_kernel<<<dim3(kChunks, B), 256, 0, stream>>>(img, H, W, L.depth, L.mm1, L.mm2);
  USAUG_CHECK_LAUNCH("cm_delta_kernel");
  cm_wconst_kernel<<<(B + 127) / 128, 128, 0, stream>>>(B, L.mm2, beta, gamma, L.wk);
  USAUG_CHECK_LAUNCH("cm_wconst_kernel");
  cm_weights_kernel<<<dim3((N + 255) / 256, B), 256, 0, stream>>>(img, H, W, L.depth, L.mm1, L.wk, L.aplane,
                                                                 L.wS, L.wE, L.wSE, L.wSW, L.wS2, L.wSE2, L.wSW2);
  USAUG_CHECK_LAUNCH("cm_weights_kernel");
  if (L.snk) {
    cm_sink_kernel<<<dim3((N + 255) / 256, B), 256, 0, stream>>>(H, W, sink_mode, sink_mask, L.snk);
    USAUG_CHECK_LAUNCH("cm_sink_kernel");
  }
  cm_factor_kernel<<<dim3(L.nblk, B), 128, 0, stream>>>(H, W, L.wS, L.wE, L.wSE, L.wSW, precond,
                                                        (flags & USAUG_CM_ZERO_GUESS) ? 1 : 0, L.tw_m, L.ip, L.c,
                                                        L.ipc, L.x64, L.bbpart, L.snk);
  USAUG_CHECK_LAUNCH("cm_factor_kernel");
  // fused kernel: the bf16 h plane lives in the z allocation with its rows padded to 8 columns; the padding is never
  // written, but the lanes beyond the last column read it (and pass it on to their neighbour's stencil with weight 0)
  if (L.fused && (W & 7)) USAUG_CUDA(cudaMemsetAsync(L.z, 0, (size_t)N * B * sizeof(float), stream));
  if (L.h4 && precond != USAUG_PRECOND_LINE) { L.h4 = nullptr; L.ipc4 = nullptr; }   // the level solve belongs to the line preconditioner
  if (L.h4) {
    // padding columns and seed rows of the level planes must read as zero
    USAUG_CUDA(cudaMemsetAsync(L.h4, 0, (size_t)B * H * L.w4p * sizeof(float), stream));
    USAUG_CUDA(cudaMemsetAsync(L.ipc4, 0, (size_t)B * H * L.w4p * sizeof(uint32_t), stream));
    cm_factor4_kernel<<<dim3(((W >> 2) + 127) / 128, B), 128, 0, stream>>>(H, W, L.w4p, L.wS, L.wE, L.wSE, L.wSW, L.tw_m, L.ipc4, L.snk);
    USAUG_CHECK_LAUNCH("cm_factor4_kernel");
  }
  if (L.cg.by) {
    coarse_assemble_kernel<<<dim3((L.cg.nc + 7) / 8, B), 256, 0, stream>>>(H, W, L.cg, L.wS, L.wE, L.wSE, L.wSW, L.Ab, L.snk);
    USAUG_CHECK_LAUNCH("coarse_assemble_kernel");
    if (L.cg.nb == 8) coarse_factor_kernel<8><<<B, 128, 0, stream>>>(L.cg, L.Ab, L.Wt);
    else if (L.cg.nb == 16) coarse_factor_kernel<16><<<B, 128, 0, stream>>>(L.cg, L.Ab, L.Wt);
    else coarse_factor_kernel<32><<<B, 128, 0, stream>>>(L.cg, L.Ab, L.Wt);
    USAUG_CHECK_LAUNCH("coarse_factor_kernel");
    // the padding columns of the coarse vectors are never written by the sweeps
    USAUG_CUDA(cudaMemsetAsync(L.rc, 0, (size_t)B * L.cg.nc * sizeof(float), stream));
    USAUG_CUDA(cudaMemsetAsync(L.zc, 0, (size_t)B * L.cg.nc * sizeof(float), stream));
  }
  if (L.spec) {
    const size_t n = (size_t)B * L.cg.nc;
    cm_rc_poison_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(L.rc, n, L.cg.nb, L.cg.nbx);
    USAUG_CHECK_LAUNCH("cm_rc_poison_kernel");
  }
  cm_init_state_kernel<<<(B + 127) / 128, 128, 0, stream>>>(B, L.nblk, L.bbpart, L.st, L.done);
  USAUG_CHECK_LAUNCH("cm_init_state_kernel");
  int rc = (flags & USAUG_CM_FP64_VECTORS) ? launch_pcg<double>(L, cm, B, H, W, rtol, maxiter, iters, relres, stream)
                                          : launch_pcg<float>(L, cm, B, H, W, rtol, maxiter, iters, relres, stream);
  if (rc == USAUG_OK) rc = L.cv.verify(stream, "usaug_confmap_pcg");
  return rc;
}
