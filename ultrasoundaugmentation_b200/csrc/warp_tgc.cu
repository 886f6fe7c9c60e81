// K1: bone-surface scan -> profile smoothing -> fused TGC + probe-pressure deformation + remap.
// Implements SURVEY.md Appendix A.1 (steps 1-5) and A.2.  The coordinate path uses explicitly
// rounded fp32 intrinsics (no FMA contraction) so nearest-neighbour label picks are bit-exact
// against the oracle (oracle/deform.py).
#include <stdlib.h>

#include <type_traits>

#include "kernels.cuh"

namespace usaug {

// ---------------------------------------------------------------------------------------------
// Surface scan: every label byte is read exactly once, fully parallel (no early-exit chains); rows whose
// bytes are all zero are skipped with one test per load.
// Block = 256 threads = 8 warps; a block covers 32*VEC columns x kScanRows rows; lane owns VEC
// adjacent columns, warps interleave rows.  Per-column first/last bone row inside the tile is
// merged through shared memory and written as one partial per (tile, column): no memset, no atomics.
// ---------------------------------------------------------------------------------------------
constexpr int kScanRows = 64;

template <int VEC>
__global__ void __launch_bounds__(256) surface_scan_kernel(const uint8_t* __restrict__ lab, int H,
                                                           int W, int* __restrict__ surf) {
  pdl_trigger();                                       // the consumer of surf may start launching (it pdl_wait()s)
  constexpr int PITCH = VEC + (VEC > 1 ? 1 : 0);      // odd lane pitch: the VEC stores of a lane hit distinct banks
  __shared__ int s_first[8][32 * PITCH];
  __shared__ int s_last[8][32 * PITCH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.z;
  const int x0 = (blockIdx.x * 32 + lane) * VEC;
  const int y0 = blockIdx.y * kScanRows;
  const uint8_t* L = lab + (size_t)b * H * W;
  int first[VEC], last[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { first[v] = kNone; last[v] = -1; }
  if (x0 < W) {
    constexpr int NW = (VEC + 3) / 4;                 // 32-bit words per load: 1 (VEC = 1, 4), 2 (VEC = 8) or 4 (VEC = 16: one 128-bit load)
    uint32_t vals[kScanRows / 8][NW];
#pragma unroll
    for (int i = 0; i < kScanRows / 8; ++i) {
      const int y = y0 + warp + 8 * i;
#pragma unroll
      for (int k = 0; k < NW; ++k) vals[i][k] = 0;
      if (y < H) {
        if (VEC == 16) {
          const uint4 q = *reinterpret_cast<const uint4*>(L + (size_t)y * W + x0);
          vals[i][0] = q.x; vals[i][NW > 1 ? 1 : 0] = q.y; vals[i][NW > 2 ? 2 : 0] = q.z; vals[i][NW > 3 ? 3 : 0] = q.w;
        } else if (VEC == 8) {
          const uint2 q = *reinterpret_cast<const uint2*>(L + (size_t)y * W + x0);
          vals[i][0] = q.x; vals[i][NW > 1 ? 1 : 0] = q.y;
        } else if (VEC == 4) {
          vals[i][0] = *reinterpret_cast<const uint32_t*>(L + (size_t)y * W + x0);
        } else {
          vals[i][0] = L[(size_t)y * W + x0];
        }
      }
    }
    // the scan is ALU-bound if every byte is examined (5 instructions per byte); almost all words are zero
    // (bone covers a few per cent of the pixels), so a row is examined only if it holds a set byte at all
#pragma unroll
    for (int i = 0; i < kScanRows / 8; ++i) {
      uint32_t any = 0;
#pragma unroll
      for (int k = 0; k < NW; ++k) any |= vals[i][k];
      if (any == 0) continue;
      const int y = y0 + warp + 8 * i;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        if ((vals[i][v / 4] >> (8 * (v & 3))) & 0xffu) {
          first[v] = min(first[v], y);
          last[v] = max(last[v], y);
        }
      }
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    s_first[warp][lane * PITCH + v] = first[v];
    s_last[warp][lane * PITCH + v] = last[v];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * VEC; c += 256) {
    int f = kNone, l = -1;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      f = min(f, s_first[w][c + (PITCH - VEC) * (c / VEC)]);
      l = max(l, s_last[w][c + (PITCH - VEC) * (c / VEC)]);
    }
    const int x = blockIdx.x * 32 * VEC + c;
    if (x < W) {   // one partial per (image, row tile, column): no memset, no atomics
      USAUG_DEV_ASSERT(c + (PITCH - VEC) * (c / VEC) < 32 * PITCH);
      int* dst = surf + (((size_t)b * gridDim.y + blockIdx.y) * W + x) * 2;
      dst[0] = f;
      dst[1] = (l >= 0) ? l : -1;
    }
  }
}

// first / last bone row of column x from the per-tile partials
__device__ __forceinline__ void column_extent(const int* __restrict__ surf, int b, int ntile, int W, int x,
                                              int& first, int& last) {
  int f = kNone, l = -1;
  for (int t = 0; t < ntile; ++t) {
    const int* p = surf + (((size_t)b * ntile + t) * W + x) * 2;
    f = min(f, p[0]);
    l = max(l, p[1]);
  }
  first = f; last = l;
}

int scan_tiles(int H) { return (H + kScanRows - 1) / kScanRows; }
size_t surface_scan_bytes(int B, int H, int W) { return align_up((size_t)B * scan_tiles(H) * W * 2 * sizeof(int)); }

int launch_surface_scan(const uint8_t* lab, int B, int H, int W, int* surf, cudaStream_t stream) {
  const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(lab) & 3u) == 0);
  const bool vec16 = (W % 16 == 0) && ((reinterpret_cast<uintptr_t>(lab) & 15u) == 0) && W >= 512;
  const bool vec8 = (W % 8 == 0) && ((reinterpret_cast<uintptr_t>(lab) & 7u) == 0) && W >= 256;
  if (vec16) {        // 128-bit loads: a warp row covers 512 columns
    dim3 grid((W + 511) / 512, (H + kScanRows - 1) / kScanRows, B);
    surface_scan_kernel<16><<<grid, 256, 0, stream>>>(lab, H, W, surf);
  } else if (vec8) {  // 64-bit loads: a warp row covers 256 columns
    dim3 grid((W + 255) / 256, (H + kScanRows - 1) / kScanRows, B);
    surface_scan_kernel<8><<<grid, 256, 0, stream>>>(lab, H, W, surf);
  } else if (vec) {
    dim3 grid((W + 127) / 128, (H + kScanRows - 1) / kScanRows, B);
    surface_scan_kernel<4><<<grid, 256, 0, stream>>>(lab, H, W, surf);
  } else {
    dim3 grid((W + 31) / 32, (H + kScanRows - 1) / kScanRows, B);
    surface_scan_kernel<1><<<grid, 256, 0, stream>>>(lab, H, W, surf);
  }
  USAUG_CHECK_LAUNCH("surface_scan_kernel");
  return USAUG_OK;
}

// ---------------------------------------------------------------------------------------------
// Profile kernel: one block per image.  Hole fill (nearest bone column, ties -> lower x),
// order-pinned smoothing, displacement clamp.  info[b] = {d_eff, y_top(as int bits)}.
// ---------------------------------------------------------------------------------------------
struct WarpInfo {
  float d_eff;
  int y_top;
};

__global__ void __launch_bounds__(256) profile_kernel(const int* __restrict__ surf, int ntile, int H, int W,
                                                      const float* __restrict__ d,
                                                      const float* __restrict__ taps, int R,
                                                      float* __restrict__ s_hat,
                                                      WarpInfo* __restrict__ info) {
  extern __shared__ int sm[];
  int* s_first = sm;                                  // [W] first bone row or -1
  float* s_fill = reinterpret_cast<float*>(sm + W);   // [W] hole-filled profile as fp32
  __shared__ int s_red_i[8];
  __shared__ float s_red_f[8];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_wait();                                          // surf comes from the scan kernel before us

  int ytop = kNone;
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    int f, l;
    column_extent(surf, b, ntile, W, x, f, l);
    s_first[x] = (f == kNone) ? -1 : f;
    ytop = min(ytop, f);
  }
  ytop = warp_min_i(ytop);
  if (lane == 0) s_red_i[warp] = ytop;
  __syncthreads();
  ytop = kNone;
#pragma unroll
  for (int w = 0; w < 8; ++w) ytop = min(ytop, s_red_i[w]);
  const bool has_bone = (ytop != kNone);

  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    int s = H - 1;
    if (has_bone) {
      s = s_first[x];
      for (int j = 1; s < 0; ++j) {            // outward search; left first => ties to lower x
        if (x - j >= 0 && s_first[x - j] >= 0) s = s_first[x - j];
        else if (x + j < W && s_first[x + j] >= 0) s = s_first[x + j];
      }
    }
    s_fill[x] = (float)s;
  }
  __syncthreads();

  float mn = 3.0e38f;
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    float acc = 0.0f;
    for (int k = -R; k <= R; ++k) {
      const int xi = min(max(x + k, 0), W - 1);
      USAUG_DEV_ASSERT(k + R >= 0 && k + R <= 2 * R && xi >= 0 && xi < W);
      acc = __fadd_rn(acc, __fmul_rn(taps[k + R], s_fill[xi]));
    }
    s_hat[(size_t)b * W + x] = acc;
    mn = fminf(mn, acc);
  }
  mn = warp_min(mn);
  if (lane == 0) s_red_f[warp] = mn;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = s_red_f[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fminf(m, s_red_f[w]);
    const float lim = __fmul_rn(0.5f, m);
    WarpInfo wi;
    wi.d_eff = fminf(fmaxf(d[b], -lim), lim);
    wi.y_top = has_bone ? ytop : -1;
    info[b] = wi;
  }
}

// ---------------------------------------------------------------------------------------------
// Fused warp: thread = one scanline (column), block = 128 scanlines x kWarpRows output rows.
// Per output pixel: 2 image taps + 1 label tap gathered along the scanline, TGC gain, clip.
// 10 algorithmic bytes per pixel (4+1 in, 4+1 out).
// ---------------------------------------------------------------------------------------------
constexpr int kWarpRows = 8;       // output rows per thread and tile
constexpr int kWarpTiles = 4;       // consecutive tiles per block

// TIn = float (image in [0,1]) or uint8_t (8-bit B-mode ingest: v/255 in correctly rounded fp32, from a
// 256-entry shared-memory table so the gather loop carries no divide).
// A block walks kWarpTiles tiles of kWarpRows rows down its 128 scanlines: the per-column quantities (s_hat, d_eff,
// 1/den) and the TGC gains are set up once per block.  With one 8-row tile per block a block lived for two dependent
// memory round trips (per-image scalars, then the gathers) and the memory pipeline idled during the first.
// Measured at B = 1024 of 512 x 512 (one box, same run): 1 tile / 48 registers 719 us, 4 tiles / 48 registers 704 us,
// 4 tiles / 64 registers (8 blocks per SM) 656 us; prefetching tile i+1 before tile i is stored
// was slower (801 us: the loads of both tiles share scoreboards, so waiting for the older tile waits for the newer too).
template <typename TIn>
__global__ void __launch_bounds__(128, 8) fused_warp_kernel(
    const TIn* __restrict__ img, const uint8_t* __restrict__ lab, float* __restrict__ out_img,
    uint8_t* __restrict__ out_lab, int H, int W, const float* __restrict__ s_hat,
    const WarpInfo* __restrict__ info, const float* __restrict__ tgc, int K, float gscale) {
  __shared__ float s_gain[kWarpRows * kWarpTiles];
  __shared__ float s_lut[std::is_same<TIn, uint8_t>::value ? 256 : 1];
  if constexpr (std::is_same<TIn, uint8_t>::value) {
    s_lut[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.0f);
    s_lut[threadIdx.x + 128] = __fdiv_rn((float)(threadIdx.x + 128), 255.0f);
  }
  auto px = [&](TIn v) -> float {
    if constexpr (std::is_same<TIn, uint8_t>::value) return s_lut[v];
    else return v;
  };
  const int b = blockIdx.z;
  const int x = blockIdx.x * 128 + threadIdx.x;
  const int yb = blockIdx.y * (kWarpRows * kWarpTiles);
  if (threadIdx.x < kWarpRows * kWarpTiles) {          // TGC gain is a function of the output row only
    const float* g = tgc + (size_t)b * K;
    const float yf = (float)(yb + threadIdx.x);
    float gain;
    if (K == 1) {
      gain = g[0];
    } else {
      const float u = __fmul_rn(yf, gscale);
      const int k0 = min((int)floorf(u), K - 2);
      const float fg = u - (float)k0;
      const float g0 = g[k0], g1 = g[k0 + 1];
      gain = g0 + fg * (g1 - g0);
    }
    s_gain[threadIdx.x] = gain;
  }
  const size_t base = (size_t)b * H * W;
  const int xs = x < W ? x : W - 1;                      // (threads beyond the last column compute, but never store)
  const TIn* __restrict__ I = img + base + xs;           // column pointers; in-image offsets fit 32 bits
  const uint8_t* __restrict__ L = lab + base + xs;
  float* __restrict__ OI = out_img + base + xs;
  uint8_t* __restrict__ OL = out_lab + base + xs;
  const float de = info[b].d_eff;
  const float den = __fsub_rn(s_hat[(size_t)b * W + xs], de);
  const unsigned uH = (unsigned)H;
  // t = y / den must be the correctly rounded quotient (label picks are bit-exact).  For 1 <= den < 2^16 and
  // integer y < min(den, 8192) it is computed as  q = RN(y r), e = RN(y - q den), t = RN(q + e r)  with
  // r = RN(1 / den) once per column -- proven equal to the IEEE division for EVERY such pair by the exhaustive run
  // of tools/div_check.cu (3.09e11 quotients, 0 mismatches; Markstein's theorem).  No slow-path call, no branch.
  const bool fastdiv = (den >= 1.0f) && (den < 65536.0f) && (H <= 8192);
  const float rden = fastdiv ? __frcp_rn(den) : 0.0f;
  __syncthreads();                                       // gains (and the 8-bit table) are in shared memory

  struct Tile {
    int i0[kWarpRows], yn[kWarpRows];
    float fr[kWarpRows];
    TIn v0[kWarpRows], v1[kWarpRows];
    unsigned lv[kWarpRows];                  // (a byte array here would be placed in local memory)
    bool interior;
  };
  // source coordinates of the tile's rows, then all its gathers back-to-back
  auto fetch = [&](Tile& t, int ya) {
#pragma unroll
    for (int r = 0; r < kWarpRows; ++r) {
      const float yf = (float)(ya + r);
      float tq;
      if (fastdiv) {
        const float q = __fmul_rn(yf, rden);
        const float e = __fmaf_rn(-q, den, yf);
        tq = (yf < den) ? __fmaf_rn(e, rden, q) : 1.0f;     // rows at / below the bone surface: t = 1
      } else {
        tq = 1.0f;
        if (yf < den) tq = __fdiv_rn(yf, den);
      }
      const float ys = __fadd_rn(yf, __fmul_rn(de, tq));
      t.i0[r] = __float2int_rd(ys);
      t.fr[r] = ys - (float)t.i0[r];
      t.yn[r] = __float2int_rd(__fadd_rn(ys, 0.5f));
    }
    // Block-uniform fast path: 0 <= y_src <= y + |d|, so a tile whose last row + |d| + 2 stays inside
    // the image needs no bounds checks at all (every tile except the bottom few).
    t.interior = (float)(ya + kWarpRows + 1) + fabsf(de) < (float)(H - 1);
    if (t.interior) {
#pragma unroll
      for (int r = 0; r < kWarpRows; ++r) {
        USAUG_DEV_ASSERT(t.i0[r] >= 0 && t.i0[r] + 1 < H && t.yn[r] >= 0 && t.yn[r] < H);   // "interior": no bounds checks
        const TIn* p0 = I + t.i0[r] * W;
        t.v0[r] = p0[0];
        t.v1[r] = p0[W];
        t.lv[r] = L[t.yn[r] * W];
      }
    } else {
#pragma unroll
      for (int r = 0; r < kWarpRows; ++r) {
        const bool in = ya + r < H;
        t.v0[r] = (in && (unsigned)t.i0[r] < uH) ? I[t.i0[r] * W] : (TIn)0;
        t.v1[r] = (in && (unsigned)(t.i0[r] + 1) < uH) ? I[(t.i0[r] + 1) * W] : (TIn)0;
        t.lv[r] = (in && (unsigned)t.yn[r] < uH) ? L[t.yn[r] * W] : 0u;
      }
    }
  };
  auto finish = [&](const Tile& t, int ya, int ti) {
    if (x >= W) return;
#pragma unroll
    for (int r = 0; r < kWarpRows; ++r) {
      const int y = ya + r;
      if (t.interior || y < H) {
        USAUG_DEV_ASSERT(y >= 0 && y < H);
        const float a0 = px(t.v0[r]), a1 = px(t.v1[r]);
        const float v = fmaf(t.fr[r], a1 - a0, a0);
        OI[y * W] = fminf(fmaxf(s_gain[ti * kWarpRows + r] * v, 0.0f), 1.0f);
        OL[y * W] = (uint8_t)t.lv[r];
      }
    }
  };
#pragma unroll 1
  for (int ti = 0; ti < kWarpTiles; ++ti) {
    const int ya = yb + ti * kWarpRows;
    if (ya >= H) break;
    Tile t;
    fetch(t, ya);
    finish(t, ya, ti);
  }
}

// ---------------------------------------------------------------------------------------------
// Bone box: reduces the surface scan to {y_top, y_bottom, x_left, x_right} per image.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bone_box_kernel(const int* __restrict__ surf, int ntile, int H, int W,
                                                       int* __restrict__ box) {
  __shared__ int s_red[4][8];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_wait();                                          // surf comes from the scan kernel before us
  int yt = kNone, yb = -1, xl = kNone, xr = -1;
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    int f, l;
    column_extent(surf, b, ntile, W, x, f, l);
    if (f != kNone) {
      yt = min(yt, f); yb = max(yb, l); xl = min(xl, x); xr = max(xr, x);
    }
  }
  yt = warp_min_i(yt); yb = warp_max_i(yb); xl = warp_min_i(xl); xr = warp_max_i(xr);
  if (lane == 0) { s_red[0][warp] = yt; s_red[1][warp] = yb; s_red[2][warp] = xl; s_red[3][warp] = xr; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      yt = min(yt, s_red[0][w]); yb = max(yb, s_red[1][w]);
      xl = min(xl, s_red[2][w]); xr = max(xr, s_red[3][w]);
    }
    int* o = box + (size_t)b * 4;
    if (yt == kNone) { o[0] = -1; o[1] = -1; o[2] = -1; o[3] = -1; }
    else { o[0] = yt; o[1] = yb; o[2] = xl; o[3] = xr; }
  }
}

}  // namespace usaug

using namespace usaug;

extern "C" size_t usaug_warp_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return surface_scan_bytes(B, H, W) + align_up((size_t)B * W * sizeof(float)) +
         align_up((size_t)B * sizeof(WarpInfo)) + 3 * kRedZone;
}

template <typename TIn>
static int fused_tgc_deform_impl(const TIn* img, const uint8_t* lab, float* out_img,
                                 uint8_t* out_lab, int B, int H, int W, const float* d,
                                 const float* tgc, int K, const float* taps, int R, void* ws,
                                 size_t ws_bytes, void* stream_) {
  if (!img || !lab || !out_img || !out_lab || !d || !tgc || !taps || !ws) {
    set_error("usaug_fused_tgc_deform: null pointer"); return USAUG_E_INVALID_ARG;
  }
  if (B <= 0 || H <= 0 || W <= 0 || W > 4096 || K < 1 || R < 0) {
    set_error("usaug_fused_tgc_deform: bad shape B=%d H=%d W=%d K=%d R=%d (need 1<=W<=4096)", B, H, W, K, R);
    return USAUG_E_INVALID_ARG;
  }
  if ((const void*)img == (const void*)out_img || lab == out_lab) {
    set_error("usaug_fused_tgc_deform: outputs must not alias inputs"); return USAUG_E_INVALID_ARG;
  }
  if (ws_bytes < usaug_warp_workspace_bytes(B, H, W)) {
    set_error("usaug_fused_tgc_deform: workspace %zu < %zu", ws_bytes, usaug_warp_workspace_bytes(B, H, W));
    return USAUG_E_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Carver cv(ws);
  int* surf = cv.take<int>((size_t)B * scan_tiles(H) * W * 2);
  float* s_hat = cv.take<float>((size_t)B * W);
  WarpInfo* info = cv.take<WarpInfo>(B);
  int rc = cv.arm(stream);
  if (rc) return rc;
  rc = launch_surface_scan(lab, B, H, W, surf, stream);
  if (rc) return rc;
  USAUG_CUDA(launch_pdl(profile_kernel, dim3(B), dim3(256), (size_t)W * 8, stream, surf, scan_tiles(H), H, W, d, taps, R,
                        s_hat, info));
  const float gscale = (H > 1 && K > 1) ? (float)((double)(K - 1) / (double)(H - 1)) : 0.0f;
  dim3 grid((W + 127) / 128, (H + kWarpRows * kWarpTiles - 1) / (kWarpRows * kWarpTiles), B);
  fused_warp_kernel<TIn><<<grid, 128, 0, stream>>>(img, lab, out_img, out_lab, H, W, s_hat, info, tgc,
                                                   (H > 1) ? K : 1, gscale);     // last kernel of the call: normal launch
  USAUG_CHECK_LAUNCH("fused_warp_kernel");
  return cv.verify(stream, "usaug_fused_tgc_deform");
}

extern "C" int usaug_fused_tgc_deform(const float* img, const uint8_t* lab, float* out_img,
                                      uint8_t* out_lab, int B, int H, int W, const float* d,
                                      const float* tgc, int K, const float* taps, int R, void* ws,
                                      size_t ws_bytes, void* stream) {
  return fused_tgc_deform_impl<float>(img, lab, out_img, out_lab, B, H, W, d, tgc, K, taps, R, ws, ws_bytes, stream);
}

extern "C" int usaug_fused_tgc_deform_u8(const uint8_t* img, const uint8_t* lab, float* out_img,
                                         uint8_t* out_lab, int B, int H, int W, const float* d,
                                         const float* tgc, int K, const float* taps, int R, void* ws,
                                         size_t ws_bytes, void* stream) {
  return fused_tgc_deform_impl<uint8_t>(img, lab, out_img, out_lab, B, H, W, d, tgc, K, taps, R, ws, ws_bytes, stream);
}

extern "C" size_t usaug_bone_box_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return surface_scan_bytes(B, H, W) + kRedZone;
}

namespace usaug {
// scan + box reduction; `followed_by_kernel`: the caller enqueues another (normally launched) kernel right after,
// so the box kernel may use a programmatic dependent launch (see common.cuh for the rule)
int launch_bone_box(const uint8_t* lab, int B, int H, int W, int* box, int* surf, cudaStream_t stream, bool followed_by_kernel) {
  int rc = launch_surface_scan(lab, B, H, W, surf, stream);
  if (rc) return rc;
  if (followed_by_kernel) {
    USAUG_CUDA(launch_pdl(bone_box_kernel, dim3(B), dim3(256), 0, stream, surf, scan_tiles(H), H, W, box));
  } else {
    bone_box_kernel<<<B, 256, 0, stream>>>(surf, scan_tiles(H), H, W, box);
    USAUG_CHECK_LAUNCH("bone_box_kernel");
  }
  return USAUG_OK;
}
}  // namespace usaug

extern "C" int usaug_bone_box(const uint8_t* lab, int B, int H, int W, int* box, void* ws,
                              size_t ws_bytes, void* stream_) {
  if (!lab || !box || !ws) { set_error("usaug_bone_box: null pointer"); return USAUG_E_INVALID_ARG; }
  if (B <= 0 || H <= 0 || W <= 0) { set_error("usaug_bone_box: bad shape"); return USAUG_E_INVALID_ARG; }
  if (ws_bytes < usaug_bone_box_workspace_bytes(B, H, W)) {
    set_error("usaug_bone_box: workspace too small"); return USAUG_E_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Carver cv(ws);
  int* surf = cv.take<int>((size_t)B * scan_tiles(H) * W * 2);
  int rc = cv.arm(stream);
  if (rc) return rc;
  rc = launch_bone_box(lab, B, H, W, box, surf, stream, false);
  if (rc) return rc;
#ifdef USAUG_BOUNDS_CHECK
  // self-test of the red zones (tests/test_bounds_checked_build.py): one stray int right behind the scan partials
  if (getenv("USAUG_DBG_FAULT")) USAUG_CUDA(cudaMemsetAsync(surf + (size_t)B * scan_tiles(H) * W * 2, 0, sizeof(int), stream));
#endif
  return cv.verify(stream, "usaug_bone_box");
}
