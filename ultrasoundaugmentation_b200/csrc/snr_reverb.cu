// K3: confidence-map SNR change fused with the reverberation artifact.
//   I1   = clip(I + (a-1) * C * S, 0, 1)                                   (skipped when C == NULL)
//          S = I, or S = sig * sig_scale[b] when a signal plane is given (monogenic local energy, 8(f) row 1)
//   out  = clip(I1 + sum_{k=1..n} rho^k * Mf[y-kD, x] * I1[y-kD, x], 0, 1)   D = first bone row
// Mf = separable Gaussian feather of the bone mask.  The ghost term only exists inside the bone
// bounding box (+R) shifted by k*D, so the 2-D tap loop runs for a small fraction of pixels.
#include <type_traits>

#include "kernels.cuh"

namespace usaug {

constexpr int kSrRows = 16;
constexpr int kSrCols = 128;
constexpr int kMaxTaps = 65;   // R <= 32

__device__ __forceinline__ float snr_px(float I, float c, float k, float S) {
  // (k*c)*S then + I, as the oracle (oracle/snr_reverb.py: snr_apply, oracle/local_energy.py: snr_apply_signal)
  return fminf(fmaxf(I + (k * c) * S, 0.0f), 1.0f);
}

// Block = 128 scanlines x 16 output rows, thread = scanline.  The ghost term only exists where the
// shifted bone box (+R) meets the tile; there the feathered mask of the 16 x 128 source tile is built
// SEPARABLY in shared memory (row pass over 16+2R label rows, then column pass: 2*(2R+1) MACs per
// pixel instead of (2R+1)^2), in the oracle's order (rows first, increasing tap index).
// TOut = float, or uint8_t: 8-bit egress q = floor(255 * v + 0.5) (two individually rounded fp32 operations).
template <typename TOut>
__global__ void __launch_bounds__(kSrCols) snr_reverb_kernel(
    const float* __restrict__ img, const float* __restrict__ cm, const uint8_t* __restrict__ lab,
    const float* __restrict__ sig, const float* __restrict__ sig_scale,
    TOut* __restrict__ out, int H, int W, const float* __restrict__ a_snr,
    const float* __restrict__ rho, const int* __restrict__ nghost, const int* __restrict__ box,
    const float* __restrict__ taps_g, int R, int vec_ok) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  __shared__ float taps[kMaxTaps];
  const int TR = kSrRows + 2 * R;              // label / row-pass rows
  const int TC = kSrCols + 2 * R;              // label columns
  float* rowp = reinterpret_cast<float*>(sm_raw);                               // [TR][kSrCols]
  uint8_t* labs = sm_raw + (size_t)TR * kSrCols * sizeof(float);               // [TR][TC]
  for (int i = threadIdx.x; i < 2 * R + 1; i += blockDim.x) taps[i] = taps_g[i];

  const int b = blockIdx.z;
  const int x0 = blockIdx.x * kSrCols, x = x0 + threadIdx.x;
  const bool xin = x < W;
  const size_t base = (size_t)b * H * W;
  const float* I = img + base;
  const float* C = cm ? cm + base : nullptr;
  const float* S = (cm && sig) ? sig + base : nullptr;
  const uint8_t* L = lab + base;
  const int ya = blockIdx.y * kSrRows;

  // Fast-path data first: the tile's pixels (thread = 4 columns x 4 rows, 128-bit accesses) are requested BEFORE the
  // per-image scalars and the bone box are read -- a block lives for two memory round trips, and with the scalars
  // first (box -> cull decision -> pixels) the memory pipeline sat idle for the first of them (0.66 of the roofline).
  const int xc = x0 + 4 * (threadIdx.x & 31), yq = ya + 4 * (threadIdx.x >> 5);
  float4 pt[4], pc[4], ps[4];
  if (vec_ok) {                                // vec_ok: W % 4 == 0 and every plane 16-byte aligned (host check)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const bool in = xc < W && yq + r < H;
      const size_t o = (size_t)(in ? yq + r : 0) * W + (in ? xc : 0);
      pt[r] = *reinterpret_cast<const float4*>(I + o);
      if (C) pc[r] = *reinterpret_cast<const float4*>(C + o);
      if (S) ps[r] = *reinterpret_cast<const float4*>(S + o);
    }
  }
  const float ss = S ? sig_scale[b] : 0.0f;
  const float ks = C ? (a_snr[b] - 1.0f) : 0.0f;
  const int4 bx = *reinterpret_cast<const int4*>(box + b * 4);
  const int ytop = bx.x, ybot = bx.y, xl = bx.z, xr = bx.w;
  const int n = nghost[b];
  const float rh = rho[b];
  const bool ghosts = (ytop >= 1) && (n >= 1);

  // Block-uniform fast path (most tiles): no ghost source meets this tile, so the result is the SNR term alone.
  bool tile_has_ghost = false;
  if (ghosts) {
    for (int k = 1; k <= n; ++k) {
      const int ys0 = ya - k * ytop;
      if (ys0 + kSrRows - 1 < 0) break;
      if (!(ys0 > ybot + R || ys0 + kSrRows - 1 < ytop - R || x0 > xr + R || x0 + kSrCols - 1 < xl - R)) tile_has_ghost = true;
    }
  }
  if (!tile_has_ghost && vec_ok) {
    if (xc >= W) return;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int y = yq + r;
      if (y >= H) break;
      const size_t o = (size_t)y * W + xc;
      float4 t = pt[r];
      if (C) {
        const float4 c = pc[r];
        float4 sg = t;
        if (S) {
          sg = ps[r];
          sg.x *= ss; sg.y *= ss; sg.z *= ss; sg.w *= ss;
        }
        t.x = snr_px(t.x, c.x, ks, sg.x); t.y = snr_px(t.y, c.y, ks, sg.y);
        t.z = snr_px(t.z, c.z, ks, sg.z); t.w = snr_px(t.w, c.w, ks, sg.w);
      }
      if (ghosts) { t.x = fminf(fmaxf(t.x, 0.f), 1.f); t.y = fminf(fmaxf(t.y, 0.f), 1.f);
                    t.z = fminf(fmaxf(t.z, 0.f), 1.f); t.w = fminf(fmaxf(t.w, 0.f), 1.f); }
      if constexpr (std::is_same<TOut, uint8_t>::value) {
        auto q8 = [](float f) { return (unsigned)__float2int_rd(__fadd_rn(__fmul_rn(f, 255.0f), 0.5f)) & 0xffu; };
        *reinterpret_cast<unsigned*>(out + base + o) = q8(t.x) | (q8(t.y) << 8) | (q8(t.z) << 16) | (q8(t.w) << 24);
      } else {
        *reinterpret_cast<float4*>(out + base + o) = t;
      }
    }
    return;
  }

  float v[kSrRows], acc[kSrRows];
#pragma unroll
  for (int r = 0; r < kSrRows; ++r) {
    const int y = ya + r;
    float t = 0.f;
    if (xin && y < H) {
      t = I[(size_t)y * W + x];
      if (C) t = snr_px(t, C[(size_t)y * W + x], ks, S ? S[(size_t)y * W + x] * ss : t);
    }
    v[r] = t; acc[r] = t;
  }
  __syncthreads();   // taps visible

  if (ghosts) {
    float coef = 1.0f;
    for (int k = 1; k <= n; ++k) {
      coef *= rh;
      const int ys0 = ya - k * ytop;                      // source row of output row ya
      if (ys0 + kSrRows - 1 < 0) break;                   // whole tile above the first ghost row
      // block-uniform cull: source tile (+R) against the bone box
      if (ys0 > ybot + R || ys0 + kSrRows - 1 < ytop - R || x0 > xr + R || x0 + kSrCols - 1 < xl - R) continue;
      __syncthreads();                                    // previous ghost done with the tiles
      for (int i = threadIdx.x; i < TR * TC; i += blockDim.x) {
        const int rr = i / TC, cc = i - rr * TC;
        const int yy = ys0 - R + rr, xx = x0 - R + cc;
        labs[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (uint8_t)(L[(size_t)yy * W + xx] != 0) : (uint8_t)0;
      }
      __syncthreads();
      for (int rr = 0; rr < TR; ++rr) {                   // row pass: thread = column
        USAUG_DEV_ASSERT(rr * TC + (int)threadIdx.x + 2 * R < TR * TC);
        const uint8_t* lr = labs + rr * TC + threadIdx.x;
        float s = 0.0f;
        for (int i = 0; i <= 2 * R; ++i) s += lr[i] ? taps[i] : 0.0f;
        rowp[rr * kSrCols + threadIdx.x] = s;
      }
      // column pass needs only this thread's own column of rowp: no barrier required
#pragma unroll
      for (int r = 0; r < kSrRows; ++r) {
        const int ys = ys0 + r;
        if (ys < 0 || ya + r >= H || !xin) continue;
        float m = 0.0f;
        USAUG_DEV_ASSERT(r + 2 * R < TR && ys < H);
        for (int j = 0; j <= 2 * R; ++j) m += taps[j] * rowp[(r + j) * kSrCols + threadIdx.x];
        if (m > 0.0f) {
          const size_t so = (size_t)ys * W + x;
          float sv = I[so];
          if (C) sv = snr_px(sv, C[so], ks, S ? S[so] * ss : sv);
          acc[r] += coef * (m * sv);
        }
      }
    }
  }
  if (xin) {
#pragma unroll
    for (int r = 0; r < kSrRows; ++r) {
      const int y = ya + r;
      if (y < H) {
        const float o = ghosts ? fminf(fmaxf(acc[r], 0.0f), 1.0f) : v[r];
        if constexpr (std::is_same<TOut, uint8_t>::value)
          out[base + (size_t)y * W + x] = (uint8_t)__float2int_rd(__fadd_rn(__fmul_rn(o, 255.0f), 0.5f));
        else
          out[base + (size_t)y * W + x] = o;
      }
    }
  }
}

}  // namespace usaug

using namespace usaug;

extern "C" size_t usaug_snr_reverb_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return surface_scan_bytes(B, H, W) + align_up((size_t)B * 4 * sizeof(int)) + 2 * kRedZone;
}

template <typename TOut>
static int snr_reverb_impl(const float* img, const float* cm, const float* sig, const float* sig_scale,
                           const uint8_t* lab, TOut* out,
                           int B, int H, int W, const float* a_snr, const float* rho,
                           const int* nghost, const float* feather_taps, int R, void* ws,
                           size_t ws_bytes, void* stream_) {
  if (!img || !lab || !out || !a_snr || !rho || !nghost || !feather_taps || !ws) {
    set_error("usaug_snr_reverb: null pointer"); return USAUG_E_INVALID_ARG;
  }
  if ((sig == nullptr) != (sig_scale == nullptr)) {
    set_error("usaug_snr_reverb: sig and sig_scale must be given together"); return USAUG_E_INVALID_ARG;
  }
  if (B <= 0 || H <= 0 || W <= 0 || R < 0 || 2 * R + 1 > kMaxTaps) {
    set_error("usaug_snr_reverb: bad shape B=%d H=%d W=%d R=%d (need R<=32)", B, H, W, R);
    return USAUG_E_INVALID_ARG;
  }
  if ((const void*)img == (const void*)out) { set_error("usaug_snr_reverb: out must not alias img"); return USAUG_E_INVALID_ARG; }
  if (ws_bytes < usaug_snr_reverb_workspace_bytes(B, H, W)) {
    set_error("usaug_snr_reverb: workspace too small"); return USAUG_E_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Carver cv(ws);
  int* surf = cv.take<int>((size_t)B * scan_tiles(H) * W * 2);
  int* box = cv.take<int>((size_t)B * 4);
  int rc = cv.arm(stream);
  if (rc) return rc;
  rc = launch_bone_box(lab, B, H, W, box, surf, stream, true);
  if (rc) return rc;
  dim3 grid((W + kSrCols - 1) / kSrCols, (H + kSrRows - 1) / kSrRows, B);
  const size_t smem = (size_t)(kSrRows + 2 * R) * kSrCols * sizeof(float) + (size_t)(kSrRows + 2 * R) * (kSrCols + 2 * R);
  if (smem > 48 * 1024)
    USAUG_CUDA(cudaFuncSetAttribute(snr_reverb_kernel<TOut>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // the 128-bit fast path needs aligned planes: a raw C-ABI caller (or a tensor view with an odd storage offset) may
  // hand in anything, and a misaligned vector access would poison the CUDA context
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const int vec_ok = (W % 4 == 0) && al16(img) && al16(cm) && al16(sig) && al16(out) ? 1 : 0;
  snr_reverb_kernel<TOut><<<grid, kSrCols, smem, stream>>>(img, cm, lab, sig, sig_scale, out, H, W, a_snr, rho, nghost,
                                                           box, feather_taps, R, vec_ok);
  USAUG_CHECK_LAUNCH("snr_reverb_kernel");
  return cv.verify(stream, "usaug_snr_reverb");
}

extern "C" int usaug_snr_reverb(const float* img, const float* cm, const uint8_t* lab, float* out,
                                int B, int H, int W, const float* a_snr, const float* rho,
                                const int* nghost, const float* feather_taps, int R, void* ws,
                                size_t ws_bytes, void* stream) {
  return snr_reverb_impl<float>(img, cm, nullptr, nullptr, lab, out, B, H, W, a_snr, rho, nghost, feather_taps, R, ws,
                                ws_bytes, stream);
}

extern "C" int usaug_snr_reverb_u8(const float* img, const float* cm, const uint8_t* lab, uint8_t* out,
                                   int B, int H, int W, const float* a_snr, const float* rho,
                                   const int* nghost, const float* feather_taps, int R, void* ws,
                                   size_t ws_bytes, void* stream) {
  return snr_reverb_impl<uint8_t>(img, cm, nullptr, nullptr, lab, out, B, H, W, a_snr, rho, nghost, feather_taps, R, ws,
                                  ws_bytes, stream);
}

extern "C" int usaug_snr_reverb_signal(const float* img, const float* cm, const float* sig, const float* sig_scale,
                                       const uint8_t* lab, void* out, int out_is_u8, int B, int H, int W,
                                       const float* a_snr, const float* rho, const int* nghost,
                                       const float* feather_taps, int R, void* ws, size_t ws_bytes, void* stream) {
  if (out_is_u8)
    return snr_reverb_impl<uint8_t>(img, cm, sig, sig_scale, lab, static_cast<uint8_t*>(out), B, H, W, a_snr, rho, nghost,
                                    feather_taps, R, ws, ws_bytes, stream);
  return snr_reverb_impl<float>(img, cm, sig, sig_scale, lab, static_cast<float*>(out), B, H, W, a_snr, rho, nghost,
                                feather_taps, R, ws, ws_bytes, stream);
}
