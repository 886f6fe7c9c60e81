// K3: confidence-map SNR change fused with the reverberation artifact.
//   I1   = clip(I + (a-1) * C * I, 0, 1)                                   (skipped when C == NULL)
//   out  = clip(I1 + sum_{k=1..n} rho^k * Mf[y-kD, x] * I1[y-kD, x], 0, 1)   D = first bone row
// Mf = separable Gaussian feather of the bone mask.  The ghost term only exists inside the bone
// bounding box (+R) shifted by k*D, so the 2-D tap loop runs for a small fraction of pixels.
#include "kernels.cuh"

namespace usaug {

constexpr int kSrRows = 16;
constexpr int kMaxTaps = 65;   // R <= 32

__device__ __forceinline__ float snr_px(float I, float c, float k) {
  // (k*c)*I then + I, as the oracle (oracle/snr_reverb.py: snr_apply)
  return fminf(fmaxf(I + (k * c) * I, 0.0f), 1.0f);
}

__global__ void __launch_bounds__(128) snr_reverb_kernel(
    const float* __restrict__ img, const float* __restrict__ cm, const uint8_t* __restrict__ lab,
    float* __restrict__ out, int H, int W, const float* __restrict__ a_snr,
    const float* __restrict__ rho, const int* __restrict__ nghost, const int* __restrict__ box,
    const float* __restrict__ taps_g, int R) {
  __shared__ float taps[kMaxTaps];
  for (int i = threadIdx.x; i < 2 * R + 1; i += blockDim.x) taps[i] = taps_g[i];
  __syncthreads();

  const int b = blockIdx.z;
  const int x = blockIdx.x * 128 + threadIdx.x;
  if (x >= W) return;
  const size_t base = (size_t)b * H * W;
  const float* I = img + base;
  const float* C = cm ? cm + base : nullptr;
  const uint8_t* L = lab + base;
  const float ks = C ? (a_snr[b] - 1.0f) : 0.0f;
  const int ytop = box[b * 4 + 0], ybot = box[b * 4 + 1], xl = box[b * 4 + 2], xr = box[b * 4 + 3];
  const int n = nghost[b];
  const float rh = rho[b];
  const bool ghosts = (ytop >= 1) && (n >= 1);
  const bool x_in = ghosts && (x >= xl - R) && (x <= xr + R);
  const int ya = blockIdx.y * kSrRows;

#pragma unroll 4
  for (int r = 0; r < kSrRows; ++r) {
    const int y = ya + r;
    if (y >= H) break;
    const size_t o = (size_t)y * W + x;
    float v = I[o];
    if (C) v = snr_px(v, C[o], ks);
    if (ghosts) {
      float acc = v;
      if (x_in) {
        float coef = 1.0f;
        for (int k = 1; k <= n; ++k) {
          coef *= rh;
          const int ys = y - k * ytop;
          if (ys < 0) break;
          if (ys < ytop - R || ys > ybot + R) continue;
          // feathered mask at (ys, x): rows then columns, increasing tap index
          float m = 0.0f;
          for (int j = -R; j <= R; ++j) {
            const int yy = ys + j;
            if (yy < 0 || yy >= H) continue;
            const uint8_t* Lr = L + (size_t)yy * W;
            float row = 0.0f;
            for (int i = -R; i <= R; ++i) {
              const int xx = x + i;
              if (xx >= 0 && xx < W && Lr[xx]) row += taps[i + R];
            }
            m += taps[j + R] * row;
          }
          if (m > 0.0f) {
            const size_t so = (size_t)ys * W + x;
            float sv = I[so];
            if (C) sv = snr_px(sv, C[so], ks);
            acc += coef * (m * sv);
          }
        }
      }
      v = fminf(fmaxf(acc, 0.0f), 1.0f);
    }
    out[base + o] = v;
  }
}

}  // namespace usaug

using namespace usaug;

extern "C" size_t usaug_snr_reverb_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return surface_scan_bytes(B, W) + align_up((size_t)B * 4 * sizeof(int));
}

extern "C" int usaug_snr_reverb(const float* img, const float* cm, const uint8_t* lab, float* out,
                                int B, int H, int W, const float* a_snr, const float* rho,
                                const int* nghost, const float* feather_taps, int R, void* ws,
                                size_t ws_bytes, void* stream_) {
  if (!img || !lab || !out || !a_snr || !rho || !nghost || !feather_taps || !ws) {
    set_error("usaug_snr_reverb: null pointer"); return USAUG_E_INVALID_ARG;
  }
  if (B <= 0 || H <= 0 || W <= 0 || R < 0 || 2 * R + 1 > kMaxTaps) {
    set_error("usaug_snr_reverb: bad shape B=%d H=%d W=%d R=%d (need R<=32)", B, H, W, R);
    return USAUG_E_INVALID_ARG;
  }
  if (img == out) { set_error("usaug_snr_reverb: out must not alias img"); return USAUG_E_INVALID_ARG; }
  if (ws_bytes < usaug_snr_reverb_workspace_bytes(B, H, W)) {
    set_error("usaug_snr_reverb: workspace too small"); return USAUG_E_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Carver cv(ws);
  int* surf = cv.take<int>((size_t)B * W * 2);
  int* box = cv.take<int>((size_t)B * 4);
  int rc = usaug_bone_box(lab, B, H, W, box, surf, surface_scan_bytes(B, W), stream_);
  if (rc) return rc;
  dim3 grid((W + 127) / 128, (H + kSrRows - 1) / kSrRows, B);
  snr_reverb_kernel<<<grid, 128, 0, stream>>>(img, cm, lab, out, H, W, a_snr, rho, nghost, box,
                                              feather_taps, R);
  USAUG_CHECK_LAUNCH("snr_reverb_kernel");
  return USAUG_OK;
}
