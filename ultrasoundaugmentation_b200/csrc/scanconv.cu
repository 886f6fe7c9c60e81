// Curvilinear (convex-probe) geometry (oracle/scanconv.py): scan conversion between the
// Cartesian display image and the polar grid of rays x samples on which the chain runs unchanged.
//   direction 0  Cartesian [H,W] -> polar [n_samples, n_rays]   (x, y) = apex + r_i (sin th_j, cos th_j)
//   direction 1  polar -> Cartesian                             (i*, j*) = ((r - r_min) / dr, (th + th_max) / dth)
// Image: 4-tap bilinear, zeros outside; label: nearest.  HBM-bound gathers: 4 + 1 B/px read, 4 + 1 written.
// Every coordinate is a fixed sequence of individually rounded fp32 operations (no FMA contraction): the ray
// sines / cosines are host tables, the arctangent is the polynomial of Abramowitz & Stegun 4.4.49 in Horner form.
#include "kernels.cuh"

namespace usaug {

struct ScGeom {
  float apex_x, apex_y, theta_max, r_min, dr, inv_dr, inv_dth;
};

__device__ __forceinline__ float spec_atan(float z) {          // 0 <= z <= 1
  const float c[8] = {-0.3333314528f, 0.1999355085f, -0.1420889944f, 0.1065626393f,
                      -0.0752896400f, 0.0429096138f, -0.0161657367f, 0.0028662257f};
  const float z2 = __fmul_rn(z, z);
  float p = c[7];
#pragma unroll
  for (int k = 6; k >= 0; --k) p = __fadd_rn(__fmul_rn(p, z2), c[k]);
  p = __fadd_rn(__fmul_rn(p, z2), 1.0f);
  return __fmul_rn(z, p);
}

__device__ __forceinline__ float spec_atan2_pos(float dx, float dy) {   // dy > 0
  const float ax = fabsf(dx);
  const bool small = ax <= dy;
  const float num = small ? ax : dy;
  float den = small ? dy : ax;
  den = (den > 0.f) ? den : 1.0f;
  float a = spec_atan(__fdiv_rn(num, den));
  if (!small) a = __fsub_rn(1.57079637f, a);                   // fp32(pi/2)
  return (dx < 0.f) ? -a : a;
}

// gathers image (bilinear) and label (nearest) of one [H,W] plane at fp32 coordinates (ys, xs)
__device__ __forceinline__ void gather(const float* __restrict__ I, const uint8_t* __restrict__ L, int H, int W,
                                       float ys, float xs, float& v, uint8_t& l) {
  const float y0f = floorf(ys), x0f = floorf(xs);
  const float fy = __fsub_rn(ys, y0f), fx = __fsub_rn(xs, x0f);
  const int y0 = (int)y0f, x0 = (int)x0f;
  auto tap = [&](int yy, int xx) -> float {
    return ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W) ? I[(size_t)yy * W + xx] : 0.f;
  };
  // coordinates far outside (|.| > 2^30) would overflow the int cast: they are outside the image anyway
  if (!(fabsf(ys) < 1e9f) || !(fabsf(xs) < 1e9f)) { v = 0.f; l = 0; return; }
  const float t00 = tap(y0, x0), t01 = tap(y0, x0 + 1), t10 = tap(y0 + 1, x0), t11 = tap(y0 + 1, x0 + 1);
  const float top = (1.0f - fx) * t00 + fx * t01, bot = (1.0f - fx) * t10 + fx * t11;
  v = (1.0f - fy) * top + fy * bot;
  const int yn = (int)floorf(__fadd_rn(ys, 0.5f)), xn = (int)floorf(__fadd_rn(xs, 0.5f));
  l = ((unsigned)yn < (unsigned)H && (unsigned)xn < (unsigned)W) ? L[(size_t)yn * W + xn] : (uint8_t)0;
}

constexpr int kScRows = 4;

// output = polar [n_s, n_r]; thread = ray, kScRows samples
__global__ void __launch_bounds__(128) sc_cart_to_polar(const float* __restrict__ img, const uint8_t* __restrict__ lab,
                                                        float* __restrict__ out_img, uint8_t* __restrict__ out_lab,
                                                        int H, int W, int ns, int nr, ScGeom g,
                                                        const float* __restrict__ sin_t, const float* __restrict__ cos_t) {
  const int j = blockIdx.x * 128 + threadIdx.x, b = blockIdx.z;
  if (j >= nr) return;
  const float s = sin_t[j], c = cos_t[j];
  const float* I = img + (size_t)b * H * W;
  const uint8_t* L = lab + (size_t)b * H * W;
#pragma unroll
  for (int k = 0; k < kScRows; ++k) {
    const int i = blockIdx.y * kScRows + k;
    if (i >= ns) break;
    const float r = __fadd_rn(g.r_min, __fmul_rn((float)i, g.dr));
    const float x = __fadd_rn(g.apex_x, __fmul_rn(r, s)), y = __fadd_rn(g.apex_y, __fmul_rn(r, c));
    float v; uint8_t l;
    gather(I, L, H, W, y, x, v, l);
    const size_t o = ((size_t)b * ns + i) * nr + j;
    out_img[o] = v; out_lab[o] = l;
  }
}

// output = Cartesian [H, W]; thread = column, kScRows rows
__global__ void __launch_bounds__(128) sc_polar_to_cart(const float* __restrict__ img, const uint8_t* __restrict__ lab,
                                                        float* __restrict__ out_img, uint8_t* __restrict__ out_lab,
                                                        int H, int W, int ns, int nr, ScGeom g) {
  const int x = blockIdx.x * 128 + threadIdx.x, b = blockIdx.z;
  if (x >= W) return;
  const float dx = __fsub_rn((float)x, g.apex_x);
  const float dx2 = __fmul_rn(dx, dx);
  const float* I = img + (size_t)b * ns * nr;
  const uint8_t* L = lab + (size_t)b * ns * nr;
#pragma unroll
  for (int k = 0; k < kScRows; ++k) {
    const int y = blockIdx.y * kScRows + k;
    if (y >= H) break;
    const float dy = __fsub_rn((float)y, g.apex_y);
    float v = 0.f; uint8_t l = 0;
    if (dy > 0.f) {
      const float r = __fsqrt_rn(__fadd_rn(dx2, __fmul_rn(dy, dy)));
      const float th = spec_atan2_pos(dx, dy);
      const float is = __fmul_rn(__fsub_rn(r, g.r_min), g.inv_dr);
      const float js = __fmul_rn(__fadd_rn(th, g.theta_max), g.inv_dth);
      gather(I, L, ns, nr, is, js, v, l);
    }
    const size_t o = ((size_t)b * H + y) * W + x;
    out_img[o] = v; out_lab[o] = l;
  }
}

}  // namespace usaug

using namespace usaug;

extern "C" int usaug_scan_convert(const float* img, const uint8_t* lab, float* out_img, uint8_t* out_lab, int B,
                                  int H, int W, int n_samples, int n_rays, int direction, float apex_x, float apex_y,
                                  float theta_max, float r_min, float dr, float inv_dr, float inv_dtheta,
                                  const float* ray_sin, const float* ray_cos, void* stream_) {
  if (!img || !lab || !out_img || !out_lab) { set_error("usaug_scan_convert: null pointer"); return USAUG_E_INVALID_ARG; }
  if (B <= 0 || H <= 0 || W <= 0 || n_samples <= 0 || n_rays <= 0 || (direction != 0 && direction != 1)) {
    set_error("usaug_scan_convert: bad shape B=%d H=%d W=%d n_samples=%d n_rays=%d direction=%d", B, H, W, n_samples,
              n_rays, direction);
    return USAUG_E_INVALID_ARG;
  }
  if (direction == 0 && (!ray_sin || !ray_cos)) {
    set_error("usaug_scan_convert: Cartesian -> polar needs the ray sine / cosine tables"); return USAUG_E_INVALID_ARG;
  }
  if ((long long)H * W > (1ll << 30) || (long long)n_samples * n_rays > (1ll << 30) || B > 65535) {
    set_error("usaug_scan_convert: plane or batch too large"); return USAUG_E_INVALID_ARG;
  }
  if ((const void*)img == (const void*)out_img || lab == out_lab) {
    set_error("usaug_scan_convert: outputs must not alias inputs"); return USAUG_E_INVALID_ARG;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const ScGeom g{apex_x, apex_y, theta_max, r_min, dr, inv_dr, inv_dtheta};
  if (direction == 0) {
    dim3 grid((n_rays + 127) / 128, (n_samples + kScRows - 1) / kScRows, B);
    sc_cart_to_polar<<<grid, 128, 0, stream>>>(img, lab, out_img, out_lab, H, W, n_samples, n_rays, g, ray_sin, ray_cos);
    USAUG_CHECK_LAUNCH("sc_cart_to_polar");
  } else {
    dim3 grid((W + 127) / 128, (H + kScRows - 1) / kScRows, B);
    sc_polar_to_cart<<<grid, 128, 0, stream>>>(img, lab, out_img, out_lab, H, W, n_samples, n_rays, g);
    USAUG_CHECK_LAUNCH("sc_polar_to_cart");
  }
  return USAUG_OK;
}
