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

// Gather plan of one output pixel: where its four bilinear taps and its nearest label tap lie in a source plane, which
// of them exist, and the two interpolation weights.  It depends on the geometry only -- not on the image -- so a thread
// builds the plans of its pixels ONCE (square root, division, arctangent polynomial, floors, bounds tests) and then
// walks a chunk of kScImgs images with them: per image a pixel costs four loads, one label load, six multiply-adds and
// two stores.  (Round 1 rebuilt the plan for every image: 0.31 / 0.45 of the 10 B/px roofline, issue-bound; with plans
// shared by 16 images and 64 registers per thread -- 8 blocks per SM -- both directions run at 0.58 / 0.59.  Measured
// alternatives, same box: 8 images 0.46, 16 images at 126 registers 0.48, at 40 registers 0.50-0.52, image loop unrolled by 2 0.34.)
struct ScPlan {
  int o00;          // offset of tap (y0, x0) in the source plane (valid even if that tap is outside: never dereferenced then)
  int ol;           // offset of the nearest label tap, -1 = outside
  float fx, fy;
  unsigned mask;    // bit 0..3: tap (y0,x0), (y0,x0+1), (y0+1,x0), (y0+1,x0+1) is inside the plane
};

__device__ __forceinline__ ScPlan make_plan(int H, int W, float ys, float xs) {
  ScPlan p;
  p.o00 = 0; p.ol = -1; p.fx = 0.f; p.fy = 0.f; p.mask = 0u;
  // coordinates far outside (|.| > 2^30) would overflow the int cast: they are outside the image anyway
  if (!(fabsf(ys) < 1e9f) || !(fabsf(xs) < 1e9f)) return p;
  const float y0f = floorf(ys), x0f = floorf(xs);
  p.fy = __fsub_rn(ys, y0f); p.fx = __fsub_rn(xs, x0f);
  const int y0 = (int)y0f, x0 = (int)x0f;
  const bool ya = (unsigned)y0 < (unsigned)H, yb = (unsigned)(y0 + 1) < (unsigned)H;
  const bool xa = (unsigned)x0 < (unsigned)W, xb = (unsigned)(x0 + 1) < (unsigned)W;
  p.mask = (ya && xa ? 1u : 0u) | (ya && xb ? 2u : 0u) | (yb && xa ? 4u : 0u) | (yb && xb ? 8u : 0u);
  // |y0|, |x0| <= 2^30 would overflow y0 * W; taps that far out have mask 0 and are never dereferenced
  p.o00 = (p.mask != 0u) ? y0 * W + x0 : 0;
  const int yn = (int)floorf(__fadd_rn(ys, 0.5f)), xn = (int)floorf(__fadd_rn(xs, 0.5f));
  p.ol = ((unsigned)yn < (unsigned)H && (unsigned)xn < (unsigned)W) ? yn * W + xn : -1;
  return p;
}

// image (bilinear, zeros outside) and label (nearest) of one source plane under a plan
__device__ __forceinline__ void apply_plan(const ScPlan& p, const float* __restrict__ I, const uint8_t* __restrict__ L, int W,
                                           float& v, uint8_t& l) {
  const float t00 = (p.mask & 1u) ? I[p.o00] : 0.f, t01 = (p.mask & 2u) ? I[p.o00 + 1] : 0.f;
  const float t10 = (p.mask & 4u) ? I[p.o00 + W] : 0.f, t11 = (p.mask & 8u) ? I[p.o00 + W + 1] : 0.f;
  l = (p.ol >= 0) ? L[p.ol] : (uint8_t)0;
  const float top = (1.0f - p.fx) * t00 + p.fx * t01, bot = (1.0f - p.fx) * t10 + p.fx * t11;
  v = (1.0f - p.fy) * top + p.fy * bot;
}

constexpr int kScRows = 4;      // output rows per thread
constexpr int kScImgs = 16;      // images per block (blockIdx.z = chunk of images)

// output = polar [n_s, n_r]; thread = ray, kScRows samples
__global__ void __launch_bounds__(128, 8) sc_cart_to_polar(const float* __restrict__ img, const uint8_t* __restrict__ lab,
                                                        float* __restrict__ out_img, uint8_t* __restrict__ out_lab,
                                                        int B, int H, int W, int ns, int nr, ScGeom g,
                                                        const float* __restrict__ sin_t, const float* __restrict__ cos_t) {
  const int j = blockIdx.x * 128 + threadIdx.x;
  if (j >= nr) return;
  const float s = sin_t[j], c = cos_t[j];
  const int i0 = blockIdx.y * kScRows;
  ScPlan plan[kScRows];
#pragma unroll
  for (int k = 0; k < kScRows; ++k) {
    const float r = __fadd_rn(g.r_min, __fmul_rn((float)(i0 + k), g.dr));
    const float x = __fadd_rn(g.apex_x, __fmul_rn(r, s)), y = __fadd_rn(g.apex_y, __fmul_rn(r, c));
    plan[k] = make_plan(H, W, y, x);
  }
  const int b1 = min(B, (int)(blockIdx.z + 1) * kScImgs);
#pragma unroll 1
  for (int b = blockIdx.z * kScImgs; b < b1; ++b) {
    const float* I = img + (size_t)b * H * W;
    const uint8_t* L = lab + (size_t)b * H * W;
    float v[kScRows]; uint8_t l[kScRows];
#pragma unroll
    for (int k = 0; k < kScRows; ++k) apply_plan(plan[k], I, L, W, v[k], l[k]);
#pragma unroll
    for (int k = 0; k < kScRows; ++k) {
      if (i0 + k < ns) {
        const size_t o = ((size_t)b * ns + i0 + k) * nr + j;
        out_img[o] = v[k]; out_lab[o] = l[k];
      }
    }
  }
}

// output = Cartesian [H, W]; thread = column, kScRows rows
__global__ void __launch_bounds__(128, 8) sc_polar_to_cart(const float* __restrict__ img, const uint8_t* __restrict__ lab,
                                                        float* __restrict__ out_img, uint8_t* __restrict__ out_lab,
                                                        int B, int H, int W, int ns, int nr, ScGeom g) {
  const int x = blockIdx.x * 128 + threadIdx.x;
  if (x >= W) return;
  const float dx = __fsub_rn((float)x, g.apex_x);
  const float dx2 = __fmul_rn(dx, dx);
  const int y0 = blockIdx.y * kScRows;
  ScPlan plan[kScRows];
#pragma unroll
  for (int k = 0; k < kScRows; ++k) {
    const float dy = __fsub_rn((float)(y0 + k), g.apex_y);
    plan[k].o00 = 0; plan[k].ol = -1; plan[k].fx = 0.f; plan[k].fy = 0.f; plan[k].mask = 0u;   // above the apex: outside the sector
    if (dy > 0.f) {
      const float r = __fsqrt_rn(__fadd_rn(dx2, __fmul_rn(dy, dy)));
      const float th = spec_atan2_pos(dx, dy);
      const float is = __fmul_rn(__fsub_rn(r, g.r_min), g.inv_dr);
      const float js = __fmul_rn(__fadd_rn(th, g.theta_max), g.inv_dth);
      plan[k] = make_plan(ns, nr, is, js);
    }
  }
  const int b1 = min(B, (int)(blockIdx.z + 1) * kScImgs);
#pragma unroll 1
  for (int b = blockIdx.z * kScImgs; b < b1; ++b) {
    const float* I = img + (size_t)b * ns * nr;
    const uint8_t* L = lab + (size_t)b * ns * nr;
    float v[kScRows]; uint8_t l[kScRows];
#pragma unroll
    for (int k = 0; k < kScRows; ++k) apply_plan(plan[k], I, L, nr, v[k], l[k]);
#pragma unroll
    for (int k = 0; k < kScRows; ++k) {
      if (y0 + k < H) {
        const size_t o = ((size_t)b * H + y0 + k) * W + x;
        out_img[o] = v[k]; out_lab[o] = l[k];
      }
    }
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
    dim3 grid((n_rays + 127) / 128, (n_samples + kScRows - 1) / kScRows, (B + kScImgs - 1) / kScImgs);
    sc_cart_to_polar<<<grid, 128, 0, stream>>>(img, lab, out_img, out_lab, B, H, W, n_samples, n_rays, g, ray_sin, ray_cos);
    USAUG_CHECK_LAUNCH("sc_cart_to_polar");
  } else {
    dim3 grid((W + 127) / 128, (H + kScRows - 1) / kScRows, (B + kScImgs - 1) / kScImgs);
    sc_polar_to_cart<<<grid, 128, 0, stream>>>(img, lab, out_img, out_lab, B, H, W, n_samples, n_rays, g);
    USAUG_CHECK_LAUNCH("sc_polar_to_cart");
  }
  return USAUG_OK;
}
