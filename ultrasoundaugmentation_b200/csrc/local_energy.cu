// Monogenic-signal local energy (oracle/local_energy.py):
//   F = FFT2(I);  even = Re IFFT2(G F);  q = IFFT2(G (v' - i u')/rho F);  E = sqrt(even^2 + |q|^2)
// with G = sum of log-Gabor band-pass filters.  Hand-written batched 2-D FFT for power-of-two sizes:
//
//   le_rows_fwd   per image row: real -> complex radix-2 DIF FFT in shared memory (natural in, BIT-REVERSED out)
//   le_cols       per tile of columns: DIF FFT along y, filter multiply (frequencies from the bit-reversed
//                 positions), two radix-2 DIT inverse FFTs along y (bit-reversed in, natural out)
//   le_rows_inv   per image row: DIT inverse FFT of both planes along x, energy, per-block maxima
//   le_scale      per image: 1 / max E (0 for an image without band-pass content)
//
// Pairing a DIF forward with a DIT inverse transform means no bit-reversal permutation is ever executed.
// Sizes that are not powers of two are mirror-extended (bottom / right, edge included) to the next power of two
// (Hp x Wp) on the way into le_rows_fwd; rows and columns beyond H x W are never written back.
// HBM traffic: 4 + 8 | 8 + 16 | 16 + 4 = 56 B/px; everything else stays in shared memory.
#include "kernels.cuh"

namespace usaug {

constexpr int kLeThreads = 256;
constexpr int kLeMaxScales = 4;
constexpr int kLeMaxN = 2048;

struct LeFilter {
  int scales;
  float ln_lambda[kLeMaxScales];
  float inv_den;     // 1 / (2 ln^2 sigma_f)
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// tw[k] = exp(-2 pi i k / n), k < n/2
__device__ __forceinline__ void make_twiddles(float2* tw, int n) {
  for (int k = threadIdx.x; k < n / 2; k += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)k / (float)n, &s, &c);
    tw[k] = make_float2(c, -s);
  }
}

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// nseq sequences of n = 1 << lg points, sequence q at data + q * pitch.  Natural order in, bit-reversed order out.
// Radix-2 decimation in frequency, two stages per shared-memory pass (a thread owns a quad of points), so an
// n-point transform takes ceil(lg / 2) passes and barriers instead of lg.
__device__ __forceinline__ void fft_dif(float2* data, int pitch, int nseq, int lg, const float2* tw) {
  int s = lg - 1;
  for (; s >= 1; s -= 2) {                       // stages s (half = 2^s) and s-1 (half = 2^(s-1))
    const int quarter = 1 << (s - 1);
    const int nquad = 1 << (lg - 2);
    for (int idx = threadIdx.x; idx < nseq * nquad; idx += blockDim.x) {
      const int q = idx >> (lg - 2), j = idx & (nquad - 1);
      const int pos = j & (quarter - 1);
      float2* x = data + q * pitch + (((j >> (s - 1)) << (s + 1)) | pos);
      const float2 a0 = x[0], a1 = x[quarter], a2 = x[2 * quarter], a3 = x[3 * quarter];
      // stage s: (0,2) and (1,3)
      const float2 b0 = cadd(a0, a2), b2 = cmul(csub(a0, a2), tw[pos << (lg - 1 - s)]);
      const float2 b1 = cadd(a1, a3), b3 = cmul(csub(a1, a3), tw[(pos + quarter) << (lg - 1 - s)]);
      // stage s-1: (0,1) and (2,3)
      const float2 w = tw[pos << (lg - s)];
      x[0] = cadd(b0, b1); x[quarter] = cmul(csub(b0, b1), w);
      x[2 * quarter] = cadd(b2, b3); x[3 * quarter] = cmul(csub(b2, b3), w);
    }
    __syncthreads();
  }
  if (s == 0) {                                  // odd lg: one plain radix-2 stage with half = 1 (twiddle 1)
    const int half_n = 1 << (lg - 1);
    for (int idx = threadIdx.x; idx < nseq * half_n; idx += blockDim.x) {
      const int q = idx >> (lg - 1), j = idx & (half_n - 1);
      float2* x = data + q * pitch + 2 * j;
      const float2 a = x[0], b = x[1];
      x[0] = cadd(a, b); x[1] = csub(a, b);
    }
    __syncthreads();
  }
}

// inverse transform (conjugate twiddles, unscaled).  Bit-reversed order in, natural order out.
// Radix-2 decimation in time, two stages per pass like fft_dif.
__device__ __forceinline__ void ifft_dit(float2* data, int pitch, int nseq, int lg, const float2* tw) {
  int s = 0;
  if (lg & 1) {                                  // odd lg: stage 0 alone (half = 1, twiddle 1)
    const int half_n = 1 << (lg - 1);
    for (int idx = threadIdx.x; idx < nseq * half_n; idx += blockDim.x) {
      const int q = idx >> (lg - 1), j = idx & (half_n - 1);
      float2* x = data + q * pitch + 2 * j;
      const float2 a = x[0], b = x[1];
      x[0] = cadd(a, b); x[1] = csub(a, b);
    }
    __syncthreads();
    s = 1;
  }
  for (; s + 1 < lg; s += 2) {                   // stages s (half = 2^s) and s+1 (half = 2^(s+1))
    const int half = 1 << s;
    const int nquad = 1 << (lg - 2);
    for (int idx = threadIdx.x; idx < nseq * nquad; idx += blockDim.x) {
      const int q = idx >> (lg - 2), j = idx & (nquad - 1);
      const int pos = j & (half - 1);
      float2* x = data + q * pitch + (((j >> s) << (s + 2)) | pos);
      const float2 w = cconj(tw[pos << (lg - 1 - s)]);
      const float2 a0 = x[0], t1 = cmul(x[half], w), a2 = x[2 * half], t3 = cmul(x[3 * half], w);
      // stage s: (0,1) and (2,3)
      const float2 b0 = cadd(a0, t1), b1 = csub(a0, t1), b2 = cadd(a2, t3), b3 = csub(a2, t3);
      // stage s+1: (0,2) and (1,3)
      const float2 u2 = cmul(b2, cconj(tw[pos << (lg - 2 - s)]));
      const float2 u3 = cmul(b3, cconj(tw[(pos + half) << (lg - 2 - s)]));
      x[0] = cadd(b0, u2); x[2 * half] = csub(b0, u2);
      x[half] = cadd(b1, u3); x[3 * half] = csub(b1, u3);
    }
    __syncthreads();
  }
}

// signed DFT frequency index of bin k (numpy.fft.fftfreq * n): the Nyquist bin is -n/2
__device__ __forceinline__ int signed_bin(int k, int n) { return (k < n / 2) ? k : k - n; }

// ---- rows, forward: block = rb rows ---------------------------------------------------------------
// img is [B][H0][W0]; spec is [B][H][W] with H, W the padded (power-of-two) sizes
__global__ void __launch_bounds__(kLeThreads) le_rows_fwd(const float* __restrict__ img, float2* __restrict__ spec,
                                                          int H0, int W0, int H, int W, int lgW, int rb, size_t nrows) {
  extern __shared__ __align__(16) unsigned char le_smem[];
  float2* tw = reinterpret_cast<float2*>(le_smem);          // [W/2]
  float2* data = tw + W / 2;                                // [rb][W]
  const size_t row0 = (size_t)blockIdx.x * rb;
  make_twiddles(tw, W);
  for (int i = threadIdx.x; i < rb * W; i += blockDim.x) {
    const size_t r = row0 + i / W;
    float v = 0.f;
    if (r < nrows) {
      const int b = (int)(r / H), yp = (int)(r % H), xp = i % W;
      const int y = (yp < H0) ? yp : 2 * H0 - 1 - yp, x = (xp < W0) ? xp : 2 * W0 - 1 - xp;   // mirror extension
      v = img[((size_t)b * H0 + y) * W0 + x];
    }
    data[i] = make_float2(v, 0.f);
  }
  __syncthreads();
  fft_dif(data, W, rb, lgW, tw);
  for (int i = threadIdx.x; i < rb * W; i += blockDim.x) {
    const size_t r = row0 + i / W;
    if (r < nrows) spec[r * W + (i % W)] = data[i];
  }
}

// ---- columns: block = tc columns of one image ---------------------------------------------------------
__global__ void __launch_bounds__(kLeThreads) le_cols(const float2* __restrict__ spec, float2* __restrict__ odd,
                                                      float2* __restrict__ even, int H0, int H, int W, int lgH, int lgW,
                                                      int tc, LeFilter flt) {
  extern __shared__ __align__(16) unsigned char le_smem[];
  const int pitch = H + 1;                                  // odd pitch: the tc columns of one row hit different banks
  float2* tw = reinterpret_cast<float2*>(le_smem);          // [H/2]
  float2* A = tw + H / 2;                                   // [tc][H+1]  spectrum, then the Riesz (odd) pair
  float2* Bv = A + tc * pitch;                              // [tc][H+1]  the even part
  const int b = blockIdx.y, x0 = blockIdx.x * tc;
  const size_t base = (size_t)b * H * W;
  make_twiddles(tw, H);
  for (int i = threadIdx.x; i < tc * H; i += blockDim.x) {
    const int c = i % tc, y = i / tc;
    A[c * pitch + y] = spec[base + (size_t)y * W + x0 + c];
  }
  __syncthreads();
  fft_dif(A, pitch, tc, lgH, tw);
  // position (py, px) holds frequency (rev(py), rev(px)): the row pass left x in bit-reversed order too
  for (int i = threadIdx.x; i < tc * H; i += blockDim.x) {
    const int c = i / H, py = i % H;
    const int kv = signed_bin((int)(__brev((unsigned)py) >> (32 - lgH)), H);
    const int ku = signed_bin((int)(__brev((unsigned)(x0 + c)) >> (32 - lgW)), W);
    const float u = (float)ku / (float)W, v = (float)kv / (float)H;
    const float rho2 = fmaf(u, u, v * v);
    float2 fo = make_float2(0.f, 0.f), fe = make_float2(0.f, 0.f);
    if (rho2 > 0.f) {
      const float lr = 0.5f * logf(rho2);
      float g = 0.f;
      for (int s = 0; s < flt.scales; ++s) {
        const float t = lr + flt.ln_lambda[s];
        g += expf(-t * t * flt.inv_den);
      }
      const float2 F = A[c * pitch + py];
      fe = make_float2(g * F.x, g * F.y);
      // (v' - i u') / rho * G * F, odd multipliers zero at the Nyquist bins
      const float gr = g * rsqrtf(rho2);
      const float uo = (2 * ku == -W) ? 0.f : u * gr, vo = (2 * kv == -H) ? 0.f : v * gr;
      fo = make_float2(fmaf(vo, F.x, uo * F.y), fmaf(vo, F.y, -uo * F.x));
    }
    A[c * pitch + py] = fo;
    Bv[c * pitch + py] = fe;
  }
  __syncthreads();
  ifft_dit(A, pitch, 2 * tc, lgH, tw);                      // A and Bv are contiguous: 2 tc sequences
  for (int i = threadIdx.x; i < tc * H0; i += blockDim.x) {   // rows of the mirror extension are not needed again
    const int c = i % tc, y = i / tc;
    const size_t o = base + (size_t)y * W + x0 + c;
    odd[o] = A[c * pitch + y];
    even[o] = Bv[c * pitch + y];
  }
}

// ---- rows, inverse + energy: block = rb rows -----------------------------------------------------------
__global__ void __launch_bounds__(kLeThreads) le_rows_inv(const float2* __restrict__ odd, const float2* __restrict__ even,
                                                          float* __restrict__ energy, float* __restrict__ part,
                                                          int H0, int W0, int H, int W, int lgW, int rb, float norm) {
  extern __shared__ __align__(16) unsigned char le_smem[];
  __shared__ float s_max[kLeThreads / 32];
  float2* tw = reinterpret_cast<float2*>(le_smem);          // [W/2]
  float2* data = tw + W / 2;                                // [2 rb][W]: odd rows, then even rows
  const int b = blockIdx.y;
  const int row0 = blockIdx.x * rb;                         // H % rb == 0; only blocks with row0 < H0 are launched
  const size_t base = ((size_t)b * H + row0) * W;
  make_twiddles(tw, W);
  for (int i = threadIdx.x; i < rb * W; i += blockDim.x) {
    data[i] = odd[base + i];
    data[rb * W + i] = even[base + i];
  }
  __syncthreads();
  ifft_dit(data, W, 2 * rb, lgW, tw);
  float m = 0.f;
  for (int i = threadIdx.x; i < rb * W; i += blockDim.x) {
    const int y = row0 + i / W, x = i % W;
    if (y >= H0 || x >= W0) continue;                       // mirror extension: not part of the result
    const float2 q = data[i];
    const float e = data[rb * W + i].x;
    const float en = sqrtf(fmaf(e, e, fmaf(q.x, q.x, q.y * q.y))) * norm;
    energy[((size_t)b * H0 + y) * W0 + x] = en;
    m = fmaxf(m, en);
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kLeThreads / 32; ++w) m = fmaxf(m, s_max[w]);
    part[(size_t)b * gridDim.x + blockIdx.x] = m;
  }
}

__global__ void le_scale(const float* __restrict__ part, int nblk, float* __restrict__ scale, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float m = 0.f;
  for (int i = 0; i < nblk; ++i) m = fmaxf(m, part[(size_t)b * nblk + i]);
  scale[b] = (m > 0.f) ? __fdiv_rn(1.0f, m) : 0.f;
}

static int ilog2_exact(int n) {
  int lg = 0;
  while ((1 << lg) < n) ++lg;
  return ((1 << lg) == n) ? lg : -1;
}

struct LeLayout {
  float2 *spec, *odd, *even;
  float* part;
  int rb_fwd, rb_inv, tc, nblk_inv;
  size_t bytes;
};

static int next_pow2_min8(int n) {
  int p = 8;
  while (p < n) p <<= 1;
  return p;
}

// H, W: padded sizes; H0: image rows
static LeLayout le_layout(void* ws, int B, int H0, int H, int W) {
  LeLayout L;
  const size_t N = (size_t)B * H * W;
  // rows per block: 32 KB of sequences (forward) / 64 KB (inverse: two planes), at least 1, at most 16
  L.rb_fwd = 4096 / W; if (L.rb_fwd < 1) L.rb_fwd = 1; if (L.rb_fwd > 16) L.rb_fwd = 16;
  L.rb_inv = 4096 / W; if (L.rb_inv < 1) L.rb_inv = 1; if (L.rb_inv > 16) L.rb_inv = 16;
  while (H % L.rb_inv) L.rb_inv >>= 1;
  // columns per block: 64 KB of sequences (two buffers) so that three blocks share an SM
  L.tc = 4096 / H; if (L.tc < 1) L.tc = 1; if (L.tc > 16) L.tc = 16; if (L.tc > W) L.tc = W;
  L.nblk_inv = (H0 + L.rb_inv - 1) / L.rb_inv;
  Carver cv(ws);
  L.spec = cv.take<float2>(N); L.odd = cv.take<float2>(N); L.even = cv.take<float2>(N);
  L.part = cv.take<float>((size_t)B * L.nblk_inv);
  L.bytes = cv.off;
  return L;
}

}  // namespace usaug

using namespace usaug;

static bool le_shape_ok(int B, int H, int W) { return B > 0 && H >= 4 && W >= 4 && H <= kLeMaxN && W <= kLeMaxN; }

extern "C" size_t usaug_local_energy_workspace_bytes(int B, int H, int W) {
  if (!le_shape_ok(B, H, W)) return 0;
  return le_layout(nullptr, B, H, next_pow2_min8(H), next_pow2_min8(W)).bytes;
}

extern "C" int usaug_local_energy(const float* img, float* energy, float* scale, int B, int H0, int W0,
                                  float lambda_min, float mult, int scales, float sigma_f, void* ws,
                                  size_t ws_bytes, void* stream_) {
  if (!img || !energy || !scale || !ws) { set_error("usaug_local_energy: null pointer"); return USAUG_E_INVALID_ARG; }
  if (!le_shape_ok(B, H0, W0)) {
    set_error("usaug_local_energy: H=%d W=%d must lie in [4, %d] (B=%d)", H0, W0, kLeMaxN, B);
    return USAUG_E_INVALID_ARG;
  }
  const int H = next_pow2_min8(H0), W = next_pow2_min8(W0);   // mirror-extended transform size
  if (scales < 1 || scales > kLeMaxScales || !(lambda_min >= 2.f) || !(mult > 0.f) || !(sigma_f > 0.f) || !(sigma_f < 1.f)) {
    set_error("usaug_local_energy: bad filter (lambda_min=%g >= 2, mult=%g > 0, 1 <= scales=%d <= %d, 0 < sigma_f=%g < 1)",
              lambda_min, mult, scales, kLeMaxScales, sigma_f);
    return USAUG_E_INVALID_ARG;
  }
  if (ws_bytes < usaug_local_energy_workspace_bytes(B, H0, W0)) {
    set_error("usaug_local_energy: workspace too small"); return USAUG_E_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const LeLayout L = le_layout(ws, B, H0, H, W);
  const int lgH = ilog2_exact(H), lgW = ilog2_exact(W);
  LeFilter flt{};
  flt.scales = scales;
  for (int s = 0; s < scales; ++s) flt.ln_lambda[s] = (float)log((double)lambda_min * pow((double)mult, s));
  const double lsf = log((double)sigma_f);
  flt.inv_den = (float)(1.0 / (2.0 * lsf * lsf));

  const size_t nrows = (size_t)B * H;
  const size_t sm_fwd = (size_t)(W / 2 + L.rb_fwd * W) * sizeof(float2);
  const size_t sm_col = (size_t)(H / 2 + 2 * L.tc * (H + 1)) * sizeof(float2);
  const size_t sm_inv = (size_t)(W / 2 + 2 * L.rb_inv * W) * sizeof(float2);
  if (sm_fwd > 48 * 1024) USAUG_CUDA(cudaFuncSetAttribute(le_rows_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_fwd));
  if (sm_col > 48 * 1024) USAUG_CUDA(cudaFuncSetAttribute(le_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_col));
  if (sm_inv > 48 * 1024) USAUG_CUDA(cudaFuncSetAttribute(le_rows_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_inv));

  le_rows_fwd<<<(unsigned)((nrows + L.rb_fwd - 1) / L.rb_fwd), kLeThreads, sm_fwd, stream>>>(img, L.spec, H0, W0, H, W, lgW,
                                                                                            L.rb_fwd, nrows);
  USAUG_CHECK_LAUNCH("le_rows_fwd");
  le_cols<<<dim3(W / L.tc, B), kLeThreads, sm_col, stream>>>(L.spec, L.odd, L.even, H0, H, W, lgH, lgW, L.tc, flt);
  USAUG_CHECK_LAUNCH("le_cols");
  le_rows_inv<<<dim3(L.nblk_inv, B), kLeThreads, sm_inv, stream>>>(L.odd, L.even, energy, L.part, H0, W0, H, W, lgW,
                                                                  L.rb_inv, (float)(1.0 / ((double)H * (double)W)));
  USAUG_CHECK_LAUNCH("le_rows_inv");
  le_scale<<<(B + 127) / 128, 128, 0, stream>>>(L.part, L.nblk_inv, scale, B);
  USAUG_CHECK_LAUNCH("le_scale");
  return USAUG_OK;
}
