// Monogenic-signal local energy (oracle/local_energy.py):
//   F = FFT2(I);  even = Re IFFT2(G F);  q = IFFT2(G (v' - i u')/rho F);  E = sqrt(even^2 + |q|^2)
// with G = sum of log-Gabor band-pass filters.  Hand-written batched 2-D FFT for power-of-two sizes on the block-level
// mixed-radix Stockham transform of fft_stockham.cuh (register radix-8 passes, swizzled shared memory, natural order
// in and out):
//
//   le_twiddles   the twiddle tables of the two transform lengths, once per call (2 blocks)
//   le_rows_fwd   per group of image rows: real -> complex FFT along x
//   le_cols       per tile of columns: FFT along y, filter multiply, two inverse FFTs along y
//   le_rows_inv   per group of image rows: inverse FFT of both planes along x, energy, per-block maxima
//   le_scale      per image: 1 / max E (0 for an image without band-pass content)
//
// Sizes that are not powers of two are mirror-extended (bottom / right, edge included) to the next power of two
// (Hp x Wp) on the way into le_rows_fwd; rows and columns beyond H x W are never written back.
// HBM traffic: 4 + 8 | 8 + 16 | 16 + 4 = 56 B/px; everything else stays in shared memory.
#include "fft_stockham.cuh"
#include "kernels.cuh"

namespace usaug {

constexpr int kLeThreads = 256;       // fft::block_fft needs blockDim.x * 8 to be a multiple of the transform length
constexpr int kLeMaxScales = 4;
constexpr int kLeMaxN = 2048;
constexpr int kLeColPad = 2;          // column sequences are H + 2 apart: the transposing copies of a tile spread over the banks

struct LeFilter {
  int scales;
  float ln_lambda[kLeMaxScales];
  float inv_den;     // 1 / (2 ln^2 sigma_f)
};

// signed DFT frequency index of bin k (numpy.fft.fftfreq * n): the Nyquist bin is -n/2
__device__ __forceinline__ int signed_bin(int k, int n) { return (k < n / 2) ? k : k - n; }

__device__ __forceinline__ void load_twiddles(float2* tw, const float2* __restrict__ twg, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) tw[i] = twg[i];
}

__global__ void __launch_bounds__(kLeThreads) le_twiddles(fft::Plan pw, fft::Plan ph, float2* __restrict__ tww,
                                                          float2* __restrict__ twh) {
  if (blockIdx.x == 0) fft::fill_twiddles(pw, tww, threadIdx.x, blockDim.x);
  else fft::fill_twiddles(ph, twh, threadIdx.x, blockDim.x);
}

// The filter of frequency (ky, kx), the same for every image of the call: filt[kx * H + ky] = (G, G u'/rho, G v'/rho, 0)
// with the odd multipliers zero at the Nyquist bins (column-major: le_cols walks ky fastest).
__global__ void __launch_bounds__(kLeThreads) le_filter(float4* __restrict__ filt, int H, int W, int lgH, LeFilter flt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const int kv = signed_bin(i & (H - 1), H), ku = signed_bin(i >> lgH, W);
  const float u = (float)ku / (float)W, v = (float)kv / (float)H;
  const float rho2 = fmaf(u, u, v * v);
  float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rho2 > 0.f) {
    const float lr = 0.5f * logf(rho2);
    float g = 0.f;
    for (int s = 0; s < flt.scales; ++s) {
      const float t = lr + flt.ln_lambda[s];
      g += expf(-t * t * flt.inv_den);
    }
    const float gr = g * rsqrtf(rho2);
    f = make_float4(g, (2 * ku == -W) ? 0.f : u * gr, (2 * kv == -H) ? 0.f : v * gr, 0.f);
  }
  filt[i] = f;
}

// ---- rows, forward: block = rb rows ---------------------------------------------------------------
// img is [B][H0][W0]; spec is [B][H][W] with H, W the padded (power-of-two) sizes.  The copies move two points per
// thread and step: x and x + 1 (x even) live at the shared-memory slots s and s ^ 1.
__global__ void __launch_bounds__(kLeThreads) le_rows_fwd(const float* __restrict__ img, float2* __restrict__ spec,
                                                          int H0, int W0, int H, int W, int lgH, fft::Plan pw,
                                                          const float2* __restrict__ twg, int rb, size_t nrows) {
  extern __shared__ __align__(16) unsigned char le_smem[];
  float2* tw = reinterpret_cast<float2*>(le_smem);          // [W]  (pw.tw_total <= W entries used)
  float2* data = tw + W;                                    // [rb][W], element x of a row at slot swz(x)
  const size_t row0 = (size_t)blockIdx.x * rb;
  const int lgW = pw.lg;
  // no mirror extension: rows are pairs of 8 bytes (if the caller's pointer is that aligned)
  const bool exact = (H0 == H) && (W0 == W) && (reinterpret_cast<size_t>(img) & 7) == 0;
  load_twiddles(tw, twg, pw.tw_total);
  for (int i = 2 * threadIdx.x; i < rb * W; i += 2 * blockDim.x) {
    const int q = i >> lgW, xp = i & (W - 1);
    const size_t r = row0 + q;
    float v0 = 0.f, v1 = 0.f;
    if (r < nrows) {
      if (exact) {
        const float2 t = *reinterpret_cast<const float2*>(img + r * W + xp);
        v0 = t.x; v1 = t.y;
      } else {
        const int b = (int)(r >> lgH), yp = (int)(r & (size_t)(H - 1));
        const int y = (yp < H0) ? yp : 2 * H0 - 1 - yp;                                        // mirror extension
        const int xa = (xp < W0) ? xp : 2 * W0 - 1 - xp, xb = (xp + 1 < W0) ? xp + 1 : 2 * W0 - 2 - xp;
        const float* row = img + ((size_t)b * H0 + y) * W0;
        v0 = row[xa]; v1 = row[xb];
      }
    }
    const int s = (q << lgW) + fft::swz(xp);
    data[s] = make_float2(v0, 0.f);
    data[s ^ 1] = make_float2(v1, 0.f);
  }
  __syncthreads();
  fft::block_fft<false>(data, W, rb, pw, tw);
  for (int i = 2 * threadIdx.x; i < rb * W; i += 2 * blockDim.x) {
    const int q = i >> lgW, xp = i & (W - 1);
    const size_t r = row0 + q;
    if (r < nrows) {
      const int s = (q << lgW) + fft::swz(xp);
      const float2 a = data[s], c = data[s ^ 1];
      *reinterpret_cast<float4*>(spec + r * W + xp) = make_float4(a.x, a.y, c.x, c.y);
    }
  }
}

// ---- columns: block = tc columns of one image ---------------------------------------------------------
__global__ void __launch_bounds__(kLeThreads) le_cols(const float2* __restrict__ spec, float2* __restrict__ odd,
                                                      float2* __restrict__ even, const float4* __restrict__ filt,
                                                      int H0, int H, int W, fft::Plan ph,
                                                      const float2* __restrict__ twg, int tc, int lg_tc) {
  extern __shared__ __align__(16) unsigned char le_smem[];
  const int pitch = H + kLeColPad, lgH = ph.lg;
  float2* tw = reinterpret_cast<float2*>(le_smem);          // [H]
  float2* A = tw + H;                                       // [tc][pitch]  spectrum, then the Riesz (odd) pair
  float2* Bv = A + tc * pitch;                              // [tc][pitch]  the even part
  const int b = blockIdx.y, x0 = blockIdx.x * tc;
  const size_t base = (size_t)b * H * W;
  load_twiddles(tw, twg, ph.tw_total);
  if (tc >= 2) {                                            // two columns per thread and step: 16-byte accesses
    for (int i = 2 * threadIdx.x; i < tc * H; i += 2 * blockDim.x) {
      const int c = i & (tc - 1), y = i >> lg_tc;
      const float4 t = *reinterpret_cast<const float4*>(spec + base + (size_t)y * W + x0 + c);
      const int s = c * pitch + fft::swz(y);
      A[s] = make_float2(t.x, t.y);
      A[s + pitch] = make_float2(t.z, t.w);
    }
  } else {
    for (int y = threadIdx.x; y < H; y += blockDim.x) A[fft::swz(y)] = spec[base + (size_t)y * W + x0];
  }
  __syncthreads();
  fft::block_fft<false>(A, pitch, tc, ph, tw);
  const float4* fcol = filt + (size_t)x0 * H;               // the tile's filter columns are contiguous
  for (int i = threadIdx.x; i < tc * H; i += blockDim.x) {
    const int c = i >> lgH, py = i & (H - 1);
    const int slot = c * pitch + fft::swz(py);
    const float4 f = fcol[i];                               // (G, G u'/rho, G v'/rho): zero at DC
    const float2 F = A[slot];
    // even: G F;  odd: (v' - i u') / rho * G * F
    Bv[slot] = make_float2(f.x * F.x, f.x * F.y);
    A[slot] = make_float2(fmaf(f.z, F.x, f.y * F.y), fmaf(f.z, F.y, -f.y * F.x));
  }
  __syncthreads();
  fft::block_fft<true>(A, pitch, 2 * tc, ph, tw);           // A and Bv are contiguous: 2 tc sequences
  // rows of the mirror extension are not needed again
  if (tc >= 2) {
    for (int i = 2 * threadIdx.x; i < tc * H0; i += 2 * blockDim.x) {
      const int c = i & (tc - 1), y = i >> lg_tc;
      const size_t o = base + (size_t)y * W + x0 + c;
      const int s = c * pitch + fft::swz(y);
      const float2 a0 = A[s], a1 = A[s + pitch], e0 = Bv[s], e1 = Bv[s + pitch];
      *reinterpret_cast<float4*>(odd + o) = make_float4(a0.x, a0.y, a1.x, a1.y);
      *reinterpret_cast<float4*>(even + o) = make_float4(e0.x, e0.y, e1.x, e1.y);
    }
  } else {
    for (int y = threadIdx.x; y < H0; y += blockDim.x) {
      const size_t o = base + (size_t)y * W + x0;
      odd[o] = A[fft::swz(y)];
      even[o] = Bv[fft::swz(y)];
    }
  }
}

// ---- rows, inverse + energy: block = rb rows -----------------------------------------------------------
__global__ void __launch_bounds__(kLeThreads) le_rows_inv(const float2* __restrict__ odd, const float2* __restrict__ even,
                                                          float* __restrict__ energy, float* __restrict__ part,
                                                          int H0, int W0, int H, int W, fft::Plan pw,
                                                          const float2* __restrict__ twg, int rb, float norm) {
  extern __shared__ __align__(16) unsigned char le_smem[];
  __shared__ float s_max[kLeThreads / 32];
  float2* tw = reinterpret_cast<float2*>(le_smem);          // [W]
  float2* data = tw + W;                                    // [2 rb][W]: odd rows, then even rows
  const int b = blockIdx.y, lgW = pw.lg;
  const int row0 = blockIdx.x * rb;                         // H % rb == 0; only blocks with row0 < H0 are launched
  const size_t base = ((size_t)b * H + row0) * W;
  load_twiddles(tw, twg, pw.tw_total);
  for (int i = 2 * threadIdx.x; i < rb * W; i += 2 * blockDim.x) {
    const int s = (i & ~(W - 1)) + fft::swz(i & (W - 1));
    const float4 o = *reinterpret_cast<const float4*>(odd + base + i);
    const float4 e = *reinterpret_cast<const float4*>(even + base + i);
    data[s] = make_float2(o.x, o.y); data[s ^ 1] = make_float2(o.z, o.w);
    data[rb * W + s] = make_float2(e.x, e.y); data[rb * W + (s ^ 1)] = make_float2(e.z, e.w);
  }
  __syncthreads();
  fft::block_fft<true>(data, W, 2 * rb, pw, tw);
  float m = 0.f;
  // even row length: x and x + 1 as one 8-byte store
  const bool pairs = (W0 & 1) == 0 && (reinterpret_cast<size_t>(energy) & 7) == 0;
  for (int i = 2 * threadIdx.x; i < rb * W; i += 2 * blockDim.x) {
    const int y = row0 + (i >> lgW), x = i & (W - 1);
    if (y >= H0 || x >= W0) continue;                       // mirror extension: not part of the result
    const int s = (i & ~(W - 1)) + fft::swz(x);
    const float2 q0 = data[s], q1 = data[s ^ 1];
    const float e0 = data[rb * W + s].x, e1 = data[rb * W + (s ^ 1)].x;
    const float en0 = sqrtf(fmaf(e0, e0, fmaf(q0.x, q0.x, q0.y * q0.y))) * norm;
    const float en1 = sqrtf(fmaf(e1, e1, fmaf(q1.x, q1.x, q1.y * q1.y))) * norm;
    float* dst = energy + ((size_t)b * H0 + y) * W0 + x;
    if (pairs) {
      *reinterpret_cast<float2*>(dst) = make_float2(en0, en1);
      m = fmaxf(m, fmaxf(en0, en1));
    } else {
      dst[0] = en0; m = fmaxf(m, en0);
      if (x + 1 < W0) { dst[1] = en1; m = fmaxf(m, en1); }
    }
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
  float2 *spec, *odd, *even, *tww, *twh;
  float4* filt;       // [W][H] filter multipliers, shared by all images of the call
  float* part;
  int rb_fwd, rb_inv, tc, nblk_inv;
  size_t bytes;
  Carver cv{nullptr};
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
  L.tww = cv.take<float2>(kLeMaxN); L.twh = cv.take<float2>(kLeMaxN);
  L.filt = cv.take<float4>((size_t)H * W);
  L.part = cv.take<float>((size_t)B * L.nblk_inv);
  L.bytes = cv.off;
  L.cv = cv;
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
  LeLayout L = le_layout(ws, B, H0, H, W);
  if (int zrc = L.cv.arm(stream)) return zrc;
  const int lgH = ilog2_exact(H), lgW = ilog2_exact(W);
  LeFilter flt{};
  flt.scales = scales;
  for (int s = 0; s < scales; ++s) flt.ln_lambda[s] = (float)log((double)lambda_min * pow((double)mult, s));
  const double lsf = log((double)sigma_f);
  flt.inv_den = (float)(1.0 / (2.0 * lsf * lsf));

  const size_t nrows = (size_t)B * H;
  const fft::Plan pw = fft::make_plan(lgW), ph = fft::make_plan(lgH);
  // twiddle table + the sequences, rounded up to whole transform steps (fft::padded_seqs)
  const size_t sm_fwd = (size_t)(W + fft::padded_seqs(L.rb_fwd, lgW, kLeThreads) * W) * sizeof(float2);
  const size_t sm_col = (size_t)(H + fft::padded_seqs(2 * L.tc, lgH, kLeThreads) * (H + kLeColPad)) * sizeof(float2);
  const size_t sm_inv = (size_t)(W + fft::padded_seqs(2 * L.rb_inv, lgW, kLeThreads) * W) * sizeof(float2);
  if (sm_fwd > 48 * 1024) USAUG_CUDA(cudaFuncSetAttribute(le_rows_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_fwd));
  if (sm_col > 48 * 1024) USAUG_CUDA(cudaFuncSetAttribute(le_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_col));
  if (sm_inv > 48 * 1024) USAUG_CUDA(cudaFuncSetAttribute(le_rows_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_inv));

  le_twiddles<<<2, kLeThreads, 0, stream>>>(pw, ph, L.tww, L.twh);
  USAUG_CHECK_LAUNCH("le_twiddles");
  le_filter<<<(unsigned)(((size_t)H * W + kLeThreads - 1) / kLeThreads), kLeThreads, 0, stream>>>(L.filt, H, W, lgH, flt);
  USAUG_CHECK_LAUNCH("le_filter");
  le_rows_fwd<<<(unsigned)((nrows + L.rb_fwd - 1) / L.rb_fwd), kLeThreads, sm_fwd, stream>>>(img, L.spec, H0, W0, H, W, lgH, pw,
                                                                                            L.tww, L.rb_fwd, nrows);
  USAUG_CHECK_LAUNCH("le_rows_fwd");
  le_cols<<<dim3(W / L.tc, B), kLeThreads, sm_col, stream>>>(L.spec, L.odd, L.even, L.filt, H0, H, W, ph, L.twh, L.tc,
                                                             ilog2_exact(L.tc));
  USAUG_CHECK_LAUNCH("le_cols");
  le_rows_inv<<<dim3(L.nblk_inv, B), kLeThreads, sm_inv, stream>>>(L.odd, L.even, energy, L.part, H0, W0, H, W, pw, L.tww,
                                                                  L.rb_inv, (float)(1.0 / ((double)H * (double)W)));
  USAUG_CHECK_LAUNCH("le_rows_inv");
  le_scale<<<(B + 127) / 128, 128, 0, stream>>>(L.part, L.nblk_inv, scale, B);
  USAUG_CHECK_LAUNCH("le_scale");
  return L.cv.verify(stream, "usaug_local_energy");
}
