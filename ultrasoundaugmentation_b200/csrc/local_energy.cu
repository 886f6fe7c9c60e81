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

// ---- in-place radix-2 FFT of nseq sequences of n = 1 << lg points (sequence q at data + q * pitch) ----------
// Stage s pairs elements half = 2^s apart; its twiddle for position pos (< half) is tw[pos << (lg - 1 - s)].
//   * stages with half >= 128 run through shared memory, two stages per pass (a thread owns a quad): the four
//     accesses of a quad are >= 64 elements apart and consecutive threads touch consecutive elements -> no bank
//     conflicts;
//   * all stages with half <= 64 run INSIDE A WARP: a warp loads a 128-point block with contiguous accesses into four
//     registers per lane (elements lane, lane + 32, lane + 64, lane + 96), does half = 64 and 32 in-thread and
//     half = 16 .. 1 with shuffles, and writes the block back.  An n = 512 transform is 2 conflict-free
//     shared-memory round trips instead of 5 passes whose small strides replayed ~57 % of the wavefronts (ncu).
// fft_dif: decimation in frequency, natural order in, bit-reversed order out.  ifft_dit: inverse transform
// (conjugate twiddles, unscaled), decimation in time, bit-reversed order in, natural order out.

__device__ __forceinline__ void smem_radix2_stage(float2* data, int pitch, int nseq, int lg, int s, const float2* tw,
                                                  bool inverse) {
  const int half = 1 << s, half_n = 1 << (lg - 1);
  for (int idx = threadIdx.x; idx < nseq * half_n; idx += blockDim.x) {
    const int q = idx >> (lg - 1), j = idx & (half_n - 1);
    const int pos = j & (half - 1);
    float2* x = data + q * pitch + (((j >> s) << (s + 1)) | pos);
    float2 w = tw[pos << (lg - 1 - s)];
    if (inverse) {
      w.y = -w.y;
      const float2 a = x[0], t = cmul(x[half], w);
      x[0] = cadd(a, t); x[half] = csub(a, t);
    } else {
      const float2 a = x[0], c = x[half];
      x[0] = cadd(a, c); x[half] = cmul(csub(a, c), w);
    }
  }
  __syncthreads();
}

// stages s (half = 2^s) and s-1 of the forward transform in one pass
__device__ __forceinline__ void smem_radix4_dif(float2* data, int pitch, int nseq, int lg, int s, const float2* tw) {
  const int quarter = 1 << (s - 1), nquad = 1 << (lg - 2);
  for (int idx = threadIdx.x; idx < nseq * nquad; idx += blockDim.x) {
    const int q = idx >> (lg - 2), j = idx & (nquad - 1);
    const int pos = j & (quarter - 1);
    float2* x = data + q * pitch + (((j >> (s - 1)) << (s + 1)) | pos);
    const float2 a0 = x[0], a1 = x[quarter], a2 = x[2 * quarter], a3 = x[3 * quarter];
    const float2 b0 = cadd(a0, a2), b2 = cmul(csub(a0, a2), tw[pos << (lg - 1 - s)]);
    const float2 b1 = cadd(a1, a3), b3 = cmul(csub(a1, a3), tw[(pos + quarter) << (lg - 1 - s)]);
    const float2 w = tw[pos << (lg - s)];
    x[0] = cadd(b0, b1); x[quarter] = cmul(csub(b0, b1), w);
    x[2 * quarter] = cadd(b2, b3); x[3 * quarter] = cmul(csub(b2, b3), w);
  }
  __syncthreads();
}

// stages s (half = 2^s) and s+1 of the inverse transform in one pass
__device__ __forceinline__ void smem_radix4_dit(float2* data, int pitch, int nseq, int lg, int s, const float2* tw) {
  const int half = 1 << s, nquad = 1 << (lg - 2);
  for (int idx = threadIdx.x; idx < nseq * nquad; idx += blockDim.x) {
    const int q = idx >> (lg - 2), j = idx & (nquad - 1);
    const int pos = j & (half - 1);
    float2* x = data + q * pitch + (((j >> s) << (s + 2)) | pos);
    const float2 w = cconj(tw[pos << (lg - 1 - s)]);
    const float2 a0 = x[0], t1 = cmul(x[half], w), a2 = x[2 * half], t3 = cmul(x[3 * half], w);
    const float2 b0 = cadd(a0, t1), b1 = csub(a0, t1), b2 = cadd(a2, t3), b3 = csub(a2, t3);
    const float2 u2 = cmul(b2, cconj(tw[pos << (lg - 2 - s)]));
    const float2 u3 = cmul(b3, cconj(tw[(pos + half) << (lg - 2 - s)]));
    x[0] = cadd(b0, u2); x[2 * half] = csub(b0, u2);
    x[half] = cadd(b1, u3); x[3 * half] = csub(b1, u3);
  }
  __syncthreads();
}

__device__ __forceinline__ float2 shfl_xor2(float2 v, int m) {
  return make_float2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}

// All stages with half <= 16 * RPL of the blocks of 32 * RPL points, in registers (RPL = 1, 2 or 4 values per lane).
template <int RPL, bool INVERSE>
__device__ __forceinline__ void warp_stages(float2* data, int pitch, int nseq, int lg, const float2* tw) {
  constexpr int BS = 32 * RPL;                    // block size; the transform has n >= BS points
  constexpr int LB = (RPL == 4) ? 7 : (RPL == 2) ? 6 : 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int nblk = nseq << (lg - LB);             // blocks in all sequences
  // The twiddles of the register stages depend on the lane only: fetched once per call.  (Read per block, the
  // strided table reads of the shuffle stages were 16-way bank conflicts.)
  float2 wsh[5], w5 = make_float2(1.f, 0.f), w6a = w5, w6b = w5;
#pragma unroll
  for (int s = 0; s <= 4; ++s) wsh[s] = tw[(lane & ((1 << s) - 1)) << (lg - 1 - s)];
  if (RPL >= 2) w5 = tw[lane << (lg - 6)];
  if (RPL == 4) { w6a = tw[lane << (lg - 7)]; w6b = tw[(lane + 32) << (lg - 7)]; }
  if (INVERSE) {
#pragma unroll
    for (int s = 0; s <= 4; ++s) wsh[s].y = -wsh[s].y;
    w5.y = -w5.y; w6a.y = -w6a.y; w6b.y = -w6b.y;
  }
  for (int blk = warp; blk < nblk; blk += nwarp) {
    const int q = blk >> (lg - LB);
    float2* x = data + q * pitch + ((blk & ((1 << (lg - LB)) - 1)) << LB);
    float2 v[RPL];
#pragma unroll
    for (int k = 0; k < RPL; ++k) v[k] = x[k * 32 + lane];
    if (!INVERSE) {
      // half = 64: (0,2) (1,3); half = 32: (0,1) (2,3); element index inside the block = k * 32 + lane
      if (RPL == 4) {
        const float2 a0 = v[0], a1 = v[1];
        v[0] = cadd(a0, v[2]); v[2] = cmul(csub(a0, v[2]), w6a);
        v[1] = cadd(a1, v[3]); v[3] = cmul(csub(a1, v[3]), w6b);
      }
      if (RPL >= 2) {
#pragma unroll
        for (int k = 0; k < RPL; k += 2) {
          const float2 a = v[k];
          v[k] = cadd(a, v[k + 1]); v[k + 1] = cmul(csub(a, v[k + 1]), w5);
        }
      }
#pragma unroll
      for (int s = 4; s >= 0; --s) {              // half = 16 .. 1 across lanes
        const int h = 1 << s;
        const bool up = lane & h;
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
          const float2 o = shfl_xor2(v[k], h);
          v[k] = up ? cmul(csub(o, v[k]), wsh[s]) : cadd(v[k], o);
        }
      }
    } else {
#pragma unroll
      for (int s = 0; s <= 4; ++s) {              // half = 1 .. 16 across lanes
        const int h = 1 << s;
        const bool up = lane & h;
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
          const float2 t = up ? cmul(v[k], wsh[s]) : v[k];  // the upper element is twiddled before the exchange
          const float2 o = shfl_xor2(t, h);
          v[k] = up ? csub(o, t) : cadd(t, o);
        }
      }
      if (RPL >= 2) {
#pragma unroll
        for (int k = 0; k < RPL; k += 2) {
          const float2 t = cmul(v[k + 1], w5), a = v[k];
          v[k] = cadd(a, t); v[k + 1] = csub(a, t);
        }
      }
      if (RPL == 4) {
        const float2 t2 = cmul(v[2], w6a), t3 = cmul(v[3], w6b);
        const float2 a0 = v[0], a1 = v[1];
        v[0] = cadd(a0, t2); v[2] = csub(a0, t2);
        v[1] = cadd(a1, t3); v[3] = csub(a1, t3);
      }
    }
#pragma unroll
    for (int k = 0; k < RPL; ++k) x[k * 32 + lane] = v[k];
  }
  __syncthreads();
}

__device__ __forceinline__ void fft_dif(float2* data, int pitch, int nseq, int lg, const float2* tw) {
  if (lg < 5) {                                   // n = 8, 16: plain shared-memory stages
    for (int s = lg - 1; s >= 0; --s) smem_radix2_stage(data, pitch, nseq, lg, s, tw, false);
    return;
  }
  const int lb = lg < 7 ? lg : 7;                 // stages below lb run in registers
  int s = lg - 1;
  if ((lg - lb) & 1) { smem_radix2_stage(data, pitch, nseq, lg, s, tw, false); --s; }
  for (; s >= lb; s -= 2) smem_radix4_dif(data, pitch, nseq, lg, s, tw);
  if (lb == 7) warp_stages<4, false>(data, pitch, nseq, lg, tw);
  else if (lb == 6) warp_stages<2, false>(data, pitch, nseq, lg, tw);
  else warp_stages<1, false>(data, pitch, nseq, lg, tw);
}

__device__ __forceinline__ void ifft_dit(float2* data, int pitch, int nseq, int lg, const float2* tw) {
  if (lg < 5) {
    for (int s = 0; s < lg; ++s) smem_radix2_stage(data, pitch, nseq, lg, s, tw, true);
    return;
  }
  const int lb = lg < 7 ? lg : 7;
  if (lb == 7) warp_stages<4, true>(data, pitch, nseq, lg, tw);
  else if (lb == 6) warp_stages<2, true>(data, pitch, nseq, lg, tw);
  else warp_stages<1, true>(data, pitch, nseq, lg, tw);
  int s = lb;
  for (; s + 1 < lg; s += 2) smem_radix4_dit(data, pitch, nseq, lg, s, tw);
  if (s < lg) smem_radix2_stage(data, pitch, nseq, lg, s, tw, true);
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
  return L.cv.verify(stream, "usaug_local_energy");
}
