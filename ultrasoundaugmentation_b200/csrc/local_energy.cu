// Monogenic-signal local energy (oracle/local_energy.py):
//   F = FFT2(I);  even = Re IFFT2(G F);  q = IFFT2(G (v' - i u')/rho F);  E = sqrt(even^2 + |q|^2)
// with G = sum of log-Gabor band-pass filters.  Hand-written batched 2-D FFT for power-of-two sizes on the block-level
// mixed-radix Stockham transform of fft_stockham.cuh (register radix-8 passes, swizzled shared memory, natural order
// in and out):
//
//   le_twiddles   the twiddle tables of the two transform lengths, once per call (2 blocks)
//   le_filter     the filter multipliers (G, G u'/rho, G v'/rho) per frequency, once per call: the same for every image
//   le_rows_fwd   per group of image rows: FFT along x, two real rows per complex transform (first pass reads the
//                 image); the two spectra are separated and their bins 0 .. W/2 stored (the others are conjugate mirrors)
//   le_cols       per tile of columns and its mirror tile: one forward FFT along y (first pass reads the columns), the
//                 filter, then the inverse FFT along y of the tile's odd (Riesz) pair and even part and of the mirror
//                 tile's odd pair, whose last pass writes the planes.  The even part is real: its columns beyond W/2 are
//                 conjugate mirrors and are neither transformed back nor stored.
//   le_rows_inv   per group of image rows: inverse FFT along x of the odd rows and of PAIRS of even rows (two real rows
//                 as real and imaginary part of one transform), energy, per-block maxima
//   le_scale      per image: 1 / max E (0 for an image without band-pass content)
//
// Sizes that are not powers of two are mirror-extended (bottom / right, edge included) to the next power of two
// (Hp x Wp) on the way into le_rows_fwd; rows and columns beyond H x W are never written back.
// HBM traffic: 4 + 4 | 4 + 8 + 4 | 8 + 4 + 4 = 40 B/px (the stage model in bench.py still counts the 56 B/px of the
// version that stored the whole even plane); everything else stays in shared memory.
#include "fft_stockham.cuh"
#include "kernels.cuh"

namespace usaug {

constexpr int kLeThreads = 256;       // fft::block_fft needs blockDim.x * 8 to be a multiple of the transform length
constexpr int kLeMaxScales = 4;
constexpr int kLeMaxN = 2048;
// column sequences are H + max(1, H / 128) apart: with the sequence-fastest lane order of le_cols' global passes the
// shared-memory accesses of a warp then fall into different banks (tests/manual/stockham_proto.py, seqfast_conflicts)
static inline int le_col_pitch(int H) { return H + ((H >> 7) > 1 ? (H >> 7) : 1); }

struct LeFilter {
  int scales;
  float ln_lambda[kLeMaxScales];
  float inv_den;     // 1 / (2 ln^2 sigma_f)
};

// signed DFT frequency index of bin k (numpy.fft.fftfreq * n): the Nyquist bin is -n/2
__device__ __forceinline__ int signed_bin(int k, int n) { return (k < n / 2) ? k : k - n; }

// sqrt.approx: one MUFU instruction, relative error 2^-23 (the energy is compared at 2e-5 of its maximum)
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

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
// img is [B][H0][W0]; spec is [B][H][W] with H, W the padded (power-of-two) sizes.  The first pass of the transform
// reads the image rows, the last one writes the spectrum rows: shared memory only carries the exchanges in between.
// Two real rows share one complex transform: sequence q carries z = row 2q + i * row 2q+1.  With Z = FFT(z) the two
// (Hermitian) row spectra are A(k) = (Z(k) + conj Z(W - k)) / 2 and B(k) = (Z(k) - conj Z(W - k)) / 2i.
struct RowsFwdIn {
  const float* __restrict__ img;
  size_t row0, nrows;                                       // nrows = B * H and the rows per block are even
  int H0, W0, H, W, lgH;
  bool exact;                                               // no mirror extension
  template <int R> __device__ __forceinline__ void load(int q, int x0, int stride, float2* v) const {
    const size_t r = row0 + 2 * (size_t)q;
#pragma unroll
    for (int k = 0; k < R; ++k) v[k] = make_float2(0.f, 0.f);
    if (r >= nrows) return;
    if (exact) {
      const float* p = img + r * W + x0;
#pragma unroll
      for (int k = 0; k < R; ++k) v[k] = make_float2(p[k * stride], p[W + k * stride]);
    } else {
      const int b = (int)(r >> lgH), yp = (int)(r & (size_t)(H - 1));                           // yp even, yp + 1 < H
      const int ya = (yp < H0) ? yp : 2 * H0 - 1 - yp, yb = (yp + 1 < H0) ? yp + 1 : 2 * H0 - 2 - yp;   // mirror extension
      const float* pa = img + ((size_t)b * H0 + ya) * W0;
      const float* pb = img + ((size_t)b * H0 + yb) * W0;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int xp = x0 + k * stride, x = (xp < W0) ? xp : 2 * W0 - 1 - xp;
        v[k] = make_float2(pa[x], pb[x]);
      }
    }
  }
};

__global__ void __launch_bounds__(kLeThreads) le_rows_fwd(const float* __restrict__ img, float2* __restrict__ spec,
                                                          int H0, int W0, int H, int W, int lgH, fft::Plan pw,
                                                          const float2* __restrict__ twg, int rb, size_t nrows) {
  extern __shared__ __align__(16) unsigned char le_smem[];
  float2* tw = reinterpret_cast<float2*>(le_smem);          // [W]  (pw.tw_total <= W entries used)
  float2* data = tw + W;                                    // [rb / 2][W], element x of a sequence at slot swz(x)
  const size_t row0 = (size_t)blockIdx.x * rb;
  const int lgW = pw.lg, np = rb >> 1;
  load_twiddles(tw, twg, pw.tw_total);
  __syncthreads();
  const RowsFwdIn in{img, row0, nrows, H0, W0, H, W, lgH, (H0 == H) && (W0 == W)};
  fft::block_fft<false, false, false>(data, W, np, pw, tw, in, fft::SmemIO{});
  // separate the two spectra; only the bins 0 .. W / 2 are stored (the rows are real: the others are conjugate mirrors)
  for (int i = threadIdx.x; i < np << (lgW - 1); i += blockDim.x) {
    const int q = i >> (lgW - 1), k = i & (W / 2 - 1);
    const size_t r = row0 + 2 * (size_t)q;
    if (r >= nrows) break;
    const float2* z = data + (q << lgW);
    const float2 zk = z[fft::swz(k)], zm = z[fft::swz((W - k) & (W - 1))];
    float2* pa = spec + r * W + k;
    pa[0] = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
    pa[W] = make_float2(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x));
    if (k == 0) {                                           // ... and the Nyquist bin, its own mirror
      const float2 zn = z[fft::swz(W / 2)];
      pa[W / 2] = make_float2(zn.x, 0.f);
      pa[W + W / 2] = make_float2(zn.y, 0.f);
    }
  }
}

// ---- columns: block = tc columns of one image AND their mirror columns ---------------------------------
// The image is real: F(ky, W - kx) = conj F(-ky, kx).  A block therefore transforms the tc columns x0 .. x0 + tc - 1
// (x0 < W / 2) forward once, and takes the spectrum of the mirror columns W - x0 - q from it by index reversal.
// Back go three groups of tc sequences: the odd (Riesz) pair and the even part of the tile, and the odd pair of the
// mirror tile (the even part is real, so le_rows_inv rebuilds its columns beyond W / 2 as conjugate mirrors; the odd
// pair has no such symmetry).  Five column transforms per (column, mirror column) instead of six.  The Nyquist column
// W / 2 is its own mirror: one extra block per image handles it alone.
// The first pass of the forward transform reads the spectrum columns, the last pass of the inverse transform writes
// the planes, both with the sequence-fastest lane order: a warp touches tc-wide contiguous pieces of consecutive rows.
struct ColsIn {
  const float2* __restrict__ src;                           // spec + image base + x0
  int W, tc;
  template <int R> __device__ __forceinline__ void load(int q, int y0, int stride, float2* v) const {
    if (q >= tc) {
#pragma unroll
      for (int k = 0; k < R; ++k) v[k] = make_float2(0.f, 0.f);
      return;
    }
    const float2* p = src + (size_t)y0 * W + q;
    const size_t step = (size_t)stride * W;
#pragma unroll
    for (int k = 0; k < R; ++k, p += step) v[k] = *p;
  }
};
// even: G F;   odd: (v' - i u') / rho * G * F;   f = (G, G u'/rho, G v'/rho), zero at DC
__device__ __forceinline__ float2 le_even(float4 f, float2 F) { return make_float2(f.x * F.x, f.x * F.y); }
__device__ __forceinline__ float2 le_odd(float4 f, float2 F) {
  return make_float2(fmaf(f.z, F.x, f.y * F.y), fmaf(f.z, F.y, -f.y * F.x));
}
// Nyquist column: the filter multiply as the last pass of the forward transform, on the way back into shared memory
struct ColsFilterOut {
  static constexpr bool kWritesSmem = true;
  float2* A;
  float2* Bv;
  const float4* __restrict__ fcol;                          // filt + x0 * H
  int H, pitch, tc;
  template <int R> __device__ __forceinline__ void store(int q, int y0, int stride, const float2* v) const {
    if (q >= tc) return;
    const float4* f = fcol + (size_t)q * H + y0;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int slot = q * pitch + fft::swz(y0 + k * stride);
      Bv[slot] = le_even(f[k * stride], v[k]);
      A[slot] = le_odd(f[k * stride], v[k]);
    }
  }
};
// sequences [0, tc): odd pair of column x0 + q; [tc, 2 tc): even part of column x0 + q - tc; [2 tc, 3 tc) (if
// `mirror`): odd pair of the mirror column W - (x0 + q - 2 tc), which does not exist for column 0.
struct ColsOut {
  static constexpr bool kWritesSmem = false;
  float2* __restrict__ odd;                                 // + image base
  float2* __restrict__ even;
  int W, tc, H0, x0;
  bool mirror;
  template <int R> __device__ __forceinline__ void store(int q, int y0, int stride, const float2* v) const {
    float2* p;
    if (q < tc) p = odd + x0 + q;
    else if (q < 2 * tc) p = even + x0 + (q - tc);
    else if (mirror && q < 3 * tc && x0 + (q - 2 * tc) != 0) p = odd + W - (x0 + (q - 2 * tc));
    else return;
    p += (size_t)y0 * W;
    const size_t step = (size_t)stride * W;
#pragma unroll
    for (int k = 0; k < R; ++k, p += step)
      if (y0 + k * stride < H0) *p = v[k];                  // rows of the mirror extension are not needed again
  }
};

__global__ void __launch_bounds__(kLeThreads, 4) le_cols(const float2* __restrict__ spec, float2* __restrict__ odd,
                                                      float2* __restrict__ even, const float4* __restrict__ filt,
                                                      int H0, int H, int W, fft::Plan ph,
                                                      const float2* __restrict__ twg, int tc, int pitch) {
  extern __shared__ __align__(16) unsigned char le_smem[];
  const int lgH = ph.lg;
  float2* tw = reinterpret_cast<float2*>(le_smem);          // [H]
  float2* A = tw + H;                                       // [tc][pitch]  spectrum, then the odd pair of the tile
  const int b = blockIdx.y;
  const size_t ibase = (size_t)b * H * W;
  load_twiddles(tw, twg, ph.tw_total);
  __syncthreads();
  if ((int)blockIdx.x * tc >= W / 2) {                      // the Nyquist column alone: [odd][even]
    const int x0 = W / 2;
    fft::block_fft<false, true, false>(A, pitch, 1, ph, tw, ColsIn{spec + ibase + x0, W, 1},
                                       ColsFilterOut{A, A + pitch, filt + (size_t)x0 * H, H, pitch, 1});
    fft::block_fft<true, false, true>(A, pitch, 2, ph, tw, fft::SmemIO{}, ColsOut{odd + ibase, even + ibase, W, 1, H0, x0, false});
    return;
  }
  float2* Bv = A + tc * pitch;                              // [tc][pitch]  the even part of the tile
  float2* Am = Bv + tc * pitch;                             // [tc][pitch]  the odd pair of the mirror tile
  const int x0 = blockIdx.x * tc;
  fft::block_fft<false, true, false>(A, pitch, tc, ph, tw, ColsIn{spec + ibase + x0, W, tc}, fft::SmemIO{});
  // filter: a thread takes the bins ky and H - ky of a column together, F'(ky) = conj F(H - ky) being the mirror
  // column's spectrum (the bins 0 and H / 2 are their own partners)
  for (int i = threadIdx.x; i < tc << (lgH - 1); i += blockDim.x) {
    const int c = i >> (lgH - 1), k1 = i & (H / 2 - 1);
    const float4* f = filt + (size_t)(x0 + c) * H;          // column x0 + c; its mirror column has u' -> -u'
    float2* a = A + c * pitch;
    float2* e = Bv + c * pitch;
    float2* m = Am + c * pitch;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      // h = 0: the pair (k1, H - k1), or bin 0 alone; h = 1: bin H / 2 alone, done by the thread of bin 0
      if (h == 1 && k1 != 0) break;
      const int ka = h ? H / 2 : k1, kb = (H - ka) & (H - 1);
      const int sa = fft::swz(ka), sb = fft::swz(kb);
      const float2 Fa = a[sa], Fb = a[sb];
      const float4 fa = f[ka];
      e[sa] = le_even(fa, Fa);
      a[sa] = le_odd(fa, Fa);
      m[sa] = le_odd(make_float4(fa.x, -fa.y, fa.z, 0.f), make_float2(Fb.x, -Fb.y));
      if (kb != ka) {
        const float4 fb = make_float4(fa.x, fa.y, -fa.z, 0.f);    // bin H - ka: v' -> -v'
        e[sb] = le_even(fb, Fb);
        a[sb] = le_odd(fb, Fb);
        m[sb] = le_odd(make_float4(fb.x, -fb.y, fb.z, 0.f), make_float2(Fa.x, -Fa.y));
      }
    }
  }
  __syncthreads();
  fft::block_fft<true, false, true>(A, pitch, 3 * tc, ph, tw, fft::SmemIO{}, ColsOut{odd + ibase, even + ibase, W, tc, H0, x0, true});
}

// ---- rows, inverse + energy: block = rb rows -----------------------------------------------------------
// sequence q < rb: the odd (Riesz) row q.  The even part of every row is real, so two of its rows share one
// transform: sequence rb + k carries even row 2k + i * even row 2k+1, and the inverse transform returns the two real
// rows as its real and imaginary parts.
struct RowsInvIn {
  const float2* __restrict__ odd;                           // + base of the block's first row
  const float2* __restrict__ even;
  int W, rb, rows_left;                                     // rows_left = H0 - row0: rows below it were never written
  template <int R> __device__ __forceinline__ void load(int q, int x0, int stride, float2* v) const {
#pragma unroll
    for (int k = 0; k < R; ++k) v[k] = make_float2(0.f, 0.f);
    if (q < rb) {
      const float2* p = odd + (size_t)q * W + x0;
#pragma unroll
      for (int k = 0; k < R; ++k) v[k] = p[k * stride];
    } else if (2 * (q - rb) < rb) {
      // even rows: stored for x <= W / 2 only, e(W - x) = conj e(x)
      const float2* p = even + (size_t)(2 * (q - rb)) * W;
      const bool two = 2 * (q - rb) + 1 < min(rb, rows_left);
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int x = x0 + k * stride;
        const bool mir = x > W / 2;
        const int xs = mir ? W - x : x;
        float2 a = p[xs];
        float2 c = two ? p[W + xs] : make_float2(0.f, 0.f);
        if (mir) { a.y = -a.y; c.y = -c.y; }
        v[k] = make_float2(a.x - c.y, a.y + c.x);
      }
    }
  }
};

__global__ void __launch_bounds__(kLeThreads) le_rows_inv(const float2* __restrict__ odd, const float2* __restrict__ even,
                                                          float* __restrict__ energy, float* __restrict__ part,
                                                          int H0, int W0, int H, int W, fft::Plan pw,
                                                          const float2* __restrict__ twg, int rb, float norm) {
  extern __shared__ __align__(16) unsigned char le_smem[];
  __shared__ float s_max[kLeThreads / 32];
  float2* tw = reinterpret_cast<float2*>(le_smem);          // [W]
  float2* data = tw + W;                                    // [rb + (rb + 1) / 2][W]: odd rows, then pairs of even rows
  const int b = blockIdx.y, lgW = pw.lg;
  const int row0 = blockIdx.x * rb;                         // H % rb == 0; only blocks with row0 < H0 are launched
  const size_t base = ((size_t)b * H + row0) * W;
  load_twiddles(tw, twg, pw.tw_total);
  __syncthreads();
  fft::block_fft<true, false, false>(data, W, rb + (rb + 1) / 2, pw, tw, RowsInvIn{odd + base, even + base, W, rb, H0 - row0}, fft::SmemIO{});
  float m = 0.f;
  // even row length: x and x + 1 as one 8-byte store
  const bool pairs = (W0 & 1) == 0 && (reinterpret_cast<size_t>(energy) & 7) == 0;
  // one item = two pixels (x, x + 1) of the two rows that share an even sequence
  const int npair = (rb + 1) >> 1;
  for (int i = threadIdx.x; i < npair << (lgW - 1); i += blockDim.x) {
    const int k = i >> (lgW - 1), x = (i & (W / 2 - 1)) << 1;
    const int y = row0 + 2 * k;
    if (y >= H0 || x >= W0) continue;                       // mirror extension: not part of the result
    const int sx = fft::swz(x);
    const float2* po = data + ((2 * k) << lgW) + sx;        // odd row 2k; row 2k + 1 follows W slots later
    const float2* pe = data + ((rb + k) << lgW) + sx;       // (even row 2k) + i (even row 2k + 1)
    const float2 e0 = pe[0], e1 = (pe - sx)[sx ^ 1];
    float* dst = energy + ((size_t)b * H0 + y) * W0 + x;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && (2 * k + 1 >= rb || y + 1 >= H0)) break;
      const float2 q0 = po[h << lgW], q1 = (po - sx)[(h << lgW) + (sx ^ 1)];
      const float a0 = h ? e0.y : e0.x, a1 = h ? e1.y : e1.x;
      const float en0 = fast_sqrt(fmaf(a0, a0, fmaf(q0.x, q0.x, q0.y * q0.y))) * norm;
      const float en1 = fast_sqrt(fmaf(a1, a1, fmaf(q1.x, q1.x, q1.y * q1.y))) * norm;
      float* d = dst + (size_t)h * W0;
      if (pairs) {
        *reinterpret_cast<float2*>(d) = make_float2(en0, en1);
        m = fmaxf(m, fmaxf(en0, en1));
      } else {
        d[0] = en0; m = fmaxf(m, en0);
        if (x + 1 < W0) { d[1] = en1; m = fmaxf(m, en1); }
      }
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
  // rows per block: 32 KB of sequences forward, 48 KB inverse (odd rows + pairs of even rows), at least 1, at most 16
  L.rb_fwd = 8192 / W; if (L.rb_fwd < 2) L.rb_fwd = 2; if (L.rb_fwd > 16) L.rb_fwd = 16;     // even: rows are transformed in pairs
  L.rb_inv = 4096 / W; if (L.rb_inv < 1) L.rb_inv = 1; if (L.rb_inv > 16) L.rb_inv = 16;
  while (H % L.rb_inv) L.rb_inv >>= 1;
  // columns per block: one transform step (2048 points) forward, two for the inverse pair: 32 KB of sequences, so that
  // six blocks share an SM (measured: 8 columns of 512 rows are 6 % slower, 8 of 1024 rows 25 %)
  L.tc = 2048 / H; if (L.tc < 2) L.tc = 2; if (L.tc > 16) L.tc = 16; if (L.tc > W / 2) L.tc = W / 2;
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
  const size_t sm_fwd = (size_t)(W + fft::padded_seqs(L.rb_fwd / 2, lgW, kLeThreads) * W) * sizeof(float2);
  const size_t sm_col = (size_t)(H + fft::padded_seqs(3 * L.tc, lgH, kLeThreads) * le_col_pitch(H)) * sizeof(float2);
  const size_t sm_inv = (size_t)(W + fft::padded_seqs(L.rb_inv + (L.rb_inv + 1) / 2, lgW, kLeThreads) * W) * sizeof(float2);
  if (sm_fwd > 32 * 1024) USAUG_CUDA(cudaFuncSetAttribute(le_rows_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_fwd));
  if (sm_col > 32 * 1024) USAUG_CUDA(cudaFuncSetAttribute(le_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_col));
  if (sm_inv > 32 * 1024) USAUG_CUDA(cudaFuncSetAttribute(le_rows_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_inv));

  le_twiddles<<<2, kLeThreads, 0, stream>>>(pw, ph, L.tww, L.twh);
  USAUG_CHECK_LAUNCH("le_twiddles");
  le_filter<<<(unsigned)(((size_t)H * W + kLeThreads - 1) / kLeThreads), kLeThreads, 0, stream>>>(L.filt, H, W, lgH, flt);
  USAUG_CHECK_LAUNCH("le_filter");
  le_rows_fwd<<<(unsigned)((nrows + L.rb_fwd - 1) / L.rb_fwd), kLeThreads, sm_fwd, stream>>>(img, L.spec, H0, W0, H, W, lgH, pw,
                                                                                            L.tww, L.rb_fwd, nrows);
  USAUG_CHECK_LAUNCH("le_rows_fwd");
  le_cols<<<dim3(W / (2 * L.tc) + 1, B), kLeThreads, sm_col, stream>>>(L.spec, L.odd, L.even, L.filt, H0, H, W, ph, L.twh, L.tc,
                                                             le_col_pitch(H));
  USAUG_CHECK_LAUNCH("le_cols");
  le_rows_inv<<<dim3(L.nblk_inv, B), kLeThreads, sm_inv, stream>>>(L.odd, L.even, energy, L.part, H0, W0, H, W, pw, L.tww,
                                                                  L.rb_inv, (float)(1.0 / ((double)H * (double)W)));
  USAUG_CHECK_LAUNCH("le_rows_inv");
  le_scale<<<(B + 127) / 128, 128, 0, stream>>>(L.part, L.nblk_inv, scale, B);
  USAUG_CHECK_LAUNCH("le_scale");
  return L.cv.verify(stream, "usaug_local_energy");
}
