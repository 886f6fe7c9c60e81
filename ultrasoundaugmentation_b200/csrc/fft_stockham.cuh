// Batched power-of-two complex FFT for one thread block, sequences resident in shared memory.
//
// Mixed-radix Stockham autosort (natural order in, natural order out), radices 8 / 4 / 2: a pass with radix R and
// Ns = product of the radices already done lets butterfly j (< n/R) read x[j + r n/R], r < R, multiply by
// exp(-+2 pi i (j mod Ns) r / (Ns R)), run an R-point DFT in registers and write y[(j div Ns) Ns R + (j mod Ns) + r Ns].
// A 512-point transform is three register radix-8 passes with two shared-memory exchanges between them (the
// radix-2 warp-shuffle version it replaces spent 84 % of its instructions in five shuffle stages).
//
// Shared-memory layout of a sequence: element i lives at slot swz(i) = i ^ ((i >> 3) & 15).  With it every read and
// every write of every pass (strides n/R, 1, 8, 64, ...) and the contiguous copies in and out touch 16 different
// 8-byte bank pairs per half warp: tests/manual/stockham_proto.py models the passes and finds 1.00x the minimal
// wavefront count for n = 64 .. 2048 (no padding: 2x to 8x).
//
// Everything that computes an index or a butterfly is __host__ __device__: tests/manual/fft_host_check.cu runs the
// same functions on the CPU against a direct DFT.
#pragma once
#include <cuda_runtime.h>

#ifndef USAUG_HD
#define USAUG_HD __host__ __device__ __forceinline__
#endif

namespace usaug {
namespace fft {

constexpr int kMaxPasses = 4;          // n <= 2048 = 8 * 8 * 8 * 4
constexpr int kPointsPerThread = 8;    // a thread holds 8 points per step of a pass: 1 radix-8, 2 radix-4 or 4 radix-2 butterflies

struct Plan {
  int lg;                    // n = 1 << lg
  int npass;
  int rlg[kMaxPasses];       // log2 of the radix of pass p (3, 2 or 1)
  int ns_lg[kMaxPasses];     // log2 of Ns (product of the earlier radices)
  int tw_off[kMaxPasses];    // offset of the pass' twiddle table: entry [(r - 1) * Ns + m] = exp(-2 pi i m r / (Ns R))
  int tw_total;              // entries of all tables (<= n)
};

inline Plan make_plan(int lg) {
  Plan p{};
  p.lg = lg;
  int left = lg, ns = 0, off = 0;
  while (left > 0) {
    // radix 8 while at least three stages are left and the remainder is not a single stage after a 4 (8.. then 4 or 2)
    const int r = (left >= 3) ? 3 : left;
    p.rlg[p.npass] = r; p.ns_lg[p.npass] = ns; p.tw_off[p.npass] = off;
    if (ns > 0) off += ((1 << r) - 1) << ns;
    ns += r; left -= r; ++p.npass;
  }
  p.tw_total = off;
  return p;
}

USAUG_HD int swz(int i) { return i ^ ((i >> 3) & 15); }

// Complex arithmetic.  On the device it uses the packed fp32x2 instructions of sm_100 (FADD2 / FMUL2 / FFMA2: one
// issue slot for both components; the transforms are issue-bound): a complex add is one instruction, a complex
// multiply three.  The host versions (tests/manual/fft_host_check.cu) are the plain formulas.
#ifdef __CUDA_ARCH__
__device__ __forceinline__ unsigned long long pk(float2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ float2 upk(unsigned long long r) {
  float2 a;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
  return a;
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)));
  return upk(r);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)));
  return upk(r);
}
// a * b = a.x * (b.x, b.y) + a.y * (-b.y, b.x)
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  unsigned long long t, r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(pk(make_float2(a.x, a.x))), "l"(pk(b)));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk(make_float2(a.y, a.y))), "l"(pk(make_float2(-b.y, b.x))), "l"(t));
  return upk(r);
}
__device__ __forceinline__ float2 cscale(float2 a, float s) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(make_float2(s, s))));
  return upk(r);
}
#else
inline float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
inline float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
inline float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
inline float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
#endif
// a * (-i) for the forward transform, a * (+i) for the inverse
template <bool INV> USAUG_HD float2 rot(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

template <bool INV> USAUG_HD void dft4(float2& b0, float2& b1, float2& b2, float2& b3) {
  const float2 c0 = cadd(b0, b2), c2 = csub(b0, b2), c1 = cadd(b1, b3), c3 = rot<INV>(csub(b1, b3));
  b0 = cadd(c0, c1); b1 = cadd(c2, c3); b2 = csub(c0, c1); b3 = csub(c2, c3);
}

// R-point DFT of v[0..R), result in natural order in v
template <int R, bool INV> USAUG_HD void dft(float2* v) {
  if (R == 2) {
    const float2 a = v[0];
    v[0] = cadd(a, v[1]); v[1] = csub(a, v[1]);
  } else if (R == 4) {
    dft4<INV>(v[0], v[1], v[2], v[3]);
  } else {
    constexpr float h = 0.70710678118654752440f;
    float2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
    float2 a4 = csub(v[0], v[4]), d5 = csub(v[1], v[5]), a6 = rot<INV>(csub(v[2], v[6])), d7 = csub(v[3], v[7]);
    // w8 = (1 -+ i) / sqrt 2 and w8^3 = (-1 -+ i) / sqrt 2
    // w8 = (1 -+ i) / sqrt 2 and w8^3 = (-1 -+ i) / sqrt 2, with rot(d) = -+ i d
    float2 a5 = cscale(cadd(d5, rot<INV>(d5)), h);
    float2 a7 = cscale(csub(rot<INV>(d7), d7), h);
    dft4<INV>(a0, a1, a2, a3);      // X0 X2 X4 X6
    dft4<INV>(a4, a5, a6, a7);      // X1 X3 X5 X7
    v[0] = a0; v[1] = a4; v[2] = a1; v[3] = a5; v[4] = a2; v[5] = a6; v[6] = a3; v[7] = a7;
  }
}

// idx[r] = s ^ swz(r << sh), r < R, from the basis bs[k] = swz(1 << (sh + k)): swz is linear over GF(2), so
// swz(x + (r << sh)) = swz(x) ^ swz(r << sh) whenever x has zeros where r << sh has its bits -- one three-input XOR
// per element instead of a shift, a mask and two XORs.
template <int R> USAUG_HD void xor_fan(int s, const int* bs, int* idx) {
  idx[0] = s; idx[1] = s ^ bs[0];
  if (R >= 4) { idx[2] = s ^ bs[1]; idx[3] = s ^ bs[0] ^ bs[1]; }
  if (R >= 8) {
    const int t = s ^ bs[2];
    idx[4] = t; idx[5] = t ^ bs[0]; idx[6] = t ^ bs[1]; idx[7] = t ^ bs[0] ^ bs[1];
  }
}
template <int R> USAUG_HD void swz_basis(int sh, int* bs) {
#pragma unroll
  for (int k = 0; (1 << k) < R; ++k) bs[k] = swz(1 << (sh + k));
}

// Butterfly j (< n / R) of a pass on the R elements v[r] = x[j + r n/R]: twiddles and transforms them in place and
// returns the element index d the r-th result goes to as d + (r << ns_lg).
template <int R, bool INV> USAUG_HD int butterfly(int j, int ns_lg, const float2* tw, float2* v) {
  constexpr int RLG = (R == 8) ? 3 : (R == 4) ? 2 : 1;
  const int m = j & ((1 << ns_lg) - 1);
  if (ns_lg > 0) {
#pragma unroll
    for (int r = 1; r < R; ++r) {
      float2 w = tw[((r - 1) << ns_lg) + m];
      if (INV) w.y = -w.y;
      v[r] = cmul(v[r], w);
    }
  }
  dft<R, INV>(v);
  return ((j >> ns_lg) << (ns_lg + RLG)) | m;
}

// The same from a sequence in (swizzled) shared memory; rbs = swz_basis of lg - log2 R.
template <int R, bool INV>
USAUG_HD int butterfly_read(const float2* seq, int j, int ns_lg, const float2* tw, const int* rbs, float2* v) {
  int idx[R];
  xor_fan<R>(swz(j), rbs, idx);            // j < n/R shares no bit with r n/R
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = seq[idx[r]];
  return butterfly<R, INV>(j, ns_lg, tw, v);
}

// wbs = swz_basis of ns_lg
template <int R> USAUG_HD void butterfly_write(float2* seq, int dst, const int* wbs, const float2* v) {
  int idx[R];
  xor_fan<R>(swz(dst), wbs, idx);          // dst has zeros where r << ns_lg has its bits
#pragma unroll
  for (int r = 0; r < R; ++r) seq[idx[r]] = v[r];
}

// Sequences a block transform touches: nseq rounded up to whole steps of `threads` * 8 points.  block_fft runs every
// thread in every step without bounds tests; the caller provides room for this many sequences (what lies beyond nseq
// is transformed too and ignored).
inline int padded_seqs(int nseq, int lg, int threads) {
  const int per_step = (threads * kPointsPerThread) >> lg;
  return (nseq + per_step - 1) / per_step * per_step;
}

#ifdef __CUDACC__
// I/O policy of the first / last pass of a block transform: SmemIO = the swizzled sequences in shared memory.  Any
// other type is a functor on NATURAL element indices, called once per butterfly --
//   In:   template <int R> void load(int q, int i0, int stride, float2* v) const    v[r] = element i0 + r stride of sequence q
//   Out:  template <int R> void store(int q, int i0, int stride, const float2* v) const
//         static constexpr bool kWritesSmem    true if store() writes the sequences the pass reads (needs the barrier)
// -- so that a transform reads its input from and writes its result to global memory directly, without staging
// copies (and without their barriers), or applies a pointwise operation on the way out.
struct SmemIO { static constexpr bool kWritesSmem = true; };
template <class T> struct is_smem { static constexpr bool value = false; };
template <> struct is_smem<SmemIO> { static constexpr bool value = true; };

constexpr int kBlockLg = 8;            // block_fft is written for 256 threads

// One pass over nseq sequences (sequence q at data + q * pitch).  The block walks the butterflies in steps of
// 256 * 8 points; that is a multiple of n, so a step covers whole sequences: the shared-memory reads of a step are
// separated from its shared-memory writes by one barrier, and steps touch different sequences.  SEQ_FAST: consecutive
// lanes take the same butterfly of consecutive sequences (a functor then sees consecutive q per warp: contiguous
// global accesses for sequences that are the columns of an image).
template <int R, bool INV, bool SEQ_FAST, class In, class Out>
__device__ __forceinline__ void block_pass(float2* data, int pitch, int nseq, int lg, int ns_lg, const float2* tw,
                                           const In& in, const Out& out) {
  constexpr int RLG = (R == 8) ? 3 : (R == 4) ? 2 : 1;
  constexpr int U = kPointsPerThread / R;
  constexpr bool GIN = !is_smem<In>::value, GOUT = !is_smem<Out>::value;
  const int nb_lg = lg - RLG, total = nseq << nb_lg;
  const int spp_lg = kBlockLg + 3 - lg;                      // sequences per step (log2)
  int rbs[RLG], wbs[RLG];
  swz_basis<R>(nb_lg, rbs);
  swz_basis<R>(ns_lg, wbs);
#pragma unroll
  for (int k = 0; k < RLG; ++k) {          // opaque: the compiler otherwise folds the basis back into per-element swz chains
    asm("" : "+r"(rbs[k]));
    asm("" : "+r"(wbs[k]));
  }
  int step = 0;
  for (int base = 0; base < total; base += (1 << kBlockLg) * U, ++step) {
    float2 v[U][R];
    int q[U], dst[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = (u << kBlockLg) + (int)threadIdx.x;     // butterfly inside the step
      int j;
      if (SEQ_FAST) { q[u] = (step << spp_lg) + (t & ((1 << spp_lg) - 1)); j = t >> spp_lg; }
      else { q[u] = (base + t) >> nb_lg; j = (base + t) & ((1 << nb_lg) - 1); }
      if constexpr (GIN) {
        in.template load<R>(q[u], j, 1 << nb_lg, v[u]);
      } else {
        int idx[R];
        xor_fan<R>(swz(j), rbs, idx);      // j < n/R shares no bit with r n/R
        const float2* seq = data + q[u] * pitch;
#pragma unroll
        for (int r = 0; r < R; ++r) v[u][r] = seq[idx[r]];
      }
      dst[u] = butterfly<R, INV>(j, ns_lg, tw, v[u]);
    }
    if (!GIN && Out::kWritesSmem) __syncthreads();
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if constexpr (GOUT) {
        out.template store<R>(q[u], dst[u], 1 << ns_lg, v[u]);
      } else {
        butterfly_write<R>(data + q[u] * pitch, dst[u], wbs, v[u]);
      }
    }
  }
  if (Out::kWritesSmem) __syncthreads();
}

// Forward (INV = false) or unscaled inverse transform of nseq sequences of n = 1 << pl.lg points each.
// tw: the plan's twiddle tables (shared memory).  256 threads; room for padded_seqs(nseq, lg, 256) sequences.
// in / out: SmemIO (in place in shared memory) or functors (see above) used by the first / last pass;
// SEQ_FAST_IN / _OUT: lane order of that pass (see block_pass).  A single-pass transform uses SEQ_FAST_IN.
template <bool INV, bool SEQ_FAST_IN, bool SEQ_FAST_OUT, class In, class Out>
__device__ __forceinline__ void block_fft(float2* data, int pitch, int nseq, const Plan& pl, const float2* tw,
                                          const In& in, const Out& out) {
  const SmemIO sm{};
#pragma unroll
  for (int p = 0; p < kMaxPasses; ++p) {            // unrolled: the plan stays in the kernel's parameter space
    if (p >= pl.npass) break;
    const float2* t = tw + pl.tw_off[p];
    const int lg = pl.lg, ns_lg = pl.ns_lg[p];
    const bool last = (p == pl.npass - 1);
    // the input functor belongs to pass 0, the output functor to the last pass
#define USAUG_FFT_PASS(R)                                                                                    \
    if (p == 0 && last) block_pass<R, INV, SEQ_FAST_IN>(data, pitch, nseq, lg, ns_lg, t, in, out);            \
    else if (p == 0) block_pass<R, INV, SEQ_FAST_IN>(data, pitch, nseq, lg, ns_lg, t, in, sm);                \
    else if (last) block_pass<R, INV, SEQ_FAST_OUT>(data, pitch, nseq, lg, ns_lg, t, sm, out);                \
    else block_pass<R, INV, false>(data, pitch, nseq, lg, ns_lg, t, sm, sm);
    if (pl.rlg[p] == 3) { USAUG_FFT_PASS(8) }
    else if (pl.rlg[p] == 2) { USAUG_FFT_PASS(4) }
    else { USAUG_FFT_PASS(2) }
#undef USAUG_FFT_PASS
  }
}
template <bool INV>
__device__ __forceinline__ void block_fft(float2* data, int pitch, int nseq, const Plan& pl, const float2* tw) {
  block_fft<INV, false, false>(data, pitch, nseq, pl, tw, SmemIO{}, SmemIO{});
}

// twiddle tables of a plan (global or shared memory), filled by `nthreads` cooperating threads
__device__ __forceinline__ void fill_twiddles(const Plan& pl, float2* __restrict__ tw, int tid, int nthreads) {
  for (int p = 0; p < pl.npass; ++p) {
    const int ns_lg = pl.ns_lg[p];
    if (ns_lg == 0) continue;
    const int R = 1 << pl.rlg[p], cnt = (R - 1) << ns_lg;
    for (int i = tid; i < cnt; i += nthreads) {
      const int r = (i >> ns_lg) + 1, m = i & ((1 << ns_lg) - 1);
      float s, c;
      sincospif(2.0f * (float)(m * r) / (float)(R << ns_lg), &s, &c);     // m r < Ns R <= 2048: exact in fp32
      tw[pl.tw_off[p] + i] = make_float2(c, -s);
    }
  }
}
#endif  // __CUDACC__

}  // namespace fft
}  // namespace usaug
