// Development aid: EXHAUSTIVE check that the 3-instruction quotient used by K1
//     r = RN(1/den);  q = RN(y r);  e = RN(y - q den) (exact, FMA);  t = RN(q + e r)
// equals the correctly rounded y / den for every integer 0 <= y < min(den, YMAX) and EVERY fp32 den in [1, 2^16)
// (Markstein's theorem says so when r is the correctly rounded reciprocal; this run proves it for the domain K1 uses).
//   nvcc -arch=sm_100a -O3 tools/div_check.cu -o /tmp/div_check && /tmp/div_check
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int YMAX = 8192;

__global__ void check(unsigned long long* bad, unsigned long long* total, uint32_t first_bits, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float den = __uint_as_float(first_bits + i);
  const float r = __frcp_rn(den);
  const int ymax = (int)fminf(ceilf(den), (float)YMAX);      // y < den
  unsigned long long nb = 0, nt = 0;
  for (int y = 0; y < ymax; ++y) {
    const float yf = (float)y;
    if (!(yf < den)) break;
    const float q = __fmul_rn(yf, r);
    const float e = __fmaf_rn(-q, den, yf);
    const float t = __fmaf_rn(e, r, q);
    nb += (__float_as_uint(t) != __float_as_uint(__fdiv_rn(yf, den)));
    ++nt;
  }
  if (nb) atomicAdd(bad, nb);
  atomicAdd(total, nt);
}

int main() {
  unsigned long long *bad, *total, h_bad = 0, h_total = 0;
  cudaMalloc(&bad, 8); cudaMalloc(&total, 8);
  cudaMemset(bad, 0, 8); cudaMemset(total, 0, 8);
  // den in [2^e, 2^(e+1)) for e = 0..15: all 2^23 mantissas each
  for (int e = 0; e < 16; ++e) {
    const uint32_t first = (uint32_t)(127 + e) << 23, count = 1u << 23;
    check<<<(count + 255) / 256, 256>>>(bad, total, first, count);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 2; }
    cudaMemcpy(&h_bad, bad, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&h_total, total, 8, cudaMemcpyDeviceToHost);
    printf("exponent %2d done: %llu quotients checked so far, %llu mismatches\n", e, h_total, h_bad);
    fflush(stdout);
  }
  // den in (0, 1): only y = 0 qualifies (0 < den); quotient must be +0
  printf("RESULT: %llu quotients, %llu mismatches\n", h_total, h_bad);
  return h_bad ? 1 : 0;
}
