// Development aid: times the single-warp coarse solve (confmap_fused.cuh: coarse_task_nb) in isolation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Iinclude -Iultrasoundaugmentation_b200/csrc \
//        tools/coarse_bench.cu ultrasoundaugmentation_b200/csrc/api.cu -o tools/bin/coarse_bench
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "confmap_setup.cuh"
#include "confmap_fused.cuh"

using namespace usaug;

template <int NB, int RING>
__global__ void __launch_bounds__(32) bench_kernel(PcgArgs<float> a, long long* cyc, double* out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x, b = blockIdx.x;
  const long long t0 = clock64();
  const double r = coarse_task_nb<NB, RING>(a, smem, b, lane);
  const long long t1 = clock64();
  if (lane == 0) { cyc[b] = t1 - t0; out[b] = r; }
}

template <int NB, int RING>
static void run(int nby, int B) {
  CoarseGeom g{};
  g.by = 2; g.by_shift = 1; g.bx_shift = 5; g.nbx = NB; g.nby = nby; g.nb = NB; g.nc = nby * NB; g.rec = NB * NB + 4 * NB;
  std::vector<float> hW((size_t)B * nby * g.rec), hr((size_t)B * g.nc);
  for (size_t i = 0; i < hW.size(); ++i) {
    const size_t k = i % g.rec;
    const int r = (int)(k / NB), c = (int)(k % NB);
    hW[i] = (k < (size_t)NB * NB) ? ((r == c) ? 0.5f : 0.01f / (1 + abs(r - c))) : -0.1f;
  }
  for (auto& v : hr) v = (float)rand() / RAND_MAX - 0.5f;
  float *dW, *drc, *dzc; long long* dcyc; double* dout;
  cudaMalloc(&dW, hW.size() * 4); cudaMalloc(&drc, hr.size() * 4); cudaMalloc(&dzc, hr.size() * 4);
  cudaMalloc(&dcyc, B * 8); cudaMalloc(&dout, B * 8);
  cudaMemcpy(dW, hW.data(), hW.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(drc, hr.data(), hr.size() * 4, cudaMemcpyHostToDevice);
  PcgArgs<float> a{};
  a.cg = g; a.Wt = dW; a.rc = drc; a.zc = dzc;
  cudaFuncSetAttribute(bench_kernel<NB, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, RING);
  for (int rep = 0; rep < 3; ++rep) bench_kernel<NB, RING><<<B, 32, RING>>>(a, dcyc, dout);
  cudaDeviceSynchronize();
  std::vector<long long> hc(B); std::vector<double> ho(B);
  cudaMemcpy(hc.data(), dcyc, B * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(ho.data(), dout, B * 8, cudaMemcpyDeviceToHost);
  double mean = 0; for (auto c : hc) mean += (double)c; mean /= B;
  printf("NB=%d ring=%d B nby=%d B=%d: %.0f cycles per solve, %.1f cycles per block row and sweep, dot[0]=%.6g  (%s)\n", NB, RING, nby, B, mean,
         mean / (2.0 * nby), ho[0], cudaGetErrorString(cudaGetLastError()));
  cudaFree(dW); cudaFree(drc); cudaFree(dzc); cudaFree(dcyc); cudaFree(dout);
}

int main() {
  run<16, 25600>(255, 148); run<16, 51200>(255, 148); run<16, 51200>(128, 148); run<16, 25600>(255, 1184);
  run<8, 25600>(127, 148); run<32, 51200>(255, 148); run<32, 25600>(255, 148);
  return 0;
}
