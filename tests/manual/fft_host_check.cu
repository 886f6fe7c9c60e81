// Host check of csrc/fft_stockham.cuh: the index and butterfly functions the kernels use, run on the CPU against a
// direct fp64 DFT.   nvcc -O2 -o /tmp/fft_host_check tests/manual/fft_host_check.cu && /tmp/fft_host_check
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../ultrasoundaugmentation_b200/csrc/fft_stockham.cuh"
using namespace usaug::fft;

template <int R, bool INV> static void pass(std::vector<float2>& d, int lg, int ns_lg, const float2* tw) {
  const int rlg = (R == 8) ? 3 : (R == 4) ? 2 : 1, nb = 1 << (lg - rlg);
  std::vector<float2> v((size_t)nb * R);
  std::vector<int> dst(nb);
  int rbs[3], wbs[3];
  swz_basis<R>(lg - rlg, rbs);
  swz_basis<R>(ns_lg, wbs);
  for (int j = 0; j < nb; ++j) dst[j] = butterfly_read<R, INV>(d.data(), j, ns_lg, tw, rbs, &v[(size_t)j * R]);
  for (int j = 0; j < nb; ++j) butterfly_write<R>(d.data(), dst[j], wbs, &v[(size_t)j * R]);
}

template <bool INV> static double run(int lg) {
  const int n = 1 << lg;
  Plan pl = make_plan(lg);
  std::vector<float2> tw(pl.tw_total + 1);
  for (int p = 0; p < pl.npass; ++p) {
    if (!pl.ns_lg[p]) continue;
    const int R = 1 << pl.rlg[p], ns_lg = pl.ns_lg[p];
    for (int i = 0; i < ((R - 1) << ns_lg); ++i) {
      const int r = (i >> ns_lg) + 1, m = i & ((1 << ns_lg) - 1);
      const double a = 2.0 * M_PI * m * r / (double)(R << ns_lg);
      tw[pl.tw_off[p] + i] = make_float2((float)cos(a), (float)-sin(a));
    }
  }
  std::vector<std::complex<double>> x(n);
  std::vector<float2> d(n);
  for (int i = 0; i < n; ++i) {
    x[i] = {rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5};
    d[swz(i)] = make_float2((float)x[i].real(), (float)x[i].imag());
  }
  for (int p = 0; p < pl.npass; ++p) {
    const float2* t = tw.data() + pl.tw_off[p];
    if (pl.rlg[p] == 3) pass<8, INV>(d, lg, pl.ns_lg[p], t);
    else if (pl.rlg[p] == 2) pass<4, INV>(d, lg, pl.ns_lg[p], t);
    else pass<2, INV>(d, lg, pl.ns_lg[p], t);
  }
  double err = 0, mx = 0;
  for (int k = 0; k < n; ++k) {
    std::complex<double> s = 0;
    for (int i = 0; i < n; ++i) s += x[i] * std::polar(1.0, (INV ? 2.0 : -2.0) * M_PI * (double)((long)i * k % n) / n);
    err = fmax(err, std::abs(s - std::complex<double>(d[swz(k)].x, d[swz(k)].y)));
    mx = fmax(mx, std::abs(s));
  }
  return err / mx;
}

int main() {
  int bad = 0;
  for (int lg = 3; lg <= 11; ++lg) {
    Plan pl = make_plan(lg);
    const double e0 = run<false>(lg), e1 = run<true>(lg);
    printf("n=%4d passes=%d tw=%4d  fwd rel err %.2e  inv rel err %.2e\n", 1 << lg, pl.npass, pl.tw_total, e0, e1);
    bad += !(e0 < 2e-6) + !(e1 < 2e-6) + (pl.tw_total > (1 << lg));
  }
  printf(bad ? "FAILED\n" : "ok\n");
  return bad;
}
