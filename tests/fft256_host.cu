// Host-side check of tfpnp_b200/csrc/fft256.cuh (the half-warp 16 x 16 transform of the PR kernels): the sixteen lanes are run in
// turn, the shared-memory exchange is a plain array.  Compiled and run by tests/test_host_logic.py (nvcc, no GPU needed): forward and
// inverse against a double-precision DFT.  Prints the two maximum errors; exit status 1 above 2e-5.
#include <cstdio>
#include <complex>
#include <vector>
#include <cstdlib>
#include "fft256.cuh"
using namespace tfpnp;
int main() {
  static float2 tw[256];
  for (int k = 0; k < 256; ++k) { tw[k].x = (float)cos(M_PI * k / 128.0); tw[k].y = (float)-sin(M_PI * k / 128.0); }
  std::vector<std::complex<double>> x(256), X(256);
  for (auto& e : x) e = {rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5};
  for (int inv = 0; inv < 2; ++inv) {
    for (int k = 0; k < 256; ++k) { std::complex<double> a = 0; for (int n = 0; n < 256; ++n) a += x[n] * std::polar(1.0, (inv ? 2 : -2) * M_PI * n * k / 256.0); X[k] = a; }
    float2 slots[kF256Slots];
    float2 v[16][16];
    for (int t = 0; t < 16; ++t) {
      for (int j = 0; j < 16; ++j) v[t][j] = make_float2((float)x[t + 16 * j].real(), (float)x[t + 16 * j].imag());
      if (inv) fft256_pre<true>(v[t], t, tw); else fft256_pre<false>(v[t], t, tw);
      for (int k1 = 0; k1 < 16; ++k1) slots[17 * k1 + t] = v[t][k1];
    }
    double err = 0, ref = 0;
    for (int t = 0; t < 16; ++t) {
      for (int tt = 0; tt < 16; ++tt) v[t][tt] = slots[17 * t + tt];
      if (inv) dft16<true>(v[t]); else dft16<false>(v[t]);
      for (int k2 = 0; k2 < 16; ++k2) {
        std::complex<double> g(v[t][k2].x, v[t][k2].y);
        err = std::max(err, std::abs(g - X[t + 16 * k2])); ref = std::max(ref, std::abs(X[t + 16 * k2]));
      }
    }
    printf("inv=%d max err %.3e (max |X| %.3f)\n", inv, err, ref);
    if (!(err <= 2e-5)) return 1;
  }
  return 0;
}
