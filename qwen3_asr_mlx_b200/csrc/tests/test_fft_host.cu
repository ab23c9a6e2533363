// CPU check of the 16x25 decomposition used by the mel kernel against a naive double DFT.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../mel_fft.cuh"
using namespace qasr::fft;
int main() {
  const int N = 400;
  double worst = 0;
  for (int trial = 0; trial < 20; ++trial) {
    std::vector<float> x(N);
    for (int i = 0; i < N; ++i) x[i] = (float)((rand() / (double)RAND_MAX) * 2 - 1);
    // reference
    std::vector<double> re(201), im(201);
    double scale = 0;
    for (int k = 0; k <= 200; ++k) {
      double sr = 0, si = 0;
      for (int n = 0; n < N; ++n) { double a = -2 * M_PI * (double)((long long)n * k % N) / N; sr += x[n] * cos(a); si += x[n] * sin(a); }
      re[k] = sr; im[k] = si; scale = fmax(scale, hypot(sr, si));
    }
    // step A+B
    static cf Y[9][25];
    for (int n2 = 0; n2 < 25; ++n2) {
      float v[16]; for (int n1 = 0; n1 < 16; ++n1) v[n1] = x[25 * n1 + n2];
      cf y[9]; rdft16(v, y);
      for (int k1 = 0; k1 <= 8; ++k1) {
        double a = -2 * M_PI * (n2 * k1) / 400.0; cf w = {(float)cos(a), (float)sin(a)};
        Y[k1][n2] = cmul(y[k1], w);
      }
    }
    std::vector<int> seen(201, 0);
    for (int k1 = 0; k1 <= 8; ++k1) {
      cf t[25]; for (int i = 0; i < 25; ++i) t[i] = Y[k1][i];
      dft25(t);
      for (int k2 = 0; k2 < 25; ++k2) {
        int k = k1 + 16 * k2; cf v = t[k2];
        int bin; double vr = v.re, vi = v.im;
        if (k <= 200) bin = k; else { if (k1 == 0 || k1 == 8) continue; bin = 400 - k; vi = -vi; }
        if (k > 200 && bin > 200) continue;
        seen[bin]++;
        worst = fmax(worst, hypot(vr - re[bin], vi - im[bin]) / scale);
      }
    }
    for (int k = 0; k <= 200; ++k) if (seen[k] != 1) { printf("bin %d seen %d times\n", k, seen[k]); return 1; }
  }
  printf("fft400 max rel err vs double DFT: %.3e\n", worst);
  return worst < 2e-6 ? 0 : 1;
}
