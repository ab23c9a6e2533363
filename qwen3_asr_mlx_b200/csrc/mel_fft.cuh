// Small fixed-size DFT building blocks for the 400-point real STFT frame transform
// (reference: np.fft.rfft(frame * window, n=400), audio.py:230-233).
//
// 400 = 16 x 25.  With n = 25*n1 + n2 and k = k1 + 16*k2:
//   X[k1 + 16 k2] = sum_{n2<25} W25^{n2 k2} * ( W400^{n2 k1} * sum_{n1<16} x[25 n1 + n2] W16^{n1 k1} )
// Step A: 25 real 16-point DFTs (only k1 = 0..8 are needed: the input is real),
// Step B: twiddle by W400^{n2 k1},
// Step C: 9 complex 25-point DFTs (as 5 x 5).  Bins above 200 are obtained by conjugate symmetry.
//
// All functions are __host__ __device__ so the exact arithmetic can be unit-tested on the CPU.
#pragma once
#include <cuda_runtime.h>

namespace qasr {
namespace fft {

struct cf {
  float re, im;
};
__host__ __device__ __forceinline__ cf cadd(cf a, cf b) { return {a.re + b.re, a.im + b.im}; }
__host__ __device__ __forceinline__ cf csub(cf a, cf b) { return {a.re - b.re, a.im - b.im}; }
__host__ __device__ __forceinline__ cf cmul(cf a, cf b) {
  return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
// multiply by -i
__host__ __device__ __forceinline__ cf mul_mi(cf a) { return {a.im, -a.re}; }

// In-place forward 5-point DFT (e^{-2 pi i nk/5}).
__host__ __device__ __forceinline__ void dft5(cf& x0, cf& x1, cf& x2, cf& x3, cf& x4) {
  const float c1 = 0.30901699437494742f;   // cos(2pi/5)
  const float c2 = -0.80901699437494742f;  // cos(4pi/5)
  const float s1 = 0.95105651629515357f;   // sin(2pi/5)
  const float s2 = 0.58778525229247313f;   // sin(4pi/5)
  cf t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
  cf m1 = {x0.re + c1 * t1.re + c2 * t2.re, x0.im + c1 * t1.im + c2 * t2.im};
  cf m2 = {x0.re + c2 * t1.re + c1 * t2.re, x0.im + c2 * t1.im + c1 * t2.im};
  cf u1 = {s1 * t3.re + s2 * t4.re, s1 * t3.im + s2 * t4.im};
  cf u2 = {s2 * t3.re - s1 * t4.re, s2 * t3.im - s1 * t4.im};
  cf r0 = {x0.re + t1.re + t2.re, x0.im + t1.im + t2.im};
  cf mi1 = mul_mi(u1), mi2 = mul_mi(u2);
  x0 = r0;
  x1 = cadd(m1, mi1);
  x4 = csub(m1, mi1);
  x2 = cadd(m2, mi2);
  x3 = csub(m2, mi2);
}

// W25^{b*c} for b,c in 1..4 (row-major [b-1][c-1]), forward sign.
__host__ __device__ __forceinline__ cf w25(int b, int c) {
  // cos/sin(2*pi*j/25), j = 0..16
  const float C[17] = {1.0f, 0.96858316112863108f, 0.87630668004386358f, 0.72896862742141155f,
                       0.53582679497899666f, 0.30901699437494742f, 0.062790519529313374f,
                       -0.18738131458572463f, -0.42577929156507272f, -0.63742398974868975f,
                       -0.80901699437494742f, -0.92977648588825146f, -0.99211470131447788f,
                       -0.99211470131447788f, -0.92977648588825146f, -0.80901699437494742f,
                       -0.63742398974868975f};
  const float S[17] = {0.0f, 0.24868988716485479f, 0.48175367410171532f, 0.68454710592868873f,
                       0.84432792550201508f, 0.95105651629515357f, 0.99802672842827156f,
                       0.98228725072868872f, 0.90482705246601958f, 0.77051324277578925f,
                       0.58778525229247313f, 0.36812455268467797f, 0.12533323356430426f,
                       -0.12533323356430426f, -0.36812455268467797f, -0.58778525229247313f,
                       -0.77051324277578925f};
  const int j = b * c;  // <= 16
  return {C[j], -S[j]};
}

// In-place forward 25-point complex DFT.  Input natural order x[n]; output X[k] natural order.
__host__ __device__ __forceinline__ void dft25(cf (&x)[25]) {
  // stage 1: for each b, 5-point DFT over a of x[5a + b]  -> t[b][c] stored at x[5c + b]
#pragma unroll
  for (int b = 0; b < 5; ++b) dft5(x[b], x[5 + b], x[10 + b], x[15 + b], x[20 + b]);
  // twiddle t[b][c] *= W25^{bc}
#pragma unroll
  for (int b = 1; b < 5; ++b)
#pragma unroll
    for (int c = 1; c < 5; ++c) x[5 * c + b] = cmul(x[5 * c + b], w25(b, c));
  // stage 2: for each c, 5-point DFT over b of t[b][c] -> X[c + 5d] stored at x[5c + d]
#pragma unroll
  for (int c = 0; c < 5; ++c) dft5(x[5 * c], x[5 * c + 1], x[5 * c + 2], x[5 * c + 3], x[5 * c + 4]);
  // reorder: y[c + 5d] = x[5c + d]
  cf y[25];
#pragma unroll
  for (int c = 0; c < 5; ++c)
#pragma unroll
    for (int d = 0; d < 5; ++d) y[c + 5 * d] = x[5 * c + d];
#pragma unroll
  for (int i = 0; i < 25; ++i) x[i] = y[i];
}

// Forward 16-point DFT of a real sequence; returns bins 0..8.
__host__ __device__ __forceinline__ void rdft16(const float (&x)[16], cf (&y)[9]) {
  const float r = 0.70710678118654752f;
  // 8-point complex FFT of z[n] = x[2n] + i x[2n+1] (radix-2 decimation in time)
  cf z[8];
#pragma unroll
  for (int n = 0; n < 8; ++n) z[n] = {x[2 * n], x[2 * n + 1]};
  // 2-point DFTs on (0,4), (2,6), (1,5), (3,7)
  cf a0 = cadd(z[0], z[4]), a1 = csub(z[0], z[4]);
  cf a2 = cadd(z[2], z[6]), a3 = csub(z[2], z[6]);
  cf b0 = cadd(z[1], z[5]), b1 = csub(z[1], z[5]);
  cf b2 = cadd(z[3], z[7]), b3 = csub(z[3], z[7]);
  // 4-point DFTs: even = {z0,z2,z4,z6}, odd = {z1,z3,z5,z7}
  cf e0 = cadd(a0, a2), e2 = csub(a0, a2);
  cf e1 = cadd(a1, mul_mi(a3)), e3 = csub(a1, mul_mi(a3));
  cf o0 = cadd(b0, b2), o2 = csub(b0, b2);
  cf o1 = cadd(b1, mul_mi(b3)), o3 = csub(b1, mul_mi(b3));
  // combine with W8^k: W8^1 = (r, -r), W8^2 = -i, W8^3 = (-r, -r)
  cf t1 = {r * (o1.re + o1.im), r * (o1.im - o1.re)};
  cf t2 = mul_mi(o2);
  cf t3 = {r * (o3.im - o3.re), -r * (o3.re + o3.im)};
  cf Z[9];
  Z[0] = cadd(e0, o0); Z[4] = csub(e0, o0);
  Z[1] = cadd(e1, t1); Z[5] = csub(e1, t1);
  Z[2] = cadd(e2, t2); Z[6] = csub(e2, t2);
  Z[3] = cadd(e3, t3); Z[7] = csub(e3, t3);
  Z[8] = Z[0];
  // real-input split: Y[k] = (Z[k] + conj Z[8-k])/2 - i W16^k (Z[k] - conj Z[8-k])/2
  const float C16[9] = {1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f, 0.0f,
                        -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f, -1.0f};
  const float S16[9] = {0.0f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f, 1.0f,
                        0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f, 0.0f};
#pragma unroll
  for (int k = 0; k <= 8; ++k) {
    cf zk = Z[k];
    cf zc = {Z[8 - k].re, -Z[8 - k].im};
    cf ev = {0.5f * (zk.re + zc.re), 0.5f * (zk.im + zc.im)};
    cf od = {0.5f * (zk.re - zc.re), 0.5f * (zk.im - zc.im)};
    cf w = {C16[k], -S16[k]};
    cf t = mul_mi(cmul(w, od));
    y[k] = cadd(ev, t);
  }
}

}  // namespace fft
}  // namespace qasr
