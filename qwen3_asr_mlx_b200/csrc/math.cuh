// Scalar math shared by the epilogues and element-wise kernels.
#pragma once
#include <cuda_runtime.h>

namespace qasr {

// GELU(x) = x * Phi(x), exact (erf) form, evaluated as 0.5 x (1 + tanh(u(x))) with
// u(x) = x (a + b x^2 + c x^4) a minimax fit of atanh(erf(x / sqrt 2)): |fit error| <= 2.6e-5
// absolute over all x (the textbook 0.044715 tanh-GELU is 4.7e-4 off).  tanh.approx.f32 adds
// <= 2^-11 relative error on the tanh, i.e. <= 2.5e-4 |x| on the result -- 8x below the bf16
// rounding (2^-9 relative) applied to every output of the call sites.  7 ALU ops + 1 MUFU.
__device__ __forceinline__ float gelu_fast(float x) {
  const float x2 = fminf(x * x, 36.0f);  // tanh is saturated beyond |x| = 6; keeps u(x) monotone
  float p = fmaf(-3.51516781e-04f, x2, 3.70056460e-02f);
  p = fmaf(p, x2, 7.97507884e-01f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * p));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// The same GELU evaluated from h = x / 2: every producer of a GELU input on this path (conv stem, fc1, proj1) has its
// weights and bias pre-scaled by 0.5 at load time (exact: a power of two), which removes the 0.5 x multiply:
// u = x p(x^2) = h (2a + 8b h^2 + 32c h^4), GELU = h (1 + tanh u).  6 ALU ops + 1 MUFU.
__device__ __forceinline__ float gelu_from_half(float h) {
  const float h2 = fminf(h * h, 9.0f);  // x^2 <= 36
  float p = fmaf(32.0f * -3.51516781e-04f, h2, 8.0f * 3.70056460e-02f);
  p = fmaf(p, h2, 2.0f * 7.97507884e-01f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h * p));
  return fmaf(h, t, h);
}

// SiLU(x) = x * sigmoid(x) = 0.5 x (1 + tanh(x / 2))  (exact identity): 3 ALU ops + 1 MUFU.
__device__ __forceinline__ float silu_fast(float x) {
  const float hx = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(hx));
  return fmaf(hx, t, hx);
}

// Exact-erf form (libdevice erff), used where the output stays fp32.
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

}  // namespace qasr
