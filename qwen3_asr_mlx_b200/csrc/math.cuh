// Scalar math shared by the epilogues and element-wise kernels.
#pragma once
#include <cuda_runtime.h>

namespace qasr {

// GELU(x) = x * Phi(x), erf form.  erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7),
// two MUFU ops (rcp, ex2) instead of the ~35-instruction erff() expansion; outputs of every
// call site are rounded to bf16 (rel. 2^-9), so this is far below the rounding already present.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = exp2f(-1.4426950408889634f * z * z);
  const float erf_abs = fmaf(-poly, e, 1.0f);          // erf(|x|/sqrt2)
  const float erf_signed = copysignf(erf_abs, x);
  return 0.5f * x * (1.0f + erf_signed);
}

// Exact-erf form (libdevice erff), used where the output stays fp32.
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

}  // namespace qasr
