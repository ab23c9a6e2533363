// Long-audio splitter (reference model.py:454-513, `_find_split_points`): per-frame RMS energy of the
// waveform and, for every multiple of chunk_samples, the lowest-energy frame within +-search_samples.
//
// The reference evaluates `np.sqrt(np.mean(frame ** 2))` per 480-sample frame in float32.  The cut is an
// argmin over those energies, so the energies are reproduced BIT-EXACTLY: squares with a separate
// round-to-nearest multiply (no FMA contraction) and the sum in numpy's pairwise order (numpy
// `pairwise_sum`: blocks of <= 128 elements are summed with 8 strided accumulators combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail; longer inputs split at n/2 rounded down to
// a multiple of 8), then `sum / n` and a correctly rounded square root.
//
// HBM-bound byte work: 4 B per sample read once, 4 B per frame written (config 4: 77 MB in, 160 KB out).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qasr {

// numpy pairwise sum of squares of a[0..n) (generic, one thread; used for frame sizes other than 480)
__device__ inline float np_pairwise_sumsq(const float* __restrict__ a, int n) {
  if (n < 8) {
    float r = 0.0f;  // numpy starts from the first element; 0 + x == x exactly (squares are never -0)
    for (int i = 0; i < n; ++i) r = __fadd_rn(r, __fmul_rn(a[i], a[i]));
    return r;
  }
  if (n <= 128) {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fmul_rn(a[j], a[j]);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], __fmul_rn(a[i + j], a[i + j]));
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, __fmul_rn(a[i], a[i]));
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  const float left = np_pairwise_sumsq(a, n2);
  return __fadd_rn(left, np_pairwise_sumsq(a + n2, n - n2));
}

__global__ void __launch_bounds__(128)
frame_rms_generic_kernel(const float* __restrict__ audio, long long n_frames, int frame, float* __restrict__ energy) {
  const long long f = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (f >= n_frames) return;
  const float s = np_pairwise_sumsq(audio + f * frame, frame);
  energy[f] = __fsqrt_rn(__fdiv_rn(s, static_cast<float>(frame)));
}

// frame == 480: one warp per frame.  numpy's order for n = 480 is ((B0 + B1) + (B2 + B3)) over four 120-element
// blocks, each block = 8 strided accumulators of 15 terms.  Lane l = 8 * block + j owns accumulator j of its block;
// the xor-butterfly (1, 2, 4, then 8, 16) adds exactly the pairs numpy adds (fp32 addition is commutative).
constexpr int kRmsWarpsPerCta = 8;
__global__ void __launch_bounds__(kRmsWarpsPerCta * 32)
frame_rms480_kernel(const float* __restrict__ audio, long long n_frames, float* __restrict__ energy) {
  const int lane = threadIdx.x & 31;
  const long long f = static_cast<long long>(blockIdx.x) * kRmsWarpsPerCta + (threadIdx.x >> 5);
  if (f >= n_frames) return;
  const float* __restrict__ a = audio + f * 480 + (lane >> 3) * 120 + (lane & 7);
  float v[15];
#pragma unroll
  for (int i = 0; i < 15; ++i) v[i] = __ldg(a + 8 * i);  // 15 independent loads in flight per lane
  float r = __fmul_rn(v[0], v[0]);
#pragma unroll
  for (int i = 1; i < 15; ++i) r = __fadd_rn(r, __fmul_rn(v[i], v[i]));
#pragma unroll
  for (int m = 1; m < 32; m <<= 1) r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, m));
  if (lane == 0) energy[f] = __fsqrt_rn(__fdiv_rn(r, 480.0f));
}

// np.argmin semantics: first index of the minimum; a NaN counts as the minimum (first NaN wins).
__device__ __forceinline__ bool argmin_better(float v, long long i, float bv, long long bi) {
  const bool vn = v != v, bn = bv != bv;
  if (vn != bn) return vn;
  if (vn) return i < bi;
  return v < bv || (v == bv && i < bi);
}

// One warp per chunk boundary (model.py:497-511).
__global__ void __launch_bounds__(32)
split_argmin_kernel(const float* __restrict__ energy, long long n_frames, long long total, long long chunk_samples,
                    long long search_samples, int frame, long long* __restrict__ points) {
  const long long boundary = (static_cast<long long>(blockIdx.x) + 1) * chunk_samples;
  if (boundary >= total) return;
  const long long centre = boundary / frame, radius = search_samples / frame;
  const long long lo = centre - radius > 0 ? centre - radius : 0;
  const long long hi = centre + radius < n_frames - 1 ? centre + radius : n_frames - 1;
  const int lane = threadIdx.x;
  if (lo >= hi) {
    if (lane == 0) points[blockIdx.x] = boundary;
    return;
  }
  float bv = __ldg(energy + lo);
  long long bi = lo;
  for (long long i = lo + lane; i <= hi; i += 32) {
    const float v = __ldg(energy + i);
    if (argmin_better(v, i, bv, bi)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, m);
    const long long oi = __shfl_xor_sync(0xffffffffu, bi, m);
    if (argmin_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) points[blockIdx.x] = bi * frame;
}

}  // namespace qasr
