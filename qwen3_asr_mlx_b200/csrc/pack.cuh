// Varlen packing of a batch of device-resident waveforms into ONE contiguous float32 buffer (the layout qasr_mel /
// qasr_encode_audio take: utterance u = [sample_offsets[u], sample_offsets[u+1])).  The reference has no counterpart (it
// handles one utterance per call, model.py:239-250); this replaces a host-side loop of B device-to-device copies (7.7 us
// each: 1024 one-second utterances cost 40x the mel kernel they feed) by one table upload and one launch.
// HBM-bound byte work: every sample is read once and written once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mel.cuh"  // upper_segment

namespace qasr {

constexpr int kPackThreads = 256;
constexpr int kPackFloatsPerCta = 16384;  // 64 KB per CTA

// block_offsets[u] = first CTA of segment u (each segment takes ceil(len / kPackFloatsPerCta) CTAs).
__global__ void __launch_bounds__(kPackThreads)
pack_segments_kernel(const float* const* __restrict__ src_ptrs, const long long* __restrict__ sample_offsets,
                     const int* __restrict__ block_offsets, int B, float* __restrict__ dst) {
  const int u = upper_segment<int>(block_offsets, B, static_cast<int>(blockIdx.x));
  const long long s0 = __ldg(sample_offsets + u);
  const long long len = __ldg(sample_offsets + u + 1) - s0;
  const long long first = static_cast<long long>(static_cast<int>(blockIdx.x) - __ldg(block_offsets + u)) * kPackFloatsPerCta;
  const int n = static_cast<int>(len - first < kPackFloatsPerCta ? len - first : kPackFloatsPerCta);
  const float* __restrict__ src = src_ptrs[u] + first;
  float* __restrict__ out = dst + s0 + first;
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const float4* __restrict__ s4 = reinterpret_cast<const float4*>(src);
    float4* __restrict__ o4 = reinterpret_cast<float4*>(out);
    const int n4 = n >> 2;
    int i = threadIdx.x;
    for (; i + 3 * kPackThreads < n4; i += 4 * kPackThreads) {  // 4 independent 16-byte loads in flight per thread
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = __ldg(s4 + i + k * kPackThreads);
#pragma unroll
      for (int k = 0; k < 4; ++k) o4[i + k * kPackThreads] = v[k];
    }
    for (; i < n4; i += kPackThreads) o4[i] = __ldg(s4 + i);
    for (int j = (n4 << 2) + threadIdx.x; j < n; j += kPackThreads) out[j] = __ldg(src + j);
  } else {
    for (int i = threadIdx.x; i < n; i += kPackThreads) out[i] = __ldg(src + i);
  }
}

}  // namespace qasr
