// Element-wise / attention kernels of the decoder PREFILL (reference decoder.py:106-253, generate.py:266-275),
// the consumer of the audio-encoding path.  The contractions run on the tcgen05 GEMM of
// gemm_sm100.cuh (q|k|v, o_proj with in-L2 residual add, gate|up with a SwiGLU epilogue, down_proj, lm_head).
//
//   cast_rows_f32_kernel     prompt embeddings (bf16 / fp32) -> fp32 residual stream
//   rmsnorm_bf16_kernel      nn.RMSNorm (decoder.py:190-192,224): x * rsqrt(mean(x^2) + eps) * w, fp32 -> bf16
//   qknorm_rope_kernel       per-head q_norm / k_norm + nn.RoPE(traditional=False) (decoder.py:150-166) in place,
//                            and the KV-cache write (decoder.py:168-169): K (post-RoPE) and V rows of every token
//   causal_attention_kernel  softmax(scale q k^T + causal mask) v with GQA (decoder.py:171-177), varlen-packed
//                            sequences, flash-style online softmax on mma.sync.m16n8k16 (head_dim 128)
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "encoder_kernels.cuh"

namespace qasr {

template <typename T>
__global__ void __launch_bounds__(256)
cast_rows_f32_kernel(const T* __restrict__ in, float* __restrict__ out, long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 v;
    if constexpr (sizeof(T) == 2) {
      const uint2 q = *reinterpret_cast<const uint2*>(in + i);
      v.x = __uint_as_float(q.x << 16); v.y = __uint_as_float(q.x & 0xFFFF0000u);
      v.z = __uint_as_float(q.y << 16); v.w = __uint_as_float(q.y & 0xFFFF0000u);
    } else {
      v = *reinterpret_cast<const float4*>(in + i);
    }
    *reinterpret_cast<float4*>(out + i) = v;
  } else {
    for (long long j = i; j < n; ++j) out[j] = static_cast<float>(in[j]);
  }
}

// One warp per row; D = 128 * VPL.  row_idx (nullable) gathers source rows (final norm of the last tokens only).
template <int VPL>
__global__ void __launch_bounds__(256)
rmsnorm_bf16_kernel(const float* __restrict__ x, const int* __restrict__ row_idx, const float* __restrict__ w,
                    __nv_bfloat16* __restrict__ y, int rows, float eps) {
  constexpr int D = 128 * VPL;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const long long src = row_idx ? row_idx[row] : row;
  const float4* __restrict__ xr = reinterpret_cast<const float4*>(x + src * D);
  float4 v[VPL];
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = xr[lane + 32 * i];
    sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float r = rsqrtf(sq * (1.0f / D) + eps);
  const float4* __restrict__ w4 = reinterpret_cast<const float4*>(w);
  uint2* __restrict__ yr = reinterpret_cast<uint2*>(y + static_cast<long long>(row) * D);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float4 g = __ldg(w4 + lane + 32 * i);
    uint2 o;
    o.x = ptx::pack_bf16x2(v[i].x * r * g.x, v[i].y * r * g.y);
    o.y = ptx::pack_bf16x2(v[i].z * r * g.z, v[i].w * r * g.w);
    yr[lane + 32 * i] = o;
  }
}

// One warp per token; head_dim = 128.  qkv row = [Hq q heads | Hkv k heads | Hkv v heads] x 128 bf16.
// Lane l owns dims {2l, 2l+1, 64+2l, 65+2l} of every head: RoPE pairs (d, d+64) (MLX traditional=False / "rotate
// half"), angle = pos * theta^(-d/64), computed once per token and reused by all heads.
__global__ void __launch_bounds__(256)
qknorm_rope_kernel(__nv_bfloat16* __restrict__ qkv, const int* __restrict__ pos, const float* __restrict__ qw,
                   const float* __restrict__ kw, int Hq, int Hkv, float eps, float log2_theta,
                   __nv_bfloat16* __restrict__ kcache, __nv_bfloat16* __restrict__ vcache, int n_tokens) {
  const int tkn = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tkn >= n_tokens) return;
  const float p = static_cast<float>(__ldg(pos + tkn));
  float c0, s0, c1, s1;
  sincosf(p * exp2f(-static_cast<float>(2 * lane) * log2_theta * (1.0f / 64.0f)), &s0, &c0);
  sincosf(p * exp2f(-static_cast<float>(2 * lane + 1) * log2_theta * (1.0f / 64.0f)), &s1, &c1);
  const int ld = (Hq + 2 * Hkv) * 128;
  __nv_bfloat16* __restrict__ row = qkv + static_cast<long long>(tkn) * ld;
  const float2 qw_lo = *reinterpret_cast<const float2*>(qw + 2 * lane), qw_hi = *reinterpret_cast<const float2*>(qw + 64 + 2 * lane);
  const float2 kw_lo = *reinterpret_cast<const float2*>(kw + 2 * lane), kw_hi = *reinterpret_cast<const float2*>(kw + 64 + 2 * lane);
  for (int hd = 0; hd < Hq + Hkv; ++hd) {
    __nv_bfloat16* __restrict__ hp = row + hd * 128;
    const uint32_t ulo = *reinterpret_cast<const uint32_t*>(hp + 2 * lane);
    const uint32_t uhi = *reinterpret_cast<const uint32_t*>(hp + 64 + 2 * lane);
    float a0 = __uint_as_float(ulo << 16), a1 = __uint_as_float(ulo & 0xFFFF0000u);
    float b0 = __uint_as_float(uhi << 16), b1 = __uint_as_float(uhi & 0xFFFF0000u);
    float sq = (a0 * a0 + a1 * a1) + (b0 * b0 + b1 * b1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float r = rsqrtf(sq * (1.0f / 128.0f) + eps);
    const bool is_q = hd < Hq;
    const float2 wl = is_q ? qw_lo : kw_lo, wh = is_q ? qw_hi : kw_hi;
    a0 *= r * wl.x; a1 *= r * wl.y; b0 *= r * wh.x; b1 *= r * wh.y;
    const uint32_t olo = ptx::pack_bf16x2(a0 * c0 - b0 * s0, a1 * c1 - b1 * s1);
    const uint32_t ohi = ptx::pack_bf16x2(b0 * c0 + a0 * s0, b1 * c1 + a1 * s1);
    *reinterpret_cast<uint32_t*>(hp + 2 * lane) = olo;
    *reinterpret_cast<uint32_t*>(hp + 64 + 2 * lane) = ohi;
    if (!is_q && kcache) {
      __nv_bfloat16* __restrict__ kc = kcache + (static_cast<long long>(tkn) * Hkv + (hd - Hq)) * 128;
      *reinterpret_cast<uint32_t*>(kc + 2 * lane) = olo;
      *reinterpret_cast<uint32_t*>(kc + 64 + 2 * lane) = ohi;
    }
  }
  if (vcache) {
    const uint2* __restrict__ vs = reinterpret_cast<const uint2*>(row + (Hq + Hkv) * 128);
    uint2* __restrict__ vd = reinterpret_cast<uint2*>(vcache + static_cast<long long>(tkn) * Hkv * 128);
    for (int i = lane; i < Hkv * 32; i += 32) vd[i] = vs[i];
  }
}

// ------------------------------------------------------------------------------ causal attention (prefill)
struct AttnTile {
  int start;  // first token (row) of the sequence
  int len;    // tokens in the sequence
  int q0;     // first query of this 64-row tile (multiple of 64)
};
constexpr int kCaThreads = 128;  // 4 warps x 16 query rows
constexpr int kCaPitch = 136;    // bf16 per smem row (128 + 8 pad: conflict-free ldmatrix)

// grid (tiles, q heads).  K / V of kv head (head / group) are read from the qkv rows at k_off / v_off.
__global__ void __launch_bounds__(kCaThreads)
causal_attention_kernel(const __nv_bfloat16* __restrict__ qkv, int ld, int k_off, int v_off, int group,
                        const AttnTile* __restrict__ tiles, __nv_bfloat16* __restrict__ out, int ldo, float scale_log2e) {
  __shared__ __align__(16) __nv_bfloat16 sK[64 * kCaPitch];
  __shared__ __align__(16) __nv_bfloat16 sV[64 * kCaPitch];
  const AttnTile tl = tiles[blockIdx.x];
  const int head = blockIdx.y, kvh = head / group;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int len = tl.len;
  const __nv_bfloat16* __restrict__ seq = qkv + static_cast<long long>(tl.start) * ld;
  const int qi_lo = tl.q0 + warp * 16 + g, qi_hi = qi_lo + 8;

  uint32_t qa[8][4];
  {
    const __nv_bfloat16* q_lo = seq + static_cast<long long>(min(qi_lo, len - 1)) * ld + head * 128;
    const __nv_bfloat16* q_hi = seq + static_cast<long long>(min(qi_hi, len - 1)) * ld + head * 128;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      qa[kk][0] = *reinterpret_cast<const uint32_t*>(q_lo + kk * 16 + 2 * t);
      qa[kk][1] = *reinterpret_cast<const uint32_t*>(q_hi + kk * 16 + 2 * t);
      qa[kk][2] = *reinterpret_cast<const uint32_t*>(q_lo + kk * 16 + 8 + 2 * t);
      qa[kk][3] = *reinterpret_cast<const uint32_t*>(q_hi + kk * 16 + 8 + 2 * t);
    }
  }
  float o[16][4];
#pragma unroll
  for (int j = 0; j < 16; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.0f; }
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.0f, l_hi = 0.0f;

  const int n_kt = tl.q0 / 64 + 1;  // causal: keys 0 .. q0 + 63
  for (int kt = 0; kt < n_kt; ++kt) {
    __syncthreads();  // the previous tile has been consumed by every warp
    for (int i = threadIdx.x; i < 64 * 16; i += kCaThreads) {
      const int r = i >> 4, c = i & 15;
      const int key = kt * 64 + r;
      uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
      if (key < len) {
        const __nv_bfloat16* src = seq + static_cast<long long>(key) * ld + kvh * 128 + c * 8;
        kv = *reinterpret_cast<const uint4*>(src + k_off);
        vv = *reinterpret_cast<const uint4*>(src + v_off);
      }
      *reinterpret_cast<uint4*>(sK + r * kCaPitch + c * 8) = kv;
      *reinterpret_cast<uint4*>(sV + r * kCaPitch + c * 8) = vv;
    }
    __syncthreads();

    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f;
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {  // 32 dims per ldmatrix_x4
        uint32_t kb[4];
        ldmatrix_x4(kb, sK + (8 * j + (lane & 7)) * kCaPitch + qd * 32 + (lane >> 3) * 8);
        mma_bf16_16816(s[j], qa[2 * qd + 0], kb[0], kb[1]);
        mma_bf16_16816(s[j], qa[2 * qd + 1], kb[2], kb[3]);
      }
    }
    // causal + length mask (only the diagonal tile can hold masked keys), running max
    float mx_lo = m_lo, mx_hi = m_hi;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c0 = kt * 64 + 8 * j + 2 * t;
      if (kt == n_kt - 1) {
        if (c0 > qi_lo || c0 >= len) s[j][0] = -INFINITY;
        if (c0 + 1 > qi_lo || c0 + 1 >= len) s[j][1] = -INFINITY;
        if (c0 > qi_hi || c0 >= len) s[j][2] = -INFINITY;
        if (c0 + 1 > qi_hi || c0 + 1 >= len) s[j][3] = -INFINITY;
      }
      mx_lo = fmaxf(mx_lo, fmaxf(s[j][0], s[j][1]));
      mx_hi = fmaxf(mx_hi, fmaxf(s[j][2], s[j][3]));
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    // key 0 is visible to every query row, so mx is finite from the first tile on
    const float al_lo = exp2f((m_lo - mx_lo) * scale_log2e), al_hi = exp2f((m_hi - mx_hi) * scale_log2e);
    m_lo = mx_lo; m_hi = mx_hi;
    l_lo *= al_lo; l_hi *= al_hi;
#pragma unroll
    for (int j = 0; j < 16; ++j) { o[j][0] *= al_lo; o[j][1] *= al_lo; o[j][2] *= al_hi; o[j][3] *= al_hi; }
    uint32_t pa[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f((s[j][0] - mx_lo) * scale_log2e), p1 = exp2f((s[j][1] - mx_lo) * scale_log2e);
      const float p2 = exp2f((s[j][2] - mx_hi) * scale_log2e), p3 = exp2f((s[j][3] - mx_hi) * scale_log2e);
      l_lo += p0 + p1;
      l_hi += p2 + p3;
      pa[j >> 1][(j & 1) * 2 + 0] = ptx::pack_bf16x2(p0, p1);
      pa[j >> 1][(j & 1) * 2 + 1] = ptx::pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jn = 0; jn < 16; jn += 2) {
        uint32_t vb[4];
        ldmatrix_x4_trans(vb, sV + (16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8) * kCaPitch + 8 * jn + (lane >> 4) * 8);
        mma_bf16_16816(o[jn], pa[kk], vb[0], vb[1]);
        mma_bf16_16816(o[jn + 1], pa[kk], vb[2], vb[3]);
      }
    }
  }
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float inv_lo = 1.0f / l_lo, inv_hi = 1.0f / l_hi;
  __nv_bfloat16* __restrict__ ob = out + static_cast<long long>(tl.start) * ldo + head * 128;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (qi_lo < len)
      *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(qi_lo) * ldo + 8 * j + 2 * t) = ptx::pack_bf16x2(o[j][0] * inv_lo, o[j][1] * inv_lo);
    if (qi_hi < len)
      *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(qi_hi) * ldo + 8 * j + 2 * t) = ptx::pack_bf16x2(o[j][2] * inv_hi, o[j][3] * inv_hi);
  }
}

// Source-dtype -> bf16 row copy used by the weight loader: dst row = map(src row).
//   mode 0: dst_row0 + r        mode 1 (gate): 64 (r / 32) + r % 32        mode 2 (up): 64 (r / 32) + 32 + r % 32
template <typename T>
__global__ void __launch_bounds__(256)
weight_rows_to_bf16_kernel(const T* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long cols, int mode, long long dst_row0) {
  const long long r = blockIdx.x;
  long long d = dst_row0 + r;
  if (mode == 1) d = 64 * (r / 32) + r % 32;
  else if (mode == 2) d = 64 * (r / 32) + 32 + r % 32;
  const T* __restrict__ s = src + r * cols;
  __nv_bfloat16* __restrict__ o = dst + d * cols;
  for (long long i = threadIdx.x; i < cols; i += blockDim.x) o[i] = __float2bfloat16_rn(static_cast<float>(s[i]));
}

}  // namespace qasr
