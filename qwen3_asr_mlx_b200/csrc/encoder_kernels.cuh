// Non-GEMM kernels of the audio encoder path:
//   conv1_gelu_kernel      Conv2d(1->C, k3, s2, p1) + bias + GELU on 100-frame mel chunks
//                          (reference encoder.py:258-273), output written straight into the
//                          parity-plane layout the conv2 implicit GEMM consumes
//   layernorm_bf16_kernel  LayerNorm(eps=1e-5, affine) fp32 -> bf16 (encoder.py:112,117,319)
//   window_attention_kernel block-diagonal (<=104-token windows) multi-head attention
//                          (encoder.py:78-85,297-311) over varlen-packed tokens; the O(n^2)
//                          additive mask of the reference is never materialised.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "math.cuh"
#include "mel.cuh"  // ordered_to_float: the per-utterance log-mel maximum is kept as an ordered uint (atomicMax)
#include "ptx.cuh"

namespace qasr {

// ------------------------------------------------------------------------------ conv1
struct ChunkDesc {
  long long mel_base;  // float offset of this utterance's (128, T) block in the packed mel buffer
  int T;               // frames in the utterance
  int frame0;          // first frame of this chunk
  int utt;             // utterance index within the call (selects the per-utterance log-mel maximum on the fused path)
  int pad_;
};

// Fused waveform -> embeddings path: the mel buffer holds the RAW log10 mel (pass 1 of mel.cuh) and the clamp / rescale of
// audio.py:275-276, max(x, utt_max - 8) then (x + 4) / 4, is applied here while conv1 stages its input -- the utterance
// maximum is final once pass 1 has ended -- so the normalise pass and its read + write of the whole mel never happen.
// Same expression as mel_normalize_kernel (bit-identical); a NaN maximum poisons the utterance, like there.
__device__ __forceinline__ float mel_clamp_rescale(float v, float thr) {
  return (thr != thr) ? thr : (fmaxf(v, thr) + 4.0f) * 0.25f;
}

constexpr int kConv1Threads = 240;    // 60 channel groups (8 ch) x 4 pixel slots (C = 480)
constexpr int kConv1RowsPerCta = 8;   // output rows per CTA  (64 / 8 = 8 CTAs per chunk)

// planes layout: [4][rows_total = G*33][26][C] bf16, plane = 2*(h&1) + (w&1), pixel (h,w) of
// the 64x50 conv1 output stored at row b*33 + h/2 + 1, column w/2 + 1 (row/column 0 = zero pad).
template <int C>
__global__ void __launch_bounds__(kConv1Threads)
conv1_gelu_kernel(const float* __restrict__ mel, const ChunkDesc* __restrict__ chunks, int chunk0,
                  const float* __restrict__ w /*[C][9]*/, const float* __restrict__ bias /*[C]*/,
                  __nv_bfloat16* __restrict__ planes, long long plane_stride, const unsigned* __restrict__ utt_max) {
  static_assert(C % 8 == 0 && (C / 8) * 4 == kConv1Threads, "thread mapping assumes C == 480");
  constexpr int IW = 104;  // smem row pitch
  __shared__ float in[2 * kConv1RowsPerCta + 1][IW];

  const int b = blockIdx.x / (64 / kConv1RowsPerCta);      // chunk within the group
  const int part = blockIdx.x % (64 / kConv1RowsPerCta);
  const int oh0 = part * kConv1RowsPerCta;
  const ChunkDesc cd = chunks[chunk0 + b];
  const float* __restrict__ src = mel + cd.mel_base;
  const int valid_w = min(100, cd.T - cd.frame0);
  const bool raw_mel = utt_max != nullptr;
  const float thr = raw_mel ? ordered_to_float(__ldg(utt_max + cd.utt)) - 8.0f : 0.0f;

  // stage input rows h = 2*oh0-1 .. 2*oh0+15, columns w = -1 .. 100 (zero outside the chunk / utterance)
  for (int i = threadIdx.x; i < (2 * kConv1RowsPerCta + 1) * 102; i += kConv1Threads) {
    const int r = i / 102, c = i - r * 102;
    const int h = 2 * oh0 - 1 + r, wv = c - 1;
    float v = 0.0f;
    if (h >= 0 && h < 128 && wv >= 0 && wv < valid_w) {
      v = __ldg(src + static_cast<long long>(h) * cd.T + cd.frame0 + wv);
      if (raw_mel) v = mel_clamp_rescale(v, thr);
    }
    in[r][c] = v;
  }

  const int cg = threadIdx.x % (C / 8);
  const int slot = threadIdx.x / (C / 8);
  float wr[8][9], br[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    br[c] = __ldg(bias + cg * 8 + c);
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[c][t] = __ldg(w + (cg * 8 + c) * 9 + t);
  }
  __syncthreads();

  for (int p = slot; p < kConv1RowsPerCta * 50; p += 4) {
    const int ohl = p / 50, ow = p - ohl * 50;
    float patch[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) patch[kh * 3 + kw] = in[2 * ohl + kh][2 * ow + kw];
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float a = br[c];
#pragma unroll
      for (int t = 0; t < 9; ++t) a = fmaf(wr[c][t], patch[t], a);
      acc[c] = gelu_from_half(a);  // w and bias are pre-scaled by 0.5
    }
    const int h = oh0 + ohl;
    const int plane = 2 * (h & 1) + (ow & 1);
    const long long pix = (static_cast<long long>(b) * 33 + (h >> 1) + 1) * 26 + (ow >> 1) + 1;
    uint4 q;
    q.x = ptx::pack_bf16x2(acc[0], acc[1]);
    q.y = ptx::pack_bf16x2(acc[2], acc[3]);
    q.z = ptx::pack_bf16x2(acc[4], acc[5]);
    q.w = ptx::pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(planes + plane * plane_stride + pix * C + cg * 8) = q;
  }
}

// ------------------------------------------------------------------------------ LayerNorm
// One warp per row; VPL float4 per lane (D = 128 * VPL); two-pass statistics in registers.
template <int VPL>
__global__ void __launch_bounds__(256)
layernorm_bf16_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                      __nv_bfloat16* __restrict__ y, int rows, float eps, int reverse = 0) {
  constexpr int D = 128 * VPL;
  ptx::grid_dep_wait();  // programmatic dependent launch (no-op otherwise): x is valid from here
  ptx::grid_dep_launch();
  const int row_fwd = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row_fwd >= rows) return;
  const int row = reverse ? rows - 1 - row_fwd : row_fwd;  // serpentine: start on the rows the producer wrote last
  const float4* __restrict__ xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * D);
  float4 v[VPL];
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = xr[lane + 32 * i];
    sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * (1.0f / D);
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    sq += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq * (1.0f / D) + eps);
  const float4* __restrict__ g4 = reinterpret_cast<const float4*>(gamma);
  const float4* __restrict__ b4 = reinterpret_cast<const float4*>(beta);
  uint2* __restrict__ yr = reinterpret_cast<uint2*>(y + static_cast<long long>(row) * D);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float4 g = __ldg(g4 + lane + 32 * i), bb = __ldg(b4 + lane + 32 * i);
    uint2 o;
    o.x = ptx::pack_bf16x2((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y);
    o.y = ptx::pack_bf16x2((v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w);
    yr[lane + 32 * i] = o;
  }
}

// ------------------------------------------------------------------------------ prompt assembly
// prepare_inputs (reference generate.py:20-81): row t of the decoder input is the text-embedding row
// of input_ids[t], except at <|audio_pad|> positions where it is the next audio embedding (cast to
// the embedding dtype).  src[t] >= 0: audio row index; src[t] < 0: embedding-table row -(src[t]+1).
template <typename TTable, typename TAudio>
__global__ void __launch_bounds__(128)
gather_prompt_rows_kernel(const int* __restrict__ src, const TTable* __restrict__ table, const TAudio* __restrict__ audio,
                          TTable* __restrict__ out, int hidden) {
  const long long t = blockIdx.x;
  const int s = __ldg(src + t);
  TTable* __restrict__ o = out + t * hidden;
  if (s >= 0) {
    const TAudio* __restrict__ a = audio + static_cast<long long>(s) * hidden;
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) o[i] = static_cast<TTable>(static_cast<float>(a[i]));
  } else {
    const TTable* __restrict__ r = table + static_cast<long long>(-(s + 1)) * hidden;
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) o[i] = r[i];
  }
}

// ------------------------------------------------------------------------------ attention
// One CTA per (window, head).  Window = up to 104 consecutive packed tokens of one utterance.
// qkv: [n_tok, 3*D] bf16 (q | k | v), head h occupies columns h*64 .. h*64+63 of each third.
// Each of 7 warps owns 16 query rows; S = Q K^T and O = P V run on mma.sync.m16n8k16 bf16
// with fp32 accumulation; the softmax is single-pass in registers (a window fits entirely).
constexpr int kAttnMaxWin = 112;   // 104 rounded up to a multiple of 16
constexpr int kAttnThreads = 224;  // 7 warps
constexpr int kAttnPitch = 72;     // bf16 per smem row (64 + 8 pad: conflict-free ldmatrix)

struct WindowDesc {
  int start;  // first token (row of qkv / out)
  int len;    // tokens in the window (1..104)
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(ptx::smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(ptx::smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------------------ conv1 on tensor cores
// Same contract as conv1_gelu_kernel, but the 9-tap contraction runs on mma.sync.m16n8k16
// (K = 9 taps zero-padded to 16).  The fp32 mel patch is split into bf16 hi + lo parts and fed
// through two MMAs, so the input keeps ~16 mantissa bits; weights are bf16 like every other
// layer.  What remains on the CUDA cores is bias + GELU + pack + store.
// Warp w owns channels [96w, 96w+96) = 3 groups of 32 = 12 n-tiles; within a group, tile j
// column n maps to channel 8*(n/2) + 2*j + (n&1), so each thread ends up holding 8 consecutive
// channels of a pixel (one 16-byte store).
constexpr int kConv1TcWarps = 5;
constexpr int kConv1TcThreads = kConv1TcWarps * 32;
#ifndef QASR_CONV1_MIN_CTAS
#define QASR_CONV1_MIN_CTAS 4  // 96 registers, 20 warps per SM: 1.87 ms vs 2.11 ms at 3 (the kernel is latency-bound)
#endif
constexpr int kConv1TcMinCtas = QASR_CONV1_MIN_CTAS;

template <int C>
__global__ void __launch_bounds__(kConv1TcThreads, kConv1TcMinCtas)
conv1_gelu_tc_kernel(const float* __restrict__ mel, const ChunkDesc* __restrict__ chunks, int chunk0,
                     const __nv_bfloat16* __restrict__ w /*[C][9] bf16*/, const float* __restrict__ bias /*[C]*/,
                     __nv_bfloat16* __restrict__ planes, long long plane_stride, const unsigned* __restrict__ utt_max) {
  static_assert(C == 96 * kConv1TcWarps, "warp/channel mapping assumes C == 480");
  constexpr int IW = 104;
  __shared__ float in[2 * kConv1RowsPerCta + 1][IW];

  ptx::grid_dep_wait();  // programmatic dependent launch (no-op otherwise): mel, maxima and chunk table are valid from here
  ptx::grid_dep_launch();
  const int b = blockIdx.x / (64 / kConv1RowsPerCta);
  const int part = blockIdx.x % (64 / kConv1RowsPerCta);
  const int oh0 = part * kConv1RowsPerCta;
  const ChunkDesc cd = chunks[chunk0 + b];
  const float* __restrict__ src = mel + cd.mel_base;
  const int valid_w = min(100, cd.T - cd.frame0);
  const bool raw_mel = utt_max != nullptr;
  const float thr = raw_mel ? ordered_to_float(__ldg(utt_max + cd.utt)) - 8.0f : 0.0f;
  // Input patch (17 mel rows x 102 frames incl. the zero border): a warp takes rows warp, warp + 5, ... and a lane the columns
  // lane + 32 k, so all (<= 16) loads of a thread are independent and in flight together (ncu before: a loop with one load
  // per iteration put 18 % of the kernel's stall samples on the first use of that load) and no index is divided.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    constexpr int kRows = 2 * kConv1RowsPerCta + 1, kRowsPerWarp = (kRows + kConv1TcWarps - 1) / kConv1TcWarps;
    float pv[kRowsPerWarp][4];
#pragma unroll
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
      const int r = warp + rr * kConv1TcWarps;
      const int h = 2 * oh0 - 1 + r;
      const bool row_ok = r < kRows && h >= 0 && h < 128;
      const float* __restrict__ rp = src + static_cast<long long>(h) * cd.T + cd.frame0 - 1;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = lane + 32 * k;
        pv[rr][k] = (row_ok && c >= 1 && c - 1 < valid_w) ? __ldg(rp + c) : 0.0f;  // c - 1 < valid_w <= 100 implies c < 102
      }
    }
#pragma unroll
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
      const int r = warp + rr * kConv1TcWarps;
      const int h = 2 * oh0 - 1 + r;
      const bool row_ok = r < kRows && h >= 0 && h < 128;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = lane + 32 * k;
        if (r < kRows && c < 102) {
          float v = pv[rr][k];
          if (raw_mel && row_ok && c >= 1 && c - 1 < valid_w) v = mel_clamp_rescale(v, thr);
          in[r][c] = v;
        }
      }
    }
  }

  const int g = lane >> 2, t = lane & 3;
  const int cwarp = warp * 96;
  // B fragments (weights) stay in registers for the whole CTA.  The bias rides in the contraction: K rows 9 and 10
  // (zero padding otherwise) hold bf16 hi and lo parts of the bias, the A operand holds 1.0 there.
  uint32_t bw0[12], bw1[12];
#pragma unroll
  for (int grp = 0; grp < 3; ++grp) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ch = cwarp + grp * 32 + 8 * (g >> 1) + 2 * j + (g & 1);  // column n = g of tile j
      const __nv_bfloat16* wr = w + ch * 9;
      const uint32_t lo = *reinterpret_cast<const unsigned short*>(wr + 2 * t);
      const uint32_t hi = *reinterpret_cast<const unsigned short*>(wr + 2 * t + 1);
      bw0[grp * 4 + j] = lo | (hi << 16);                                                   // k = 2t, 2t+1
      const float bv = __ldg(bias + ch);
      const __nv_bfloat16 b_hi = __float2bfloat16_rn(bv);
      const __nv_bfloat16 b_lo = __float2bfloat16_rn(bv - __bfloat162float(b_hi));
      uint32_t k89 = 0u;                                                                    // k = 2t+8, 2t+9
      if (t == 0) k89 = static_cast<uint32_t>(*reinterpret_cast<const unsigned short*>(wr + 8)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b_hi)) << 16);
      else if (t == 1) k89 = static_cast<uint32_t>(__bfloat16_as_ushort(b_lo));
      bw1[grp * 4 + j] = k89;
    }
  }
  __syncthreads();

  const int kh0 = (2 * t) / 3, kw0 = (2 * t) % 3, kh1 = (2 * t + 1) / 3, kw1 = (2 * t + 1) % 3;
  for (int mb = 0; mb < (kConv1RowsPerCta * 50) / 16; ++mb) {
    const int p_lo = mb * 16 + g, p_hi = p_lo + 8;
    const int oh_lo = p_lo / 50, ow_lo = p_lo - oh_lo * 50;
    const int oh_hi = p_hi / 50, ow_hi = p_hi - oh_hi * 50;
    const float v0 = in[2 * oh_lo + kh0][2 * ow_lo + kw0], v1 = in[2 * oh_lo + kh1][2 * ow_lo + kw1];
    const float v2 = in[2 * oh_hi + kh0][2 * ow_hi + kw0], v3 = in[2 * oh_hi + kh1][2 * ow_hi + kw1];
    const float v4 = (t == 0) ? in[2 * oh_lo + 2][2 * ow_lo + 2] : 0.0f;
    const float v5 = (t == 0) ? in[2 * oh_hi + 2][2 * ow_hi + 2] : 0.0f;
    uint32_t ahi[4], alo[4];
    {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1), h2 = __float2bfloat16_rn(v2),
                          h3 = __float2bfloat16_rn(v3), h4 = __float2bfloat16_rn(v4), h5 = __float2bfloat16_rn(v5);
      ahi[0] = ptx::pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
      ahi[1] = ptx::pack_bf16x2(__bfloat162float(h2), __bfloat162float(h3));
      // k = 2t+8, 2t+9: tap 8 and the constant 1.0 that multiplies bias_hi (t == 0); 1.0 for bias_lo (t == 1)
      ahi[2] = ptx::pack_bf16x2(t == 1 ? 1.0f : __bfloat162float(h4), t == 0 ? 1.0f : 0.0f);
      ahi[3] = ptx::pack_bf16x2(t == 1 ? 1.0f : __bfloat162float(h5), t == 0 ? 1.0f : 0.0f);
      alo[0] = ptx::pack_bf16x2(v0 - __bfloat162float(h0), v1 - __bfloat162float(h1));
      alo[1] = ptx::pack_bf16x2(v2 - __bfloat162float(h2), v3 - __bfloat162float(h3));
      alo[2] = ptx::pack_bf16x2(v4 - __bfloat162float(h4), 0.0f);
      alo[3] = ptx::pack_bf16x2(v5 - __bfloat162float(h5), 0.0f);
    }
    const int h_lo = oh0 + oh_lo, h_hi = oh0 + oh_hi;
    __nv_bfloat16* d_lo = planes + (2 * (h_lo & 1) + (ow_lo & 1)) * plane_stride +
                          ((static_cast<long long>(b) * 33 + (h_lo >> 1) + 1) * 26 + (ow_lo >> 1) + 1) * C + cwarp + 8 * t;
    __nv_bfloat16* d_hi = planes + (2 * (h_hi & 1) + (ow_hi & 1)) * plane_stride +
                          ((static_cast<long long>(b) * 33 + (h_hi >> 1) + 1) * 26 + (ow_hi >> 1) + 1) * C + cwarp + 8 * t;
#pragma unroll
    for (int grp = 0; grp < 3; ++grp) {
      float acc[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;
        mma_bf16_16816(acc[j], ahi, bw0[grp * 4 + j], bw1[grp * 4 + j]);
        mma_bf16_16816(acc[j], alo, bw0[grp * 4 + j], bw1[grp * 4 + j]);
      }
      uint4 q_lo, q_hi;
      q_lo.x = ptx::pack_bf16x2(gelu_from_half(acc[0][0]), gelu_from_half(acc[0][1]));
      q_lo.y = ptx::pack_bf16x2(gelu_from_half(acc[1][0]), gelu_from_half(acc[1][1]));
      q_lo.z = ptx::pack_bf16x2(gelu_from_half(acc[2][0]), gelu_from_half(acc[2][1]));
      q_lo.w = ptx::pack_bf16x2(gelu_from_half(acc[3][0]), gelu_from_half(acc[3][1]));
      q_hi.x = ptx::pack_bf16x2(gelu_from_half(acc[0][2]), gelu_from_half(acc[0][3]));
      q_hi.y = ptx::pack_bf16x2(gelu_from_half(acc[1][2]), gelu_from_half(acc[1][3]));
      q_hi.z = ptx::pack_bf16x2(gelu_from_half(acc[2][2]), gelu_from_half(acc[2][3]));
      q_hi.w = ptx::pack_bf16x2(gelu_from_half(acc[3][2]), gelu_from_half(acc[3][3]));
      *reinterpret_cast<uint4*>(d_lo + grp * 32) = q_lo;
      *reinterpret_cast<uint4*>(d_hi + grp * 32) = q_hi;
    }
  }
}

__global__ void __launch_bounds__(kAttnThreads, 3)
window_attention_kernel(const __nv_bfloat16* __restrict__ qkv, const WindowDesc* __restrict__ windows,
                        __nv_bfloat16* __restrict__ out, int D, float scale_log2e) {
  __shared__ __align__(16) __nv_bfloat16 sK[kAttnMaxWin * kAttnPitch];
  __shared__ __align__(16) __nv_bfloat16 sV[kAttnMaxWin * kAttnPitch];

  const WindowDesc wd = windows[blockIdx.x];
  const int head = blockIdx.y;
  const int len = wd.len;
  const int nk16 = (len + 15) >> 4;  // 16-key blocks
  const long long ld = 3LL * D;
  const __nv_bfloat16* __restrict__ base = qkv + static_cast<long long>(wd.start) * ld + head * 64;

  // stage K and V (zero rows beyond len so that masked probabilities multiply finite values)
  for (int i = threadIdx.x; i < nk16 * 16 * 8; i += kAttnThreads) {
    const int r = i >> 3, c = i & 7;  // 8 x 16-byte vectors per 64-wide row
    uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
    if (r < len) {
      kv = *reinterpret_cast<const uint4*>(base + r * ld + D + c * 8);
      vv = *reinterpret_cast<const uint4*>(base + r * ld + 2 * D + c * 8);
    }
    *reinterpret_cast<uint4*>(sK + r * kAttnPitch + c * 8) = kv;
    *reinterpret_cast<uint4*>(sV + r * kAttnPitch + c * 8) = vv;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int row0 = warp * 16;
  if (row0 >= len) return;

  // Q fragments straight from global (each row is used by exactly one warp)
  uint32_t qa[4][4];
  {
    const int r_lo = row0 + g, r_hi = row0 + g + 8;
    const __nv_bfloat16* q_lo = base + static_cast<long long>(min(r_lo, len - 1)) * ld;
    const __nv_bfloat16* q_hi = base + static_cast<long long>(min(r_hi, len - 1)) * ld;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      qa[kk][0] = *reinterpret_cast<const uint32_t*>(q_lo + kk * 16 + 2 * t);
      qa[kk][1] = *reinterpret_cast<const uint32_t*>(q_hi + kk * 16 + 2 * t);
      qa[kk][2] = *reinterpret_cast<const uint32_t*>(q_lo + kk * 16 + 8 + 2 * t);
      qa[kk][3] = *reinterpret_cast<const uint32_t*>(q_hi + kk * 16 + 8 + 2 * t);
    }
  }

  // S = Q K^T : 14 key tiles of 8
  float s[14][4];
#pragma unroll
  for (int j = 0; j < 14; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f; }
#pragma unroll
  for (int j = 0; j < 14; ++j) {
    if (j < 2 * nk16) {
      // matrices: (keys 8j.., dims 0-7), (dims 8-15), (dims 16-23), (dims 24-31) then dims 32-63
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t kb[4];
        const int mrow = 8 * j + (lane & 7);
        const int mcol = half * 32 + (lane >> 3) * 8;
        ldmatrix_x4(kb, sK + mrow * kAttnPitch + mcol);
        mma_bf16_16816(s[j], qa[2 * half + 0], kb[0], kb[1]);
        mma_bf16_16816(s[j], qa[2 * half + 1], kb[2], kb[3]);
      }
    }
  }

  // softmax over keys (rows g and g+8 of this warp's slab)
  float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
  for (int j = 0; j < 14; ++j) {
    const int c0 = 8 * j + 2 * t;
    if (c0 >= len) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
    if (c0 + 1 >= len) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
    mx_lo = fmaxf(mx_lo, fmaxf(s[j][0], s[j][1]));
    mx_hi = fmaxf(mx_hi, fmaxf(s[j][2], s[j][3]));
  }
  mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
  mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
  mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
  mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
  float sum_lo = 0.0f, sum_hi = 0.0f;
  uint32_t pa[7][4];
#pragma unroll
  for (int j = 0; j < 14; ++j) {
    const float p0 = exp2f((s[j][0] - mx_lo) * scale_log2e);
    const float p1 = exp2f((s[j][1] - mx_lo) * scale_log2e);
    const float p2 = exp2f((s[j][2] - mx_hi) * scale_log2e);
    const float p3 = exp2f((s[j][3] - mx_hi) * scale_log2e);
    sum_lo += p0 + p1;
    sum_hi += p2 + p3;
    pa[j >> 1][(j & 1) * 2 + 0] = ptx::pack_bf16x2(p0, p1);
    pa[j >> 1][(j & 1) * 2 + 1] = ptx::pack_bf16x2(p2, p3);
  }
  sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1);
  sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
  sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1);
  sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);

  // O = P V : 8 dim tiles of 8
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.0f; }
#pragma unroll
  for (int kk = 0; kk < 7; ++kk) {
    if (kk < nk16) {
#pragma unroll
      for (int jn = 0; jn < 8; jn += 2) {
        // matrices: (keys 16kk+0..7, dims 8jn..), (keys +8..15, dims 8jn..), same for jn+1
        uint32_t vb[4];
        const int mrow = 16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int mcol = 8 * jn + (lane >> 4) * 8;
        ldmatrix_x4_trans(vb, sV + mrow * kAttnPitch + mcol);
        mma_bf16_16816(o[jn], pa[kk], vb[0], vb[1]);
        mma_bf16_16816(o[jn + 1], pa[kk], vb[2], vb[3]);
      }
    }
  }

  const float inv_lo = 1.0f / sum_lo, inv_hi = 1.0f / sum_hi;
  const int r_lo = row0 + g, r_hi = row0 + g + 8;
  __nv_bfloat16* __restrict__ obase = out + static_cast<long long>(wd.start) * D + head * 64;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (r_lo < len)
      *reinterpret_cast<uint32_t*>(obase + static_cast<long long>(r_lo) * D + 8 * j + 2 * t) =
          ptx::pack_bf16x2(o[j][0] * inv_lo, o[j][1] * inv_lo);
    if (r_hi < len)
      *reinterpret_cast<uint32_t*>(obase + static_cast<long long>(r_hi) * D + 8 * j + 2 * t) =
          ptx::pack_bf16x2(o[j][2] * inv_hi, o[j][3] * inv_hi);
  }
}

}  // namespace qasr
