// Windowed multi-head attention on tcgen05 (reference: encoder.py:78-85 inside the block-diagonal
// windows of encoder.py:297-311; head_dim 64, windows of <= 104 packed tokens).
//
// Persistent CTAs, one (window, head) work item at a time; loads run up to 4 items ahead
// (4 smem stages hide the L2/HBM latency), two items share the TMEM (2 slots):
//   warp 0        TMA producer: Q, K, V tiles (128 tokens x 64 dims, 128B swizzle) of item j -> smem stage j%4
//   warp 1        MMA issuer:   S = Q K^T   SS mode, M=128 N=112 K=64  -> TMEM slot j&1, columns [0,112)
//                               O = P V     TS mode (P read from TMEM), V is the MN-major B operand,
//                                           M=128 N=64 K=16*ceil(len/16) -> columns [192,256)
//   warps 2-5     softmax + epilogue group 0 (even items, TMEM slot 0), one thread per query row:
//   warps 6-9     softmax + epilogue group 1 (odd items,  TMEM slot 1)
//                 tcgen05.ld S -> fp32 softmax in registers -> bf16 P via tcgen05.st into columns
//                 [128,184) -> ... -> tcgen05.ld O, scale by 1/sum, bf16, swizzled-smem transpose,
//                 16-byte coalesced stores.  Two groups keep two warps per scheduler busy (the
//                 single-group version was bound by instruction latency: ncu 4.3 cycles / issue).
// The MMA thread polls (test_wait: never suspends) its two kinds of pending work, S = QK^T of the next item and
// O = PV of the oldest item, so neither can block the other.
// The O(n^2) additive mask of the reference never exists: keys beyond the window length get
// probability exactly 0.  Rows of the 128-token tiles that lie beyond the window belong to neighbouring windows /
// utterances (or are stale workspace / TMA zero fill) and may hold ANYTHING, including the NaNs of a poisoned
// utterance: masked score columns are overwritten with -inf before use, the V rows [len, 16*ceil(len/16)) that
// the PV MMA multiplies by those exact zeros are zeroed in shared memory first (0 * NaN would be NaN), and query
// rows beyond the window are never stored.  Utterances therefore stay isolated, like the reference's loop of singles.
#pragma once
#include "encoder_kernels.cuh"
#include "ptx.cuh"

namespace qasr {

constexpr int kAtThreads = 320;
constexpr int kAtTileBytes = 128 * 128;            // 128 tokens x 64 bf16
constexpr int kAtStageBytes = 3 * kAtTileBytes;    // Q, K, V
constexpr int kAtStagingBytes = 8 * 32 * 128;      // per softmax warp: 32 rows x 128 B
constexpr int kAtStages = 4;
constexpr int kAtSmemBytes = kAtStages * kAtStageBytes + kAtStagingBytes + 256 + 1024;
constexpr int kAtKeysPadded = 112;                 // 104 rounded up to a multiple of 16
constexpr int kAtSlotCols = 256;                   // TMEM columns per in-flight item
constexpr int kAtPCol = 128, kAtOCol = 192;

__global__ void __launch_bounds__(kAtThreads, 1)
window_attention_sm100(const __grid_constant__ CUtensorMap tmap_qkv, const WindowDesc* __restrict__ windows,
                       int num_windows, int num_heads, int D, __nv_bfloat16* __restrict__ out, float scale_log2e,
                       int reverse = 0 /* 1: last window first (serpentine row order between consecutive kernels) */,
                       unsigned long long qkv_policy = 0 /* L2 eviction priority of the Q/K/V loads (0: none) */) {
  extern __shared__ uint8_t at_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;                                   // kAtStages x (Q | K | V)
  uint8_t* staging = smem + kAtStages * kAtStageBytes;          // output transpose buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kAtStagingBytes);
  uint64_t* full = bars;                      // [kAtStages] TMA landed
  uint64_t* sempty = bars + kAtStages;        // [kAtStages] smem stage consumed by the MMAs
  uint64_t* s_full = bars + 2 * kAtStages;    // [2] S ready in TMEM
  uint64_t* p_ready = s_full + 2;             // [2] P written to TMEM
  uint64_t* o_full = s_full + 4;              // [2] O ready in TMEM
  uint64_t* tfree = s_full + 6;               // [2] TMEM slot drained
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(s_full + 8);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = num_windows * num_heads;

  if (warp_idx == 0 && lane == 0) ptx::prefetch_tmap(&tmap_qkv);
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < kAtStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&sempty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&p_ready[i], 128);
      ptx::mbar_init(&o_full[i], 1);
      ptx::mbar_init(&tfree[i], 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) {
    ptx::tmem_alloc<1>(tmem_ptr_smem, 512);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  ptx::grid_dep_wait();  // programmatic dependent launch: q/k/v are valid from here
  ptx::grid_dep_launch();

  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int j = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++j) {
        const int st = j % kAtStages;
        const uint32_t ph = (j / kAtStages) & 1;
        const int it = reverse ? num_items - 1 - item : item;
        const WindowDesc wd = windows[it / num_heads];
        const int head = it % num_heads;
        ptx::mbar_wait(&sempty[st], ph ^ 1);
        ptx::mbar_expect_tx(&full[st], kAtStageBytes);
        uint8_t* sb = stage_base + st * kAtStageBytes;
        if (qkv_policy) {
          ptx::tma_load_2d_hint(sb, &tmap_qkv, &full[st], head * 64, wd.start, qkv_policy);
          ptx::tma_load_2d_hint(sb + kAtTileBytes, &tmap_qkv, &full[st], D + head * 64, wd.start, qkv_policy);
          ptx::tma_load_2d_hint(sb + 2 * kAtTileBytes, &tmap_qkv, &full[st], 2 * D + head * 64, wd.start, qkv_policy);
        } else {
          ptx::tma_load_2d(sb, &tmap_qkv, &full[st], head * 64, wd.start);
          ptx::tma_load_2d(sb + kAtTileBytes, &tmap_qkv, &full[st], D + head * 64, wd.start);
          ptx::tma_load_2d(sb + 2 * kAtTileBytes, &tmap_qkv, &full[st], 2 * D + head * 64, wd.start);
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(128, kAtKeysPadded);
      constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major
      int n_mine = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) ++n_mine;
      int next_qk = 0, next_pv = 0;
      while (next_pv < n_mine) {
        // S = Q K^T of item next_qk: needs its smem stage and a drained TMEM slot (at most 2 items ahead of PV)
        if (next_qk < n_mine && next_qk < next_pv + 2) {
          const int j = next_qk, st = j % kAtStages, slot = j & 1;
          if (ptx::mbar_test_wait(&full[st], (j / kAtStages) & 1) && ptx::mbar_test_wait(&tfree[slot], ((j >> 1) & 1) ^ 1)) {
            ptx::tc_fence_after();
            const uint8_t* sb = stage_base + st * kAtStageBytes;
            const uint64_t qd = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sb));
            const uint64_t kd = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sb + kAtTileBytes));
            const uint32_t tmem_s = tmem_base + slot * kAtSlotCols;
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::umma_bf16_ss<1>(tmem_s, qd + 2 * k, kd + 2 * k, idesc_qk, k != 0);
            ptx::umma_commit(&s_full[slot]);
            ++next_qk;
            continue;
          }
        }
        // O = P V of item next_pv: needs the probabilities written by its softmax group
        if (next_pv < next_qk) {
          const int j = next_pv, st = j % kAtStages, slot = j & 1;
          if (ptx::mbar_test_wait(&p_ready[slot], (j >> 1) & 1)) {
            ptx::tc_fence_after();
            const int item = blockIdx.x + j * gridDim.x;
            const int len = windows[(reverse ? num_items - 1 - item : item) / num_heads].len;
            const int ksteps = (len + 15) >> 4;
            const uint8_t* sb = stage_base + st * kAtStageBytes;
            // V tile: 64 dims (one 128-byte swizzle row) per key, 8-key groups 1024 B apart
            const uint64_t vd = ptx::make_sw128_mnmajor_desc(ptx::smem_u32(sb + 2 * kAtTileBytes), 1024, 1024);
            const uint32_t tmem_p = tmem_base + slot * kAtSlotCols + kAtPCol;
            const uint32_t tmem_o = tmem_base + slot * kAtSlotCols + kAtOCol;
            for (int k = 0; k < ksteps; ++k)  // 16 keys per step: 8 packed TMEM columns of P, 2048 B of V
              ptx::umma_bf16_ts(tmem_o, tmem_p + 8 * k, vd + static_cast<uint64_t>(128 * k), idesc_pv, k != 0);
            ptx::umma_commit(&o_full[slot]);
            ptx::umma_commit(&sempty[st]);
            ++next_pv;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue (warps 2..5)
    const int quarter = warp_idx & 3;            // TMEM lane quarter accessible to this warp
    const int group = (warp_idx - 2) >> 2;       // 0: even items / slot 0, 1: odd items / slot 1
    uint4* stg = reinterpret_cast<uint4*>(staging) + (warp_idx - 2) * (32 * 8);
    for (int j = group; blockIdx.x + static_cast<long long>(j) * gridDim.x < num_items; j += 2) {
      const int item = blockIdx.x + j * gridDim.x;
      const int st = group;                      // TMEM slot
      const uint32_t ph = (j >> 1) & 1;
      const int it = reverse ? num_items - 1 - item : item;
      const WindowDesc wd = windows[it / num_heads];
      const int head = it % num_heads;
      const int len = wd.len;
      const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
      const uint32_t tmem_s = tmem_base + lane_addr + st * kAtSlotCols;

      ptx::mbar_wait(&s_full[st], ph);
      ptx::tc_fence_after();
      uint32_t s0[32], s1[32], s2[32], s3[16];
      ptx::tmem_ld_32x32(tmem_s, s0);
      ptx::tmem_ld_32x32(tmem_s + 32, s1);
      ptx::tmem_ld_32x32(tmem_s + 64, s2);
      ptx::tmem_ld_32x16(tmem_s + 96, s3);
      ptx::tmem_ld_wait();
      // Columns >= len are masked.  `len` is uniform, so fully valid 32-column blocks take a mask-free path.
      auto block_max = [&](uint32_t (&s)[32], int c0) {
        float m = -INFINITY;
        if (len >= c0 + 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(s[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (c0 + i >= len) s[i] = 0xff800000u;  // -inf -> probability exactly 0
            m = fmaxf(m, __uint_as_float(s[i]));
          }
        }
        return m;
      };
      float mx = fmaxf(fmaxf(block_max(s0, 0), block_max(s1, 32)), block_max(s2, 64));
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (96 + i >= len) s3[i] = 0xff800000u;
        mx = fmaxf(mx, __uint_as_float(s3[i]));
      }
      const float moff = mx * scale_log2e;
      auto ex2 = [&](uint32_t sv) {
        float r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(__uint_as_float(sv), scale_log2e, -moff)));
        return r;
      };
      float sum = 0.0f;
      uint32_t p0[16], p1[16], p2[16], p3[8];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a0 = ex2(s0[2 * i]), a1 = ex2(s0[2 * i + 1]);
        const float b0 = ex2(s1[2 * i]), b1 = ex2(s1[2 * i + 1]);
        const float c0 = ex2(s2[2 * i]), c1 = ex2(s2[2 * i + 1]);
        sum += (a0 + a1) + (b0 + b1) + (c0 + c1);
        p0[i] = ptx::pack_bf16x2(a0, a1);
        p1[i] = ptx::pack_bf16x2(b0, b1);
        p2[i] = ptx::pack_bf16x2(c0, c1);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a0 = ex2(s3[2 * i]), a1 = ex2(s3[2 * i + 1]);
        sum += a0 + a1;
        p3[i] = ptx::pack_bf16x2(a0, a1);
      }
      if (len & 15) {
        // zero the V rows the PV MMA reads beyond the window (row r of the 128B-swizzled tile is bytes [128 r, 128 r + 128))
        uint8_t* vt = stage_base + (j % kAtStages) * kAtStageBytes + 2 * kAtTileBytes;
        const int n16 = (16 - (len & 15)) * 8;  // 16-byte vectors, <= 120
        const int t = quarter * 32 + lane;      // the 4 warps of a group cover all four quarters
        if (t < n16) reinterpret_cast<uint4*>(vt + len * 128)[t] = make_uint4(0u, 0u, 0u, 0u);
        ptx::fence_proxy_async_smem();          // generic-proxy writes -> visible to the MMA's async-proxy reads
      }
      const uint32_t tmem_p = tmem_base + lane_addr + st * kAtSlotCols + kAtPCol;
      ptx::tmem_st_32x16(tmem_p, p0);
      ptx::tmem_st_32x16(tmem_p + 16, p1);
      ptx::tmem_st_32x16(tmem_p + 32, p2);
      {
        uint32_t p3w[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) { p3w[i] = p3[i]; p3w[8 + i] = 0u; }  // columns 56..63: zero padding
        ptx::tmem_st_32x16(tmem_p + 48, p3w);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_ready[st]);

      // ---- O = P V is running; then read it back
      ptx::mbar_wait(&o_full[st], ph);
      ptx::tc_fence_after();
      uint32_t o0[32], o1[32];
      const uint32_t tmem_o = tmem_base + lane_addr + st * kAtSlotCols + kAtOCol;
      ptx::tmem_ld_32x32(tmem_o, o0);
      ptx::tmem_ld_32x32(tmem_o + 32, o1);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tfree[st]);  // the slot may be overwritten by the QK^T of item j+2

      const float inv = 1.0f / sum;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint4 q;
        q.x = ptx::pack_bf16x2(__uint_as_float(o0[8 * u + 0]) * inv, __uint_as_float(o0[8 * u + 1]) * inv);
        q.y = ptx::pack_bf16x2(__uint_as_float(o0[8 * u + 2]) * inv, __uint_as_float(o0[8 * u + 3]) * inv);
        q.z = ptx::pack_bf16x2(__uint_as_float(o0[8 * u + 4]) * inv, __uint_as_float(o0[8 * u + 5]) * inv);
        q.w = ptx::pack_bf16x2(__uint_as_float(o0[8 * u + 6]) * inv, __uint_as_float(o0[8 * u + 7]) * inv);
        stg[lane * 8 + (u ^ (lane & 7))] = q;
        q.x = ptx::pack_bf16x2(__uint_as_float(o1[8 * u + 0]) * inv, __uint_as_float(o1[8 * u + 1]) * inv);
        q.y = ptx::pack_bf16x2(__uint_as_float(o1[8 * u + 2]) * inv, __uint_as_float(o1[8 * u + 3]) * inv);
        q.z = ptx::pack_bf16x2(__uint_as_float(o1[8 * u + 4]) * inv, __uint_as_float(o1[8 * u + 5]) * inv);
        q.w = ptx::pack_bf16x2(__uint_as_float(o1[8 * u + 6]) * inv, __uint_as_float(o1[8 * u + 7]) * inv);
        stg[lane * 8 + ((u + 4) ^ (lane & 7))] = q;
      }
      __syncwarp();
      // transposed write-out: 8 lanes cover one token's 128-byte head slice, 4 tokens per instruction
      const int u = lane & 7, rsub = lane >> 3;
      __nv_bfloat16* obase = out + static_cast<long long>(wd.start) * D + head * 64 + 8 * u;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + rsub;
        const int tok = quarter * 32 + rr;
        const uint4 q = stg[rr * 8 + (u ^ (rr & 7))];
        if (tok < len) *reinterpret_cast<uint4*>(obase + static_cast<long long>(tok) * D) = q;
      }
      __syncwarp();
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, 512);
  }
}

}  // namespace qasr
