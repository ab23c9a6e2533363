// Causal grouped-query attention of the decoder prefill on tcgen05 (reference decoder.py:150-177: q/k already
// normalised and rotated, softmax(scale q k^T + causal mask) v with each KV head shared by n_heads / n_kv_heads query
// heads; head_dim 128; varlen-packed prompts, attention never crosses a prompt).
//
// Persistent CTAs; a work item is (prompt, 128-query tile, query head), its KV loop runs over the 128-key tiles
// 0 .. q0/128 (causal).  Flash-style online softmax with every matrix in tensor memory:
//   warp 0      TMA producer: Q tile (2 boxes of 64 dims) per item; K and V tiles (2 boxes each) per KV step into a
//               2-stage ring
//   warp 1      MMA issuer:   S = Q K^T   SS mode, M=128 N=128 K=128  -> TMEM S slot (kt & 1)      (QK of step kt+1 is
//                             issued before PV of step kt, so it overlaps the softmax of step kt)
//                             O (+)= P V  TS mode (P from TMEM), V as MN-major B operand, two N=64 halves
//   warps 2-9   softmax, TWO threads per query row (warps 2-5: keys / dims [0,64), warps 6-9: [64,128) of the tile):
//               tcgen05.ld the 64 scores once, row max exchanged through shared memory, exp2 + pack, rescale this
//               half of the running O in TMEM by exp2(m_old - m_new) once the previous PV has retired, park bf16 P
//               in TMEM; after the last step O / l -> bf16 -> global.  (One thread per row was latency-bound: 4.2 us
//               per KV step.)
// TMEM columns: S0 [0,128) | S1 [128,256) | P [256,320) (bf16 pairs) | O [320,448).
// Keys beyond the prompt (rows of the next prompt, stale workspace or TMA zero fill -- possibly NaN) and keys after
// the query get probability exactly 0: their score columns are overwritten with -inf, and on a prompt's last, partial
// KV tile the V rows beyond the prompt are zeroed in shared memory before the PV MMA (0 * NaN would be NaN); query rows
// beyond the prompt are never stored.  Prompts therefore stay isolated, like the reference's loop of singles.
#pragma once
#include "decoder_kernels.cuh"
#include "ptx.cuh"

namespace qasr {

constexpr int kCtThreads = 320;                     // 10 warps
constexpr int kCtBoxBytes = 128 * 128;              // 128 tokens x 64 dims bf16
constexpr int kCtQBytes = 2 * kCtBoxBytes;          // Q tile: 128 x 128
constexpr int kCtKvStageBytes = 4 * kCtBoxBytes;    // K (2 boxes) + V (2 boxes)
constexpr int kCtKvStages = 2;
constexpr int kCtXchgBytes = 3 * 2 * 128 * 4;        // [step parity 0/1 | epilogue][column half][row] floats
constexpr int kCtSmemBytes = 2 * kCtQBytes + kCtKvStages * kCtKvStageBytes + kCtXchgBytes + 256 + 1024;
constexpr bool kCtPvSingle = true;                  // one N=128 PV MMA per 16 keys (V's two 64-dim boxes 16 KB apart) instead of two N=64
constexpr int kCtPCol = 256, kCtOCol = 320;

__global__ void __launch_bounds__(kCtThreads, 1)
causal_attention_sm100(const __grid_constant__ CUtensorMap tmap_qkv, const AttnTile* __restrict__ tiles, int num_tiles,
                       int num_heads, int group, int k_off, int v_off, __nv_bfloat16* __restrict__ out, int ldo,
                       float scale_log2e) {
  extern __shared__ uint8_t ct_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ct_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_base = smem;                                  // 2 x Q tile
  uint8_t* kv_base = smem + 2 * kCtQBytes;                 // kCtKvStages x (K | V)
  float* xchg = reinterpret_cast<float*>(kv_base + kCtKvStages * kCtKvStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_base + kCtKvStages * kCtKvStageBytes + kCtXchgBytes);
  uint64_t* q_full = bars;         // [2]
  uint64_t* q_empty = bars + 2;    // [2]
  uint64_t* kv_full = bars + 4;    // [2]
  uint64_t* kv_empty = bars + 6;   // [2]
  uint64_t* s_full = bars + 8;     // [2]
  uint64_t* p_ready = bars + 10;   // [1]  P written (and O rescaled) by all 128 rows
  uint64_t* o_done = bars + 11;    // [1]  a PV step retired
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = num_tiles * num_heads;

  if (warp_idx == 0 && lane == 0) ptx::prefetch_tmap(&tmap_qkv);
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&q_full[i], 1);
      ptx::mbar_init(&q_empty[i], 1);
      ptx::mbar_init(&kv_full[i], 1);
      ptx::mbar_init(&kv_empty[i], 1);
      ptx::mbar_init(&s_full[i], 1);
    }
    ptx::mbar_init(p_ready, 256);
    ptx::mbar_init(o_done, 1);
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

  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int ji = 0;        // items processed by this CTA
      long long jk = 0;  // KV steps processed by this CTA
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++ji) {
        const AttnTile tl = tiles[item / num_heads];
        const int head = item % num_heads, kvh = head / group;
        const int qs = ji & 1;
        ptx::mbar_wait(&q_empty[qs], ((ji >> 1) & 1) ^ 1);
        ptx::mbar_expect_tx(&q_full[qs], kCtQBytes);
        uint8_t* qb = q_base + qs * kCtQBytes;
        ptx::tma_load_2d(qb, &tmap_qkv, &q_full[qs], head * 128, tl.start + tl.q0);
        ptx::tma_load_2d(qb + kCtBoxBytes, &tmap_qkv, &q_full[qs], head * 128 + 64, tl.start + tl.q0);
        const int n_kt = tl.q0 / 128 + 1;
        for (int kt = 0; kt < n_kt; ++kt, ++jk) {
          const int st = static_cast<int>(jk % kCtKvStages);
          ptx::mbar_wait(&kv_empty[st], static_cast<uint32_t>((jk / kCtKvStages) & 1) ^ 1);
          ptx::mbar_expect_tx(&kv_full[st], kCtKvStageBytes);
          uint8_t* sb = kv_base + st * kCtKvStageBytes;
          const int row = tl.start + kt * 128;
          ptx::tma_load_2d(sb, &tmap_qkv, &kv_full[st], k_off + kvh * 128, row);
          ptx::tma_load_2d(sb + kCtBoxBytes, &tmap_qkv, &kv_full[st], k_off + kvh * 128 + 64, row);
          ptx::tma_load_2d(sb + 2 * kCtBoxBytes, &tmap_qkv, &kv_full[st], v_off + kvh * 128, row);
          ptx::tma_load_2d(sb + 3 * kCtBoxBytes, &tmap_qkv, &kv_full[st], v_off + kvh * 128 + 64, row);
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(128, 128);
      constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major
      int ji = 0;
      long long jk = 0;  // global KV-step counter: KV ring stage, S slot and barrier phases derive from it
      auto issue_qk = [&](const uint8_t* qb, long long step) {
        const int st = static_cast<int>(step % kCtKvStages), slot = static_cast<int>(step & 1);
        ptx::mbar_wait(&kv_full[st], static_cast<uint32_t>((step / kCtKvStages) & 1));
        ptx::tc_fence_after();
        const uint8_t* sb = kv_base + st * kCtKvStageBytes;
        const uint32_t tmem_s = tmem_base + slot * 128;
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // 128 dims = two 64-dim swizzle atoms
          const uint64_t qd = ptx::make_sw128_kmajor_desc(ptx::smem_u32(qb + (k >> 2) * kCtBoxBytes)) + 2 * (k & 3);
          const uint64_t kd = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sb + (k >> 2) * kCtBoxBytes)) + 2 * (k & 3);
          ptx::umma_bf16_ss<1>(tmem_s, qd, kd, idesc_qk, k != 0);
        }
        ptx::umma_commit(&s_full[slot]);
      };
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++ji) {
        const AttnTile tl = tiles[item / num_heads];
        const int n_kt = tl.q0 / 128 + 1;
        const int qs = ji & 1;
        const uint8_t* qb = q_base + qs * kCtQBytes;
        ptx::mbar_wait(&q_full[qs], (ji >> 1) & 1);
        issue_qk(qb, jk);
        for (int kt = 0; kt < n_kt; ++kt) {
          const long long step = jk + kt;
          if (kt + 1 < n_kt) issue_qk(qb, step + 1);      // overlaps the softmax of this step
          else ptx::umma_commit(&q_empty[qs]);             // all QK^T of the item issued: Q stage free when they retire
          ptx::mbar_wait(p_ready, static_cast<uint32_t>(step & 1));
          ptx::tc_fence_after();
          const int st = static_cast<int>(step % kCtKvStages);
          const uint8_t* vb = kv_base + st * kCtKvStageBytes + 2 * kCtBoxBytes;
          const uint32_t tmem_p = tmem_base + kCtPCol, tmem_o = tmem_base + kCtOCol;
          if constexpr (kCtPvSingle) {
            constexpr uint32_t idesc_pv128 = ptx::make_idesc_bf16(128, 128, 0, 1);
            // MN-major B operand: 128 dims = two 64-dim swizzle blocks kCtBoxBytes apart (LBO), 8-key groups 1024 B apart (SBO)
            const uint64_t vd = ptx::make_sw128_mnmajor_desc(ptx::smem_u32(vb), kCtBoxBytes, 1024);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              ptx::umma_bf16_ts(tmem_o, tmem_p + 8 * k, vd + static_cast<uint64_t>(128 * k), idesc_pv128, (kt | k) != 0);
          } else {
#pragma unroll
            for (int half = 0; half < 2; ++half) {  // dims [0,64) and [64,128)
              const uint64_t vd = ptx::make_sw128_mnmajor_desc(ptx::smem_u32(vb + half * kCtBoxBytes), 1024, 1024);
#pragma unroll
              for (int k = 0; k < 8; ++k)  // 16 keys per step: 8 packed TMEM columns of P, 2048 B of V
                ptx::umma_bf16_ts(tmem_o + half * 64, tmem_p + 8 * k, vd + static_cast<uint64_t>(128 * k), idesc_pv, (kt | k) != 0);
            }
          }
          ptx::umma_commit(o_done);
          ptx::umma_commit(&kv_empty[st]);
        }
        jk += n_kt;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue (warps 2..9)
    const int quarter = warp_idx & 3;        // TMEM lane quarter accessible to this warp
    const int ch = (warp_idx - 2) >> 2;      // column half: keys / dims [64 ch, 64 ch + 64) of the tile
    const int r = quarter * 32 + lane;       // query row within the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    long long jk = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const AttnTile tl = tiles[item / num_heads];
      const int head = item % num_heads;
      const int n_kt = tl.q0 / 128 + 1;
      const int qi = tl.q0 + r;  // query index within the prompt
      float m = -INFINITY, l = 0.0f;  // l: this thread's half of the row sum
      for (int kt = 0; kt < n_kt; ++kt) {
        const long long step = jk + kt;
        const int slot = static_cast<int>(step & 1);
        const uint32_t tmem_s = tmem_base + lane_addr + slot * 128 + ch * 64;
        ptx::mbar_wait(&s_full[slot], static_cast<uint32_t>((step >> 1) & 1));
        ptx::tc_fence_after();
        const bool diag = kt == n_kt - 1;
        const int key0 = kt * 128 + ch * 64;
        uint32_t s0[32], s1[32];
        ptx::tmem_ld_32x32(tmem_s, s0);
        ptx::tmem_ld_32x32(tmem_s + 32, s1);
        ptx::tmem_ld_wait();
        if (diag) {  // masked keys (after the query, or beyond the prompt) -> -inf -> probability exactly 0
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (!(key0 + i <= qi && key0 + i < tl.len)) s0[i] = 0xff800000u;
            if (!(key0 + 32 + i <= qi && key0 + 32 + i < tl.len)) s1[i] = 0xff800000u;
          }
        }
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(s0[i]), __uint_as_float(s1[i])));
        // exchange the half-row maxima (double-buffered by step parity: no reader of step k can race a writer of k+1)
        float* xm = xchg + (step & 1) * 256;
        xm[ch * 128 + r] = mx;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mx = fmaxf(m, fmaxf(mx, xm[(ch ^ 1) * 128 + r]));
        // key 0 is visible to every row, so mx is finite from the first step on
        const float alpha = exp2f((m - mx) * scale_log2e);
        const float moff = mx * scale_log2e;
        m = mx;
        uint32_t pk[32];
        float sum = 0.0f;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float a0, a1, b0, b1;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(a0) : "f"(fmaf(__uint_as_float(s0[i]), scale_log2e, -moff)));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(a1) : "f"(fmaf(__uint_as_float(s0[i + 1]), scale_log2e, -moff)));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(b0) : "f"(fmaf(__uint_as_float(s1[i]), scale_log2e, -moff)));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(b1) : "f"(fmaf(__uint_as_float(s1[i + 1]), scale_log2e, -moff)));
          sum += (a0 + a1) + (b0 + b1);
          pk[i >> 1] = ptx::pack_bf16x2(a0, a1);
          pk[16 + (i >> 1)] = ptx::pack_bf16x2(b0, b1);
        }
        l = l * alpha + sum;
        const uint32_t tmem_o = tmem_base + lane_addr + kCtOCol + ch * 64;
        if (kt > 0) {
          // the previous PV (which read P and accumulated into O) must have retired
          ptx::mbar_wait(o_done, static_cast<uint32_t>((step - 1) & 1));
          ptx::tc_fence_after();
          if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t o[32];
              ptx::tmem_ld_32x32(tmem_o + 32 * c, o);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              ptx::tmem_st_32x32(tmem_o + 32 * c, o);
            }
          }
        }
        if (diag && tl.len < (kt + 1) * 128) {
          // zero V rows [len - 128 kt, 128) of both 64-dim boxes (row r of a 128B-swizzled box is bytes [128 r, 128 r + 128))
          uint8_t* vb = kv_base + static_cast<int>(step % kCtKvStages) * kCtKvStageBytes + 2 * kCtBoxBytes;
          const int r0 = tl.len - kt * 128;                     // 1 .. 127
          const int n16 = (128 - r0) * 8;                       // 16-byte vectors per box
          for (int t = ch * 128 + r; t < 2 * n16; t += 256) {
            const int box = t >= n16;
            reinterpret_cast<uint4*>(vb + box * kCtBoxBytes + r0 * 128)[t - box * n16] = make_uint4(0u, 0u, 0u, 0u);
          }
          ptx::fence_proxy_async_smem();
        }
        const uint32_t tmem_p = tmem_base + lane_addr + kCtPCol + ch * 32;  // 64 keys = 32 packed columns
        {
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = pk[i];
          ptx::tmem_st_32x16(tmem_p, w);
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = pk[16 + i];
          ptx::tmem_st_32x16(tmem_p + 16, w);
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(p_ready);
      }
      // ---- epilogue: O / l -> bf16 -> this thread's 128-byte half of the row's head slice
      const long long last = jk + n_kt - 1;
      float* xl = xchg + 512;  // its own buffer: a bar.sync of the next item's first step separates consecutive uses
      xl[ch * 128 + r] = l;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float inv = 1.0f / (l + xl[(ch ^ 1) * 128 + r]);
      ptx::mbar_wait(o_done, static_cast<uint32_t>(last & 1));
      ptx::tc_fence_after();
      __nv_bfloat16* orow = out + (static_cast<long long>(tl.start) + qi) * ldo + head * 128 + ch * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        ptx::tmem_ld_32x32(tmem_base + lane_addr + kCtOCol + ch * 64 + 32 * c, o);
        ptx::tmem_ld_wait();
        if (qi < tl.len) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 q;
            q.x = ptx::pack_bf16x2(__uint_as_float(o[8 * u + 0]) * inv, __uint_as_float(o[8 * u + 1]) * inv);
            q.y = ptx::pack_bf16x2(__uint_as_float(o[8 * u + 2]) * inv, __uint_as_float(o[8 * u + 3]) * inv);
            q.z = ptx::pack_bf16x2(__uint_as_float(o[8 * u + 4]) * inv, __uint_as_float(o[8 * u + 5]) * inv);
            q.w = ptx::pack_bf16x2(__uint_as_float(o[8 * u + 6]) * inv, __uint_as_float(o[8 * u + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + 32 * c + 8 * u) = q;
          }
        }
      }
      ptx::tc_fence_before();
      jk += n_kt;
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
