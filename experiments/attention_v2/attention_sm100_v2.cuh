// Windowed multi-head attention on tcgen05, THREE work items in flight per SM (reference: encoder.py:78-85 inside the
// block-diagonal windows of encoder.py:297-311; head_dim 64, windows of <= 104 packed tokens).
//
// attention_sm100.cuh (two items in flight, 168 registers, ~1000 instructions per row and item) sits at 12 % tensor pipe:
// ncu shows 0.38 IPC per scheduler with two softmax warps each -- one long dependent chain per item
// (S -> max -> exp -> P -> PV -> O) -- and the XU (ex2) pipe at 36 %.  The bounds of the algorithm on this machine are
// tensor-memory READ bandwidth (64 B/clk/SM: S 104 x 128 x 4 B + O 64 x 128 x 4 B = 86 KB -> 1350 clk per item) and
// XU (104 x 128 ex2 / 16 per clk = 830 clk).  (A first attempt with four items in flight and a TWO-pass softmax, which
// re-read S to save registers, ran into exactly that TMEM read bound and was slower.)  This kernel keeps ONE pass and
//   * cuts the instruction count per row and item roughly in half: 3-input max, packed FFMA2 / FADD2 (fma.rn.f32x2) for
//     the scale-and-subtract and the row sum, only the 104 columns a window can have, output rows stored straight from
//     registers (64 dims = one 128-byte line per thread) instead of a shared-memory transpose;
//   * fits in 144 registers, so THREE softmax groups (12 warps, 3 chains per scheduler) run beside the producer / MMA warps;
//   * uses 128-column TMEM slots: the bf16 probabilities overwrite the score columns their own thread has already
//     consumed (cols [0,56)), O lands in cols [64,128);
//   * loads Q / K / V as 104-row TMA boxes (13 KB, 13 KB and a 14 KB region for V): the MMAs still read 128 Q rows and
//     112 K rows -- the rows beyond a tile are the first rows of the next one (valid shared memory, results never
//     used) -- which leaves room for 5 stages (two items of load-ahead beyond the 3 in flight).
//   * gives the two kinds of MMA work their OWN issuing threads, each sleeping on its own mbarriers (try_wait, ~60 clk
//     wake-up).  Cycle accounting of the single polling issuer (test_wait ~150 clk a probe, two or three probes, a global
//     load of the window length and an integer division per iteration) showed 3300 clk per item in that ONE thread --
//     the whole kernel ran at its pace, the softmax warps idle 70 % of the time.
//   warp 0      TMA producer (also publishes the window length of every stage in shared memory)
//   warp 1      QK^T issuer: S = Q K^T of item j as soon as its stage has landed and its TMEM slot is drained
//   warp 2      PV issuer:   O = P V of item j as soon as its softmax group has written P
//   warps 3-14  softmax + epilogue, group g = (warp - 3) / 4 serves items j with j % 3 == g in TMEM slot g.
// Masking and isolation exactly as in attention_sm100.cuh: score columns >= len get probability 0, V rows
// [len, 16 ceil(len/16)) are zeroed in shared memory before the PV MMA, query rows >= len are never stored.
#pragma once
#include "encoder_kernels.cuh"
#include "ptx.cuh"

namespace qasr {

constexpr int kAt2Groups = 3;
constexpr int kAt2Threads = 96 + kAt2Groups * 128;  // 480: producer, QK issuer, PV issuer, 12 softmax warps
constexpr int kAt2BoxRows = 104;                    // window_tokens of the 1.7B config (13 * 8)
constexpr int kAt2QBytes = kAt2BoxRows * 128;       // 13312
constexpr int kAt2VBytes = 112 * 128;               // 14336: rows 104..111 are zeroed, never loaded
constexpr int kAt2StageBytes = 2 * kAt2QBytes + kAt2VBytes;  // 40960
constexpr int kAt2TxBytes = 3 * kAt2BoxRows * 128;  // bytes one stage's three TMA boxes deliver
constexpr int kAt2Stages = 5;
constexpr int kAt2SmemBytes = kAt2Stages * kAt2StageBytes + 512 + 1024;
constexpr int kAt2SlotCols = 128;
constexpr int kAt2OCol = 64;
static_assert(kAt2StageBytes % 1024 == 0 && kAt2QBytes % 1024 == 0, "SW128 tiles need 1024-byte aligned bases");

// -DQASR_AT2_DBG=1: cycle accounting of CTA 0 (one softmax thread of group 0, the MMA thread, the producer), read back with
// qasr_debug_counters().  Compiled out of the product library.
#ifndef QASR_AT2_DBG
#define QASR_AT2_DBG 0
#endif
#if QASR_AT2_DBG
__device__ unsigned long long g_at2_dbg[32];
#define AT2_T(var) const long long var = clock64()
#define AT2_ADD(i, v) atomicAdd(&g_at2_dbg[i], static_cast<unsigned long long>(v))
#else
#define AT2_T(var)
#define AT2_ADD(i, v)
#endif

__device__ __forceinline__ float at2_max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float at2_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// (d0, d1) = (a0, a1) * b + c with one FFMA2; (s0, s1) += (a0, a1) with one FADD2 (IEEE, same results as the scalar ops)
__device__ __forceinline__ void at2_fma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm("{ .reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%4};\n mov.b64 rc, {%5,%5};\n"
      " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0,%1}, rd; }"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void at2_add2(float& s0, float& s1, float a0, float a1) {
  asm("{ .reg .b64 ra, rb, rd;\n mov.b64 ra, {%0,%1};\n mov.b64 rb, {%2,%3};\n add.f32x2 rd, ra, rb;\n mov.b64 {%0,%1}, rd; }"
      : "+f"(s0), "+f"(s1) : "f"(a0), "f"(a1));
}
__device__ __forceinline__ void at2_mul2(float& d0, float& d1, float a0, float a1, float b) {
  asm("{ .reg .b64 ra, rb, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%4};\n mul.f32x2 rd, ra, rb;\n mov.b64 {%0,%1}, rd; }"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b));
}

__global__ void __launch_bounds__(kAt2Threads, 1)
window_attention_sm100_v2(const __grid_constant__ CUtensorMap tmap_qkv /* box {64, 104} */, const WindowDesc* __restrict__ windows,
                          int num_windows, int num_heads, int D, __nv_bfloat16* __restrict__ out, float scale_log2e) {
  extern __shared__ uint8_t at2_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at2_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;  // kAt2Stages x (Q | K | V)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAt2Stages * kAt2StageBytes);
  uint64_t* full = bars;                       // [kAt2Stages] TMA landed
  uint64_t* sempty = bars + kAt2Stages;        // [kAt2Stages] stage consumed by the PV MMA
  uint64_t* s_full = bars + 2 * kAt2Stages;    // [4] S ready in TMEM
  uint64_t* p_ready = s_full + kAt2Groups;     // [4] P written to TMEM
  uint64_t* o_full = s_full + 2 * kAt2Groups;  // [4] O ready in TMEM
  uint64_t* tfree = s_full + 3 * kAt2Groups;   // [4] TMEM slot drained
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(s_full + 4 * kAt2Groups);
  int* s_len = reinterpret_cast<int*>(tmem_ptr_smem + 2);  // [kAt2Stages] window length of the item in each stage

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = num_windows * num_heads;

  if (warp_idx == 0 && lane == 0) ptx::prefetch_tmap(&tmap_qkv);
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < kAt2Stages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&sempty[i], 1);
    }
    for (int i = 0; i < kAt2Groups; ++i) {
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

  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int j = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++j) {
        const int st = j % kAt2Stages;
        const uint32_t ph = (j / kAt2Stages) & 1;
        const WindowDesc wd = windows[item / num_heads];
        const int head = item % num_heads;
        AT2_T(tp0);
        ptx::mbar_wait(&sempty[st], ph ^ 1);
        AT2_T(tp1);
        if (blockIdx.x == 0) AT2_ADD(11, tp1 - tp0);
        s_len[st] = wd.len;  // read by the PV issuer (ordered by the full -> s_full -> p_ready barrier chain)
        ptx::mbar_expect_tx(&full[st], kAt2TxBytes);
        uint8_t* sb = stage_base + st * kAt2StageBytes;
        ptx::tma_load_2d(sb, &tmap_qkv, &full[st], head * 64, wd.start);
        ptx::tma_load_2d(sb + kAt2QBytes, &tmap_qkv, &full[st], D + head * 64, wd.start);
        ptx::tma_load_2d(sb + 2 * kAt2QBytes, &tmap_qkv, &full[st], 2 * D + head * 64, wd.start);
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------------ QK^T issuer
    if (lane == 0) {
      constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(128, 112);
      int j = 0;
      AT2_T(tm0);
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++j) {
        const int st = j % kAt2Stages, slot = j % kAt2Groups;
        ptx::mbar_wait(&full[st], (j / kAt2Stages) & 1);                 // Q, K (and V) have landed
#ifdef QASR_AT2_STAGGER
        if (j > 0 && j < kAt2Groups) __nanosleep(j * QASR_AT2_STAGGER);    // experiment: de-phase the three softmax groups
#endif
        ptx::mbar_wait(&tfree[slot], ((j / kAt2Groups) & 1) ^ 1);        // the slot's previous item has been read out
        ptx::tc_fence_after();
        const uint8_t* sb = stage_base + st * kAt2StageBytes;
        const uint64_t qd = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sb));
        const uint64_t kd = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sb + kAt2QBytes));
        const uint32_t tmem_s = tmem_base + slot * kAt2SlotCols;
#pragma unroll
        AT2_T(tq0);
        for (int k = 0; k < 4; ++k) ptx::umma_bf16_ss<1>(tmem_s, qd + 2 * k, kd + 2 * k, idesc_qk, k != 0);
        ptx::umma_commit(&s_full[slot]);
#if QASR_AT2_DBG
        if (blockIdx.x == 0) {  // issue -> completion latency of the four QK^T MMAs (debug build only: serialises the issuer)
          ptx::mbar_wait(&s_full[slot], (j / kAt2Groups) & 1);
          AT2_ADD(13, clock64() - tq0);
        }
#endif
      }
      AT2_T(tm1);
      if (blockIdx.x == 0) { AT2_ADD(8, tm1 - tm0); AT2_ADD(9, j); }
    }
  } else if (warp_idx == 2) {
    // ------------------------------------------------------------------ PV issuer
    if (lane == 0) {
      constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major
      int j = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++j) {
        const int st = j % kAt2Stages, slot = j % kAt2Groups;
        ptx::mbar_wait(&p_ready[slot], (j / kAt2Groups) & 1);           // the softmax group has written P (and zeroed V's tail)
        ptx::tc_fence_after();
        const int ksteps = (s_len[st] + 15) >> 4;
        const uint8_t* sb = stage_base + st * kAt2StageBytes;
        // V tile: 64 dims (one 128-byte swizzle row) per key, 8-key groups 1024 B apart
        const uint64_t vd = ptx::make_sw128_mnmajor_desc(ptx::smem_u32(sb + 2 * kAt2QBytes), 1024, 1024);
        const uint32_t tmem_p = tmem_base + slot * kAt2SlotCols;
        const uint32_t tmem_o = tmem_base + slot * kAt2SlotCols + kAt2OCol;
        AT2_T(tv0);
        for (int k = 0; k < ksteps; ++k)  // 16 keys per step: 8 packed TMEM columns of P, 2048 B of V
          ptx::umma_bf16_ts(tmem_o, tmem_p + 8 * k, vd + static_cast<uint64_t>(128 * k), idesc_pv, k != 0);
        ptx::umma_commit(&o_full[slot]);
        ptx::umma_commit(&sempty[st]);
#if QASR_AT2_DBG
        if (blockIdx.x == 0) {
          ptx::mbar_wait(&o_full[slot], (j / kAt2Groups) & 1);
          AT2_ADD(12, clock64() - tv0);
        }
#endif
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue (warps 3..14)
    const int quarter = warp_idx & 3;             // TMEM lane quarter accessible to this warp
    const int group = (warp_idx - 3) >> 2;        // serves items j % 3 == group in TMEM slot `group`
    const int row = quarter * 32 + lane;          // query row within the window == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tmem_s = tmem_base + lane_addr + group * kAt2SlotCols;
    for (int j = group; blockIdx.x + static_cast<long long>(j) * gridDim.x < num_items; j += kAt2Groups) {
      const int item = blockIdx.x + j * gridDim.x;
      const uint32_t ph = (j / kAt2Groups) & 1;
      const WindowDesc wd = windows[item / num_heads];
      const int head = item % num_heads;
      const int len = wd.len;

      const bool dbg_me = QASR_AT2_DBG && blockIdx.x == 0 && threadIdx.x == 96;
      (void)dbg_me;
      AT2_T(t0);
      ptx::mbar_wait(&s_full[group], ph);
      ptx::tc_fence_after();
      AT2_T(t1);
      // ---- one pass: all 104 score columns of this row into registers
      uint32_t s[kAt2BoxRows];
      {
        uint32_t (&s0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[0]);
        uint32_t (&s1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[32]);
        uint32_t (&s2)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[64]);
        uint32_t (&s3)[8] = *reinterpret_cast<uint32_t (*)[8]>(&s[96]);
        ptx::tmem_ld_32x32(tmem_s, s0);
        ptx::tmem_ld_32x32(tmem_s + 32, s1);
        ptx::tmem_ld_32x32(tmem_s + 64, s2);
        ptx::tmem_ld_32x8(tmem_s + 96, s3);
        ptx::tmem_ld_wait();
      }
      AT2_T(t2);
      uint32_t p[kAt2BoxRows / 2 + 4];  // bf16 pairs; 4 zero words pad the last 16-key step (keys 104..111)
      float sum0 = 0.0f, sum1 = 0.0f;
      if (len == kAt2BoxRows) {
        // full window (the common case): no masks
        float mx = at2_max3(__uint_as_float(s[0]), __uint_as_float(s[1]), __uint_as_float(s[2]));
        float mx2 = at2_max3(__uint_as_float(s[3]), __uint_as_float(s[4]), __uint_as_float(s[5]));
#pragma unroll
        for (int i = 6; i + 3 < kAt2BoxRows; i += 4) {
          mx = at2_max3(mx, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
          mx2 = at2_max3(mx2, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
        }
        mx = at2_max3(mx, mx2, fmaxf(__uint_as_float(s[102]), __uint_as_float(s[103])));
        const float moff = -mx * scale_log2e;
#pragma unroll
        for (int i = 0; i < kAt2BoxRows / 2; ++i) {
          float x0, x1;
          at2_fma2(x0, x1, __uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]), scale_log2e, moff);
          const float e0 = at2_ex2(x0), e1 = at2_ex2(x1);
          at2_add2(sum0, sum1, e0, e1);
          p[i] = ptx::pack_bf16x2(e0, e1);
        }
      } else {
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < kAt2BoxRows; ++i)
          if (i < len) mx = fmaxf(mx, __uint_as_float(s[i]));
        const float moff = -mx * scale_log2e;
#pragma unroll
        for (int i = 0; i < kAt2BoxRows / 2; ++i) {
          float x0, x1;
          at2_fma2(x0, x1, __uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]), scale_log2e, moff);
          const float e0 = (2 * i < len) ? at2_ex2(x0) : 0.0f;      // masked keys: probability exactly 0
          const float e1 = (2 * i + 1 < len) ? at2_ex2(x1) : 0.0f;
          at2_add2(sum0, sum1, e0, e1);
          p[i] = ptx::pack_bf16x2(e0, e1);
        }
      }
      const float sum = sum0 + sum1;
      AT2_T(t3);
#pragma unroll
      for (int i = 0; i < 4; ++i) p[kAt2BoxRows / 2 + i] = 0u;
      {  // P (bf16 pairs) over the consumed score columns [0,56)
        ptx::tmem_st_32x32(tmem_s, *reinterpret_cast<const uint32_t (*)[32]>(&p[0]));
        ptx::tmem_st_32x16(tmem_s + 32, *reinterpret_cast<const uint32_t (*)[16]>(&p[32]));
        ptx::tmem_st_32x8(tmem_s + 48, *reinterpret_cast<const uint32_t (*)[8]>(&p[48]));
      }
      if (len & 15) {
        // zero the V rows the PV MMA reads beyond the window (row r of the 128B-swizzled tile is bytes [128 r, 128 r + 128))
        uint8_t* vt = stage_base + (j % kAt2Stages) * kAt2StageBytes + 2 * kAt2QBytes;
        const int n16 = (16 - (len & 15)) * 8;  // 16-byte vectors, <= 120
        if (row < n16) reinterpret_cast<uint4*>(vt + len * 128)[row] = make_uint4(0u, 0u, 0u, 0u);
        ptx::fence_proxy_async_smem();          // generic-proxy writes -> visible to the MMA's async-proxy reads
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_ready[group]);
      AT2_T(t4);

      // ---- O = P V is running; then read it back, scale by 1 / sum and store the row's 128-byte head slice
      ptx::mbar_wait(&o_full[group], ph);
      ptx::tc_fence_after();
      AT2_T(t5);
      uint32_t o0[32], o1[32];
      ptx::tmem_ld_32x32(tmem_s + kAt2OCol, o0);
      ptx::tmem_ld_32x32(tmem_s + kAt2OCol + 32, o1);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tfree[group]);  // the slot may be overwritten by the QK^T of item j + 3
      AT2_T(t6);

      if (row < len) {
        const float inv = 1.0f / sum;
        uint4* orow = reinterpret_cast<uint4*>(out + (static_cast<long long>(wd.start) + row) * D + head * 64);
        auto pack2 = [&](uint32_t a, uint32_t b) {
          float x0, x1;
          at2_mul2(x0, x1, __uint_as_float(a), __uint_as_float(b), inv);
          return ptx::pack_bf16x2(x0, x1);
        };
#pragma unroll
        for (int u = 0; u < 4; ++u)
          orow[u] = make_uint4(pack2(o0[8 * u], o0[8 * u + 1]), pack2(o0[8 * u + 2], o0[8 * u + 3]), pack2(o0[8 * u + 4], o0[8 * u + 5]),
                               pack2(o0[8 * u + 6], o0[8 * u + 7]));
#pragma unroll
        for (int u = 0; u < 4; ++u)
          orow[4 + u] = make_uint4(pack2(o1[8 * u], o1[8 * u + 1]), pack2(o1[8 * u + 2], o1[8 * u + 3]), pack2(o1[8 * u + 4], o1[8 * u + 5]),
                                   pack2(o1[8 * u + 6], o1[8 * u + 7]));
      }
#if QASR_AT2_DBG
      if (dbg_me) {
        const long long t7 = clock64();
        AT2_ADD(0, t1 - t0); AT2_ADD(1, t2 - t1); AT2_ADD(2, t3 - t2); AT2_ADD(3, t4 - t3); AT2_ADD(4, t5 - t4);
        AT2_ADD(5, t6 - t5); AT2_ADD(6, t7 - t6); AT2_ADD(7, 1);
      }
#endif
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
