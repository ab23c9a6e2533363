// Persistent warp-specialised bf16 GEMM for sm_100a:  D[M,N] = A[M,K] * W[N,K]^T  (+ epilogue)
//
//   warp 0      TMA producer      (one elected lane; cp.async.bulk.tensor -> 128B-swizzled smem ring)
//   warp 1      MMA issuer        (one elected lane; tcgen05.mma kind::f16, fp32 accumulators in TMEM)
//   warp 2      TMEM allocator
//   warp 3      idle
//   warps 4-11  epilogue          (tcgen05.ld -> registers -> swizzled smem transpose -> bias / GELU /
//                                  residual / scatter -> coalesced global accesses)
//
// Epilogue data path: tcgen05.ld hands each thread one accumulator ROW (32 lanes x 32 columns per
// warp).  Storing rows straight from registers makes every warp-level access touch 32 different
// cache lines; instead each warp transposes its 32 x 32 (or 32 x 16) block through a private,
// XOR-swizzled shared-memory buffer so that a warp instruction covers whole 64-128 byte row
// segments (4-16 lines per instruction instead of 32).
//
// Two accumulator stages (2 x 256 TMEM columns) let the epilogue of tile i overlap the
// main loop of tile i+1.  The same kernel serves every dense contraction on the encoder
// path (reference call sites: encoder.py:74-76,86 q/k/v/out_proj; :118-119 fc1/fc2;
// :279 conv_out; :320-321 proj1/proj2) and, with kAMode == A_CONV, the stride-2 3x3
// convolutions conv2d2/conv2d3 (encoder.py:274-275) as an implicit GEMM whose A operand is
// gathered by 4-D TMA boxes from a parity-split ("space-to-depth") activation layout, so
// that every filter tap is a unit-stride box and zero padding is a physical zero border.
#pragma once
#include "math.cuh"
#include "ptx.cuh"

namespace qasr {

enum AMode : int { A_ROWS = 0, A_CONV = 1 };
enum EpiMode : int {
  EPI_STORE_BF16 = 0,    // out_bf16[m,n] = acc + bias
  EPI_GELU_BF16 = 1,     // out_bf16[m,n] = gelu(2 (acc + bias)): W and bias are pre-scaled by 0.5 (math.cuh gelu_from_half)
  EPI_RESID_F32 = 2,     // out_f32[m,n] += acc + bias          (residual stream, in place)
  EPI_STORE_F32 = 3,     // out_f32[m,n] = acc + bias
  EPI_CONV_PLANES = 4,   // gelu(2 (acc+bias)) -> bf16, scattered into the next conv's parity planes (W, bias pre-scaled by 0.5)
  EPI_CONV_FLAT = 5,     // gelu(2 (acc+bias)) -> bf16, [(chunk*OW + ow)*OH + oh][ch]  (conv_out's A operand)
  EPI_CONVOUT_PACK = 6,  // out_f32[row_map[m], n] = acc + pe[m % period, n]  (PE add + strip padding + pack)
  EPI_GELU_F32 = 7,      // out_f32[m,n] = gelu(acc + bias)
  // micro-benchmark-only modes (qasr_bench_gemm); never instantiated on the product path
  EPI_DISCARD = 8,       // accumulators are read from TMEM and dropped (main-loop ceiling)
  EPI_MATH_ONLY = 9,     // bias + GELU + pack, nothing written
  // decoder prefill (reference decoder.py:88-99): the weight rows interleave gate and up projections in blocks of 32
  // (rows 64b..64b+31 = gate[32b..], rows 64b+32..64b+63 = up[32b..]); out_bf16[m, n/2] = silu(gate) * up
  EPI_SWIGLU_BF16 = 10
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 384;
constexpr int kNumEpiWarps = 8;
constexpr int kAccumStride = 256;  // TMEM columns per accumulator stage

struct GemmParams {
  int M, N, K;
  int num_m_tiles, num_n_tiles, num_k_blocks;
  void* out;
  long long ldo;  // elements
  const float* bias;
  // implicit-GEMM convolution geometry (kAMode == A_CONV)
  int conv_OW;             // output width (pixels per output row)
  int conv_OH;             // real output rows per chunk
  int conv_OHp;            // padded output rows per chunk (= OH + 1, the extra row is a dummy)
  int conv_rows_per_tile;  // output rows per M tile (rows_per_tile * OW <= 128)
  int conv_kc_per_tap;     // K blocks per filter tap (ceil(C / 64))
  int conv_tail_k16;       // K=16 MMA steps that carry data in the LAST K block of a tap (C % 64 / 16, 0 -> all 4)
  int conv_chunks;         // number of real chunks
  int a_tx_bytes;          // bytes one A-operand TMA box delivers (0 -> full stage: 128 rows x 128 B)
  // EPI_CONV_PLANES destination geometry
  int out_Hp, out_Wp;             // padded rows per chunk / padded width of destination planes
  long long out_plane_stride;     // elements between destination planes
  int out_C;                      // channels per pixel in the destination
  // EPI_CONVOUT_PACK
  const int* row_map;  // [M] destination row or -1
  const float* pe;     // [period, N]
  int pe_period;
  // 1: tiles are taken from the last M block downwards.  Consecutive kernels of a layer alternate the direction
  // ("serpentine"), so each starts on the rows its producer wrote last -- the ones still in L2.
  int reverse_tiles;
  // L2 eviction priorities (ptx::kL2Evict*; 0 = no hint) of the A-operand loads, the B-operand loads and the fp32 residual
  // reduction (A_ROWS kernels): activations that are read for the last time leave L2 first, the residual stream -- touched
  // by four of the seven kernels of a layer -- and the weights stay.
  unsigned long long a_policy, b_policy, out_policy;
};

// kCta = 2: a pair of CTAs (thread-block cluster of 2, one per SM of a TPC) computes a 256 x BLOCK_N
// tile with tcgen05.mma.cta_group::2; each CTA stages its own 128 A rows and HALF of the B rows.
template <int BLOCK_N, int kStages, int kCta = 1>
struct GemmSmem {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = (BLOCK_N / kCta) * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarrierBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kEpiStageBytes = 8 * 32 * 32 * 4;  // 8 epilogue warps x (32 rows x 128 B)
  static constexpr int kEpiRowBytes = 8 * 32 * 4;         // per-warp destination row table
  static constexpr int kTotal = kStages * kStageBytes + kEpiStageBytes + kEpiRowBytes + kBarrierBytes + 1024;
};

// 16-byte-unit XOR swizzle of a per-warp staging buffer with U units per row (U = 8, 4 or 2):
// conflict-free for both the row-owner writes and the transposed reads.
template <int U>
__device__ __forceinline__ int stage_unit(int row, int unit) {
  return row * U + (unit ^ ((row / (8 / U)) % U));
}

template <int kEpi>
constexpr bool epi_is_bf16() {
  return kEpi == EPI_STORE_BF16 || kEpi == EPI_GELU_BF16 || kEpi == EPI_CONV_PLANES || kEpi == EPI_CONV_FLAT ||
         kEpi == EPI_SWIGLU_BF16;
}

// kDbg = true (micro-benchmark only) makes the MMA-issuing thread account its waiting cycles into the buffer passed
// in p.row_map.  It is a COMPILE-TIME switch on purpose: even a run-time-gated clock64() in that loop cost 6 % of GEMM
// throughput in the full step (the single issuing thread is latency-critical).
template <int BLOCK_N, int kStages, int kAMode, int kEpi, int kCta = 1, bool kDbg = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_sm100(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_out, const GemmParams p) {
  using L = GemmSmem<BLOCK_N, kStages, kCta>;
  static_assert(kCta == 1 || kCta == 2, "cta_group must be 1 or 2");
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N <= 256, "invalid UMMA N");
  static_assert((L::kBBytes % 1024) == 0, "B stage must keep 1024-byte alignment");
  constexpr int kChunk = (BLOCK_N % 32 == 0) ? 32 : 16;  // epilogue column granularity
  constexpr int kNumChunks = BLOCK_N / kChunk;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * L::kABytes;
  uint8_t* epi_base = smem + kStages * L::kStageBytes;  // 1024-byte aligned (TMA-store swizzle atoms)
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_base + L::kEpiStageBytes + L::kEpiRowBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full_bar = bars + 2 * kStages;
  uint64_t* tmem_empty_bar = bars + 2 * kStages + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Tile scheduler: a "tile" is kCta vertically adjacent 128-row blocks x one BLOCK_N column block.
  const uint32_t cta_rank = (kCta == 2) ? ptx::cluster_ctarank() : 0u;
  const int sched_id = static_cast<int>(blockIdx.x) / kCta;
  const int sched_n = static_cast<int>(gridDim.x) / kCta;
  const int num_tiles = ((p.num_m_tiles + kCta - 1) / kCta) * p.num_n_tiles;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full_bar[i], 1);
      ptx::mbar_init(&tmem_empty_bar[i], kNumEpiWarps * kCta);  // the leader collects both CTAs' epilogue warps
    }
    ptx::fence_barrier_init();
  }
  if (warp_idx == 2) {
    ptx::tmem_alloc<kCta>(tmem_ptr_smem, 2 * kAccumStride);
    ptx::tmem_relinquish<kCta>();
  }
  ptx::tc_fence_before();
  if constexpr (kCta == 2) ptx::cluster_sync_all();  // peer barriers must be initialised before remote signals
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // programmatic dependent launch: the prologue above ran beside the previous kernel's tail; operands are valid from here
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();

  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // a conv box covers rows_per_tile * OW (< 128) pixels: the tail rows of the stage are never written
      const uint32_t a_tx = p.a_tx_bytes > 0 ? static_cast<uint32_t>(p.a_tx_bytes) : static_cast<uint32_t>(L::kABytes);
      for (int tile = sched_id; tile < num_tiles; tile += sched_n) {
        const int tsel = p.reverse_tiles ? num_tiles - 1 - tile : tile;
        const int m_blk = (tsel / p.num_n_tiles) * kCta + static_cast<int>(cta_rank);
        const int n_blk = tsel % p.num_n_tiles;
        const int b_row0 = n_blk * BLOCK_N + static_cast<int>(cta_rank) * (BLOCK_N / kCta);  // this CTA's B rows
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          void* sa = smem_a + stage * L::kABytes;
          void* sb = smem_b + stage * L::kBBytes;
          if constexpr (kCta == 1) {
            ptx::mbar_expect_tx(&full_bar[stage], a_tx + L::kBBytes);
            if constexpr (kAMode == A_ROWS) {
              if (p.a_policy) ptx::tma_load_2d_hint(sa, &tmap_a, &full_bar[stage], kb * kBlockK, m_blk * kBlockM, p.a_policy);
              else ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kBlockK, m_blk * kBlockM);
              if (p.b_policy) ptx::tma_load_2d_hint(sb, &tmap_b, &full_bar[stage], kb * kBlockK, b_row0, p.b_policy);
              else ptx::tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * kBlockK, b_row0);
            } else {
              const int tap = kb / p.conv_kc_per_tap;
              const int kc = kb - tap * p.conv_kc_per_tap;
              const int kh = tap / 3, kw = tap - kh * 3;
              const int plane = 2 * (kh != 1) + (kw != 1);
              const int row0 = m_blk * p.conv_rows_per_tile + (kh != 0);
              ptx::tma_load_4d(sa, &tmap_a, &full_bar[stage], kc * kBlockK, (kw != 0), row0, plane);
              ptx::tma_load_3d(sb, &tmap_b, &full_bar[stage], kc * kBlockK, tap, b_row0);
            }
          } else {
            // both CTAs' boxes complete on the LEADER's barrier, which expects the bytes of the pair
            const uint32_t lead_bar = ptx::mapa_u32(ptx::smem_u32(&full_bar[stage]), 0);
            if (cta_rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * (a_tx + L::kBBytes));
            if constexpr (kAMode == A_ROWS) {
              if (p.a_policy) ptx::tma_load_2d_cg2_hint(sa, &tmap_a, lead_bar, kb * kBlockK, m_blk * kBlockM, p.a_policy);
              else ptx::tma_load_2d_cg2(sa, &tmap_a, lead_bar, kb * kBlockK, m_blk * kBlockM);
              if (p.b_policy) ptx::tma_load_2d_cg2_hint(sb, &tmap_b, lead_bar, kb * kBlockK, b_row0, p.b_policy);
              else ptx::tma_load_2d_cg2(sb, &tmap_b, lead_bar, kb * kBlockK, b_row0);
            } else {
              const int tap = kb / p.conv_kc_per_tap;
              const int kc = kb - tap * p.conv_kc_per_tap;
              const int kh = tap / 3, kw = tap - kh * 3;
              const int plane = 2 * (kh != 1) + (kw != 1);
              const int row0 = m_blk * p.conv_rows_per_tile + (kh != 0);
              ptx::tma_load_4d_cg2(sa, &tmap_a, lead_bar, kc * kBlockK, (kw != 0), row0, plane);
              ptx::tma_load_3d_cg2(sb, &tmap_b, lead_bar, kc * kBlockK, tap, b_row0);
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && cta_rank == 0) {  // with cta_group::2 only the leader CTA issues the pair's MMAs
      constexpr uint32_t idesc = ptx::make_idesc_bf16(kBlockM * kCta, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      [[maybe_unused]] long long t_wait_acc = 0, t_wait_full = 0, t_tiles = 0, t0 = 0, t_begin = 0;
      if constexpr (kDbg) t_begin = clock64();
      for (int tile = sched_id; tile < num_tiles; tile += sched_n) {
        if constexpr (kDbg) t0 = clock64();
        ptx::mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
        if constexpr (kDbg) { t_wait_acc += clock64() - t0; ++t_tiles; }
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kAccumStride;
        int kc = 0;  // K block within the current filter tap (A_CONV)
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          if constexpr (kDbg) t0 = clock64();
          ptx::mbar_wait(&full_bar[stage], phase);
          if constexpr (kDbg) t_wait_full += clock64() - t0;
          ptx::tc_fence_after();
          const uint64_t adesc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem_a + stage * L::kABytes));
          const uint64_t bdesc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem_b + stage * L::kBBytes));
          // The last K block of a conv tap is zero-filled beyond C channels (480 = 7.5 x 64): MMAs over the
          // all-zero K=16 slices are skipped (they would add nothing and cost tensor-pipe time and power).
          int nk = kBlockK / kUmmaK;
          if constexpr (kAMode == A_CONV) {
            if (++kc == p.conv_kc_per_tap) { kc = 0; if (p.conv_tail_k16 > 0) nk = p.conv_tail_k16; }
          }
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance 16 elements (32 bytes) along K inside the swizzle row: +2 in the >>4 address field
            if (k < nk) ptx::umma_bf16_ss<kCta>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          // frees the smem slot (in both CTAs of a pair) when these MMAs retire
          if constexpr (kCta == 1) ptx::umma_commit(&empty_bar[stage]);
          else ptx::umma_commit_cg2(&empty_bar[stage], 0x3);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if constexpr (kCta == 1) ptx::umma_commit(&tmem_full_bar[as]);
        else ptx::umma_commit_cg2(&tmem_full_bar[as], 0x3);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
      if constexpr (kDbg) {  // the counter buffer travels in row_map (unused by the dense epilogues that are benchmarked)
        long long* d = reinterpret_cast<long long*>(const_cast<int*>(p.row_map)) + static_cast<long long>(blockIdx.x) * 4;
        d[0] = clock64() - t_begin; d[1] = t_wait_acc; d[2] = t_wait_full; d[3] = t_tiles;
      }
    }
  } else if (warp_idx >= 4) {
    // ------------------------------------------------------------------ epilogue
    constexpr bool kBf16Out = epi_is_bf16<kEpi>();
    // Dense fp32 outputs leave through the TMA: a plain tile store, or for the residual stream an
    // in-L2 reduction (x += tile), so the SM never reads the old residual values.
    constexpr bool kTmaOut = (kAMode == A_ROWS) && (kEpi == EPI_RESID_F32 || kEpi == EPI_STORE_F32);
    constexpr int U = kBf16Out ? kChunk / 8 : kChunk / 4;  // 16-byte units per staged row
    static_assert(kBf16Out || kChunk == 32, "fp32 epilogues assume 32-column chunks");
    const int ew = warp_idx - 4;
    const int quarter = warp_idx & 3;  // TMEM lane quarter this warp may access
    const int half = ew >> 2;          // interleaved column chunks
    const int r = quarter * 32 + lane;  // tile row == TMEM lane owned by this thread
    uint4* stg = reinterpret_cast<uint4*>(epi_base) + ew * (32 * 8);
    int* rowdst = reinterpret_cast<int*>(epi_base + L::kEpiStageBytes) + ew * 32;
    const long long row_stride = (kAMode == A_ROWS) ? p.ldo : static_cast<long long>(p.out_C);
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = sched_id; tile < num_tiles; tile += sched_n) {
      const int tsel = p.reverse_tiles ? num_tiles - 1 - tile : tile;
      const int m_blk = (tsel / p.num_n_tiles) * kCta + static_cast<int>(cta_rank);
      const int n_blk = tsel % p.num_n_tiles;
      const int n0 = n_blk * BLOCK_N;

      // Destination row (or pixel) index of this thread's accumulator row, -1 if it is not stored.
      int dst = -1;
      if constexpr (kAMode == A_ROWS) {
        const int m = m_blk * kBlockM + r;
        if (m < p.M) dst = (kEpi == EPI_CONVOUT_PACK) ? __ldg(p.row_map + m) : m;
      } else {
        // tile row r -> (padded output row, ow) -> (chunk, oh, ow)
        const int rr = r / p.conv_OW;
        const int ow = r - rr * p.conv_OW;
        if (rr < p.conv_rows_per_tile) {
          const int gg = m_blk * p.conv_rows_per_tile + rr;
          const int b = gg / p.conv_OHp;
          const int oh = gg - b * p.conv_OHp;
          if (b < p.conv_chunks && oh < p.conv_OH) {
            if constexpr (kEpi == EPI_CONV_PLANES) {
              const int plane = 2 * (oh & 1) + (ow & 1);
              dst = static_cast<int>(plane * (p.out_plane_stride / p.out_C)) +
                    ((b * p.out_Hp + (oh >> 1) + 1) * p.out_Wp + (ow >> 1) + 1);
            } else {  // EPI_CONV_FLAT
              dst = (b * p.conv_OW + ow) * p.conv_OH + oh;
            }
          }
        }
      }
      __syncwarp();  // previous tile's readers are done with rowdst / stg
      rowdst[lane] = dst;

      ptx::mbar_wait(&tmem_full_bar[as], aphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * kAccumStride;
      if constexpr (kEpi == EPI_SWIGLU_BF16) {
        static_assert(kEpi != EPI_SWIGLU_BF16 || (kChunk == 32 && kNumChunks % 2 == 0), "SwiGLU epilogue needs 64-column pairs");
#pragma unroll 1
        for (int pj = half; pj < kNumChunks / 2; pj += 2) {
          uint32_t ag[32], au[32];
          ptx::tmem_ld_32x32(taddr + (2 * pj) * 32, ag);
          ptx::tmem_ld_32x32(taddr + (2 * pj + 1) * 32, au);
          const int n = n0 + 2 * pj * 32;  // first absolute (interleaved) column of the gate chunk
          const bool n_ok = n < p.N;
          ptx::tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = silu_fast(__uint_as_float(ag[8 * u + i])) * __uint_as_float(au[8 * u + i]);
            uint4 q;
            q.x = ptx::pack_bf16x2(v[0], v[1]);
            q.y = ptx::pack_bf16x2(v[2], v[3]);
            q.z = ptx::pack_bf16x2(v[4], v[5]);
            q.w = ptx::pack_bf16x2(v[6], v[7]);
            stg[stage_unit<4>(lane, u)] = q;
          }
          __syncwarp();
          const int u = lane & 3, rsub = lane >> 2;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + rsub;
            const int d = rowdst[rr];
            const uint4 q = stg[stage_unit<4>(rr, u)];
            if (d >= 0 && n_ok)
              *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + d * row_stride + (n >> 1) + 8 * u) = q;
          }
          __syncwarp();
        }
      } else {
      uint32_t acc[kChunk];
      if (half < kNumChunks) {
        if constexpr (kChunk == 32) ptx::tmem_ld_32x32(taddr + half * kChunk, acc);
        else ptx::tmem_ld_32x16(taddr + half * kChunk, acc);
      }
#pragma unroll 1
      for (int j = half; j < kNumChunks; j += 2) {
        const int n = n0 + j * kChunk;  // first absolute column of this chunk
        const bool n_ok = n < p.N;
        ptx::tmem_ld_wait();
        if constexpr (kEpi == EPI_MATH_ONLY) {
          float v[kChunk];
#pragma unroll
          for (int i = 0; i < kChunk; ++i) v[i] = __uint_as_float(acc[i]);
          if (p.bias != nullptr && n_ok) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
            for (int i = 0; i < kChunk / 4; ++i) {
              const float4 t = __ldg(b4 + i);
              v[4 * i + 0] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
            }
          }
          uint32_t keep = 0;
#pragma unroll
          for (int i = 0; i < kChunk; i += 2) keep |= ptx::pack_bf16x2(gelu_fast(v[i]), gelu_fast(v[i + 1]));
          if (keep == 0x7fc12345u && n_ok) static_cast<float*>(p.out)[0] = 1.0f;  // never true: keeps the math alive
          if (j + 2 < kNumChunks) {
            if constexpr (kChunk == 32) ptx::tmem_ld_32x32(taddr + (j + 2) * kChunk, acc);
            else ptx::tmem_ld_32x16(taddr + (j + 2) * kChunk, acc);
          }
          continue;
        }
        if constexpr (kEpi == EPI_DISCARD) {
          uint32_t keep = 0;
#pragma unroll
          for (int i = 0; i < kChunk; ++i) keep |= acc[i];
          if (keep == 0x7fc12345u && n_ok) static_cast<float*>(p.out)[0] = 1.0f;  // never true: keeps the loads alive
          if (j + 2 < kNumChunks) {
            if constexpr (kChunk == 32) ptx::tmem_ld_32x32(taddr + (j + 2) * kChunk, acc);
            else ptx::tmem_ld_32x16(taddr + (j + 2) * kChunk, acc);
          }
          continue;
        }
        // ---- row-owner phase: registers -> swizzled staging buffer
        if constexpr (kBf16Out) {
          float v[kChunk];
#pragma unroll
          for (int i = 0; i < kChunk; ++i) v[i] = __uint_as_float(acc[i]);
          if (p.bias != nullptr && n_ok) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
            for (int i = 0; i < kChunk / 4; ++i) {
              const float4 t = __ldg(b4 + i);
              v[4 * i + 0] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
            }
          }
          if constexpr (kEpi != EPI_STORE_BF16) {
#pragma unroll
            for (int i = 0; i < kChunk; ++i) v[i] = gelu_from_half(v[i]);  // weights and bias carry the 0.5 (see math.cuh)
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            uint4 q;
            q.x = ptx::pack_bf16x2(v[8 * u + 0], v[8 * u + 1]);
            q.y = ptx::pack_bf16x2(v[8 * u + 2], v[8 * u + 3]);
            q.z = ptx::pack_bf16x2(v[8 * u + 4], v[8 * u + 5]);
            q.w = ptx::pack_bf16x2(v[8 * u + 6], v[8 * u + 7]);
            stg[stage_unit<U>(lane, u)] = q;
          }
        } else if constexpr (kTmaOut) {
          float v[kChunk];
#pragma unroll
          for (int i = 0; i < kChunk; ++i) v[i] = __uint_as_float(acc[i]);
          if (p.bias != nullptr && n_ok) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
            for (int i = 0; i < kChunk / 4; ++i) {
              const float4 t = __ldg(b4 + i);
              v[4 * i + 0] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
            }
          }
          // the previous chunk's bulk copy must have finished reading the staging buffer
          if (lane == 0) ptx::bulk_wait_group_read<0>();
          __syncwarp();
#pragma unroll
          for (int u = 0; u < U; ++u)
            stg[stage_unit<U>(lane, u)] = make_uint4(__float_as_uint(v[4 * u + 0]), __float_as_uint(v[4 * u + 1]),
                                                     __float_as_uint(v[4 * u + 2]), __float_as_uint(v[4 * u + 3]));
        } else {
#pragma unroll
          for (int u = 0; u < U; ++u)
            stg[stage_unit<U>(lane, u)] = make_uint4(acc[4 * u + 0], acc[4 * u + 1], acc[4 * u + 2], acc[4 * u + 3]);
        }
        // the accumulator registers are free again: fetch the next chunk while this one is written out
        if (j + 2 < kNumChunks) {
          if constexpr (kChunk == 32) ptx::tmem_ld_32x32(taddr + (j + 2) * kChunk, acc);
          else ptx::tmem_ld_32x16(taddr + (j + 2) * kChunk, acc);
        }
        if constexpr (kTmaOut) {
          // staging layout == SWIZZLE_128B box {32 cols, 32 rows}; rows beyond M are clipped by the tensor map
          ptx::fence_proxy_async_smem();
          __syncwarp();
          const int row0 = m_blk * kBlockM + quarter * 32;
          if (lane == 0 && n_ok && row0 < p.M) {
            if constexpr (kEpi == EPI_RESID_F32) {
              if (p.out_policy) ptx::tma_reduce_add_2d_hint(&tmap_out, stg, n, row0, p.out_policy);
              else ptx::tma_reduce_add_2d(&tmap_out, stg, n, row0);
            }
            else ptx::tma_store_2d(&tmap_out, stg, n, row0);
            ptx::bulk_commit_group();
          }
          continue;
        }
        __syncwarp();
        // ---- transposed phase: each instruction covers 32/U rows x (16 U) contiguous bytes
        constexpr int kRowsPerInstr = 32 / U;
        const int u = lane % U;
        const int rsub = lane / U;
        if constexpr (kBf16Out) {
#pragma unroll
          for (int i = 0; i < U; ++i) {
            const int rr = i * kRowsPerInstr + rsub;
            const int d = rowdst[rr];
            const uint4 q = stg[stage_unit<U>(rr, u)];
            if (d >= 0 && n_ok)
              *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + d * row_stride + n + 8 * u) = q;
          }
        } else {
          const int nc = n + 4 * u;  // this lane's 4 columns
          float4 add = make_float4(0.f, 0.f, 0.f, 0.f);
          if constexpr (kEpi != EPI_CONVOUT_PACK) {
            if (p.bias != nullptr && n_ok) add = __ldg(reinterpret_cast<const float4*>(p.bias + nc));
          }
          float4* dptr[U];
          float4 old[U];
#pragma unroll
          for (int i = 0; i < U; ++i) {
            const int rr = i * kRowsPerInstr + rsub;
            const int d = rowdst[rr];
            dptr[i] = (d >= 0 && n_ok) ? reinterpret_cast<float4*>(static_cast<float*>(p.out) + d * row_stride + nc) : nullptr;
            if constexpr (kEpi == EPI_RESID_F32) {
              if (dptr[i] != nullptr) old[i] = *dptr[i];
            }
          }
#pragma unroll
          for (int i = 0; i < U; ++i) {
            const int rr = i * kRowsPerInstr + rsub;
            const uint4 q = stg[stage_unit<U>(rr, u)];
            float4 v = make_float4(__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w));
            if constexpr (kEpi == EPI_CONVOUT_PACK) {
              const int m = m_blk * kBlockM + quarter * 32 + rr;
              if (dptr[i] != nullptr) add = __ldg(reinterpret_cast<const float4*>(p.pe + static_cast<long long>(m % p.pe_period) * p.N + nc));
            }
            v.x += add.x; v.y += add.y; v.z += add.z; v.w += add.w;
            if constexpr (kEpi == EPI_GELU_F32) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
            if constexpr (kEpi == EPI_RESID_F32) {
              if (dptr[i] != nullptr) { v.x += old[i].x; v.y += old[i].y; v.z += old[i].z; v.w += old[i].w; }
            }
            if (dptr[i] != nullptr) *dptr[i] = v;
          }
        }
        __syncwarp();  // staging buffer is rewritten by the next chunk
      }
      }  // generic (non-SwiGLU) epilogues
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kCta == 1) ptx::mbar_arrive(&tmem_empty_bar[as]);
        else ptx::mbar_arrive_cluster(&tmem_empty_bar[as], 0);  // the leader's MMA warp waits for both CTAs
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if constexpr (kTmaOut) {
      if (lane == 0) ptx::bulk_wait_group_read<0>();  // smem must outlive the last bulk copies
    }
  }

  ptx::tc_fence_before();
  if constexpr (kCta == 2) ptx::cluster_sync_all();  // no CTA may exit while its pair still signals / reads it
  else __syncthreads();
  if (warp_idx == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kCta>(tmem_base, 2 * kAccumStride);
  }
}

}  // namespace qasr
