// Host side of the tcgen05 GEMM: TMA tensor-map construction (driver entry point resolved at
// run time, so libqasr.so does not link libcuda) and typed launch wrappers.
#pragma once
#include <stdlib.h>
#include <string>

#include "gemm_sm100.cuh"

namespace qasr {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled get_tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

// Tiled tensor map, dims innermost-first, strides in BYTES for dims 1..rank-1, zero OOB fill.
inline bool make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, CUtensorMapDataType dtype, CUtensorMapSwizzle swizzle, std::string* err) {
  PFN_tmapEncodeTiled enc = get_tmap_encoder();
  if (enc == nullptr) {
    if (err) *err = "cuTensorMapEncodeTiled entry point not found";
    return false;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(out, dtype, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r));
    return false;
  }
  return true;
}

inline bool make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, std::string* err) {
  return make_tmap(out, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                   CU_TENSOR_MAP_SWIZZLE_128B, err);
}

// Row-major fp32 [rows, cols] OUTPUT matrix for the epilogue's TMA store / reduce: box {32 cols = 128 B, 32 rows}.
inline bool make_tmap_out_f32(CUtensorMap* out, void* base, uint64_t rows, uint64_t cols, uint64_t ld, std::string* err) {
  uint64_t dims[2] = {cols, rows};
  uint64_t str[1] = {ld * 4};
  uint32_t box[2] = {32, 32};
  return make_tmap(out, base, 2, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B, err);
}

// Row-major [rows, cols] bf16 matrix (cols contiguous, leading dimension ld elements); box {64, box_rows}.
inline bool make_tmap_rows(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                           uint32_t box_rows, std::string* err) {
  uint64_t dims[2] = {cols, rows};
  uint64_t str[1] = {ld * 2};
  uint32_t box[2] = {static_cast<uint32_t>(kBlockK), box_rows};
  return make_tmap_bf16(out, base, 2, dims, str, box, err);
}

// Parity-plane activation tensor [4 planes][rows_total][Wp][C] bf16; box {64, OW, rows_per_tile, 1}.
inline bool make_tmap_conv_act(CUtensorMap* out, const void* base, int C, int Wp, uint64_t rows_total, int OW,
                               int rows_per_tile, std::string* err) {
  uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(Wp), rows_total, 4};
  uint64_t str[3] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(C) * Wp * 2,
                     static_cast<uint64_t>(C) * Wp * rows_total * 2};
  uint32_t box[4] = {static_cast<uint32_t>(kBlockK), static_cast<uint32_t>(OW),
                     static_cast<uint32_t>(rows_per_tile), 1};
  return make_tmap_bf16(out, base, 4, dims, str, box, err);
}

// Convolution weights [O][9 taps][C] bf16 (the reference's (O,kH,kW,I) layout flattened); box {64, 1, block_n}.
inline bool make_tmap_conv_w(CUtensorMap* out, const void* base, int C, int O, uint32_t block_n, std::string* err) {
  uint64_t dims[3] = {static_cast<uint64_t>(C), 9, static_cast<uint64_t>(O)};
  uint64_t str[2] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(C) * 9 * 2};
  uint32_t box[3] = {static_cast<uint32_t>(kBlockK), 1, block_n};
  return make_tmap_bf16(out, base, 3, dims, str, box, err);
}

inline int gemm_num_sms() {
  static int n[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& v = n[dev & 63];
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
  }
  return v;
}

// Programmatic dependent launch (QASR_PDL, default on): kernels that call ptx::grid_dep_wait() are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so their prologue overlaps the tail of the previous kernel in the stream
// (also inside captured graphs, where the edge becomes a programmatic dependency).
inline bool& pdl_flag() {
  static bool on = !(getenv("QASR_PDL") && atoi(getenv("QASR_PDL")) == 0);
  return on;
}
inline bool pdl_enabled() { return pdl_flag(); }
// re-read QASR_PDL (every qasr_create does: the switch is process-wide because launch_gemm is shared with the decoder)
inline void pdl_refresh_from_env() { pdl_flag() = !(getenv("QASR_PDL") && atoi(getenv("QASR_PDL")) == 0); }
// <<<grid, block, smem, stream>>> with the PDL attribute; only for kernels that execute grid_dep_wait() before they touch
// global data.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

template <int BLOCK_N, int kStages, int kAMode, int kEpi, int kCta = 1, bool kDbg = false>
inline cudaError_t launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, GemmParams p, cudaStream_t stream,
                               const CUtensorMap* tout = nullptr, int max_ctas = 0) {
  using L = GemmSmem<BLOCK_N, kStages, kCta>;
  auto kern = gemm_bf16_sm100<BLOCK_N, kStages, kAMode, kEpi, kCta, kDbg>;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {  // the attribute is per device
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return e;
    configured[dev & 63] = true;
  }
  p.num_n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int tiles = ((p.num_m_tiles + kCta - 1) / kCta) * p.num_n_tiles;  // tiles of kCta x 128 rows
  if (tiles <= 0 || p.num_k_blocks <= 0) return cudaSuccess;
  int grid = gemm_num_sms();
  if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
  grid = (grid / kCta) * kCta;
  if (tiles * kCta < grid) grid = tiles * kCta;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if constexpr (kCta == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = kCta;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, ta, tb, tout ? *tout : ta, p);
}

// Plain dense GEMM parameter block: A [M,K] row-major, W [N,K] row-major.
inline GemmParams dense_params(int M, int N, int K, void* out, long long ldo, const float* bias) {
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.num_m_tiles = (M + kBlockM - 1) / kBlockM;
  p.num_k_blocks = (K + kBlockK - 1) / kBlockK;
  p.out = out; p.ldo = ldo; p.bias = bias;
  return p;
}

}  // namespace qasr
