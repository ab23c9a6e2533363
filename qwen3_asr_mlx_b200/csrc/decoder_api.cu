// libqasr: decoder prefill (include/qasr_decoder.h).  Handle, weights, workspace and the launch sequence of
// TextDecoder.__call__ for a varlen-packed batch of prompts.  Every compute step is a kernel from
// decoder_kernels.cuh or the tcgen05 GEMM of gemm_sm100.cuh; there is no CPU fallback.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <stdlib.h>

#include <algorithm>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../include/qasr_decoder.h"
#include "causal_attention_sm100.cuh"
#include "decoder_kernels.cuh"
#include "gemm_host.cuh"

using namespace qasr;

namespace {

thread_local std::string g_dec_error;
constexpr int kDecStages = 6;  // CTA-pair kernels: 32 KB / stage

struct DBuf {
  void* p = nullptr;
  size_t bytes = 0;
};
struct WMaps {
  CUtensorMap m2;  // CTA-pair kernel: each CTA loads half of the B rows (box = 128 rows)
};
struct DecLayer {
  __nv_bfloat16 *wqkv = nullptr, *wo = nullptr, *wgu = nullptr, *wd = nullptr;
  float *ln1 = nullptr, *ln2 = nullptr, *qn = nullptr, *kn = nullptr;
  WMaps tm_wqkv, tm_wo, tm_wgu, tm_wd;
};

}  // namespace

struct qasr_decoder {
  int device = 0;
  qasr_decoder_config cfg{};
  std::string err;
  bool finalized = false;
  qasr_stats stats{};
  bool attn_tc = true;  // QASR_DEC_ATTN_TC=0 selects the mma.sync causal-attention kernel instead of the tcgen05 one
  __nv_bfloat16* embed = nullptr;
  float* norm_w = nullptr;
  WMaps tm_embed;
  std::vector<DecLayer> layers;
  std::set<std::string> loaded;
  std::vector<void*> allocs;
  DBuf stage;  // device staging for host-supplied weights
  // workspace (grow-only)
  long long cap_tokens = 0, cap_batch = 0, cap_tiles = 0;
  DBuf x, xn, qkv, attn, hbuf, xl, d_pos, d_tiles, d_last;
  CUtensorMap tm_xn, tm_attn, tm_h, tm_xl, tm_qkv;
  uint8_t* pin = nullptr;
  size_t pin_bytes = 0;
  cudaEvent_t pin_event = nullptr;
  bool pin_pending = false;
};

namespace {

int dfail(qasr_decoder* d, int code, const std::string& msg) {
  if (d) d->err = msg;
  g_dec_error = msg;
  return code;
}
#define DCUDA(d, expr)                                                                              \
  do {                                                                                              \
    cudaError_t e__ = (expr);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return dfail(d, QASR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));          \
  } while (0)

int dalloc(qasr_decoder* d, DBuf& b, size_t bytes, bool zero) {
  if (bytes <= b.bytes) return QASR_OK;
  if (b.p) {
    DCUDA(d, cudaDeviceSynchronize());
    DCUDA(d, cudaFree(b.p));
    d->stats.workspace_bytes -= b.bytes;
    b.p = nullptr;
    b.bytes = 0;
  }
  cudaError_t e = cudaMalloc(&b.p, bytes);
  if (e != cudaSuccess) {
    b.p = nullptr;
    return dfail(d, QASR_ERR_NOMEM, "cudaMalloc(" + std::to_string(bytes) + " bytes): " + cudaGetErrorString(e));
  }
  b.bytes = bytes;
  d->stats.workspace_bytes += bytes;
  if (zero) DCUDA(d, cudaMemset(b.p, 0, bytes));
  return QASR_OK;
}

template <typename T>
int walloc(qasr_decoder* d, T** out, size_t count) {
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, count * sizeof(T));
  if (e != cudaSuccess) return dfail(d, QASR_ERR_NOMEM, std::string("cudaMalloc (weights): ") + cudaGetErrorString(e));
  d->allocs.push_back(p);
  d->stats.weight_bytes += count * sizeof(T);
  *out = static_cast<T*>(p);
  return QASR_OK;
}

int q_dim(const qasr_decoder_config& c) { return c.num_attention_heads * c.head_dim; }
int kv_dim(const qasr_decoder_config& c) { return c.num_key_value_heads * c.head_dim; }

int validate(qasr_decoder* d, const qasr_decoder_config& c) {
  if (c.head_dim != 128) return dfail(d, QASR_ERR_UNSUPPORTED, "head_dim must be 128");
  if (c.hidden_size < 128 || c.hidden_size > 4096 || c.hidden_size % 128 != 0)
    return dfail(d, QASR_ERR_UNSUPPORTED, "hidden_size must be a multiple of 128 in [128, 4096]");
  if (c.num_attention_heads <= 0 || c.num_key_value_heads <= 0 || c.num_attention_heads % c.num_key_value_heads != 0)
    return dfail(d, QASR_ERR_INVALID, "num_attention_heads must be a positive multiple of num_key_value_heads");
  if (c.intermediate_size < 64 || c.intermediate_size % 64 != 0) return dfail(d, QASR_ERR_UNSUPPORTED, "intermediate_size must be a multiple of 64");
  if (c.vocab_size < 32 || c.vocab_size % 32 != 0) return dfail(d, QASR_ERR_UNSUPPORTED, "vocab_size must be a multiple of 32");
  if (c.num_hidden_layers < 0 || c.num_hidden_layers > 512) return dfail(d, QASR_ERR_INVALID, "bad num_hidden_layers");
  if (!(c.rms_norm_eps > 0.0f) || !(c.rope_theta > 1.0f)) return dfail(d, QASR_ERR_INVALID, "bad rms_norm_eps / rope_theta");
  return QASR_OK;
}

template <int VPL>
void launch_rms(const float* x, const int* idx, const float* w, __nv_bfloat16* y, int rows, float eps, cudaStream_t st) {
  rmsnorm_bf16_kernel<VPL><<<(rows + 7) / 8, 256, 0, st>>>(x, idx, w, y, rows, eps);
}
int rmsnorm(qasr_decoder* d, const float* x, const int* idx, const float* w, __nv_bfloat16* y, int rows, cudaStream_t st) {
  const float eps = d->cfg.rms_norm_eps;
  switch (d->cfg.hidden_size / 128) {
#define RMS_CASE(V) case V: launch_rms<V>(x, idx, w, y, rows, eps, st); break;
    RMS_CASE(1) RMS_CASE(2) RMS_CASE(3) RMS_CASE(4) RMS_CASE(5) RMS_CASE(6) RMS_CASE(7) RMS_CASE(8)
    RMS_CASE(9) RMS_CASE(10) RMS_CASE(11) RMS_CASE(12) RMS_CASE(13) RMS_CASE(14) RMS_CASE(15) RMS_CASE(16)
    RMS_CASE(20) RMS_CASE(24) RMS_CASE(28) RMS_CASE(32)
#undef RMS_CASE
    default: return dfail(d, QASR_ERR_UNSUPPORTED, "hidden_size / 128 not instantiated");
  }
  DCUDA(d, cudaGetLastError());
  d->stats.kernel_launches++;
  return QASR_OK;
}

template <int EPI>
int gemm(qasr_decoder* d, const CUtensorMap& ta, const WMaps& tw, int M, int N, int K, void* out, long long ldo, cudaStream_t st) {
  GemmParams p = dense_params(M, N, K, out, ldo, nullptr);
  CUtensorMap tout;
  const CUtensorMap* toutp = nullptr;
  if (EPI == EPI_RESID_F32 || EPI == EPI_STORE_F32) {
    std::string e;
    if (!make_tmap_out_f32(&tout, out, M, N, ldo, &e)) return dfail(d, QASR_ERR_CUDA, e);
    toutp = &tout;
  }
  DCUDA(d, (launch_gemm<256, kDecStages, A_ROWS, EPI, 2>(ta, tw.m2, p, st, toutp)));
  d->stats.kernel_launches++;
  return QASR_OK;
}

int ensure_ws(qasr_decoder* d, long long n, long long batch, long long tiles) {
  const qasr_decoder_config& c = d->cfg;
  const int H = c.hidden_size, Q = q_dim(c), KV = kv_dim(c), I = c.intermediate_size;
  int rc;
  std::string e;
  if (n > d->cap_tokens) {
    if ((rc = dalloc(d, d->x, static_cast<size_t>(n) * H * 4, false))) return rc;
    if ((rc = dalloc(d, d->xn, static_cast<size_t>(n) * H * 2, true))) return rc;
    if ((rc = dalloc(d, d->qkv, static_cast<size_t>(n) * (Q + 2 * KV) * 2, true))) return rc;
    if ((rc = dalloc(d, d->attn, static_cast<size_t>(n) * Q * 2, true))) return rc;
    if ((rc = dalloc(d, d->hbuf, static_cast<size_t>(n) * I * 2, true))) return rc;
    if ((rc = dalloc(d, d->d_pos, static_cast<size_t>(n) * 4, false))) return rc;
    if (!make_tmap_rows(&d->tm_xn, d->xn.p, n, H, H, kBlockM, &e) || !make_tmap_rows(&d->tm_attn, d->attn.p, n, Q, Q, kBlockM, &e) ||
        !make_tmap_rows(&d->tm_h, d->hbuf.p, n, I, I, kBlockM, &e) || !make_tmap_rows(&d->tm_qkv, d->qkv.p, n, Q + 2 * KV, Q + 2 * KV, 128, &e))
      return dfail(d, QASR_ERR_CUDA, e);
    d->cap_tokens = n;
  }
  if (batch > d->cap_batch) {
    if ((rc = dalloc(d, d->xl, static_cast<size_t>(batch) * H * 2, true))) return rc;
    if ((rc = dalloc(d, d->d_last, static_cast<size_t>(batch) * 4, false))) return rc;
    if (!make_tmap_rows(&d->tm_xl, d->xl.p, batch, H, H, kBlockM, &e)) return dfail(d, QASR_ERR_CUDA, e);
    d->cap_batch = batch;
  }
  if (tiles > d->cap_tiles) {
    if ((rc = dalloc(d, d->d_tiles, static_cast<size_t>(tiles) * sizeof(AttnTile), false))) return rc;
    d->cap_tiles = tiles;
  }
  return QASR_OK;
}

// Destination of a named parameter.
struct Slot {
  void* dst = nullptr;   // device buffer (bf16 matrix or fp32 vector)
  bool is_matrix = true;
  long long rows = 0, cols = 0;
  int mode = 0;          // weight_rows_to_bf16_kernel row mapping
  long long row0 = 0;
};

bool find_slot(qasr_decoder* d, const std::string& name, Slot& s) {
  const qasr_decoder_config& c = d->cfg;
  const long long H = c.hidden_size, Q = q_dim(c), KV = kv_dim(c), I = c.intermediate_size;
  if (name == "embed_tokens.weight") { s = {d->embed, true, c.vocab_size, H, 0, 0}; return true; }
  if (name == "norm.weight") { s = {d->norm_w, false, H, 1, 0, 0}; return true; }
  if (name.rfind("layers.", 0) != 0) return false;
  const size_t dot = name.find('.', 7);
  if (dot == std::string::npos) return false;
  int li = -1;
  try { li = std::stoi(name.substr(7, dot - 7)); } catch (...) { return false; }
  if (li < 0 || li >= static_cast<int>(d->layers.size())) return false;
  DecLayer& L = d->layers[li];
  const std::string leaf = name.substr(dot + 1);
  if (leaf == "input_layernorm.weight") s = {L.ln1, false, H, 1, 0, 0};
  else if (leaf == "post_attention_layernorm.weight") s = {L.ln2, false, H, 1, 0, 0};
  else if (leaf == "self_attn.q_norm.weight") s = {L.qn, false, c.head_dim, 1, 0, 0};
  else if (leaf == "self_attn.k_norm.weight") s = {L.kn, false, c.head_dim, 1, 0, 0};
  else if (leaf == "self_attn.q_proj.weight") s = {L.wqkv, true, Q, H, 0, 0};
  else if (leaf == "self_attn.k_proj.weight") s = {L.wqkv, true, KV, H, 0, Q};
  else if (leaf == "self_attn.v_proj.weight") s = {L.wqkv, true, KV, H, 0, Q + KV};
  else if (leaf == "self_attn.o_proj.weight") s = {L.wo, true, H, Q, 0, 0};
  else if (leaf == "mlp.gate_proj.weight") s = {L.wgu, true, I, H, 1, 0};
  else if (leaf == "mlp.up_proj.weight") s = {L.wgu, true, I, H, 2, 0};
  else if (leaf == "mlp.down_proj.weight") s = {L.wd, true, H, I, 0, 0};
  else return false;
  return true;
}

}  // namespace

extern "C" {

void qasr_decoder_default_config(qasr_decoder_config* c) {
  if (!c) return;
  c->hidden_size = 2048; c->num_hidden_layers = 28; c->num_attention_heads = 16; c->num_key_value_heads = 8; c->head_dim = 128;
  c->intermediate_size = 6144; c->vocab_size = 151936; c->rms_norm_eps = 1e-6f; c->rope_theta = 1e6f;
}

const char* qasr_decoder_last_error(const qasr_decoder* d) { return d ? d->err.c_str() : g_dec_error.c_str(); }

int qasr_decoder_create(int device, const qasr_decoder_config* cfg, qasr_decoder** out) {
  if (!cfg || !out) return dfail(nullptr, QASR_ERR_INVALID, "qasr_decoder_create: null argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return dfail(nullptr, QASR_ERR_UNSUPPORTED, std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return dfail(nullptr, QASR_ERR_INVALID, "device index out of range");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return dfail(nullptr, QASR_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10) return dfail(nullptr, QASR_ERR_UNSUPPORTED, "libqasr is built for sm_100a only");
  qasr_decoder* d = new qasr_decoder();
  d->device = device;
  d->cfg = *cfg;
  int rc = validate(d, *cfg);
  if (rc) { g_dec_error = d->err; delete d; return rc; }
  if (cudaSetDevice(device) != cudaSuccess) { delete d; return dfail(nullptr, QASR_ERR_CUDA, "cudaSetDevice failed"); }
  const size_t H = cfg->hidden_size, Q = q_dim(*cfg), KV = kv_dim(*cfg), I = cfg->intermediate_size;
  d->layers.resize(cfg->num_hidden_layers);
  rc = walloc(d, &d->embed, static_cast<size_t>(cfg->vocab_size) * H);
  if (!rc) rc = walloc(d, &d->norm_w, H);
  for (auto& L : d->layers) {
    if (rc) break;
    if ((rc = walloc(d, &L.wqkv, (Q + 2 * KV) * H)) || (rc = walloc(d, &L.wo, H * Q)) || (rc = walloc(d, &L.wgu, 2 * I * H)) ||
        (rc = walloc(d, &L.wd, H * I)) || (rc = walloc(d, &L.ln1, H)) || (rc = walloc(d, &L.ln2, H)) ||
        (rc = walloc(d, &L.qn, static_cast<size_t>(cfg->head_dim))) || (rc = walloc(d, &L.kn, static_cast<size_t>(cfg->head_dim))))
      break;
  }
  if (const char* at = getenv("QASR_DEC_ATTN_TC")) d->attn_tc = atoi(at) != 0;
  if (!rc && cudaFuncSetAttribute(causal_attention_sm100, cudaFuncAttributeMaxDynamicSharedMemorySize, kCtSmemBytes) != cudaSuccess)
    rc = dfail(d, QASR_ERR_CUDA, "cudaFuncSetAttribute(causal_attention_sm100) failed");
  if (!rc && cudaEventCreateWithFlags(&d->pin_event, cudaEventDisableTiming) != cudaSuccess) rc = dfail(d, QASR_ERR_CUDA, "event creation failed");
  if (rc) { g_dec_error = d->err; qasr_decoder_destroy(d); return rc; }
  *out = d;
  return QASR_OK;
}

void qasr_decoder_destroy(qasr_decoder* d) {
  if (!d) return;
  cudaSetDevice(d->device);
  cudaDeviceSynchronize();
  for (void* p : d->allocs) cudaFree(p);
  for (DBuf* b : {&d->stage, &d->x, &d->xn, &d->qkv, &d->attn, &d->hbuf, &d->xl, &d->d_pos, &d->d_tiles, &d->d_last})
    if (b->p) cudaFree(b->p);
  if (d->pin) cudaFreeHost(d->pin);
  if (d->pin_event) cudaEventDestroy(d->pin_event);
  delete d;
}

int qasr_decoder_set_weight(qasr_decoder* d, const char* name, const void* data, int dtype, int ndim, const int64_t* shape) {
  if (!d || !name || !data || ndim < 1 || ndim > 2 || !shape) return dfail(d, QASR_ERR_INVALID, "qasr_decoder_set_weight: bad argument");
  if (d->finalized) return dfail(d, QASR_ERR_STATE, "weights already finalised");
  DCUDA(d, cudaSetDevice(d->device));
  const bool on_device = (dtype & QASR_DEVICE_PTR) != 0;
  const int dt = dtype & ~QASR_DEVICE_PTR;
  if (dt != QASR_F32 && dt != QASR_BF16) return dfail(d, QASR_ERR_INVALID, "bad dtype");
  Slot s;
  if (!find_slot(d, name, s)) return dfail(d, QASR_ERR_INVALID, std::string("unexpected parameter ") + name);
  const long long rows = shape[0], cols = ndim == 2 ? shape[1] : 1;
  if (rows != s.rows || cols != s.cols)
    return dfail(d, QASR_ERR_INVALID, std::string("parameter ") + name + " has shape (" + std::to_string(rows) + ", " + std::to_string(cols) +
                                          "), expected (" + std::to_string(s.rows) + ", " + std::to_string(s.cols) + ")");
  const size_t esz = dt == QASR_BF16 ? 2 : 4;
  const size_t bytes = static_cast<size_t>(rows) * cols * esz;
  const void* src = data;
  if (!on_device) {
    int rc;
    if ((rc = dalloc(d, d->stage, bytes, false))) return rc;
    DCUDA(d, cudaMemcpy(d->stage.p, data, bytes, cudaMemcpyHostToDevice));
    src = d->stage.p;
  }
  if (s.is_matrix) {
    auto* dst = static_cast<__nv_bfloat16*>(s.dst);
    if (dt == QASR_BF16) weight_rows_to_bf16_kernel<<<static_cast<unsigned>(rows), 256>>>(static_cast<const __nv_bfloat16*>(src), dst, cols, s.mode, s.row0);
    else weight_rows_to_bf16_kernel<<<static_cast<unsigned>(rows), 256>>>(static_cast<const float*>(src), dst, cols, s.mode, s.row0);
  } else {
    const long long n = rows * cols;
    if (dt == QASR_BF16) cast_rows_f32_kernel<<<static_cast<unsigned>((n + 1023) / 1024), 256>>>(static_cast<const __nv_bfloat16*>(src), static_cast<float*>(s.dst), n);
    else DCUDA(d, cudaMemcpy(s.dst, src, bytes, cudaMemcpyDeviceToDevice));
  }
  DCUDA(d, cudaGetLastError());
  DCUDA(d, cudaDeviceSynchronize());  // the staging buffer / caller's buffer may be reused right away
  d->loaded.insert(name);
  return QASR_OK;
}

int qasr_decoder_finalize(qasr_decoder* d) {
  if (!d) return dfail(nullptr, QASR_ERR_INVALID, "null decoder");
  if (d->finalized) return QASR_OK;
  DCUDA(d, cudaSetDevice(d->device));
  const qasr_decoder_config& c = d->cfg;
  std::vector<std::string> need = {"embed_tokens.weight", "norm.weight"};
  for (int i = 0; i < c.num_hidden_layers; ++i) {
    const std::string p = "layers." + std::to_string(i) + ".";
    for (const char* leaf : {"input_layernorm.weight", "post_attention_layernorm.weight", "self_attn.q_norm.weight", "self_attn.k_norm.weight",
                             "self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight", "self_attn.o_proj.weight",
                             "mlp.gate_proj.weight", "mlp.up_proj.weight", "mlp.down_proj.weight"})
      need.push_back(p + leaf);
  }
  for (const auto& n : need)
    if (!d->loaded.count(n)) return dfail(d, QASR_ERR_STATE, "parameter " + n + " (missing)");
  const uint64_t H = c.hidden_size, Q = q_dim(c), KV = kv_dim(c), I = c.intermediate_size;
  std::string e;
  bool ok = make_tmap_rows(&d->tm_embed.m2, d->embed, c.vocab_size, H, H, 128, &e);
  for (auto& L : d->layers)
    ok = ok && make_tmap_rows(&L.tm_wqkv.m2, L.wqkv, Q + 2 * KV, H, H, 128, &e) && make_tmap_rows(&L.tm_wo.m2, L.wo, H, Q, Q, 128, &e) &&
         make_tmap_rows(&L.tm_wgu.m2, L.wgu, 2 * I, H, H, 128, &e) && make_tmap_rows(&L.tm_wd.m2, L.wd, H, I, I, 128, &e);
  if (!ok) return dfail(d, QASR_ERR_CUDA, e);
  if (d->stage.p) {
    cudaFree(d->stage.p);
    d->stats.workspace_bytes -= d->stage.bytes;
    d->stage = DBuf{};
  }
  d->finalized = true;
  return QASR_OK;
}

int qasr_decoder_embed_table(const qasr_decoder* d, const void** table_dev, int* dtype) {
  if (!d || !table_dev || !dtype) return dfail(nullptr, QASR_ERR_INVALID, "qasr_decoder_embed_table: null argument");
  *table_dev = d->embed;
  *dtype = QASR_BF16;
  return QASR_OK;
}

int qasr_decoder_get_stats(const qasr_decoder* d, qasr_stats* out) {
  if (!d || !out) return dfail(nullptr, QASR_ERR_INVALID, "null argument");
  *out = d->stats;
  return QASR_OK;
}

int qasr_decoder_prefill(qasr_decoder* d, const void* embeds_dev, int embed_dtype, const int64_t* seq_offsets, int32_t B,
                         float* last_logits_dev, float* all_logits_dev, float* hidden_dev, void* k_cache_dev, void* v_cache_dev,
                         void* stream) {
  if (!d) return dfail(nullptr, QASR_ERR_INVALID, "null decoder");
  if (!d->finalized) return dfail(d, QASR_ERR_STATE, "weights not finalised (call qasr_decoder_finalize)");
  if (!embeds_dev || !seq_offsets || B <= 0) return dfail(d, QASR_ERR_INVALID, "qasr_decoder_prefill: bad argument");
  if (embed_dtype != QASR_F32 && embed_dtype != QASR_BF16) return dfail(d, QASR_ERR_INVALID, "bad embed_dtype");
  if (seq_offsets[0] != 0) return dfail(d, QASR_ERR_INVALID, "seq_offsets[0] must be 0");
  DCUDA(d, cudaSetDevice(d->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const qasr_decoder_config& c = d->cfg;
  const int H = c.hidden_size, Q = q_dim(c), KV = kv_dim(c), I = c.intermediate_size, V = c.vocab_size;
  const long long n = seq_offsets[B];
  if (n <= 0 || n > 0x7FFFFF00LL) return dfail(d, QASR_ERR_INVALID, "bad total token count");

  // ---- host-side tables: position of every token in its prompt, 64-query attention tiles, last-token rows
  std::vector<int> pos(static_cast<size_t>(n)), last(B);
  std::vector<AttnTile> tiles;
  for (int u = 0; u < B; ++u) {
    const long long a = seq_offsets[u], b = seq_offsets[u + 1];
    if (b <= a) return dfail(d, QASR_ERR_INVALID, "empty prompt in batch");
    for (long long t = a; t < b; ++t) pos[static_cast<size_t>(t)] = static_cast<int>(t - a);
    const int qt = d->attn_tc ? 128 : 64;  // query rows per attention tile
    for (long long q0 = 0; q0 < b - a; q0 += qt) tiles.push_back(AttnTile{static_cast<int>(a), static_cast<int>(b - a), static_cast<int>(q0)});
    last[u] = static_cast<int>(b - 1);
  }
  // longest KV loops first: the persistent CTAs of the tcgen05 kernel take items round-robin
  if (d->attn_tc) std::stable_sort(tiles.begin(), tiles.end(), [](const AttnTile& x, const AttnTile& y) { return x.q0 > y.q0; });
  int rc;
  if ((rc = ensure_ws(d, n, B, static_cast<long long>(tiles.size())))) return rc;
  const size_t bp = pos.size() * 4, bt = tiles.size() * sizeof(AttnTile), bl = last.size() * 4;
  if (d->pin_pending) {
    DCUDA(d, cudaEventSynchronize(d->pin_event));
    d->pin_pending = false;
  }
  if (bp + bt + bl + 64 > d->pin_bytes) {
    if (d->pin) cudaFreeHost(d->pin);
    d->pin = nullptr;
    d->pin_bytes = 0;
    void* pp = nullptr;
    DCUDA(d, cudaMallocHost(&pp, 2 * (bp + bt + bl) + 64));
    d->pin = static_cast<uint8_t*>(pp);
    d->pin_bytes = 2 * (bp + bt + bl) + 64;
  }
  const size_t o_t = (bp + 15) & ~static_cast<size_t>(15), o_l = (o_t + bt + 15) & ~static_cast<size_t>(15);
  memcpy(d->pin, pos.data(), bp);
  memcpy(d->pin + o_t, tiles.data(), bt);
  memcpy(d->pin + o_l, last.data(), bl);
  DCUDA(d, cudaMemcpyAsync(d->d_pos.p, d->pin, bp, cudaMemcpyHostToDevice, st));
  DCUDA(d, cudaMemcpyAsync(d->d_tiles.p, d->pin + o_t, bt, cudaMemcpyHostToDevice, st));
  DCUDA(d, cudaMemcpyAsync(d->d_last.p, d->pin + o_l, bl, cudaMemcpyHostToDevice, st));
  DCUDA(d, cudaEventRecord(d->pin_event, st));
  d->pin_pending = true;

  float* x = static_cast<float*>(d->x.p);
  auto* xn = static_cast<__nv_bfloat16*>(d->xn.p);
  auto* qkv = static_cast<__nv_bfloat16*>(d->qkv.p);
  auto* attn = static_cast<__nv_bfloat16*>(d->attn.p);
  auto* hb = static_cast<__nv_bfloat16*>(d->hbuf.p);
  const int ni = static_cast<int>(n);
  const long long nel = n * H;
  {
    const unsigned grid = static_cast<unsigned>((nel + 1023) / 1024);
    if (embed_dtype == QASR_BF16) cast_rows_f32_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(embeds_dev), x, nel);
    else cast_rows_f32_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(embeds_dev), x, nel);
    DCUDA(d, cudaGetLastError());
    d->stats.kernel_launches++;
  }
  const float scale_log2e = (1.0f / sqrtf(static_cast<float>(c.head_dim))) * 1.4426950408889634f;
  const float log2_theta = log2f(c.rope_theta);
  const int group = c.num_attention_heads / c.num_key_value_heads;
  for (size_t li = 0; li < d->layers.size(); ++li) {
    DecLayer& L = d->layers[li];
    if ((rc = rmsnorm(d, x, nullptr, L.ln1, xn, ni, st))) return rc;
    if ((rc = gemm<EPI_STORE_BF16>(d, d->tm_xn, L.tm_wqkv, ni, Q + 2 * KV, H, qkv, Q + 2 * KV, st))) return rc;
    __nv_bfloat16* kc = k_cache_dev ? static_cast<__nv_bfloat16*>(k_cache_dev) + static_cast<long long>(li) * n * KV : nullptr;
    __nv_bfloat16* vc = v_cache_dev ? static_cast<__nv_bfloat16*>(v_cache_dev) + static_cast<long long>(li) * n * KV : nullptr;
    qknorm_rope_kernel<<<(ni + 7) / 8, 256, 0, st>>>(qkv, static_cast<const int*>(d->d_pos.p), L.qn, L.kn, c.num_attention_heads,
                                                     c.num_key_value_heads, c.rms_norm_eps, log2_theta, kc, vc, ni);
    DCUDA(d, cudaGetLastError());
    if (d->attn_tc) {
      const long long items = static_cast<long long>(tiles.size()) * c.num_attention_heads;
      const int grid = static_cast<int>(items < gemm_num_sms() ? items : gemm_num_sms());
      causal_attention_sm100<<<grid, kCtThreads, kCtSmemBytes, st>>>(d->tm_qkv, static_cast<const AttnTile*>(d->d_tiles.p),
                                                                    static_cast<int>(tiles.size()), c.num_attention_heads, group, Q, Q + KV,
                                                                    attn, Q, scale_log2e);
    } else {
      causal_attention_kernel<<<dim3(static_cast<unsigned>(tiles.size()), c.num_attention_heads), kCaThreads, 0, st>>>(
          qkv, Q + 2 * KV, Q, Q + KV, group, static_cast<const AttnTile*>(d->d_tiles.p), attn, Q, scale_log2e);
    }
    DCUDA(d, cudaGetLastError());
    d->stats.kernel_launches += 2;
    if ((rc = gemm<EPI_RESID_F32>(d, d->tm_attn, L.tm_wo, ni, H, Q, x, H, st))) return rc;
    if ((rc = rmsnorm(d, x, nullptr, L.ln2, xn, ni, st))) return rc;
    if ((rc = gemm<EPI_SWIGLU_BF16>(d, d->tm_xn, L.tm_wgu, ni, 2 * I, H, hb, I, st))) return rc;
    if ((rc = gemm<EPI_RESID_F32>(d, d->tm_h, L.tm_wd, ni, H, I, x, H, st))) return rc;
  }
  if (hidden_dev) DCUDA(d, cudaMemcpyAsync(hidden_dev, x, static_cast<size_t>(nel) * 4, cudaMemcpyDeviceToDevice, st));
  // ---- final norm + tied lm_head (decoder.py:251-253); generate() reads the last position only (generate.py:278)
  if (last_logits_dev) {
    if ((rc = rmsnorm(d, x, static_cast<const int*>(d->d_last.p), d->norm_w, static_cast<__nv_bfloat16*>(d->xl.p), B, st))) return rc;
    if ((rc = gemm<EPI_STORE_F32>(d, d->tm_xl, d->tm_embed, B, V, H, last_logits_dev, V, st))) return rc;
  }
  if (all_logits_dev) {
    if ((rc = rmsnorm(d, x, nullptr, d->norm_w, xn, ni, st))) return rc;
    if ((rc = gemm<EPI_STORE_F32>(d, d->tm_xn, d->tm_embed, ni, V, H, all_logits_dev, V, st))) return rc;
  }
  return QASR_OK;
}

}  // extern "C"
