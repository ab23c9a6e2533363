/* libqasr — C ABI of the decoder PREFILL, the stage that consumes the audio-encoding path's output
 * (SURVEY.md 8f rank 4).  Same library (libqasr.so), same conventions as qasr.h: plain pointers and sizes, 0 or a
 * negative qasr_status, nothing throws, no CPU fallback.
 *
 *   qasr_decoder_create      TextDecoder.__init__(config)                  src/qwen3_asr_mlx/decoder.py:203-221
 *   qasr_decoder_set_weight  load_decoder_weights / decoder.load_weights   src/qwen3_asr_mlx/decoder.py:257-291
 *   qasr_decoder_prefill     the prefill call of generate():               src/qwen3_asr_mlx/generate.py:266-275
 *                            logits = decoder(embeddings, cache=KVCache(), is_embeds=True), i.e.
 *                            TextDecoder.__call__ (decoder.py:223-253) over DecoderLayer (:181-200), Attention (:106-178:
 *                            q/k/v_proj, q_norm/k_norm, RoPE, KV-cache append, causal SDPA with GQA, o_proj) and MLP (:88-99)
 *   qasr_decoder_embed_table decoder.embed_tokens (decoder.py:218), the table qasr_prepare_inputs gathers from
 *
 * The token-by-token loop of generate() (generate.py:289-313) is NOT part of this library.
 *
 * Batched layout: B prompts are concatenated ("varlen packing"), prompt u = rows [seq_offsets[u], seq_offsets[u+1]) of the
 * [n, hidden] embedding matrix; attention is causal within a prompt and never crosses prompts; positions restart at 0 for
 * every prompt (the reference runs one prompt at a time with cache.offset = 0, generate.py:268-270).  Results equal a
 * per-prompt loop over the reference.
 *
 * KV cache written by the prefill (reference KVCache.update, decoder.py:31-61, keys/values of shape
 * (1, n_kv_heads, T, head_dim) per layer): k_cache / v_cache are [num_hidden_layers][n][num_key_value_heads * head_dim]
 * bf16, token-major; reference keys[layer][0, h, t, :] of prompt u == k_cache[layer][seq_offsets[u] + t][h * head_dim ...].
 * Keys are stored after q/k-norm and RoPE, exactly what the reference caches (decoder.py:165-169).
 */
#ifndef QASR_DECODER_H_
#define QASR_DECODER_H_

#include "qasr.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qasr_decoder qasr_decoder;

/* Mirrors TextDecoderConfig (src/qwen3_asr_mlx/config.py:61-76); mrope_section / rope_interleaved are ignored by the
 * reference decoder (plain nn.RoPE, decoder.py:128) and therefore absent. */
typedef struct {
  int32_t hidden_size;          /* 2048 (multiple of 128) */
  int32_t num_hidden_layers;    /* 28 */
  int32_t num_attention_heads;  /* 16 */
  int32_t num_key_value_heads;  /* 8 */
  int32_t head_dim;             /* 128 (fixed) */
  int32_t intermediate_size;    /* 6144 (multiple of 64) */
  int32_t vocab_size;           /* 151936 (multiple of 32) */
  float rms_norm_eps;           /* 1e-6 */
  float rope_theta;             /* 1e6 */
} qasr_decoder_config;

void qasr_decoder_default_config(qasr_decoder_config* cfg);
int qasr_decoder_create(int device, const qasr_decoder_config* cfg, qasr_decoder** out);
void qasr_decoder_destroy(qasr_decoder* d);
const char* qasr_decoder_last_error(const qasr_decoder* d);

/* Names are the reference's parameter paths after the "model." prefix is stripped (decoder.py:282-288):
 *   embed_tokens.weight (vocab, hidden)          norm.weight (hidden)
 *   layers.{i}.input_layernorm.weight            layers.{i}.post_attention_layernorm.weight
 *   layers.{i}.self_attn.{q,k,v,o}_proj.weight   layers.{i}.self_attn.{q,k}_norm.weight (head_dim)
 *   layers.{i}.mlp.{gate,up,down}_proj.weight    (Linear weights are (out, in), y = x W^T, no biases)
 * data: row-major array of dtype QASR_F32 / QASR_BF16 in HOST memory, or in DEVICE memory when QASR_DEVICE_PTR is or-ed
 * into dtype.  Matrices are stored as bf16 on the device, norm weights as fp32. */
#define QASR_DEVICE_PTR 0x100
int qasr_decoder_set_weight(qasr_decoder* d, const char* name, const void* data, int dtype, int ndim, const int64_t* shape);
/* Checks that every parameter has been supplied and builds the TMA descriptors. */
int qasr_decoder_finalize(qasr_decoder* d);

/* Device pointer and dtype (always QASR_BF16) of embed_tokens.weight, for qasr_prepare_inputs. */
int qasr_decoder_embed_table(const qasr_decoder* d, const void** table_dev, int* dtype);

/* Prefill.  embeds_dev: [n, hidden] of embed_dtype (QASR_F32 / QASR_BF16), n = seq_offsets[B]; seq_offsets: HOST, B + 1 entries.
 * Outputs (device pointers, any may be NULL):
 *   last_logits_dev  fp32 [B, vocab]   logits at the LAST position of every prompt (what generate() samples from, generate.py:278)
 *   all_logits_dev   fp32 [n, vocab]   logits at every position (the reference's full return value; parity tests, small n only)
 *   hidden_dev       fp32 [n, hidden]  residual stream after the last layer, before the final norm (parity tests)
 *   k_cache_dev, v_cache_dev           see the layout above */
int qasr_decoder_prefill(qasr_decoder* d, const void* embeds_dev, int embed_dtype, const int64_t* seq_offsets, int32_t batch,
                         float* last_logits_dev, float* all_logits_dev, float* hidden_dev, void* k_cache_dev, void* v_cache_dev,
                         void* stream);

/* Kernel launches issued so far, workspace and weight bytes held. */
int qasr_decoder_get_stats(const qasr_decoder* d, qasr_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* QASR_DECODER_H_ */
