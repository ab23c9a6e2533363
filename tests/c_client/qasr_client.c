/* A plain-C client of libqasr.so: proof that the drop-in boundary is the C ABI of include/qasr.h and nothing else
 * (no Python, no torch, no C++ types).  It does what the reference's call site does at model.py:331-335 --
 *   mel = log_mel_spectrogram(samples); out = self._encoder(mel); mx.eval(out)
 * -- through qasr_create / qasr_set_weight / qasr_finalize_weights / qasr_encode_audio_host.
 *
 *   qasr_client <weights.bin> <audio.f32> <out.f32>
 *
 * weights.bin (written by tests/test_c_abi.py): int32 config[10]; int32 n_params; then per parameter
 *   int32 name_len, name bytes, int32 ndim, int64 shape[ndim], float32 data[prod(shape)].
 * audio.f32: int32 batch, int64 sample_offsets[batch + 1], float32 samples[sample_offsets[batch]].
 * out.f32:   int64 token_offsets[batch + 1], float32 embeddings[token_offsets[batch] * output_dim].
 * Exit code 0 on success; on any failure prints qasr_last_error() and returns 1. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qasr.h"

static int die(const qasr_handle* h, const char* what) {
  fprintf(stderr, "qasr_client: %s: %s\n", what, qasr_last_error(h));
  return 1;
}

int main(int argc, char** argv) {
  if (argc != 4) { fprintf(stderr, "usage: qasr_client weights.bin audio.f32 out.f32\n"); return 2; }
  FILE* fw = fopen(argv[1], "rb");
  FILE* fa = fopen(argv[2], "rb");
  if (!fw || !fa) { fprintf(stderr, "qasr_client: cannot open inputs\n"); return 2; }

  qasr_config cfg;
  int32_t raw_cfg[10], n_params = 0;
  if (fread(raw_cfg, sizeof(int32_t), 10, fw) != 10 || fread(&n_params, sizeof(int32_t), 1, fw) != 1) return 2;
  memcpy(&cfg, raw_cfg, sizeof(cfg)); /* qasr_config is ten int32 in declaration order */

  qasr_handle* h = NULL;
  if (qasr_create(0, &cfg, &h) != QASR_OK) return die(NULL, "qasr_create");
  for (int32_t i = 0; i < n_params; ++i) {
    int32_t name_len = 0, ndim = 0;
    char name[256];
    int64_t shape[8], count = 1;
    if (fread(&name_len, 4, 1, fw) != 1 || name_len <= 0 || name_len >= (int32_t)sizeof(name)) return 2;
    if (fread(name, 1, (size_t)name_len, fw) != (size_t)name_len) return 2;
    name[name_len] = '\0';
    if (fread(&ndim, 4, 1, fw) != 1 || ndim < 1 || ndim > 8) return 2;
    if (fread(shape, 8, (size_t)ndim, fw) != (size_t)ndim) return 2;
    for (int32_t d = 0; d < ndim; ++d) count *= shape[d];
    float* data = (float*)malloc((size_t)count * sizeof(float));
    if (!data || fread(data, sizeof(float), (size_t)count, fw) != (size_t)count) return 2;
    if (qasr_set_weight(h, name, data, QASR_F32, ndim, shape) != QASR_OK) return die(h, name);
    free(data);
  }
  fclose(fw);
  if (qasr_finalize_weights(h) != QASR_OK) return die(h, "qasr_finalize_weights");

  int32_t batch = 0;
  if (fread(&batch, 4, 1, fa) != 1 || batch <= 0) return 2;
  int64_t* soffs = (int64_t*)malloc((size_t)(batch + 1) * sizeof(int64_t));
  int64_t* toffs = (int64_t*)calloc((size_t)(batch + 1), sizeof(int64_t));
  if (!soffs || !toffs || fread(soffs, 8, (size_t)(batch + 1), fa) != (size_t)(batch + 1)) return 2;
  float* audio = (float*)malloc((size_t)soffs[batch] * sizeof(float));
  if (!audio || fread(audio, sizeof(float), (size_t)soffs[batch], fa) != (size_t)soffs[batch]) return 2;
  fclose(fa);

  int64_t n_tokens = 0; /* output size from the token rule, before any GPU work (encoder.py:197-207,258-268) */
  for (int32_t u = 0; u < batch; ++u) {
    int64_t frames = 0, tok = 0;
    if (qasr_count_frames(soffs[u + 1] - soffs[u], &frames) != QASR_OK) return die(NULL, "qasr_count_frames");
    if (qasr_count_tokens(h, frames, &tok) != QASR_OK) return die(h, "qasr_count_tokens");
    n_tokens += tok;
  }
  float* emb = (float*)malloc((size_t)n_tokens * (size_t)cfg.output_dim * sizeof(float));
  if (!emb) return 2;
  if (qasr_encode_audio_host(h, audio, soffs, batch, emb, QASR_F32, toffs) != QASR_OK) return die(h, "qasr_encode_audio_host");
  if (toffs[batch] != n_tokens) { fprintf(stderr, "qasr_client: token rule mismatch\n"); return 1; }

  /* error convention: fewer than 160 samples is QASR_ERR_INVALID, like the reference's ValueError (audio.py:275) */
  int64_t bad[2] = {0, 100};
  if (qasr_encode_audio_host(h, audio, bad, 1, emb, QASR_F32, toffs) != QASR_ERR_INVALID) {
    fprintf(stderr, "qasr_client: a 100-sample utterance must be rejected with QASR_ERR_INVALID\n");
    return 1;
  }
  /* recompute after the failed call: the handle stays usable */
  if (qasr_encode_audio_host(h, audio, soffs, batch, emb, QASR_F32, toffs) != QASR_OK) return die(h, "second encode");

  qasr_stats st;
  if (qasr_get_stats(h, &st) != QASR_OK) return die(h, "qasr_get_stats");
  FILE* fo = fopen(argv[3], "wb");
  if (!fo) return 2;
  fwrite(toffs, 8, (size_t)(batch + 1), fo);
  fwrite(emb, sizeof(float), (size_t)n_tokens * (size_t)cfg.output_dim, fo);
  fclose(fo);
  printf("qasr_client ok: %d utterances, %lld tokens, %llu kernel launches\n", (int)batch, (long long)n_tokens,
         (unsigned long long)st.kernel_launches);
  qasr_destroy(h);
  free(audio); free(emb); free(soffs); free(toffs);
  return 0;
}
