/* libqasr — C ABI of the B200-native Qwen3-ASR audio-encoding hot path (log-mel frontend + audio encoder).
 *
 * The reference (gabrimatic/qwen3-asr-mlx) has no FFI: its "operator interface" for this path is two
 * in-process Python callables.  Each entry point below names the reference call it replaces:
 *
 *   qasr_mel*            log_mel_spectrogram(audio)                 src/qwen3_asr_mlx/audio.py:238-278
 *   qasr_encode*         AudioEncoder.__call__(mel)                 src/qwen3_asr_mlx/encoder.py:235-323
 *   qasr_encode_audio*   the back-to-back call site                 src/qwen3_asr_mlx/model.py:331-335, 418-420
 *   qasr_encode_audio_hidden + qasr_project_rows   the same call split in front of the projector (encoder.py:235-317 | 319-321)
 *   qasr_create          AudioEncoder.__init__(config)              src/qwen3_asr_mlx/encoder.py:142-191
 *   qasr_set_weight      load_encoder_weights / model.load_weights  src/qwen3_asr_mlx/encoder.py:330-359
 *   qasr_count_tokens    AudioEncoder._conv_output_length + chunking src/qwen3_asr_mlx/encoder.py:197-207,258-268
 *   qasr_destroy         Qwen3ASR.close()                           src/qwen3_asr_mlx/model.py:261-269
 *   qasr_find_split_points   _find_split_points (long-audio feeder) src/qwen3_asr_mlx/model.py:454-513
 *   qasr_prepare_inputs  prepare_inputs (consumer of the output)    src/qwen3_asr_mlx/generate.py:20-81
 *   qasr_pack_audio      varlen packing of a batch of waveforms (no reference counterpart: it encodes one utterance per call)
 *   qasr_scatter_rows_to_peers   final gather of the data-parallel launcher over NVLink peer memory (no reference counterpart)
 *   (decoder prefill: include/qasr_decoder.h)
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * qasr_status; nothing throws or exits; qasr_last_error() gives the message for the last failure.
 * A handle is bound to one CUDA device and is not thread-safe (one handle per GPU per worker).
 * All kernels are launched on the caller-supplied stream (cudaStream_t passed as void*).
 * There is no CPU fallback: without an sm_100 device qasr_create fails.
 *
 * Batched layout ("varlen packing"): B utterances are concatenated.
 *   audio : float32[sample_offsets[B]]            utterance u = [sample_offsets[u], sample_offsets[u+1])
 *   mel   : float32[128 * frame_offsets[B]]       utterance u = row-major (128, T_u) block at 128*frame_offsets[u],
 *                                                 T_u = N_u / 160  (exactly the reference's (n_mels, T) array)
 *   emb   : [token_offsets[B], output_dim]        utterance u = rows [token_offsets[u], token_offsets[u+1])
 * Results equal a per-utterance loop over the reference (which processes one utterance at a time).
 */
#ifndef QASR_H_
#define QASR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qasr_handle qasr_handle;

typedef enum {
  QASR_OK = 0,
  QASR_ERR_INVALID = -1,      /* bad argument (maps to ValueError in the Python host) */
  QASR_ERR_CUDA = -2,         /* CUDA runtime / driver failure */
  QASR_ERR_UNSUPPORTED = -3,  /* configuration outside what the kernels implement, or no sm_100 device */
  QASR_ERR_STATE = -4,        /* call sequence error (e.g. weights not finalised) */
  QASR_ERR_NOMEM = -5
} qasr_status;

typedef enum { QASR_F32 = 0, QASR_BF16 = 1 } qasr_dtype;

/* Mirrors AudioEncoderConfig (src/qwen3_asr_mlx/config.py:14-29). */
typedef struct {
  int32_t d_model;                 /* 1024 */
  int32_t encoder_layers;          /* 24 */
  int32_t encoder_attention_heads; /* 16 (head_dim must be 64) */
  int32_t encoder_ffn_dim;         /* 4096 */
  int32_t num_mel_bins;            /* 128 (fixed) */
  int32_t max_source_positions;    /* 1500 */
  int32_t output_dim;              /* 2048 */
  int32_t n_window;                /* 50  -> chunk_size = 100 frames (fixed) */
  int32_t n_window_infer;          /* 800 -> attention window = 13 * (800 / 100) = 104 tokens */
  int32_t downsample_hidden_size;  /* 480 (fixed) */
} qasr_config;

typedef struct {
  uint64_t kernel_launches;  /* kernels launched by this handle since creation */
  uint64_t workspace_bytes;  /* device memory currently owned by the handle */
  uint64_t weight_bytes;
} qasr_stats;

/* Per-kernel-category timing, measured with CUDA events recorded on the launch stream. */
enum {
  QASR_PROF_MEL_LOGMEL = 0, QASR_PROF_MEL_NORM, QASR_PROF_CONV1, QASR_PROF_CONV2, QASR_PROF_CONV3, QASR_PROF_CONV_OUT,
  QASR_PROF_LAYERNORM, QASR_PROF_GEMM_QKV, QASR_PROF_ATTENTION, QASR_PROF_GEMM_OPROJ, QASR_PROF_GEMM_FC1,
  QASR_PROF_GEMM_FC2, QASR_PROF_GEMM_PROJ, QASR_PROF_CATEGORIES
};
typedef struct {
  double ms[QASR_PROF_CATEGORIES];         /* summed device time of the category's launches */
  double flops[QASR_PROF_CATEGORIES];      /* ALGORITHMIC flops of those launches (SURVEY.md 8d counting) */
  double bytes[QASR_PROF_CATEGORIES];      /* ALGORITHMIC bytes (memory-bound categories) */
  uint64_t launches[QASR_PROF_CATEGORIES];
} qasr_profile;

void qasr_default_config(qasr_config* cfg);

int qasr_create(int device, const qasr_config* cfg, qasr_handle** out);
void qasr_destroy(qasr_handle* h);
/* h may be NULL: returns the message of the last failed call on this thread that had no handle. */
const char* qasr_last_error(const qasr_handle* h);

/* name: reference parameter name without the "audio_tower." prefix, e.g. "layers.3.fc1.weight".
 * data: HOST pointer, dtype QASR_F32 or QASR_BF16, shape as stored by the reference
 * (Linear (out,in); Conv2d (O,kH,kW,I); LayerNorm (d,)). */
int qasr_set_weight(qasr_handle* h, const char* name, const void* data, int dtype, int ndim, const int64_t* shape);
/* Checks that every parameter was provided and builds the fused / re-laid-out device copies. */
int qasr_finalize_weights(qasr_handle* h);

/* T = N / 160 frames; fails with QASR_ERR_INVALID for N < 160 (the reference raises ValueError there). */
int qasr_count_frames(int64_t n_samples, int64_t* n_frames);
/* n = 13 * (T / 100) + f3(T % 100), f(L) = (L - 1) / 2 + 1. */
int qasr_count_tokens(const qasr_handle* h, int64_t n_frames, int64_t* n_tokens);

/* Optional: pre-size the workspace so that later calls with total_frames <= this allocate nothing. */
int qasr_reserve(qasr_handle* h, int64_t total_frames, int32_t batch);

/* ---- device-pointer entry points (inputs/outputs resident in HBM; offsets are HOST arrays) ----
 * emb_dev / mel_dev must be 16-byte aligned (the epilogues store tiles through TMA); cudaMalloc'd buffers always are. */
int qasr_mel(qasr_handle* h, const float* audio_dev, const int64_t* sample_offsets, int32_t batch, float* mel_dev,
             void* stream);
int qasr_encode(qasr_handle* h, const float* mel_dev, const int64_t* frame_offsets, int32_t batch, void* emb_dev,
                int out_dtype, int64_t* token_offsets_out, void* stream);
/* mel + encoder back to back (the mel lives in handle-owned scratch). */
int qasr_encode_audio(qasr_handle* h, const float* audio_dev, const int64_t* sample_offsets, int32_t batch,
                      void* emb_dev, int out_dtype, int64_t* token_offsets_out, void* stream);

/* The same call split in two for callers that ship the embeddings elsewhere while they are being produced (the
 * data-parallel launcher pushes each finished row block to its NVLink peers): qasr_encode_audio_hidden runs the mel, the
 * conv stem and the transformer layers (encoder.py:235-317) and leaves the final hidden states in the handle;
 * qasr_project_rows then applies ln_post -> proj1 -> GELU -> proj2 (encoder.py:319-321) to rows [row0, row0 + n_rows) of
 * that call, in any blocking and order, writing row `row0` at emb_dev.  Results are bit-identical to qasr_encode_audio.
 * The hidden states are valid until the next mel / encode call on the handle (QASR_ERR_STATE otherwise). */
int qasr_encode_audio_hidden(qasr_handle* h, const float* audio_dev, const int64_t* sample_offsets, int32_t batch,
                             int64_t* token_offsets_out, void* stream);
int qasr_project_rows(qasr_handle* h, int64_t row0, int64_t n_rows, void* emb_dev, int out_dtype, void* stream);

/* ---- host-pointer entry points (H2D / D2H inside the call, synchronous on return) ---- */
int qasr_mel_host(qasr_handle* h, const float* audio_host, const int64_t* sample_offsets, int32_t batch,
                  float* mel_host);
int qasr_encode_host(qasr_handle* h, const float* mel_host, const int64_t* frame_offsets, int32_t batch,
                     void* emb_host, int out_dtype, int64_t* token_offsets_out);
int qasr_encode_audio_host(qasr_handle* h, const float* audio_host, const int64_t* sample_offsets, int32_t batch,
                           void* emb_host, int out_dtype, int64_t* token_offsets_out);

/* Double-buffered host pipeline: slot 0/1 each own device staging buffers; the H2D copy, the kernels and the D2H copy
 * of a submission run on three streams, so that the copies of one submission overlap the kernels of the other slot's.
 * token_offsets_out is filled before the call returns; emb_host is valid after qasr_host_wait(slot). */
int qasr_encode_audio_host_async(qasr_handle* h, int32_t slot, const float* audio_host, const int64_t* sample_offsets,
                                 int32_t batch, void* emb_host, int out_dtype, int64_t* token_offsets_out);
int qasr_host_wait(qasr_handle* h, int32_t slot);

/* Prompt assembly, the consumer of this path (reference prepare_inputs, src/qwen3_asr_mlx/generate.py:20-81):
 * out[t] = embed_table[input_ids[t]], except at audio_pad_id positions, where out[t] = the next audio embedding row cast to
 * the table dtype.  input_ids is a HOST array; tables / embeddings / out are device pointers; out has n_ids x hidden elements of
 * table_dtype.  A pad count different from n_audio (and from 0) is QASR_ERR_INVALID, like the reference's ValueError. */
int qasr_prepare_inputs(qasr_handle* h, const int32_t* input_ids, int64_t n_ids, const void* embed_table_dev, int table_dtype,
                        int64_t vocab, int32_t hidden, const void* audio_emb_dev, int audio_dtype, int64_t n_audio,
                        int32_t audio_pad_id, void* out_dev, void* stream);

/* Long-audio splitter, the feeder of this path (reference _find_split_points, src/qwen3_asr_mlx/model.py:454-513, called at
 * model.py:400-403): float32 RMS energy of every frame_samples-long frame (bit-identical to the reference's
 * np.sqrt(np.mean(frame ** 2)), numpy pairwise summation order included) and, for every multiple of chunk_samples below
 * n_samples, the first lowest-energy frame within +-search_samples (cut = frame start; the boundary itself when the search
 * window is degenerate).  audio_dev is a DEVICE pointer; points_out is a HOST array of max_points entries; the call
 * synchronises `stream`.  energy_out_dev (n_samples / frame_samples floats, device) may be NULL.
 * Returns QASR_ERR_INVALID when more than max_points cuts are needed (n_points_out then holds the required count). */
int qasr_find_split_points(qasr_handle* h, const float* audio_dev, int64_t n_samples, int64_t chunk_samples, int64_t search_samples,
                           int32_t frame_samples, int64_t* points_out, int32_t max_points, int32_t* n_points_out,
                           float* energy_out_dev, void* stream);

/* Varlen packing of device-resident waveforms (the batched layout above has no reference counterpart: the reference handles
 * one utterance per call, model.py:239-250).  segments_dev: HOST array of `batch` DEVICE pointers, segment u holding
 * sample_offsets[u+1] - sample_offsets[u] floats; packed_dev: device buffer of sample_offsets[batch] floats.  One table upload
 * and one kernel, whatever the batch size. */
int qasr_pack_audio(qasr_handle* h, const float* const* segments_dev, const int64_t* sample_offsets, int32_t batch,
                    float* packed_dev, void* stream);

/* Final gather over NVLink peer memory (the only communication of the data-parallel path; the reference is
 * single-device and has no counterpart).  local_dev: this rank's packed rows [n_rows, row_bytes]; dst_rows_dev: DEVICE array,
 * final row index of every local row in the gathered matrix; peer_ptrs: HOST array of n_peers DEVICE pointers to every
 * rank's output buffer (P2P-mapped, e.g. torch symmetric memory), this rank's own included.  row_bytes % 16 == 0.
 * The caller synchronises the ranks before (buffers free) and after (writes landed). */
int qasr_scatter_rows_to_peers(const void* local_dev, int64_t n_rows, int32_t row_bytes, const int64_t* dst_rows_dev,
                               void* const* peer_ptrs, int32_t n_peers, void* stream);

/* ---- constant tables, as the library builds them (for parity tests) ---- */
int qasr_mel_filterbank(float* out_128x201);
int qasr_hann_window(float* out_400);
int qasr_positional_embedding(const qasr_handle* h, int32_t rows, float* out_rows_x_dmodel);

int qasr_get_stats(const qasr_handle* h, qasr_stats* out);

/* Enable/disable event bracketing of every launch (resets the accumulators). */
int qasr_set_profile(qasr_handle* h, int enabled);
/* Synchronises the device and returns the totals accumulated since qasr_set_profile. */
int qasr_get_profile(qasr_handle* h, qasr_profile* out);
const char* qasr_profile_name(int category);

/* ---- test hooks ---- */
/* When enabled, qasr_encode keeps copies of intermediate activations for qasr_debug_read. */
int qasr_set_debug(qasr_handle* h, int enabled);
/* what: "stem" (fp32 [n_tok, d_model] after conv stem + PE + packing),
 *       "layer0" (fp32 [n_tok, d_model] after the first transformer layer),
 *       "hidden" (fp32 [n_tok, d_model] after the last layer, before ln_post). */
int qasr_debug_read(qasr_handle* h, const char* what, float* host_out, size_t n_floats);
/* Stand-alone dense GEMM through the encoder's tcgen05 kernel: out[M,N] = a[M,K] * w[N,K]^T (+bias),
 * a/w bf16 HOST arrays, out fp32 HOST array.  mode: 0 store, 1 gelu, 2 residual (out += result, out is also an input); +16 selects the CTA-pair (cta_group::2) kernel. */
int qasr_test_gemm(int device, const uint16_t* a_bf16, const uint16_t* w_bf16, const float* bias, int32_t M,
                   int32_t N, int32_t K, int32_t mode, float* out);

/* The GELU the epilogues evaluate (math.cuh gelu_from_half, applied to x / 2 like the kernels do), elementwise on a HOST array:
 * lets a test bound its deviation from the reference's exact-erf nn.gelu (encoder.py:118,273-275,320) over every bf16 input. */
int qasr_test_gelu(int device, const float* x_host, int32_t n, float* out_host);

/* Micro-benchmark of the dense tcgen05 GEMM on device-resident pseudo-random bf16 operands (CUDA-event timed, back to back).
 * mode: 0 bf16 store, 1 bias+GELU bf16 store, 2 fp32 residual (TMA reduce-add), 3 accumulators discarded (main-loop ceiling);
 * +16 selects the CTA-pair kernel. */
int qasr_bench_gemm(int device, int32_t M, int32_t N, int32_t K, int32_t mode, int32_t iters, float* ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* QASR_H_ */
