/* asr_b200.h — C ABI of the B200-native per-chunk compute path of Naiscorp-Robotics/ASR-streaming's
 * "lightspeech" streaming decoder (PCM -> log-mel -> Emformer chunk forward with per-session K/V caches ->
 * CTC log-softmax -> greedy / prefix-beam decode), batched across many websocket sessions.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The reference has no FFI of its own
 * (it is pure Python); each entry point below names the reference Python interface it replaces, with file:line
 * relative to the reference repo root.  The Python host mirror (asr_streaming_b200.recognition.LightningASR)
 * binds these with ctypes; INTEGRATION.md shows the stub a maintainer adds on the reference side.
 *
 * Conventions: every function returns 0 on success, non-zero on failure, and never throws; the message of the
 * last failure on the calling thread is asr_last_error().  One engine per GPU; calls on one engine are serialised
 * by an internal mutex.  Device memory is owned by the engine; host buffers are owned by the caller and are only
 * read / written during the call.
 */
#ifndef ASR_B200_H_
#define ASR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASR_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define ASR_API __attribute__((visibility("default")))
#else
#define ASR_API
#endif

typedef struct AsrEngine AsrEngine;

/* Precision of the dense projections.  FAST: bf16 operands, fp32 accumulate (north-star tolerance: max-abs <= 1e-2
 * on log-probs).  EXACT: split-bf16 (hi + lo) operands, three tcgen05 products per GEMM, fp32 K/V cache
 * (max-abs ~1e-5; used for the token-exact gate against the reference). */
enum { ASR_PRECISION_FAST = 0, ASR_PRECISION_EXACT = 1 };
enum { ASR_PCM_I16 = 0, ASR_PCM_F32 = 1 };
enum { ASR_FBANK_MELSPEC128 = 0, ASR_FBANK_KALDI80 = 1 };

/* Geometry = AudioConfig (streaming_decoder/utils.py:9-23, config/asr-online.yaml:112-118) + model hyper-parameters
 * (lightspeech/modules/encoder.py:73-117, lightspeech/models/recognition.py:207-217). */
typedef struct AsrConfig {
  int32_t abi_version;     /* ASR_B200_ABI_VERSION */
  int32_t sample_rate;     /* 16000 */
  int32_t hop;             /* 160  samples */
  int32_t n_fft;           /* 800 */
  int32_t win;             /* 400 */
  int32_t n_mels;          /* 128 */
  int32_t segment_size;    /* 64 frames  (32 in the low-latency mode) */
  int32_t context_size;    /* 16 frames */
  int32_t bias;            /* 4 frames */
  int32_t stride;          /* 4 */
  int32_t d_model;         /* 512 */
  int32_t n_heads;         /* 8 */
  int32_t ffn_dim;         /* 2048 */
  int32_t n_layers;        /* 20 */
  int32_t left_context;    /* 32 rows */
  int32_t ctc_hidden;      /* 512 */
  int32_t vocab;           /* 804 */
  int32_t precision;       /* ASR_PRECISION_* */
  int32_t max_sessions;    /* K/V slots to allocate */
  int32_t max_batch;       /* max stream-chunks per step */
} AsrConfig;

typedef struct AsrStepOut {     /* all pointers nullable, host memory, n = streams in the step, S = segment rows (16) */
  int32_t* argmax_ids;          /* [n*S]  per-frame argmax of the log-probs (recognition.py:36)                       */
  int32_t* new_tokens;          /* [n*S]  ids appended this chunk after unique_consecutive + blank drop (:44-45)     */
  int32_t* n_new;               /* [n]                                                                                */
  int32_t* blank_frames;        /* [n]    frames since the last id > 1, or all frames of the segment if none (:38-43) */
  int32_t* has_token;           /* [n]                                                                                */
  float* logprobs;              /* [n*S*vocab]  the reference's `emission` (recognition.py:203-204)                   */
  /* CTC prefix beam search (only when asr_set_beam enabled it): best hypothesis of the utterance so far               */
  int32_t* beam_tokens;         /* [n*ASR_BEAM_MAX_LEN]                                                               */
  int32_t* beam_len;            /* [n]                                                                                */
  float* beam_score;            /* [n]  log P(best prefix)                                                            */
} AsrStepOut;

#define ASR_BEAM_MAX_LEN 256

typedef struct AsrStats {
  uint64_t steps;               /* asr_step calls                                  */
  uint64_t stream_chunks;       /* stream-chunks processed                         */
  uint64_t kernel_launches;     /* CUDA kernels launched by this engine            */
  double step_ms_p50, step_ms_p99, step_ms_max, step_ms_mean;   /* host wall time of asr_step, last <= 4096 steps */
} AsrStats;

ASR_API const char* asr_last_error(void);
ASR_API int asr_abi_version(void);

/* Fills *cfg with the canonical geometry of SURVEY.md §8a (chunk_size 16).  low_latency != 0 -> segment_size 32. */
ASR_API int asr_default_config(AsrConfig* cfg, int low_latency);
/* Number of fp32 values asr_engine_create expects in `weights` for this config (layout: see weights.py / DESIGN.md). */
ASR_API int asr_weights_count(const AsrConfig* cfg, uint64_t* n_floats);
/* Samples per stream-chunk (AudioConfig.chunk_length, utils.py:22) and output rows per chunk. */
ASR_API int asr_chunk_geometry(const AsrConfig* cfg, int32_t* chunk_length, int32_t* segment_length, int32_t* seg_rows);

/* Replaces LightningASR.__init__ / _load_checkpoint (recognition.py:137-159): weights are the fp32 tensors of the
 * checkpoint's state_dict["encoder"|"decoder"], packed in the documented order. */
ASR_API int asr_engine_create(const AsrConfig* cfg, const float* weights, uint64_t n_floats, int device, AsrEngine** out);
ASR_API int asr_engine_destroy(AsrEngine* e);

/* Replaces LightningASR.init_state (recognition.py:207-217) and `stream.state = state_init`
 * (streaming_server.py:324-326, :530): a session owns one K/V ring slot + greedy carry. */
ASR_API int asr_session_open(AsrEngine* e, int32_t* slot_out);
ASR_API int asr_session_reset(AsrEngine* e, int32_t slot);           /* endpoint: state := init, emission := []  (:514-515, :530) */
ASR_API int asr_session_close(AsrEngine* e, int32_t slot);
/* Endpoints of one tick in one launch.  Asynchronous and stream-ordered: takes effect after every step already submitted. */
ASR_API int asr_session_reset_many(AsrEngine* e, int32_t n, const int32_t* slots);

/* Native batch assembly for the scheduler: for i < n copy chunk_length int16 samples from base[rows[i]*row_stride + offsets[i]]
 * into row i of the pinned staging buffer of the NEXT step (multi-threaded); *pinned_out = that buffer (pass it as `pcm`). */
ASR_API int asr_gather_pcm(AsrEngine* e, int32_t n, const int16_t* base, int64_t row_stride, const int32_t* rows, const int64_t* offsets,
                           void** pinned_out);
/* Host helper for the scheduler's energy gate (stand-in for the WebRTC VAD of stream.py:166-189, whose library is absent):
 * peaks[i] = max |x| over samples [from, to) of the chunk at base[rows[i]*row_stride + offsets[i]].  Multi-threaded, no GPU work. */
ASR_API int asr_pcm_peaks(int32_t n, const int16_t* base, int64_t row_stride, const int32_t* rows, const int64_t* offsets,
                          int32_t from, int32_t to, int32_t* peaks);

/* Replaces LightningASR.stream (recognition.py:191-204) + greedy_search (recognition.py:33-57) for n sessions with
 * arbitrary, different progress.  pcm: packed [n, chunk_length] int16 (as received from the websocket,
 * streaming_server.py:362) or float32 in [-1,1).  A session may appear at most once per step. */
ASR_API int asr_step(AsrEngine* e, int32_t n, const int32_t* slots, const void* pcm, int32_t pcm_format, const AsrStepOut* out);

/* Pipelined form: asr_submit enqueues a step (H2D on a copy stream, kernels + D2H of the results on the compute stream) and
 * returns a ticket without waiting; asr_collect waits for that ticket and delivers its results.  Up to two tickets may be
 * in flight, so the input copy of step k+1 overlaps the kernels of step k.  Steps execute in submission order. */
ASR_API int asr_submit(AsrEngine* e, int32_t n, const int32_t* slots, const void* pcm, int32_t pcm_format, int32_t want_logprobs, int32_t* ticket);
ASR_API int asr_collect(AsrEngine* e, int32_t ticket, const AsrStepOut* out);
/* asr_submit with the batch assembled by the GPU: chunk i is read straight out of the sessions' audio rings at
 * base[rows[i]*row_stride + offsets[i]] (int16; `base` must be pinned host memory from asr_host_alloc) by a gather kernel on the copy
 * stream — replaces the torch.cat / slicing of Stream.audio_stream (stream.py:78-87, :159-160) without a host-side copy.  The rings
 * may be appended to at any time; rewriting a region a submitted step reads (buffer compaction) needs asr_wait_inputs first. */
ASR_API int asr_submit_rings(AsrEngine* e, int32_t n, const int32_t* slots, const int16_t* base, int64_t row_stride, const int32_t* rows,
                             const int64_t* offsets, int32_t want_logprobs, int32_t* ticket);
ASR_API int asr_wait_inputs(AsrEngine* e);
ASR_API void* asr_host_alloc(uint64_t bytes);
ASR_API int asr_host_free(void* p);

/* Same, split for pipelining / device-resident timing: stage = H2D of inputs; run = kernels only (async on the
 * engine stream); fetch = D2H of results + synchronise. */
ASR_API int asr_stage(AsrEngine* e, int32_t n, const int32_t* slots, const void* pcm, int32_t pcm_format);
ASR_API int asr_run_staged(AsrEngine* e, int32_t n, int32_t want_logprobs);
ASR_API int asr_fetch(AsrEngine* e, int32_t n, const AsrStepOut* out);
ASR_API int asr_sync(AsrEngine* e);
/* Zero-copy input: the pinned host staging buffer the NEXT asr_step / asr_submit will read (two alternate; capacity
 * max_batch * chunk_length * 4 bytes).  A caller that assembles its batch directly in this buffer passes the returned pointer
 * as `pcm` and the host-side copy is skipped (the websocket receive path can write chunks straight into pinned memory). */
ASR_API void* asr_pinned_pcm(AsrEngine* e, uint64_t* capacity_bytes);
ASR_API void* asr_stream_handle(AsrEngine* e);      /* cudaStream_t the engine launches on (for CUDA-event timing) */

/* Replaces extract_filterbank (lightspeech/datas/audio.py:9-30) [MELSPEC128 -> out [n, frames, 128]] and provides the
 * north-star Kaldi front-end (torchaudio.compliance.kaldi.fbank, 80 bins, dither 0) [KALDI80 -> out [n, frames, 80],
 * frames = 1 + (n_samples - 400) / 160; subtract_mean != 0 applies per-utterance CMVN over the frames of the call]. */
ASR_API int asr_fbank(AsrEngine* e, int32_t kind, int32_t n, const void* pcm, int32_t pcm_format, int32_t n_samples, int32_t subtract_mean,
              float* out);
ASR_API int asr_fbank_staged(AsrEngine* e, int32_t kind, int32_t n, int32_t pcm_format, int32_t n_samples);   /* kernels only, inputs from asr_stage_raw */
ASR_API int asr_stage_raw(AsrEngine* e, const void* pcm, uint64_t bytes);

/* CTC prefix beam search as part of every step (north-star; BASELINE config #4: beam = 10).  beam = 0 disables it.
 * beam <= 16, cand_k (extension candidates per frame) <= 8.  State is per session and is cleared by asr_session_reset.
 * The reference has no CTC prefix beam (its final pass is the flashlight lexicon decoder, recognition.py:220-300):
 * semantics = oracle/ctc_beam_oracle.py, parity unpinned. */
ASR_API int asr_set_beam(AsrEngine* e, int32_t beam, int32_t cand_k);

ASR_API int asr_get_stats(AsrEngine* e, AsrStats* out);

/* Per-kernel-family device time, measured with CUDA events recorded on the engine stream around every launch while
 * enabled.  asr_profile_read synchronises, returns accumulated milliseconds / launch counts per family since the last
 * read, and clears them.  (bench.py's roofline leg; off by default.) */
enum { ASR_PROF_FBANK = 0, ASR_PROF_GEMM_IN, ASR_PROF_LN, ASR_PROF_GEMM_QKV, ASR_PROF_ATTN, ASR_PROF_GEMM_OUT, ASR_PROF_GEMM_FFN1,
       ASR_PROF_GEMM_FFN2, ASR_PROF_GEMM_CTC1, ASR_PROF_GEMM_CTC2, ASR_PROF_CTC, ASR_PROF_BEAM, ASR_PROF_COUNT };
ASR_API int asr_profile_enable(AsrEngine* e, int32_t on);
ASR_API int asr_profile_read(AsrEngine* e, double* ms, uint64_t* launches);

/* ---- diagnostics used by tests (not part of the serving path) ---- */
/* Runs the step but stops after `n_layers` encoder layers (no CTC, no state advance); buffers readable below. */
ASR_API int asr_debug_step_partial(AsrEngine* e, int32_t n, const int32_t* slots, const void* pcm, int32_t pcm_format, int32_t n_layers);
/* which: 0 = x (layer output / input_linear output) [n*rows, d]; 1 = x1; 2 = x2; 3 = q; 4 = logits [n*S, vocab] */
ASR_API int asr_debug_read(AsrEngine* e, int32_t which, float* out, uint64_t n_floats);
/* Reads the K (which=0) / V (which=1) left context of `layer` of a session in the reference's layout
 * [left_context, d] (oldest row first, zero rows where past_length < left_context) and past_length. */
ASR_API int asr_debug_read_state(AsrEngine* e, int32_t slot, int32_t layer, int32_t which, float* out, int32_t* past_length);
/* Stand-alone GEMM C[M,N] = A[M,K] * B[N,K]^T (+bias): impl 0 = tcgen05 kernel, 1 = CUDA-core cross-check. */
ASR_API int asr_debug_gemm(int32_t impl, int32_t M, int32_t N, int32_t K, int32_t split, int32_t bn, const float* A, const float* B, const float* bias,
                   float* C, int device);

/* Mean milliseconds per launch of the tcgen05 GEMM on operands resident in HBM (microbenchmark; bn = 512 selects the CTA-pair
 * kernel).  epi_kind: 0 plain fp32 store, 1 bias + fp32 residual, 2 bias + GELU -> bf16. */
/* ---- Streaming convolution module with per-session cache (csrc/convmod.cu): streaming form of ConvolutionBlock
 * (lightspeech/layers/block.py:129-171: pre_norm -> pointwise_conv1 -> SiLU -> depthwise_conv(k) -> BatchNorm1d(eval) -> SiLU ->
 * pointwise_conv2).  Output = the reference block's output on the whole sequence, delayed by (kernel-1)/2 frames.  Sessions are
 * caller-numbered slots in [0, max_sessions).  weights: fp32 blob in the order documented at asr_convmod_create in convmod.cu. */
typedef struct AsrConvModule AsrConvModule;
ASR_API int asr_convmod_weights_count(int32_t d_model, int32_t kernel, uint64_t* n_floats);
ASR_API int asr_convmod_create(int32_t d_model, int32_t kernel, int32_t rows_per_chunk, int32_t max_sessions, int32_t max_batch, int32_t precision,
                               const float* weights, uint64_t n_floats, int32_t device, AsrConvModule** out);
ASR_API int asr_convmod_destroy(AsrConvModule* m);
ASR_API int asr_convmod_reset(AsrConvModule* m, int32_t n, const int32_t* slots);
ASR_API int asr_convmod_step(AsrConvModule* m, int32_t n, const int32_t* slots, const float* x, float* y);

/* Device time (CUDA events on the engine stream: kernel chain + result D2H) of the pipelined steps collected so far: with the wall time of
 * the same ticks it says how much of a tick the GPU was busy.  reset != 0 clears the counters. */
ASR_API int asr_pipeline_gpu_time(AsrEngine* e, double* total_ms, uint64_t* n_steps, int32_t reset);

/* Diagnostic: the GEMM (N = 512) with residual add + LayerNorm(s) fused into the epilogue (csrc/gemm_ln.cu) on host operands. */
ASR_API int asr_debug_gemm_ln(int32_t M, int32_t K, int32_t split, const float* A, const float* W, const float* bias, const float* res,
                              const float* g1, const float* b1, const float* g2, const float* b2, int32_t f32_normed, int32_t compact_rows,
                              int32_t compact_seg, float* out_f32, float* out_op_f32, int32_t iters, float* ms_out, int32_t pair, int device);
/* Diagnostic: mean ms per launch of the tcgen05 GEMM on operands already in HBM.  bn: 64 / 128 / 256 = 1-CTA tile width, 512 = cta_group::2
 * pair (256 x 256), 513 = pair with 256 x 128 tiles and four accumulator stages, 514 = pair with the A tile resident in shared memory (epi 2
 * only).  epi_kind: 0 fp32 store, 1 + bias + fp32 residual, 2 bias + GELU -> bf16 operand, 3 none (accumulator dropped: the mainloop alone). */
ASR_API int asr_debug_gemm_time(int32_t M, int32_t N, int32_t K, int32_t split, int32_t bn, int32_t epi_kind, int32_t iters, float* ms_out, int device);

#ifdef __cplusplus
}
#endif
#endif /* ASR_B200_H_ */
