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

#define ASR_B200_ABI_VERSION 2

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

#define ASR_BEAM_MAX_LEN 1024        /* tokens per beam hypothesis: rule4 force-ends an utterance at 40 s = 1000 frames (asr-online.yaml:103-107) */
enum { ASR_FLAG_BEAM_TRUNCATED = 1 };  /* AsrStepOut.flags: a beam hypothesis reached ASR_BEAM_MAX_LEN - 1 tokens and could not be extended */

typedef struct AsrStepOut {     /* all pointers nullable, host memory, n = streams in the step, S = segment rows (16) */
  int32_t* argmax_ids;          /* [n*S]  per-frame argmax of the log-probs (recognition.py:36)                       */
  int32_t* new_tokens;          /* [n*S]  ids appended this chunk after unique_consecutive + blank drop (:44-45)     */
  int32_t* n_new;               /* [n]                                                                                */
  int32_t* blank_frames;        /* [n]    frames since the last id > 1, or all frames of the segment if none (:38-43) */
  int32_t* has_token;           /* [n]    an id > 1 exists in the segment: `len(tokens_idx)` (recognition.py:40-41)    */
  float* logprobs;              /* [n*S*vocab]  the reference's `emission` (recognition.py:203-204)                   */
  /* CTC prefix beam search (only when asr_set_beam enabled it): best hypothesis of the utterance so far               */
  int16_t* beam_tokens;         /* [n*ASR_BEAM_MAX_LEN]  row i valid in [0, beam_len[i])                              */
  int32_t* beam_len;            /* [n]                                                                                */
  float* beam_score;            /* [n]  log P(best prefix)                                                            */
  int32_t* has_text;            /* [n]    greedy_search's rendered text of the segment is non-empty: the `if text:` of
                                          Stream.update_stream (stream.py:121); differs from has_token for ids that render to ""
                                          (asr_set_silent_ids)                                                        */
  int32_t* flags;               /* [n]    ASR_FLAG_* bits                                                             */
} AsrStepOut;

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
/* open / reset never wait for the device: the state is cleared by a stream-ordered kernel, i.e. after every step already submitted
 * (a reset of a session that rides a step in flight applies to the steps submitted afterwards).  close fails while a submitted,
 * uncollected step carries the session. */
ASR_API int asr_session_open(AsrEngine* e, int32_t* slot_out);
ASR_API int asr_session_reset(AsrEngine* e, int32_t slot);           /* endpoint: state := init, emission := []  (:514-515, :530) */
ASR_API int asr_session_close(AsrEngine* e, int32_t slot);
/* Endpoints of one tick in one launch.  Asynchronous and stream-ordered: takes effect after every step already submitted. */
ASR_API int asr_session_reset_many(AsrEngine* e, int32_t n, const int32_t* slots);
/* Vocabulary ids whose string renders to "" in greedy_search (recognition.py:47-52 strips '<<', '>>', '-' and turns '|' into a space that
 * strip() removes): they do not make a segment's text non-empty (AsrStepOut.has_text).  Default {0, 1}; the reference vocabulary
 * (lightspeech/corpus/vocab.txt) also has 792 '<<' and 793 '>>'. */
ASR_API int asr_set_silent_ids(AsrEngine* e, int32_t n, const int32_t* ids);

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
 * streaming_server.py:362) or float32 in [-1,1).  A session may appear at most once per step (duplicates are rejected). */
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
 * engine stream); fetch = D2H of results + synchronise.  These (and asr_fbank, asr_stage_raw) use staging buffer 0 and fail while an
 * asr_submit ticket is in flight. */
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

/* ---- Native session scheduler (csrc/sched.cu): the per-connection loop of the reference server for thousands of sessions at once.
 * Replaces, per session, Stream (streaming_decoder/stream.py: :23-26 initial zero buffer, :78-87 accept_waveform, :110-125
 * update_stream, :127-163 endpoint_detected, :159-160 advance by segment_length, :166-189 VAD skip), the chunk loop of
 * handle_connection_impl (streaming_server.py:367-546), the rule evaluation of online_endpoint.py:42-94 and the v1 batcher
 * StreamingE2E.process (streaming_decoder_v1/streaming_asr.py:41-119).  Session state is struct-of-arrays owned by the library
 * (asr_sched_arrays hands out the pointers; Python wraps them as numpy views).  A tick with the engine attached is two calls,
 * asr_sched_submit / asr_sched_collect, with up to two ticks in flight; plan / commit / update / endpoints are the same bookkeeping
 * without an engine (CPU tests, or a caller that interposes its own VAD or language-model cost between them). */
typedef struct AsrScheduler AsrScheduler;
typedef struct AsrSchedConfig {
  int32_t capacity;          /* session rows */
  int32_t chunk_length;      /* samples per chunk            (AudioConfig.chunk_length, utils.py:22) */
  int32_t segment_length;    /* samples consumed per chunk   (utils.py:18, stream.py:159-160)        */
  int32_t buffer_length;     /* leading zeros / overlap      (utils.py:20, stream.py:23)             */
  int32_t seg_rows;          /* output frames per chunk */
  int32_t sample_rate;
  int32_t max_batch;         /* sessions per tick at most */
  int32_t backlog_chunks;    /* ring capacity per session = chunk_length + backlog_chunks * segment_length samples */
  int32_t max_tokens;        /* greedy tokens kept per utterance segment (overflow is reported, never silent) */
  int32_t device_gather;     /* != 0: rings in pinned, device-mapped memory, the GPU gathers each tick's chunks itself */
  double relative_cost;      /* default LM relative cost fed to the endpoint rules (utils.py:126-139; the ARPA LM is absent) */
} AsrSchedConfig;
typedef struct AsrSchedArrays {     /* one entry (row) per session; valid for the scheduler's lifetime */
  int16_t* audio; int64_t audio_row_samples;
  int64_t *rd, *wr;                 /* read / write positions inside the session's ring */
  uint8_t *active, *inflight;
  int32_t* slot;
  int32_t* tok; int32_t* ntok;      /* [capacity][max_tokens] greedy tokens of the current segment */
  int64_t *n_frames, *chunk_processed, *chunk_processed_total;
  double* trailing;                 /* trailing_blank_duration (stream.py:122-125) */
  uint8_t* contain_token;           /* is_contain_token (stream.py:123) */
  int64_t *segment, *last_served;
  double* relative_cost;            /* per session; a language model writes here before the endpoint rules run */
  uint8_t* overflow;                /* the segment lost tokens (max_tokens) or its beam hypotheses were truncated */
} AsrSchedArrays;
typedef struct AsrSchedPlan {       /* the sessions of the tick being assembled (valid until the tick is collected) */
  int32_t tick, n;
  const int32_t *rows, *slots; const int64_t* offsets;     /* chunk i = audio[rows[i]][offsets[i] : + chunk_length] */
  int32_t n_skipped; const int32_t* skipped;               /* VAD-gated sessions: consumed, not run */
} AsrSchedPlan;
typedef struct AsrSchedResult {     /* outcome of a tick; pointers valid until two more ticks were submitted */
  int32_t n; const int32_t* rows;                          /* sessions run through the model, batch order */
  const int32_t *n_new, *new_tokens;                       /* [n], [n * seg_rows] */
  const uint8_t* final_flags; const int32_t* final_rule;   /* [n] an endpoint fired after this chunk / index of the rule, -1 */
  const uint8_t* overflow;                                 /* [n] */
  int32_t n_skipped; const int32_t* skipped;
  int32_t n_final;                                         /* endpoints of the tick (served and skipped sessions) */
  const int32_t *final_rows, *final_rule_of, *final_ntok, *final_tok_off; const double* final_utt; const int32_t* final_tok;
  /* the engine's step outputs (asr_sched_collect only; pinned memory) */
  const int32_t *argmax_ids, *blank_frames, *has_token, *has_text, *flags;
  const int16_t* beam_tokens; const int32_t* beam_len; const float* beam_score; const float* logprobs;
} AsrSchedResult;
ASR_API int asr_sched_create(const AsrSchedConfig* cfg, AsrEngine* engine /* nullable */, AsrScheduler** out);
ASR_API int asr_sched_destroy(AsrScheduler* s);
ASR_API int asr_sched_arrays(AsrScheduler* s, AsrSchedArrays* out);
/* Rule table in evaluation order (online_endpoint.py:4-21, asr-online.yaml:31-104); n = 0 disables endpointing. */
ASR_API int asr_sched_set_rules(AsrScheduler* s, int32_t n, const uint8_t* must_contain_nonsilence, const double* min_trailing_silence,
                                const double* min_utterance_length, const double* max_relative_cost);
ASR_API int asr_sched_open(AsrScheduler* s, int32_t row, int32_t slot /* < 0: open an engine session */);
ASR_API int asr_sched_close(AsrScheduler* s, int32_t row);
ASR_API int asr_sched_reset_rows(AsrScheduler* s, int32_t n, const int32_t* rows);
ASR_API int asr_sched_accept(AsrScheduler* s, int32_t row, const int16_t* pcm, int64_t n);      /* returns 1 when the backlog does not fit */
ASR_API int asr_sched_accept_block(AsrScheduler* s, int32_t n, const int32_t* rows, const int16_t* block, int64_t samples_per_row);
ASR_API int asr_sched_ready(AsrScheduler* s, int32_t max_rows, const int32_t** rows, int32_t* n);
/* gate_threshold >= 0: energy gate (peak of the chunk's new samples) for sessions without text in their segment; keep != NULL: the
 * caller's own decision for the rows asr_sched_ready returned. */
ASR_API int asr_sched_plan(AsrScheduler* s, int32_t max_rows, int32_t gate_threshold, const uint8_t* keep, AsrSchedPlan* plan);
ASR_API int asr_sched_commit(AsrScheduler* s, int32_t tick, AsrSchedResult* res);
ASR_API int asr_sched_update(AsrScheduler* s, int32_t tick, const AsrStepOut* out);
ASR_API int asr_sched_endpoints(AsrScheduler* s, int32_t tick, AsrSchedResult* res);
ASR_API int asr_sched_abort(AsrScheduler* s, int32_t tick);
ASR_API int asr_sched_submit(AsrScheduler* s, int32_t max_rows, int32_t gate_threshold, const uint8_t* keep, int32_t want_logprobs, AsrSchedResult* res,
                             int32_t* tick);
ASR_API int asr_sched_collect(AsrScheduler* s, int32_t tick, int32_t run_endpoints, AsrSchedResult* res);
/* One-tick-per-pass pipelining: gather + H2D of every buffered chunk (also of the sessions still in flight) NOW, overlapping the running
 * tick; the next asr_sched_submit decides which of them run and launches on that subset.  See csrc/sched.cu. */
ASR_API int asr_sched_prestage(AsrScheduler* s, int32_t gate_threshold, int32_t* n_staged);

/* ---- diagnostics used by tests (not part of the serving path) ---- */
/* Runs the step but stops after `n_layers` encoder layers (no CTC, no state advance); buffers readable below. */
ASR_API int asr_debug_step_partial(AsrEngine* e, int32_t n, const int32_t* slots, const void* pcm, int32_t pcm_format, int32_t n_layers);
/* The decode stage alone on caller-supplied CTC logits [n*S, vocab] (host memory): log_softmax + argmax + incremental greedy collapse
 * (+ the per-frame extension candidates and the prefix beam search when asr_set_beam enabled it), session carries advanced exactly as
 * by asr_step.  Lets tests drive the decoders with peaked, tied or degenerate posteriors that the random-init encoder never produces. */
ASR_API int asr_debug_decode_logits(AsrEngine* e, int32_t n, const int32_t* slots, const float* logits, const AsrStepOut* out);
/* which: 0 = x (layer output / input_linear output) [n*rows, d]; 1 = x1; 2 = x2; 3 = q; 4 = logits [n*S, vocab]; with the prefix beam enabled
 * also 5 = log-probs of the per-frame extension candidates [n*S, 8]; 6 = their ids (int32 bit patterns); 7 = (max logit, lse) per row [n*S, 2] */
ASR_API int asr_debug_read(AsrEngine* e, int32_t which, float* out, uint64_t n_floats);
/* Reads the K (which=0) / V (which=1) left context of `layer` of a session in the reference's layout
 * [left_context, d] (oldest row first, zero rows where past_length < left_context) and past_length. */
ASR_API int asr_debug_read_state(AsrEngine* e, int32_t slot, int32_t layer, int32_t which, float* out, int32_t* past_length);
/* Stand-alone GEMM C[M,N] = A[M,K] * B[N,K]^T (+bias) through the tcgen05 kernel (impl must be 0). */
ASR_API int asr_debug_gemm(int32_t impl, int32_t M, int32_t N, int32_t K, int32_t split, int32_t bn, const float* A, const float* B, const float* bias,
                   float* C, int device);

/* Stand-alone act(A B^T + bias) -> bf16 operand (returned as fp32).  bn: 64 / 128 / 256 one-CTA tiles, 512 = cta_group::2 pair with the
 * LSU epilogue, 515 = pair with the TMA-store epilogue.  act: 0 none, 1 GELU, 2 SiLU. */
ASR_API int asr_debug_gemm_operand(int32_t M, int32_t N, int32_t K, int32_t bn, int32_t act, const float* A, const float* B, const float* bias,
                                   float* out, int device);

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
 * pair (256 x 256), 515 = pair with the TMA-store epilogue (epi 2 only).  epi_kind: 0 fp32 store, 1 + bias + fp32 residual, 2 bias + GELU ->
 * bf16 operand, 3 none (accumulator dropped: the mainloop alone). */
ASR_API int asr_debug_gemm_time(int32_t M, int32_t N, int32_t K, int32_t split, int32_t bn, int32_t epi_kind, int32_t iters, float* ms_out, int device);

#ifdef __cplusplus
}
#endif
#endif /* ASR_B200_H_ */
