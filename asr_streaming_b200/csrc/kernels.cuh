// Launch interfaces of the memory-bound kernels (fbank, LayerNorm, chunk attention, CTC decode).
#pragma once
#include "common.cuh"

namespace asr {

// ---------------------------------------------------------------- fbank
struct FbankParams {
  const void* pcm;        // [n_streams, pcm_stride] int16 or float
  const int* row_index;   // nullable: stream b reads row row_index[b] of `pcm` (batch assembled earlier than the batch was decided)
  int pcm_is_f32;
  int pcm_stride;         // samples between consecutive streams
  int n_samples;          // samples staged per stream
  int n_frames, hop, frame_len, frame_off;
  int blk;                // gcd(hop, frame_len): samples per block sum of the DC-offset pass (filled in by fbank_launch)
  int nc;                 // complex FFT size: 400 (800-pt real) or 256 (512-pt real)
  int kaldi;              // remove-DC + pre-emphasis path
  float in_scale;         // 1/32768 for int16 -> [-1,1) (streaming_server.py:362-363), 1 otherwise
  float preemph;
  float log_floor;
  const float* window;    // [frame_len]
  const float2* tw;       // [nc]      exp(-2*pi*i*m/nc)
  const float2* w2;       // [nc+1]    exp(-2*pi*i*k/(2*nc))
  const int* mel_start;   // [n_mels]
  const int* mel_cnt;     // [n_mels]
  const int* mel_off;     // [n_mels]
  const float* mel_w;     // packed triangle weights
  int n_mels;
  float* out_f32;         // nullable [n_streams, n_frames, n_mels]
  bf16* out_op;           // nullable A operand [n_streams*n_frames, op_ld]
  int op_ld, op_lo_off;
};
int fbank_launch(const FbankParams& P, int n_streams, cudaStream_t st);

// ---------------------------------------------------------------- LayerNorm (d = 512)
// y = LN(x; g, b)  ->  A operand (bf16 hi [+lo])                         (TA:emformer.py:427-434, pos_ff[0])
int ln_to_operand(const float* x, const float* g, const float* b, bf16* out, int ld, int lo_off, int M, int d, cudaStream_t st);
// y = LN(x2; g1, b1) -> fp32 y ; then (g2 != null) LN(y; g2, b2) -> A operand           (layer_norm_output + next layer_norm_input)
// or (g2 == null) the segment rows of y -> compact A operand [B*seg_rows, ld]           (TA:emformer.py:803 drops rc rows)
int ln_out_fused(const float* x2, const float* g1, const float* b1, float* y, const float* g2, const float* b2, bf16* out, int ld,
                 int lo_off, int M, int d, int rows, int seg_rows, cudaStream_t st);

// ---------------------------------------------------------------- chunk attention over the ring KV cache
template <typename T>
struct AttnParams {
  const T* q;              // [M, d], already scaled by d_h^-0.5
  const T* cache_layer;    // cache + layer*(2*ring*d)
  size_t slot_stride;
  const T* rc;             // [B, 2, rc_rows, d]
  const int* slots;        // [B]
  const int* past_len;     // [n_slots]  (value before this step)
  bf16* out;               // A operand [M, ld]
  int ld, lo_off;
  int rows, seg_rows, rc_rows, ring, left, d, n_heads;
  // streaming bf16 kernel only (host pointers, read by the launcher): 3D head-major tensor maps of the whole K/V cache (box of
  // seg_rows rows) and of the right-context scratch (box of rc_rows rows); cache_row0 = first cache row of this layer's slab.
  const CUtensorMap* h_tm_cache = nullptr;
  const CUtensorMap* h_tm_rc = nullptr;
  long long cache_row0 = 0;
  long long slot_rows = 0;   // cache rows per session inside a layer slab (2 * ring)
};
template <typename T> int attention_launch(const AttnParams<T>& P, int n_streams, cudaStream_t st);

constexpr int BEAM_MAX = 16;         // beam width limit
constexpr int BEAM_CAND_MAX = 8;     // extension candidates per frame limit

// ---------------------------------------------------------------- CTC log-softmax + argmax + incremental greedy
struct CtcParams {
  const float* logits;     // [B*seg_rows, vocab]
  int vocab, seg_rows;
  const int* slots;        // [B]
  // per-slot carry of greedy_search's unique_consecutive / last_blank across chunks (recognition.py:33-57)
  int* prev_id;            // last argmax id of the segment so far (-1: none)
  int* n_frames;           // frames accumulated in the current utterance segment
  int* last_tok_frame;     // index of last frame with id > 1 (-1: none)
  int* seg_has_text;       // the segment so far holds an id whose vocabulary string survives greedy_search's stripping (recognition.py:47-52)
  const uint32_t* silent_mask;   // [ceil(vocab / 32)] bit = id renders to the empty string ('-', '|', '<<', '>>' in the reference vocabulary)
  int* past_len;           // advanced by seg_rows here (end of step)
  // outputs per stream
  int* argmax_ids;         // [B*seg_rows]
  int* new_tokens;         // [B*seg_rows] collapsed, blank-dropped ids appended this chunk
  int* n_new;              // [B]
  int* blank_frames;       // [B]  frames since last token (or all frames if none)
  int* has_token;          // [B]  an id > 1 exists in the segment (the `tokens_idx` test of recognition.py:40)
  int* has_text;           // [B]  the rendered text of the segment is non-empty (the `if text:` test of stream.py:121)
  int* flags;              // [B]  cleared here (ASR_FLAG_* bits are OR-ed in by later kernels of the step)
  float* logprobs;         // nullable [B*seg_rows, vocab]
  // prefix beam search only (cand_k > 0): per row the cand_k best non-blank ids (value desc, id asc; BEAM_CAND_MAX entries per row) and
  // the row's (max logit, log-sum-exp) so that the beam kernel evaluates log-probs of single ids as (logit - max) - lse: the same two
  // fp32 operations as here, without the [rows, vocab] log-prob array ever being written
  int cand_k = 0;
  int* cand_tok = nullptr;       // [B*seg_rows, BEAM_CAND_MAX]
  float* cand_lp = nullptr;      // [B*seg_rows, BEAM_CAND_MAX]
  float* row_stat = nullptr;     // [B*seg_rows, 2]
};
int ctc_greedy_launch(const CtcParams& P, int n_streams, cudaStream_t st);

// ---------------------------------------------------------------- CTC prefix beam search (warp per stream)
constexpr int BEAM_MAX_LEN = 1024;   // tokens per hypothesis: an utterance is force-ended at 40 s = 1000 frames (asr-online.yaml:103-107) and CTC emits
                                     // at most one token per frame, so the cap is never reached under the reference's rules; if it is, the step says so
struct BeamParams {
  const float* logits;     // [n*seg_rows, vocab] of the current step (CTC head output before log_softmax)
  const float* row_stat;   // [n*seg_rows, 2]  (max logit, log-sum-exp) per row: log-prob(id) = (logit[id] - max) - lse
  const int* cand_tok;     // [n*seg_rows, BEAM_CAND_MAX]  extension candidates per frame, from ctc_greedy_kernel
  const float* cand_lp;    // [n*seg_rows, BEAM_CAND_MAX]
  const int* slots;        // [n]
  int n, seg_rows, vocab, beam, cand_k, max_len;
  // per-slot state
  int* n_beam; int* cur;               // [slots]
  int* len; int* last;                 // [slots*BEAM_MAX]
  float* pb; float* pnb;               // [slots*BEAM_MAX]
  unsigned long long* hash;            // [slots*BEAM_MAX]
  int16_t* tokens;                     // [slots][2][BEAM_MAX][BEAM_MAX_LEN]
  // outputs of the step: best hypothesis so far per stream
  int16_t* out_tokens;     // [n*BEAM_MAX_LEN]
  int* out_len;            // [n]
  float* out_score;        // [n]
  int* out_flags;          // [n] |= ASR_FLAG_BEAM_TRUNCATED when a hypothesis could not be extended because it is BEAM_MAX_LEN - 1 tokens long
};
int beam_launch(const BeamParams& P, cudaStream_t st);
int beam_reset_many_launch(const BeamParams& P, const int* d_slots, int n, cudaStream_t st);
int beam_reset_launch(const BeamParams& P, int slot /* -1: all */, int n_slots_all, cudaStream_t st);

// fp32 [n] -> bf16 hi (+ lo) weight conversion at engine creation
int convert_weight(const float* src, bf16* dst, int rows, int cols, int ld, int lo_off, cudaStream_t st);
int fill_i32(int* p, int v, size_t n, cudaStream_t st);
// device-side batch assembly from pinned host rings (zero-copy reads): dst[i] = base[src_off[i] .. + chunk_len)
int gather_rings_launch(const int16_t* base_dev, const long long* src_off, int16_t* dst, int n, int chunk_len, cudaStream_t st);
// dst block i = src block idx[i], blocks of block_bytes (multiple of 16): compaction of pre-computed per-stream operand blocks
int gather_blocks_launch(const void* src, void* dst, const int* idx, int n, size_t block_bytes, cudaStream_t st);
// endpoint on many sessions at once: past_len = n_frames = has_text = 0, prev_id = last_tok = -1 for the listed slots
int reset_slots_launch(const int* slots, int n, int* past_len, int* n_frames, int* prev_id, int* last_tok, int* has_text, cudaStream_t st);
// per-utterance CMVN over the frames of one call (TA:compliance/kaldi.py:603-606, subtract_mean)
int subtract_mean_launch(float* x, int n_streams, int n_frames, int n_mels, cudaStream_t st);

}  // namespace asr
