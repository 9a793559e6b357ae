// Engine + C ABI (include/asr_b200.h): device-resident session state (K/V rings, greedy carry), weight
// packing, the per-step kernel chain, and the host<->device staging of one ragged batch of stream-chunks.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/asr_b200.h"
#include "gemm.cuh"
#include "kernels.cuh"
#include "sched_hooks.h"

namespace asr {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static thread_local bool g_pdl_active = true;
void pdl_set_active(bool on) { g_pdl_active = on; }
bool pdl_enabled() {
  static const bool on = [] { const char* v = getenv("ASR_B200_NO_PDL"); return !(v && v[0] == '1'); }();
  return on && g_pdl_active;
}

namespace {

inline size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

template <class F>
void parallel_rows(int n, int max_threads, F&& work) {          // work(first, last) over [0, n) on a few host threads
  // ASR_B200_HOST_THREADS caps the helper threads (one process per GPU on a shared host: cores / ranks)
  static const int env_cap = [] { const char* v = getenv("ASR_B200_HOST_THREADS"); return v ? std::max(1, atoi(v)) : 1 << 20; }();
  const int hw = (int)std::thread::hardware_concurrency();
  const int nt = std::max(1, std::min({max_threads, env_cap, hw > 0 ? hw : 1, n / 64 + 1}));
  if (nt == 1) { work(0, n); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t) th.emplace_back([&, t] { work((int)((long long)n * t / nt), (int)((long long)n * (t + 1) / nt)); });
  for (auto& x : th) x.join();
}

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int alloc(size_t n) {
    bytes = n;
    ASR_CUDA_OK(cudaMalloc(&p, n ? n : 16));
    return 0;
  }
  void free() { if (p) cudaFree(p); p = nullptr; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct WeightMat {            // one nn.Linear weight as a tcgen05 B operand
  bf16* w = nullptr;          // [N, ld]  (ld = K or 2K)
  int N = 0, K = 0, ld = 0;
  CUtensorMap tm[3];          // box rows 64 / 128 / 256
};

struct LayerW {
  WeightMat qkv, o, w1, w2;
  TsMaps ts_qkv, ts_ffn1;                 // tensor maps + bias (kernel-parameter copy) of the TMA-store epilogues
  const float *bqkv, *bo, *ln_in_g, *ln_in_b, *ln_ff_g, *ln_ff_b, *b1, *b2, *ln_out_g, *ln_out_b;
  const float* ln_out_consts = nullptr;   // LnEpilogue::y_consts of layer_norm_output (device), see gemm.cuh
};

struct FbankPlan {
  int nc = 0, frame_len = 0, n_mels = 0;
  DevBuf window, tw, w2, mel_start, mel_cnt, mel_off, mel_w;
};

struct Operand {              // a GEMM A operand buffer + its tensor maps
  DevBuf buf;
  int rows = 0, K = 0, ld = 0, lo_off = 0;
  CUtensorMap tm;             // load map: box {64, 128}
  CUtensorMap tm_st;          // store map of the GEMM that produces it (TMA-store epilogue): box {64, 32}
};

}  // namespace
}  // namespace asr

using namespace asr;

struct AsrEngine {
  AsrConfig cfg;
  Geo geo;
  int device = 0, num_sms = 148;
  int no_pair = 0;
  int tma_store = 1;            // TMA-store epilogues of the bf16-output pair GEMMs (ASR_B200_NO_TMA_STORE=1: LSU epilogue, for A/B)
  TsMaps ts_ctc1;               // CTC1 (Linear + SiLU): store map of a_ctc + bias
  CUtensorMap tm_a_ln_seg;      // a_ln as [stream][row][K], segment-row box (stream-tiled Q | K | V projection)
  bool qkv_ts_ready = false;
  int no_pair_ln = 0;           // ASR_B200_NO_PAIR_LN=1: gemm_ln always in the 2-CTA shape
  int no_fuse2 = 0;             // ASR_B200_NO_LN_FUSE2=1: second LayerNorm statistics by their own TMEM pass
  // CUDA graphs of the per-step kernel chain for small batches (launch-bound: ~290 launches of a few us each).  Key = (streams, staging
  // buffer, pcm format, log-probs wanted, beam): captured the second time a key is seen, replayed afterwards.  Opt-in (ASR_B200_GRAPHS=1):
  // capture + instantiation of a 146-node graph costs milliseconds per NEW key, which a scheduler with a different batch size every tick
  // pays in its tail latency (real-time simulation, 10,240 streams: p99 3.5 -> 6.1 ms, max 5.6 -> 27.8 ms) for a 4-5 % shorter step.
  struct StepGraph { cudaGraphExec_t exec = nullptr; uint64_t launches = 0; int seen = 0; };
  std::unordered_map<uint64_t, StepGraph> graphs;
  int use_graphs = 0, graph_max_streams = 128, graph_max_entries = 256;
  int pair_ln_min_tiles = 34;   // 256-row tiles needed before gemm_ln takes the cta_group::2 shape (34 clusters of 4 fit on 148 SMs)
  int quad_ln_max_tiles = 34;   // 128-row tiles up to which gemm_ln takes the four-column-quarter shape: more would need a second wave of clusters (measured: 3200 rows 11.7 vs 15.2 us, 5120 rows 16.6 vs 15.9 us)
  int fused_ln = 1;             // LayerNorm fused into the out_proj / FFN2 epilogues (gemm_ln.cu); ASR_B200_NO_FUSED_LN=1 -> separate passes
  int fused_ln_min_streams = 96;    // below this batch the separate LN passes win (measured with the quad shape: 64 streams 1.31 vs 1.33 ms,
                                    // 128: 1.53 vs 1.48, 1024: 5.99 vs 5.77)
  int pdl_max_streams = 1536;   // programmatic dependent launch below this batch size (see common.cuh)
  int staged_fmt = 0;
  cudaStream_t stream = nullptr;
  std::mutex mu;

  DevBuf w_f32, w_bf16, ln_consts;
  WeightMat w_in, ctc1, ctc2;
  const float *ctc_b1 = nullptr, *ctc_b2 = nullptr;
  std::vector<LayerW> layers;
  FbankPlan mel128, kaldi80;

  // per-step activations
  DevBuf d_pcm, d_slots, x, x1, x2, q, rc_kv, logits, fb_f32;
  Operand a_fb, a_ln, a_attn, a_h, a_enc, a_ctc;
  // per-session state
  DevBuf kv_cache, past_len, prev_id, n_frames, last_tok, seg_has_text, silent_mask;
  CUtensorMap tm_kv, tm_rc;      // head-major 3D views of the K/V cache / right-context scratch for the streaming attention kernel (bf16 only)
  bool attn_tma = false;
  size_t slot_stride = 0;        // elements between two sessions' rings inside one layer's slab
  size_t layer_stride = 0;       // elements between two layers' slabs: K/V cache layout is [layer][slot][K|V][ring][d]
  std::vector<int> free_slots;
  std::vector<uint8_t> slot_open;
  std::vector<uint8_t> slot_inflight;     // submitted, uncollected steps that carry the slot (0..2)
  std::vector<uint32_t> slot_stamp;       // duplicate detection inside one step: epoch of the last step that listed the slot
  uint32_t stamp_epoch = 0;
  // outputs
  DevBuf d_argmax, d_newtok, d_nnew, d_blank, d_hastok, d_hastext, d_flags, d_logprobs;
  // prefix beam search (optional)
  int beam = 0, cand_k = 0;
  DevBuf bm_n, bm_cur, bm_len, bm_last, bm_pb, bm_pnb, bm_hash, bm_tokens, d_beam_tok, d_beam_len, d_beam_score;
  DevBuf bm_cand_tok, bm_cand_lp, bm_row_stat;   // per row of the step: extension candidates and (max, lse), written by ctc_greedy_kernel
  // Double-buffered staging so that step k+1's H2D overlaps step k's kernels (asr_submit / asr_collect):
  // pinned host [pcm | slots | results] x 2, device pcm/slots x 2 (buffer 0 = d_pcm / d_slots above), a copy stream.
  void* h_stage = nullptr;      // == h_buf[0]
  void* h_buf[2] = {nullptr, nullptr};
  size_t h_stage_bytes = 0;
  size_t h_out_off = 0;
  DevBuf d_pcm2, d_slots2;
  DevBuf d_src_off[2];          // asr_submit_rings: element offsets of the step's chunks inside the pinned session rings
  DevBuf d_row_index[2];        // pre-staged batches: staged row of every batch element
  int prestaged_rows[2] = {-1, -1};   // rows gathered + copied ahead of the decision which of them run (engine_prestage)
  // speculative front-end: the fbank of every pre-staged chunk does not depend on which of them will run, so engine_prestage queues it
  // behind the running step (it executes in the collect -> submit gap); the step then only compacts the rows that run into a_fb
  Operand a_fb_stage;
  int fb_staged[2] = {0, 0};
  bool spec_fbank = true;             // ASR_B200_NO_SPEC_FBANK=1: fbank inside the step through the row index (A/B)
  bool act_fb_stage = false, act_fb_identity = false;
  cudaEvent_t ev_pre[2] = {nullptr, nullptr};
  const int* act_row_index = nullptr;
  void* act_pcm = nullptr;      // input buffers the kernels of the step being enqueued read
  int* act_slots = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  cudaEvent_t ev_t0[2] = {nullptr, nullptr}, ev_t1[2] = {nullptr, nullptr};     // timing: kernel chain + D2H of a pipelined step on the engine stream
  double pipe_gpu_ms = 0.0; uint64_t pipe_gpu_n = 0;
  // asr_session_reset_many: ring of pinned slot lists (+ the event of each list's H2D) so that back-to-back calls never wait
  static constexpr int kResetRing = 4;
  cudaEvent_t ev_reset[kResetRing] = {nullptr, nullptr, nullptr, nullptr};
  int32_t* h_reset = nullptr;
  DevBuf d_reset;
  int reset_pos = 0;
  struct Pending { int n = 0; int want_lp = 0; int active = 0; std::chrono::steady_clock::time_point t0; std::vector<int32_t> slots; } pend[2];
  int cur_buf = 0;

  // per-kernel-family CUDA-event profiling (bench.py roofline): pairs recorded on the launching stream
  int prof_on = 0;
  struct ProfRec { int cat; cudaEvent_t a, b; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[ASR_PROF_COUNT] = {0};
  uint64_t prof_n[ASR_PROF_COUNT] = {0};

  // stats
  uint64_t steps = 0, stream_chunks = 0, launches = 0;
  std::vector<float> step_ms;
  size_t step_ms_pos = 0;
};

namespace {

// ------------------------------------------------------------------------------------------ geometry
int fill_geo(const AsrConfig& c, Geo* g) {
  if (c.abi_version != ASR_B200_ABI_VERSION) { set_error("AsrConfig.abi_version %d != %d", c.abi_version, ASR_B200_ABI_VERSION); return -1; }
  g->hop = c.hop; g->n_fft = c.n_fft; g->win = c.win; g->n_mels = c.n_mels;
  g->chunk_len = (c.segment_size + c.context_size + c.bias) * c.hop;         // utils.py:18-22
  g->frames = 1 + (g->chunk_len - c.n_fft) / c.hop;                          // torch.stft center=False
  g->stride = c.stride; g->d_model = c.d_model; g->n_heads = c.n_heads; g->ffn = c.ffn_dim; g->n_layers = c.n_layers;
  g->seg_rows = c.segment_size / c.stride; g->rc_rows = c.context_size / c.stride; g->rows = g->seg_rows + g->rc_rows;
  g->left = c.left_context; g->ring = g->left + g->seg_rows;
  g->ctc_hidden = c.ctc_hidden; g->vocab = c.vocab; g->split = c.precision == ASR_PRECISION_EXACT;
  if (g->frames != g->rows * g->stride) { set_error("geometry: %d fbank frames != rows %d * stride %d (Emformer.infer size check)", g->frames, g->rows, g->stride); return -1; }
  if (c.n_fft != 800 || c.win != 400 || (c.win & 1)) { set_error("fbank kernel is built for n_fft 800 / win 400"); return -1; }
  if (c.d_model % c.stride || c.d_model != c.n_heads * 64) { set_error("d_model must be n_heads * 64 and divisible by stride"); return -1; }
  if (c.n_mels % 64 || c.d_model % 64 || c.ffn_dim % 64 || c.ctc_hidden % 64) { set_error("GEMM K dims must be multiples of 64"); return -1; }
  if (c.max_batch <= 0 || c.max_sessions <= 0) { set_error("max_batch / max_sessions must be positive"); return -1; }
  return 0;
}

uint64_t weights_count(const Geo& g) {
  const uint64_t d = g.d_model, f = g.ffn, din = d / g.stride;
  uint64_t n = din * g.n_mels;
  n += (uint64_t)g.n_layers * (3 * d * d + 3 * d + d * d + d + 4 * d + f * d + f + d * f + d + 2 * d);
  n += (uint64_t)g.ctc_hidden * d + g.ctc_hidden + (uint64_t)g.vocab * g.ctc_hidden + g.vocab;
  return n;
}

// ------------------------------------------------------------------------------------------ fbank tables
float linspace_f32(float start, float end, int steps, int i) {          // torch.linspace (fp32, symmetric evaluation)
  const float step = (end - start) / (float)(steps - 1);
  return i < steps / 2 ? start + step * (float)i : end - step * (float)(steps - 1 - i);
}

int upload(DevBuf* b, const void* src, size_t bytes) {
  if (b->alloc(bytes)) return -1;
  ASR_CUDA_OK(cudaMemcpy(b->p, src, bytes, cudaMemcpyHostToDevice));
  return 0;
}

int upload_sparse_mel(FbankPlan* pl, const std::vector<float>& fb /*[n_freqs][n_mels]*/, int n_freqs, int n_mels) {
  std::vector<int> start(n_mels), cnt(n_mels), off(n_mels);
  std::vector<float> w;
  for (int m = 0; m < n_mels; ++m) {
    int lo = -1, hi = -1;
    for (int k = 0; k < n_freqs; ++k)
      if (fb[(size_t)k * n_mels + m] != 0.f) { if (lo < 0) lo = k; hi = k; }
    start[m] = lo < 0 ? 0 : lo; cnt[m] = lo < 0 ? 0 : hi - lo + 1; off[m] = (int)w.size();
    for (int k = 0; k < cnt[m]; ++k) w.push_back(fb[(size_t)(start[m] + k) * n_mels + m]);
  }
  if (w.empty()) w.push_back(0.f);
  pl->n_mels = n_mels;
  return upload(&pl->mel_start, start.data(), 4 * n_mels) || upload(&pl->mel_cnt, cnt.data(), 4 * n_mels) ||
         upload(&pl->mel_off, off.data(), 4 * n_mels) || upload(&pl->mel_w, w.data(), 4 * w.size());
}

int upload_twiddles(FbankPlan* pl, int nc) {
  std::vector<float2> tw(nc), w2(nc + 2);
  for (int i = 0; i < nc; ++i) { const double a = -2.0 * M_PI * i / nc; tw[i] = make_float2((float)cos(a), (float)sin(a)); }
  for (int i = 0; i <= nc; ++i) { const double a = -2.0 * M_PI * i / (2.0 * nc); w2[i] = make_float2((float)cos(a), (float)sin(a)); }
  w2[nc + 1] = make_float2(0.f, 0.f);
  pl->nc = nc;
  return upload(&pl->tw, tw.data(), sizeof(float2) * nc) || upload(&pl->w2, w2.data(), sizeof(float2) * (nc + 2));
}

// torchaudio.functional.melscale_fbanks(n_freqs, 0, sr/2, n_mels, sr, norm=None, mel_scale="htk")  (TA:functional.py:492-587)
int build_melspec_plan(AsrEngine* e) {
  const AsrConfig& c = e->cfg;
  FbankPlan* pl = &e->mel128;
  pl->frame_len = c.win;
  std::vector<float> win(c.win);
  for (int n = 0; n < c.win; ++n) win[n] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * n / c.win));        // hann periodic
  const int n_freqs = c.n_fft / 2 + 1, n_mels = c.n_mels;
  const float f_max = (float)(c.sample_rate / 2);
  const double m_min = 2595.0 * log10(1.0 + 0.0 / 700.0), m_max = 2595.0 * log10(1.0 + (double)f_max / 700.0);
  std::vector<float> f_pts(n_mels + 2);
  for (int i = 0; i < n_mels + 2; ++i) {
    const float m = linspace_f32((float)m_min, (float)m_max, n_mels + 2, i);
    f_pts[i] = 700.0f * (powf(10.0f, m / 2595.0f) - 1.0f);
  }
  std::vector<float> fb((size_t)n_freqs * n_mels);
  for (int k = 0; k < n_freqs; ++k) {
    const float fr = linspace_f32(0.f, f_max, n_freqs, k);
    for (int m = 0; m < n_mels; ++m) {
      const float down = (-1.0f * (f_pts[m] - fr)) / (f_pts[m + 1] - f_pts[m]);
      const float up = (f_pts[m + 2] - fr) / (f_pts[m + 2] - f_pts[m + 1]);
      fb[(size_t)k * n_mels + m] = fmaxf(0.f, fminf(down, up));
    }
  }
  return upload(&pl->window, win.data(), 4 * c.win) || upload_twiddles(pl, c.n_fft / 2) || upload_sparse_mel(pl, fb, n_freqs, n_mels);
}

// torchaudio.compliance.kaldi.get_mel_banks / _feature_window_function (TA:compliance/kaldi.py:104-124, :374-449)
int build_kaldi_plan(AsrEngine* e) {
  FbankPlan* pl = &e->kaldi80;
  const int frame_len = 400, padded = 512, n_mels = 80, n_bins = padded / 2;
  pl->frame_len = frame_len;
  std::vector<float> win(frame_len);
  for (int n = 0; n < frame_len; ++n) win[n] = (float)pow(0.5 - 0.5 * cos(2.0 * M_PI * n / (frame_len - 1)), 0.85);   // povey
  const float sr = (float)e->cfg.sample_rate, low = 20.f, high = 0.5f * sr;
  const float bin_w = sr / (float)padded;
  const float mel_low = 1127.0f * logf(1.0f + low / 700.0f), mel_high = 1127.0f * logf(1.0f + high / 700.0f);
  const float delta = (mel_high - mel_low) / (float)(n_mels + 1);
  std::vector<float> fb((size_t)(n_bins + 1) * n_mels, 0.f);
  for (int m = 0; m < n_mels; ++m) {
    const float left = mel_low + (float)m * delta, center = mel_low + (float)(m + 1) * delta, right = mel_low + (float)(m + 2) * delta;
    for (int k = 0; k < n_bins; ++k) {
      const float mel = 1127.0f * logf(1.0f + (bin_w * (float)k) / 700.0f);
      const float up = (mel - left) / (center - left), down = (right - mel) / (right - center);
      fb[(size_t)k * n_mels + m] = fmaxf(0.f, fminf(up, down));
    }
  }
  return upload(&pl->window, win.data(), 4 * frame_len) || upload_twiddles(pl, padded / 2) || upload_sparse_mel(pl, fb, n_bins + 1, n_mels);
}

// ------------------------------------------------------------------------------------------ weights / operands
int make_weight(AsrEngine* e, WeightMat* w, const float* src_f32_dev, bf16*& cursor, int N, int K) {
  const int split = e->geo.split;
  w->N = N; w->K = K; w->ld = split ? 2 * K : K;
  w->w = cursor;
  cursor += round_up((size_t)N * w->ld, 64);
  if (convert_weight(src_f32_dev, w->w, N, K, w->ld, split ? K : 0, e->stream)) return -1;
  const uint32_t boxes[3] = {64, 128, 256};
  for (int i = 0; i < 3; ++i)
    if (make_tmap_bf16_2d(&w->tm[i], w->w, (uint64_t)w->ld, (uint64_t)N, (uint64_t)w->ld, boxes[i])) return -1;
  return 0;
}

int make_operand(AsrEngine* e, Operand* a, int rows, int K) {
  const int split = e->geo.split;
  a->rows = (int)round_up(rows, 128); a->K = K; a->ld = split ? 2 * K : K; a->lo_off = split ? K : 0;
  if (a->buf.alloc((size_t)a->rows * a->ld * sizeof(bf16))) return -1;
  ASR_CUDA_OK(cudaMemset(a->buf.p, 0, a->buf.bytes));
  return make_tmap_bf16_2d(&a->tm, a->buf.p, (uint64_t)a->ld, (uint64_t)a->rows, (uint64_t)a->ld, 128) ||
         make_tmap_bf16_2d(&a->tm_st, a->buf.p, (uint64_t)a->ld, (uint64_t)a->rows, (uint64_t)a->ld, 32);
}

int pick_bn(const AsrEngine* e, int M, int N) {
  // N-tile width by a two-term cost model fitted to the B200 sweeps (profiles/r01_gemm_sweep_16warp_epilogue.txt):
  // time ~ rounds(bn) * (bn + 64), rounds = ceil(tiles / SMs); the +64 is the per-tile fixed cost (pipeline fill, epilogue
  // drain).  Large problems get 256-wide tiles, a single stream gets 64-wide ones so its latency spreads over more SMs.
  // Ragged N (vocab 804) is fine: TMA zero-fills, the epilogue masks.
  const int mt = (M + 127) / 128;
  int best = 64; long best_cost = -1;
  for (int bn : {256, 128, 64}) {
    if (bn > 64 && bn / 2 >= N) continue;
    const long tiles = (long)mt * ((N + bn - 1) / bn);
    const long cost = ((tiles + e->num_sms - 1) / e->num_sms) * (bn + 64);
    if (best_cost < 0 || cost < best_cost) { best = bn; best_cost = cost; }
  }
  return best;
}

cudaEvent_t prof_event(AsrEngine* e) {
  if (!e->prof_pool.empty()) { cudaEvent_t ev = e->prof_pool.back(); e->prof_pool.pop_back(); return ev; }
  cudaEvent_t ev = nullptr;
  cudaEventCreate(&ev);
  return ev;
}

struct ProfScope {       // brackets one kernel launch with events when profiling is on; counts the launch always
  AsrEngine* e; int cat; cudaEvent_t a = nullptr;
  ProfScope(AsrEngine* e_, int cat_) : e(e_), cat(cat_) {
    ++e->launches;
    if (e->prof_on) { a = prof_event(e); cudaEventRecord(a, e->stream); }
  }
  ~ProfScope() {
    if (a) { cudaEvent_t b = prof_event(e); cudaEventRecord(b, e->stream); e->prof_recs.push_back({cat, a, b}); }
  }
};

// ts: tensor maps + bias of the TMA-store epilogue of the pair kernel, when this GEMM has one (bf16 operand outputs)
template <class Epi>
int run_gemm(AsrEngine* e, int cat, const Operand& a, const WeightMat& w, int M, const Epi& epi, const TsMaps* ts = nullptr, int n_streams = 0) {
  const GemmProblem p = make_problem(M, w.N, w.K, e->geo.split);
  ProfScope ps(e, cat);
  int bn = pick_bn(e, M, w.N);
  // CTA pairs (cta_group::2, 256 x 256 tile, half the B-operand traffic per SM: 32 instead of 48 KB per k-block through L2 -> shared
  // memory -> tensor core).  Measured on B200 (profiles/r01_gemm_sweep_*.txt): pair wins at K = 2048 (FFN2 127 vs 138 us) and also
  // at K = 512 for large M (FFN1 150 vs 157 us, QKV 116 vs 121 us); at small M the coupled epilogues lose.
  const int k_eff = w.K * (e->geo.split ? 3 : 1);
  const long pair_tiles = (long)((M + 255) / 256) * (w.N / 256);
  if (bn == 256 && !e->no_pair && w.N % 256 == 0 && ((k_eff >= 1024 && pair_tiles >= e->num_sms / 2) || pair_tiles >= 4 * (e->num_sms / 2))) bn = kPairTile;
  const CUtensorMap& tmB = w.tm[bn == 64 ? 0 : (bn == 128 || bn == kPairTile ? 1 : 2)];
  if (bn == kPairTile && e->tma_store && ts && w.N <= kTsBiasMax) {
    if constexpr (Epi::kStreamTiles) {
      if (e->qkv_ts_ready && n_streams > 0 && &a == &e->a_ln) {
        const StreamTiling stl = make_stream_tiling(n_streams, e->geo.seg_rows, e->geo.rc_rows);
        return gemm_tc<Epi>(e->tm_a_ln_seg, tmB, p, epi, kPairTileQKV, e->num_sms, e->stream, ts, &stl);
      }
    } else if constexpr (Epi::kBf16Rows) {
      if (epi.bf16_rows()) return gemm_tc<Epi>(a.tm, tmB, p, epi, kPairTileTS, e->num_sms, e->stream, ts);
    }
  }
  return gemm_tc<Epi>(a.tm, tmB, p, epi, bn, e->num_sms, e->stream);
}

int run_gemm_ln(AsrEngine* e, int cat, const Operand& a, const WeightMat& w, int M, const LnEpilogue& ep) {
  const GemmProblem p = make_problem(M, w.N, w.K, e->geo.split);
  ProfScope ps(e, cat);
  // CTA pairs when the mainloop is long enough to be L2-bound (FFN2, K = 2048) and there are enough 256-row tiles to fill the clusters
  int shape = 0;
  if (!e->no_pair_ln && w.K * (e->geo.split ? 3 : 1) >= 1024 && (M + 255) / 256 >= e->pair_ln_min_tiles) shape = 1;
  else if ((M + 127) / 128 <= e->quad_ln_max_tiles) shape = 2;       // few row tiles: four column quarters put twice the CTAs to work
  return gemm_ln(a.tm, w.tm[2], w.tm[1], p, ep, shape, e->num_sms, e->stream);
}

// ------------------------------------------------------------------------------------------ the per-step kernel chain
template <typename T>
int run_layers(AsrEngine* e, int n, int n_layers_to_run) {
  const Geo& g = e->geo;
  const int M = n * g.rows, d = g.d_model;
  const int* slots = e->act_slots;
  for (int l = 0; l < n_layers_to_run; ++l) {
    const LayerW& L = e->layers[l];
    T* cache_layer = e->kv_cache.as<T>() + (size_t)l * e->layer_stride;
    EpiQKV<T> eq;
    eq.q = e->q.as<T>(); eq.cache_layer = cache_layer; eq.slot_stride = e->slot_stride; eq.rc = e->rc_kv.as<T>();
    eq.bias = L.bqkv; eq.slots = slots; eq.past_len = e->past_len.as<int>();
    eq.rows = g.rows; eq.seg_rows = g.seg_rows; eq.rc_rows = g.rc_rows; eq.ring = g.ring; eq.d = d;
    eq.qscale = 1.0f / sqrtf((float)(d / g.n_heads));                                   // TA:emformer.py:108
    eq.kv_row0 = (long long)l * (long long)(e->layer_stride / d); eq.slot_rows = (int)(e->slot_stride / d);
    if (run_gemm(e, ASR_PROF_GEMM_QKV, e->a_ln, L.qkv, M, eq, &L.ts_qkv, n)) return -1;

    AttnParams<T> ap;
    ap.q = e->q.as<T>(); ap.cache_layer = cache_layer; ap.slot_stride = e->slot_stride; ap.rc = e->rc_kv.as<T>();
    ap.slots = slots; ap.past_len = e->past_len.as<int>(); ap.out = e->a_attn.buf.as<bf16>(); ap.ld = e->a_attn.ld; ap.lo_off = e->a_attn.lo_off;
    if (e->attn_tma) { ap.h_tm_cache = &e->tm_kv; ap.h_tm_rc = &e->tm_rc; ap.cache_row0 = (long long)l * (long long)(e->layer_stride / d); ap.slot_rows = (long long)(e->slot_stride / d); }
    ap.rows = g.rows; ap.seg_rows = g.seg_rows; ap.rc_rows = g.rc_rows; ap.ring = g.ring; ap.left = g.left; ap.d = d; ap.n_heads = g.n_heads;
    { ProfScope ps(e, ASR_PROF_ATTN); if (attention_launch<T>(ap, n, e->stream)) return -1; }

    const bool last = l == g.n_layers - 1;
    if (e->fused_ln && n >= e->fused_ln_min_streams) {
      // out_proj + residual (pre-LN input) + LN_ff -> x1 (fp32) and the FFN1 operand
      LnEpilogue eo{L.bo, e->x.as<float>(), L.ln_ff_g, L.ln_ff_b, nullptr, nullptr, e->x1.as<float>(), e->a_ln.buf.as<bf16>(), e->a_ln.ld, e->a_ln.lo_off, 0, 0, 0};
      if (run_gemm_ln(e, ASR_PROF_GEMM_OUT, e->a_attn, L.o, M, eo)) return -1;
      EpiOperand e1{e->a_h.buf.as<bf16>(), L.b1, e->a_h.ld, e->a_h.lo_off, ACT_GELU};
      if (run_gemm(e, ASR_PROF_GEMM_FFN1, e->a_ln, L.w1, M, e1, &L.ts_ffn1)) return -1;
      // FFN2 + residual + LN_out -> x (fp32) and LN_in of the next layer -> QKV operand (last layer: segment rows -> CTC operand)
      LnEpilogue e2{L.b2, e->x1.as<float>(), L.ln_out_g, L.ln_out_b, nullptr, nullptr, e->x.as<float>(), nullptr, 0, 0, 0, 0, 0};
      if (last) {
        e2.out_op = e->a_enc.buf.as<bf16>(); e2.op_ld = e->a_enc.ld; e2.op_lo_off = e->a_enc.lo_off;
        e2.f32_normed = 1; e2.compact_rows = g.rows; e2.compact_seg = g.seg_rows;
      } else {
        const LayerW& Nx = e->layers[l + 1];
        e2.g2 = Nx.ln_in_g; e2.b2 = Nx.ln_in_b;
        e2.y_consts = e->no_fuse2 ? nullptr : L.ln_out_consts;
        e2.out_op = e->a_ln.buf.as<bf16>(); e2.op_ld = e->a_ln.ld; e2.op_lo_off = e->a_ln.lo_off;
      }
      if (run_gemm_ln(e, ASR_PROF_GEMM_FFN2, e->a_h, L.w2, M, e2)) return -1;
      continue;
    }
    EpiF32 eo{e->x1.as<float>(), L.bo, e->x.as<float>(), d, d};                          // out_proj + residual (pre-LN input)
    if (run_gemm(e, ASR_PROF_GEMM_OUT, e->a_attn, L.o, M, eo)) return -1;
    { ProfScope ps(e, ASR_PROF_LN); if (ln_to_operand(e->x1.as<float>(), L.ln_ff_g, L.ln_ff_b, e->a_ln.buf.as<bf16>(), e->a_ln.ld, e->a_ln.lo_off, M, d, e->stream)) return -1; }
    EpiOperand e1{e->a_h.buf.as<bf16>(), L.b1, e->a_h.ld, e->a_h.lo_off, ACT_GELU};
    if (run_gemm(e, ASR_PROF_GEMM_FFN1, e->a_ln, L.w1, M, e1, &L.ts_ffn1)) return -1;
    EpiF32 e2{e->x2.as<float>(), L.b2, e->x1.as<float>(), d, d};
    if (run_gemm(e, ASR_PROF_GEMM_FFN2, e->a_h, L.w2, M, e2)) return -1;
    ProfScope ps_ln(e, ASR_PROF_LN);
    if (last) {
      if (ln_out_fused(e->x2.as<float>(), L.ln_out_g, L.ln_out_b, e->x.as<float>(), nullptr, nullptr, e->a_enc.buf.as<bf16>(), e->a_enc.ld,
                       e->a_enc.lo_off, M, d, g.rows, g.seg_rows, e->stream)) return -1;
    } else {
      const LayerW& Nx = e->layers[l + 1];
      if (ln_out_fused(e->x2.as<float>(), L.ln_out_g, L.ln_out_b, e->x.as<float>(), Nx.ln_in_g, Nx.ln_in_b, e->a_ln.buf.as<bf16>(), e->a_ln.ld,
                       e->a_ln.lo_off, M, d, g.rows, g.seg_rows, e->stream)) return -1;
    }
  }
  return 0;
}

int run_fbank_melspec(AsrEngine* e, int n, int pcm_format, float* out_f32, bool to_operand, const Operand* dst = nullptr, const void* pcm = nullptr) {
  const Geo& g = e->geo;
  const FbankPlan& pl = e->mel128;
  const Operand& afb = dst ? *dst : e->a_fb;
  FbankParams P;
  memset(&P, 0, sizeof(P));
  P.pcm = pcm ? pcm : e->act_pcm; P.row_index = pcm ? nullptr : e->act_row_index; P.pcm_is_f32 = pcm_format == ASR_PCM_F32; P.pcm_stride = g.chunk_len; P.n_samples = g.chunk_len;
  P.n_frames = g.frames; P.hop = g.hop; P.frame_len = g.win; P.frame_off = (g.n_fft - g.win) / 2; P.nc = pl.nc; P.kaldi = 0;
  P.in_scale = P.pcm_is_f32 ? 1.0f : 1.0f / 32768.0f;                                  // streaming_server.py:362-363
  P.preemph = 0.f; P.log_floor = 1e-5f;                                                // audio.py:25 clamp(1e-5)
  P.window = pl.window.as<float>(); P.tw = pl.tw.as<float2>(); P.w2 = pl.w2.as<float2>();
  P.mel_start = pl.mel_start.as<int>(); P.mel_cnt = pl.mel_cnt.as<int>(); P.mel_off = pl.mel_off.as<int>(); P.mel_w = pl.mel_w.as<float>();
  P.n_mels = g.n_mels; P.out_f32 = out_f32;
  P.out_op = to_operand ? afb.buf.as<bf16>() : nullptr; P.op_ld = afb.ld; P.op_lo_off = afb.lo_off;
  ProfScope ps(e, ASR_PROF_FBANK);
  return fbank_launch(P, n, e->stream);
}

BeamParams beam_params(AsrEngine* e, int n) {
  BeamParams P;
  P.logits = e->logits.as<float>(); P.row_stat = e->bm_row_stat.as<float>(); P.cand_tok = e->bm_cand_tok.as<int>(); P.cand_lp = e->bm_cand_lp.as<float>();
  P.slots = e->act_slots;
  P.n = n; P.seg_rows = e->geo.seg_rows; P.vocab = e->geo.vocab; P.beam = e->beam; P.cand_k = e->cand_k; P.max_len = BEAM_MAX_LEN - 1;
  P.n_beam = e->bm_n.as<int>(); P.cur = e->bm_cur.as<int>(); P.len = e->bm_len.as<int>(); P.last = e->bm_last.as<int>();
  P.pb = e->bm_pb.as<float>(); P.pnb = e->bm_pnb.as<float>(); P.hash = e->bm_hash.as<unsigned long long>(); P.tokens = e->bm_tokens.as<int16_t>();
  P.out_tokens = e->d_beam_tok.as<int16_t>(); P.out_len = e->d_beam_len.as<int>(); P.out_score = e->d_beam_score.as<float>(); P.out_flags = e->d_flags.as<int>();
  return P;
}

// log_softmax + argmax + incremental greedy collapse (+ the per-frame extension candidates) over e->logits, then the prefix beam search
int run_decode(AsrEngine* e, int n, bool want_logprobs) {
  const Geo& g = e->geo;
  CtcParams cp;
  cp.logits = e->logits.as<float>(); cp.vocab = g.vocab; cp.seg_rows = g.seg_rows; cp.slots = e->act_slots;
  cp.prev_id = e->prev_id.as<int>(); cp.n_frames = e->n_frames.as<int>(); cp.last_tok_frame = e->last_tok.as<int>(); cp.past_len = e->past_len.as<int>();
  cp.seg_has_text = e->seg_has_text.as<int>(); cp.silent_mask = e->silent_mask.as<uint32_t>();
  cp.argmax_ids = e->d_argmax.as<int>(); cp.new_tokens = e->d_newtok.as<int>(); cp.n_new = e->d_nnew.as<int>();
  cp.blank_frames = e->d_blank.as<int>(); cp.has_token = e->d_hastok.as<int>(); cp.has_text = e->d_hastext.as<int>(); cp.flags = e->d_flags.as<int>();
  cp.logprobs = want_logprobs ? e->d_logprobs.as<float>() : nullptr;
  if (e->beam > 0) { cp.cand_k = e->cand_k; cp.cand_tok = e->bm_cand_tok.as<int>(); cp.cand_lp = e->bm_cand_lp.as<float>(); cp.row_stat = e->bm_row_stat.as<float>(); }
  { ProfScope ps(e, ASR_PROF_CTC); if (ctc_greedy_launch(cp, n, e->stream)) return -1; }
  if (e->beam > 0) { ProfScope ps(e, ASR_PROF_BEAM); if (beam_launch(beam_params(e, n), e->stream)) return -1; }
  return 0;
}

int run_pipeline(AsrEngine* e, int n, int pcm_format, int n_layers_to_run, bool with_ctc, bool want_logprobs) {
  pdl_set_active(n <= e->pdl_max_streams);
  const Geo& g = e->geo;
  const Operand* afb = &e->a_fb;
  if (e->act_fb_stage) {                                     // the features exist already (engine_prestage): keep the rows that run
    if (e->act_fb_identity) afb = &e->a_fb_stage;
    else {
      ProfScope ps(e, ASR_PROF_FBANK);
      if (gather_blocks_launch(e->a_fb_stage.buf.p, e->a_fb.buf.p, e->act_row_index, n, (size_t)g.frames * e->a_fb.ld * sizeof(bf16), e->stream)) return -1;
    }
  } else if (run_fbank_melspec(e, n, pcm_format, nullptr, true)) return -1;
  // input_linear (encoder.py:142, no bias); its [n*frames, d/stride] output *is* the time-reduced [n*rows, d] (common.py:118-119)
  EpiF32 ein{e->x.as<float>(), nullptr, nullptr, g.d_model / g.stride, g.d_model / g.stride};
  if (run_gemm(e, ASR_PROF_GEMM_IN, *afb, e->w_in, n * g.frames, ein)) return -1;
  const int M = n * g.rows;
  { ProfScope ps(e, ASR_PROF_LN);
    if (ln_to_operand(e->x.as<float>(), e->layers[0].ln_in_g, e->layers[0].ln_in_b, e->a_ln.buf.as<bf16>(), e->a_ln.ld, e->a_ln.lo_off, M, g.d_model,
                      e->stream)) return -1; }
  if (g.split ? run_layers<float>(e, n, n_layers_to_run) : run_layers<bf16>(e, n, n_layers_to_run)) return -1;
  if (!with_ctc) return 0;
  const int Mc = n * g.seg_rows;
  EpiOperand ec1{e->a_ctc.buf.as<bf16>(), e->ctc_b1, e->a_ctc.ld, e->a_ctc.lo_off, ACT_SILU};    // decoder.py:67
  if (run_gemm(e, ASR_PROF_GEMM_CTC1, e->a_enc, e->ctc1, Mc, ec1, &e->ts_ctc1)) return -1;
  EpiF32 ec2{e->logits.as<float>(), e->ctc_b2, nullptr, g.vocab, g.vocab};                        // decoder.py:68
  if (run_gemm(e, ASR_PROF_GEMM_CTC2, e->a_ctc, e->ctc2, Mc, ec2)) return -1;
  return run_decode(e, n, want_logprobs);
}

void drop_step_graphs(AsrEngine* e) {
  for (auto& kv : e->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  e->graphs.clear();
}

// The full per-step chain, replayed from a CUDA graph when the batch is small enough to be launch-bound.
int run_step_chain(AsrEngine* e, int n, int pcm_format, bool want_logprobs) {
  const Geo& g = e->geo;
  if (!e->use_graphs || e->prof_on || n > e->graph_max_streams || e->act_fb_stage) return run_pipeline(e, n, pcm_format, g.n_layers, true, want_logprobs);
  const char* asm_env = getenv("ASR_B200_ATTN_STREAM_MIN");           // read per launch by the attention dispatch: part of what a graph froze
  const uint64_t key = ((uint64_t)n << 8) | ((uint64_t)(e->act_slots == e->d_slots2.as<int>()) << 0) | ((uint64_t)(pcm_format == ASR_PCM_F32) << 1) |
                       ((uint64_t)want_logprobs << 2) | ((uint64_t)(e->beam > 0) << 3) | ((uint64_t)(e->act_row_index != nullptr) << 4) | ((uint64_t)((asm_env ? atoi(asm_env) : 148) & 0xffff) << 32);
  auto it = e->graphs.find(key);
  if (it == e->graphs.end()) {
    if ((int)e->graphs.size() >= e->graph_max_entries) return run_pipeline(e, n, pcm_format, g.n_layers, true, want_logprobs);
    it = e->graphs.emplace(key, AsrEngine::StepGraph()).first;
  }
  AsrEngine::StepGraph& sg = it->second;
  if (sg.exec) {
    ASR_CUDA_OK(cudaGraphLaunch(sg.exec, e->stream));
    e->launches += sg.launches;
    return 0;
  }
  if (sg.seen++ == 0 || sg.seen < 0) return run_pipeline(e, n, pcm_format, g.n_layers, true, want_logprobs);   // first sight: direct (also warms every per-device cache)
  // second sight: capture, instantiate, launch
  const uint64_t l0 = e->launches;
  if (cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); sg.seen = -1000000; return run_pipeline(e, n, pcm_format, g.n_layers, true, want_logprobs); }
  const int rc = run_pipeline(e, n, pcm_format, g.n_layers, true, want_logprobs);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
  const uint64_t captured = e->launches - l0;
  e->launches = l0;
  if (rc || ce != cudaSuccess || !graph) {
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    sg.seen = -1000000;                                   // never try this key again
    if (rc) return -1;
    return run_pipeline(e, n, pcm_format, g.n_layers, true, want_logprobs);
  }
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess || !exec) { cudaGetLastError(); sg.seen = -1000000; return run_pipeline(e, n, pcm_format, g.n_layers, true, want_logprobs); }
  sg.exec = exec; sg.launches = captured;
  ASR_CUDA_OK(cudaGraphLaunch(sg.exec, e->stream));
  e->launches += sg.launches;
  return 0;
}

int check_step_args(AsrEngine* e, int n, const int32_t* slots) {
  if (!e) { set_error("null engine"); return -1; }
  if (n < 0 || n > e->cfg.max_batch) { set_error("n = %d outside [0, max_batch = %d]", n, e->cfg.max_batch); return -1; }
  if (n && !slots) { set_error("null slots"); return -1; }
  if (++e->stamp_epoch == 0) { std::fill(e->slot_stamp.begin(), e->slot_stamp.end(), 0u); e->stamp_epoch = 1; }
  for (int i = 0; i < n; ++i) {
    if (slots[i] < 0 || slots[i] >= e->cfg.max_sessions || !e->slot_open[slots[i]]) { set_error("slot %d (index %d) is not an open session", slots[i], i); return -1; }
    // a session may appear at most once per step: two rows of one slot would race on its K/V ring rows and greedy carry
    if (e->slot_stamp[slots[i]] == e->stamp_epoch) { set_error("slot %d appears twice in one step (index %d)", slots[i], i); return -1; }
    e->slot_stamp[slots[i]] = e->stamp_epoch;
  }
  return 0;
}

void mark_inflight(AsrEngine* e, int b, int n, const int32_t* slots) {
  e->pend[b].slots.assign(slots, slots + n);
  for (int i = 0; i < n; ++i) ++e->slot_inflight[slots[i]];
}
void clear_inflight(AsrEngine* e, int b) {
  for (int32_t s : e->pend[b].slots) if (e->slot_inflight[s]) --e->slot_inflight[s];
  e->pend[b].slots.clear();
}
bool ticket_pending(const AsrEngine* e) { return e->pend[0].active || e->pend[1].active; }

size_t pcm_bytes(const AsrEngine* e, int n, int fmt) { return (size_t)n * e->geo.chunk_len * (fmt == ASR_PCM_F32 ? 4 : 2); }

size_t slots_off(const AsrEngine* e) { return round_up(pcm_bytes(e, e->cfg.max_batch, ASR_PCM_F32), 256); }
void* dev_pcm(AsrEngine* e, int b) { return b ? e->d_pcm2.p : e->d_pcm.p; }
int* dev_slots(AsrEngine* e, int b) { return b ? e->d_slots2.as<int>() : e->d_slots.as<int>(); }
void use_buffer(AsrEngine* e, int b) { e->act_pcm = dev_pcm(e, b); e->act_slots = dev_slots(e, b); e->act_row_index = nullptr; }

// host -> pinned (skipped when the caller assembled the batch in the pinned buffer) -> device, on `st`
int stage_inputs(AsrEngine* e, int b, int n, const int32_t* slots, const void* pcm, int fmt, cudaStream_t st) {
  if (check_step_args(e, n, slots)) return -1;
  if (n && !pcm) { set_error("null pcm"); return -1; }
  if (fmt != ASR_PCM_I16 && fmt != ASR_PCM_F32) { set_error("bad pcm_format %d", fmt); return -1; }
  if (!n) return 0;
  ASR_CUDA_OK(cudaSetDevice(e->device));
  const size_t pb = pcm_bytes(e, n, fmt);
  uint8_t* hs = reinterpret_cast<uint8_t*>(e->h_buf[b]);
  const uint8_t* src = reinterpret_cast<const uint8_t*>(pcm);
  const bool pinned_src = (src == reinterpret_cast<uint8_t*>(e->h_buf[0]) || src == reinterpret_cast<uint8_t*>(e->h_buf[1]));
  if (!pinned_src) { memcpy(hs, pcm, pb); src = hs; }        // asr_pinned_pcm callers filled a staging buffer themselves
  memcpy(hs + slots_off(e), slots, 4 * (size_t)n);
  ASR_CUDA_OK(cudaMemcpyAsync(dev_pcm(e, b), src, pb, cudaMemcpyHostToDevice, st));
  ASR_CUDA_OK(cudaMemcpyAsync(dev_slots(e, b), hs + slots_off(e), 4 * (size_t)n, cudaMemcpyHostToDevice, st));
  return 0;
}

struct OutItem { const DevBuf* src; size_t bytes, hoff; int field; };

// fixed layout of the pinned result area of one staging buffer
std::vector<OutItem> out_layout(AsrEngine* e, int n, bool want_lp) {
  const Geo& g = e->geo;
  const size_t nS = (size_t)n * g.seg_rows, B = e->cfg.max_batch, BS = B * g.seg_rows;
  std::vector<OutItem> v;
  size_t off = 0;
  auto add = [&](const DevBuf& src, size_t bytes, size_t cap, int field, bool on) {
    if (on && bytes) v.push_back({&src, bytes, off, field});
    off += round_up(cap, 256);
  };
  add(e->d_argmax, 4 * nS, 4 * BS, 0, true);
  add(e->d_newtok, 4 * nS, 4 * BS, 1, true);
  add(e->d_nnew, 4 * (size_t)n, 4 * B, 2, true);
  add(e->d_blank, 4 * (size_t)n, 4 * B, 3, true);
  add(e->d_hastok, 4 * (size_t)n, 4 * B, 4, true);
  add(e->d_hastext, 4 * (size_t)n, 4 * B, 9, true);
  add(e->d_flags, 4 * (size_t)n, 4 * B, 10, true);
  add(e->d_beam_tok, 2 * (size_t)n * BEAM_MAX_LEN, 2 * B * BEAM_MAX_LEN, 6, e->beam > 0);
  add(e->d_beam_len, 4 * (size_t)n, 4 * B, 7, e->beam > 0);
  add(e->d_beam_score, 4 * (size_t)n, 4 * B, 8, e->beam > 0);
  add(e->d_logprobs, 4 * nS * g.vocab, 4 * BS * g.vocab, 5, want_lp);
  return v;
}

void* out_field(const AsrStepOut* o, int f) {
  switch (f) {
    case 0: return o->argmax_ids; case 1: return o->new_tokens; case 2: return o->n_new; case 3: return o->blank_frames;
    case 4: return o->has_token; case 5: return o->logprobs; case 6: return o->beam_tokens; case 7: return o->beam_len; case 8: return o->beam_score;
    case 9: return o->has_text; case 10: return o->flags;
  }
  return nullptr;
}

int enqueue_d2h(AsrEngine* e, int b, int n, bool want_lp) {
  uint8_t* ho = reinterpret_cast<uint8_t*>(e->h_buf[b]) + e->h_out_off;
  for (auto& it : out_layout(e, n, want_lp))
    ASR_CUDA_OK(cudaMemcpyAsync(ho + it.hoff, it.src->p, it.bytes, cudaMemcpyDeviceToHost, e->stream));
  return 0;
}

void deliver(AsrEngine* e, int b, int n, bool want_lp, const AsrStepOut* out) {
  if (!out) return;
  const uint8_t* ho = reinterpret_cast<const uint8_t*>(e->h_buf[b]) + e->h_out_off;
  const auto items = out_layout(e, n, want_lp);
  const int32_t* beam_len = nullptr;
  for (auto& it : items) if (it.field == 7) beam_len = reinterpret_cast<const int32_t*>(ho + it.hoff);
  for (auto& it : items) {
    void* dst = out_field(out, it.field);
    if (!dst) continue;
    if (it.field == 6 && beam_len) {                        // hypotheses: only the valid prefix of every row (2 KB rows, mostly short)
      const int16_t* src = reinterpret_cast<const int16_t*>(ho + it.hoff);
      int16_t* d = reinterpret_cast<int16_t*>(dst);
      for (int i = 0; i < n; ++i) memcpy(d + (size_t)i * BEAM_MAX_LEN, src + (size_t)i * BEAM_MAX_LEN, 2 * (size_t)std::min(beam_len[i], (int32_t)BEAM_MAX_LEN));
    } else {
      memcpy(dst, ho + it.hoff, it.bytes);
    }
  }
}

// enqueue one step on buffer b: H2D on the copy stream, kernels + D2H on the compute stream
int submit_step(AsrEngine* e, int n, const int32_t* slots, const void* pcm, int fmt, bool want_lp, int* ticket) {
  const int b = e->cur_buf;
  if (e->pend[b].active) { set_error("two steps are already in flight: asr_collect the oldest ticket first"); return -1; }
  e->prestaged_rows[b] = -1; const int fb_was_staged = e->fb_staged[b]; e->fb_staged[b] = 0; (void)fb_was_staged;
  const auto t0 = std::chrono::steady_clock::now();
  if (stage_inputs(e, b, n, slots, pcm, fmt, e->copy_stream)) return -1;
  if (n) {
    ASR_CUDA_OK(cudaEventRecord(e->ev_in[b], e->copy_stream));
    ASR_CUDA_OK(cudaStreamWaitEvent(e->stream, e->ev_in[b], 0));
    use_buffer(e, b);
    if (e->ev_t0[b]) cudaEventRecord(e->ev_t0[b], e->stream);
    if (run_step_chain(e, n, fmt, want_lp)) return -1;
    if (enqueue_d2h(e, b, n, want_lp)) return -1;
    if (e->ev_t1[b]) cudaEventRecord(e->ev_t1[b], e->stream);
    ASR_CUDA_OK(cudaEventRecord(e->ev_done[b], e->stream));
  }
  e->pend[b].n = n; e->pend[b].want_lp = want_lp; e->pend[b].active = 1; e->pend[b].t0 = t0;
  mark_inflight(e, b, n, slots);
  e->cur_buf ^= 1;
  if (ticket) *ticket = b;
  return 0;
}

// Same as submit_step, but the batch is assembled by the GPU: the chunks are read straight out of the sessions' pinned audio rings
// (zero-copy over PCIe, gather_rings_kernel on the copy stream) — no host-side gather, no staging copy.
int submit_rings(AsrEngine* e, int n, const int32_t* slots, const int16_t* base, int64_t row_stride, const int32_t* rows, const int64_t* offsets,
                 bool want_lp, int* ticket) {
  const int b = e->cur_buf;
  if (e->pend[b].active) { set_error("two steps are already in flight: asr_collect the oldest ticket first"); return -1; }
  if (check_step_args(e, n, slots)) return -1;
  if (n && (!base || !rows || !offsets)) { set_error("null argument"); return -1; }
  const auto t0 = std::chrono::steady_clock::now();
  if (n) {
    ASR_CUDA_OK(cudaSetDevice(e->device));
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, base) != cudaSuccess || pa.type != cudaMemoryTypeHost) {
      cudaGetLastError();
      set_error("asr_submit_rings: the audio rings must live in pinned host memory (asr_host_alloc)"); return -1;
    }
    const int16_t* base_dev = reinterpret_cast<const int16_t*>(pa.devicePointer);
    uint8_t* hs = reinterpret_cast<uint8_t*>(e->h_buf[b]);
    long long* h_off = reinterpret_cast<long long*>(hs);                         // the PCM area of the staging buffer is unused on this path
    for (int i = 0; i < n; ++i) h_off[i] = (long long)rows[i] * row_stride + offsets[i];
    memcpy(hs + slots_off(e), slots, 4 * (size_t)n);
    ASR_CUDA_OK(cudaMemcpyAsync(e->d_src_off[b].p, h_off, 8 * (size_t)n, cudaMemcpyHostToDevice, e->copy_stream));
    ASR_CUDA_OK(cudaMemcpyAsync(dev_slots(e, b), hs + slots_off(e), 4 * (size_t)n, cudaMemcpyHostToDevice, e->copy_stream));
    if (gather_rings_launch(base_dev, e->d_src_off[b].as<long long>(), reinterpret_cast<int16_t*>(dev_pcm(e, b)), n, e->geo.chunk_len, e->copy_stream)) return -1;
    ASR_CUDA_OK(cudaEventRecord(e->ev_in[b], e->copy_stream));
    ASR_CUDA_OK(cudaStreamWaitEvent(e->stream, e->ev_in[b], 0));
    use_buffer(e, b);
    if (e->ev_t0[b]) cudaEventRecord(e->ev_t0[b], e->stream);
    if (run_step_chain(e, n, ASR_PCM_I16, want_lp)) return -1;
    if (enqueue_d2h(e, b, n, want_lp)) return -1;
    if (e->ev_t1[b]) cudaEventRecord(e->ev_t1[b], e->stream);
    ASR_CUDA_OK(cudaEventRecord(e->ev_done[b], e->stream));
  }
  e->pend[b].n = n; e->pend[b].want_lp = want_lp; e->pend[b].active = 1; e->pend[b].t0 = t0;
  mark_inflight(e, b, n, slots);
  e->cur_buf ^= 1;
  if (ticket) *ticket = b;
  return 0;
}

int collect_step(AsrEngine* e, int ticket, const AsrStepOut* out) {
  if (ticket < 0 || ticket > 1 || !e->pend[ticket].active) { set_error("asr_collect: ticket %d is not in flight", ticket); return -1; }
  auto& pd = e->pend[ticket];
  if (pd.n) {
    const cudaError_t ce = cudaSetDevice(e->device) == cudaSuccess ? cudaEventSynchronize(e->ev_done[ticket]) : cudaErrorInvalidDevice;
    if (ce != cudaSuccess) {                                // the ticket is gone either way: its sessions must not stay locked
      pd.active = 0; clear_inflight(e, ticket);
      set_error("asr_collect: step failed on the device: %s", cudaGetErrorString(ce));
      return -1;
    }
    if (e->ev_t0[ticket] && e->ev_t1[ticket]) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, e->ev_t0[ticket], e->ev_t1[ticket]) == cudaSuccess) { e->pipe_gpu_ms += ms; ++e->pipe_gpu_n; } else cudaGetLastError();
    }
    deliver(e, ticket, pd.n, pd.want_lp, out);
  }
  pd.active = 0;
  clear_inflight(e, ticket);
  return 0;
}

// legacy split path (asr_stage / asr_run_staged / asr_fetch) works on buffer 0 and the compute stream only
int fetch_outputs(AsrEngine* e, int n, const AsrStepOut* out, bool) {
  if (n && out) {
    if (enqueue_d2h(e, 0, n, out->logprobs != nullptr)) return -1;
  }
  ASR_CUDA_OK(cudaStreamSynchronize(e->stream));
  if (n && out) deliver(e, 0, n, out->logprobs != nullptr, out);
  return 0;
}

void record_step(AsrEngine* e, int n, double ms) {
  ++e->steps; e->stream_chunks += n;
  if (e->step_ms.size() < 4096) e->step_ms.push_back((float)ms);
  else { e->step_ms[e->step_ms_pos] = (float)ms; e->step_ms_pos = (e->step_ms_pos + 1) % 4096; }
}

// Stream-ordered, asynchronous reset of the listed sessions (endpoint / open): takes effect after every step already enqueued and
// never waits for the device — the pinned slot list is one of kResetRing, so only that list's previous H2D copy is waited for.
int reset_slots_async(AsrEngine* e, int n, const int32_t* slots) {
  if (n <= 0) return 0;
  const int r = e->reset_pos;
  e->reset_pos = (r + 1) % AsrEngine::kResetRing;
  int32_t* hs = e->h_reset + (size_t)r * e->cfg.max_sessions;
  int* ds = e->d_reset.as<int>() + (size_t)r * e->cfg.max_sessions;
  ASR_CUDA_OK(cudaEventSynchronize(e->ev_reset[r]));
  memcpy(hs, slots, 4 * (size_t)n);
  ASR_CUDA_OK(cudaMemcpyAsync(ds, hs, 4 * (size_t)n, cudaMemcpyHostToDevice, e->stream));
  ASR_CUDA_OK(cudaEventRecord(e->ev_reset[r], e->stream));
  if (reset_slots_launch(ds, n, e->past_len.as<int>(), e->n_frames.as<int>(), e->prev_id.as<int>(), e->last_tok.as<int>(), e->seg_has_text.as<int>(), e->stream)) return -1;
  if (e->beam > 0 && beam_reset_many_launch(beam_params(e, 0), ds, n, e->stream)) return -1;
  return 0;
}

void destroy_engine(AsrEngine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (auto& kv : e->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  e->graphs.clear();
  DevBuf* bufs[] = {&e->w_f32, &e->w_bf16, &e->ln_consts, &e->d_pcm, &e->d_slots, &e->d_pcm2, &e->d_slots2, &e->d_src_off[0], &e->d_src_off[1], &e->d_row_index[0], &e->d_row_index[1], &e->x, &e->x1, &e->x2, &e->q, &e->rc_kv, &e->logits, &e->fb_f32,
                    &e->a_fb.buf, &e->a_fb_stage.buf, &e->a_ln.buf, &e->a_attn.buf, &e->a_h.buf, &e->a_enc.buf, &e->a_ctc.buf, &e->kv_cache, &e->past_len,
                    &e->prev_id, &e->n_frames, &e->last_tok, &e->seg_has_text, &e->silent_mask, &e->d_argmax, &e->d_newtok, &e->d_nnew, &e->d_blank, &e->d_hastok, &e->d_hastext, &e->d_flags, &e->d_logprobs,
                    &e->bm_n, &e->bm_cur, &e->bm_len, &e->bm_last, &e->bm_pb, &e->bm_pnb, &e->bm_hash, &e->bm_tokens, &e->d_beam_tok, &e->d_beam_len, &e->d_beam_score, &e->bm_cand_tok, &e->bm_cand_lp, &e->bm_row_stat};
  for (DevBuf* b : bufs) b->free();
  for (FbankPlan* pl : {&e->mel128, &e->kaldi80}) {
    pl->window.free(); pl->tw.free(); pl->w2.free(); pl->mel_start.free(); pl->mel_cnt.free(); pl->mel_off.free(); pl->mel_w.free();
  }
  for (auto& r : e->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto ev : e->prof_pool) cudaEventDestroy(ev);
  for (int i = 0; i < 2; ++i) { if (e->h_buf[i]) cudaFreeHost(e->h_buf[i]); if (e->ev_in[i]) cudaEventDestroy(e->ev_in[i]); if (e->ev_done[i]) cudaEventDestroy(e->ev_done[i]); if (e->ev_pre[i]) cudaEventDestroy(e->ev_pre[i]);
                                if (e->ev_t0[i]) cudaEventDestroy(e->ev_t0[i]); if (e->ev_t1[i]) cudaEventDestroy(e->ev_t1[i]); }
  if (e->h_reset) cudaFreeHost(e->h_reset);
  for (auto ev : e->ev_reset) if (ev) cudaEventDestroy(ev);
  e->d_reset.free();
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int create_engine(const AsrConfig* cfg, const float* weights, uint64_t n_floats, int device, AsrEngine** out) {
  if (!cfg || !weights || !out) { set_error("null argument"); return -1; }
  Geo g;
  if (fill_geo(*cfg, &g)) return -1;
  if (weights_count(g) != n_floats) { set_error("weights blob has %llu floats, config needs %llu", (unsigned long long)n_floats, (unsigned long long)weights_count(g)); return -1; }
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { set_error("no CUDA device: the B200 path has no CPU fallback"); return -1; }
  if (device < 0 || device >= n_dev) { set_error("device %d out of range (%d devices)", device, n_dev); return -1; }
  ASR_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  ASR_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) { set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor); return -1; }

  AsrEngine* e = new AsrEngine();
  e->cfg = *cfg; e->geo = g; e->device = device; e->num_sms = prop.multiProcessorCount;
  const char* np_ = getenv("ASR_B200_NO_PAIR_GEMM");
  e->no_pair = np_ && np_[0] == '1';
  if (const char* nf = getenv("ASR_B200_NO_FUSED_LN")) e->fused_ln = !(nf[0] == '1');
  if (cfg->d_model != 512) e->fused_ln = 0;
  if (const char* pl = getenv("ASR_B200_NO_PAIR_LN")) e->no_pair_ln = pl[0] == '1';
  if (const char* nf2 = getenv("ASR_B200_NO_LN_FUSE2")) e->no_fuse2 = nf2[0] == '1';
  if (const char* pt = getenv("ASR_B200_PAIR_LN_MIN_TILES")) e->pair_ln_min_tiles = atoi(pt);
  if (const char* nt = getenv("ASR_B200_NO_TMA_STORE")) e->tma_store = !(nt[0] == '1');
  if (const char* ns = getenv("ASR_B200_NO_SPEC_FBANK")) e->spec_fbank = !(ns[0] == '1');
  if (const char* ug = getenv("ASR_B200_GRAPHS")) e->use_graphs = ug[0] == '1';
  if (const char* gm = getenv("ASR_B200_GRAPH_MAX_STREAMS")) e->graph_max_streams = atoi(gm);
  if (const char* qt = getenv("ASR_B200_QUAD_LN_MAX_TILES")) e->quad_ln_max_tiles = atoi(qt);
  if (const char* fm = getenv("ASR_B200_FUSED_LN_MIN_STREAMS")) e->fused_ln_min_streams = atoi(fm);
  if (const char* pm = getenv("ASR_B200_PDL_MAX_STREAMS")) e->pdl_max_streams = atoi(pm);
  int rc = -1;
  do {
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); break; }
    if (e->w_f32.alloc(n_floats * 4)) break;
    if (cudaMemcpy(e->w_f32.p, weights, n_floats * 4, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("weights H2D failed"); break; }
    // ---- carve the fp32 blob and convert the matrices to bf16 (hi|lo) B operands
    const int d = g.d_model, f = g.ffn, din = d / g.stride, kmul = g.split ? 2 : 1;
    size_t bf_elems = round_up((size_t)din * g.n_mels * kmul, 64);
    bf_elems += (size_t)g.n_layers * (round_up((size_t)3 * d * d * kmul, 64) + round_up((size_t)d * d * kmul, 64) + 2 * round_up((size_t)f * d * kmul, 64));
    bf_elems += round_up((size_t)g.ctc_hidden * d * kmul, 64) + round_up((size_t)g.vocab * g.ctc_hidden * kmul, 64);
    if (e->w_bf16.alloc(bf_elems * sizeof(bf16))) break;
    bf16* cur = e->w_bf16.as<bf16>();
    const float* wp = e->w_f32.as<float>();
    auto take = [&](size_t n) { const float* p = wp; wp += n; return p; };
    bool ok = true;
    ok = ok && !make_weight(e, &e->w_in, take((size_t)din * g.n_mels), cur, din, g.n_mels);
    e->layers.resize(g.n_layers);
    for (int l = 0; l < g.n_layers && ok; ++l) {
      LayerW& L = e->layers[l];
      ok = ok && !make_weight(e, &L.qkv, take((size_t)3 * d * d), cur, 3 * d, d);
      L.bqkv = take(3 * d);
      ok = ok && !make_weight(e, &L.o, take((size_t)d * d), cur, d, d);
      L.bo = take(d);
      L.ln_in_g = take(d); L.ln_in_b = take(d); L.ln_ff_g = take(d); L.ln_ff_b = take(d);
      ok = ok && !make_weight(e, &L.w1, take((size_t)f * d), cur, f, d);
      L.b1 = take(f);
      ok = ok && !make_weight(e, &L.w2, take((size_t)d * f), cur, d, f);
      L.b2 = take(d);
      L.ln_out_g = take(d); L.ln_out_b = take(d);
    }
    if (ok && d == 512) {
      // constants that let gemm_ln derive the statistics of LN_out's output in one pass (LnEpilogue::y_consts), from the host blob
      const int per = (d + 3 * (d / 32) + 2 + 3) / 4 * 4;          // float4 loads: keep every layer's block 16-byte aligned
      std::vector<float> hc((size_t)g.n_layers * per);
      for (int l = 0; l < g.n_layers; ++l) {
        const float* gam = weights + (e->layers[l].ln_out_g - e->w_f32.as<float>());
        const float* bet = weights + (e->layers[l].ln_out_b - e->w_f32.as<float>());
        float* c = hc.data() + (size_t)l * per;
        double mb = 0.0, vb = 0.0;
        for (int j = 0; j < d; ++j) mb += bet[j];
        mb /= d;
        for (int j = 0; j < d; ++j) vb += (bet[j] - mb) * (bet[j] - mb);
        vb /= d;
        for (int ch = 0; ch < d / 32; ++ch) {
          double g1 = 0.0, g2 = 0.0, g3 = 0.0;
          for (int j = ch * 32; j < ch * 32 + 32; ++j) {
            const float gc = (float)(gam[j] * (bet[j] - mb));
            c[j] = gc;
            g1 += gam[j]; g2 += (double)gam[j] * gam[j]; g3 += gc;
          }
          c[d + ch] = (float)g1; c[d + d / 32 + ch] = (float)g2; c[d + 2 * (d / 32) + ch] = (float)g3;
        }
        c[d + 3 * (d / 32)] = (float)mb; c[d + 3 * (d / 32) + 1] = (float)vb;
      }
      if (upload(&e->ln_consts, hc.data(), hc.size() * 4)) break;
      for (int l = 0; l < g.n_layers; ++l) e->layers[l].ln_out_consts = e->ln_consts.as<float>() + (size_t)l * per;
    }
    ok = ok && !make_weight(e, &e->ctc1, take((size_t)g.ctc_hidden * d), cur, g.ctc_hidden, d);
    e->ctc_b1 = take(g.ctc_hidden);
    ok = ok && !make_weight(e, &e->ctc2, take((size_t)g.vocab * g.ctc_hidden), cur, g.vocab, g.ctc_hidden);
    e->ctc_b2 = take(g.vocab);
    if (!ok) break;
    if (build_melspec_plan(e) || build_kaldi_plan(e)) break;

    // ---- activations
    const int B = cfg->max_batch, M = B * g.rows, Mc = B * g.seg_rows;
    const size_t esz = g.split ? 4 : 2;
    if (e->d_pcm.alloc(pcm_bytes(e, B, ASR_PCM_F32)) || e->d_slots.alloc(4 * (size_t)B) || e->d_pcm2.alloc(pcm_bytes(e, B, ASR_PCM_F32)) || e->d_slots2.alloc(4 * (size_t)B) || e->d_src_off[0].alloc(8 * (size_t)B) || e->d_src_off[1].alloc(8 * (size_t)B) || e->d_row_index[0].alloc(4 * (size_t)B) || e->d_row_index[1].alloc(4 * (size_t)B) || e->x.alloc(4 * (size_t)M * d) || e->x1.alloc(4 * (size_t)M * d) ||
        e->x2.alloc(4 * (size_t)M * d) || e->q.alloc(4 * (size_t)M * d) || e->rc_kv.alloc(esz * (size_t)B * 2 * g.rc_rows * d) ||
        e->logits.alloc(4 * (size_t)Mc * g.vocab) || e->d_logprobs.alloc(4 * (size_t)Mc * g.vocab) || e->d_argmax.alloc(4 * (size_t)Mc) ||
        e->d_newtok.alloc(4 * (size_t)Mc) || e->d_nnew.alloc(4 * (size_t)B) || e->d_blank.alloc(4 * (size_t)B) || e->d_hastok.alloc(4 * (size_t)B) ||
        e->d_hastext.alloc(4 * (size_t)B) || e->d_flags.alloc(4 * (size_t)B)) break;
    if (make_operand(e, &e->a_fb, B * g.frames, g.n_mels) || make_operand(e, &e->a_ln, M, d) || make_operand(e, &e->a_attn, M, d) ||
        make_operand(e, &e->a_h, M, f) || make_operand(e, &e->a_enc, Mc, d) || make_operand(e, &e->a_ctc, Mc, g.ctc_hidden)) break;
    // ---- sessions
    // Layer-major: the kernels of one layer touch one contiguous slab (max_sessions x 96 KB) instead of one 96 KB piece out of every
    // session's 1.97 MB block (TLB reach, DRAM page locality).
    const size_t S = cfg->max_sessions;
    e->slot_stride = (size_t)2 * g.ring * d;
    e->layer_stride = S * e->slot_stride;
    if (e->kv_cache.alloc(esz * S * g.n_layers * 2 * g.ring * d) || e->past_len.alloc(4 * S) || e->prev_id.alloc(4 * S) || e->n_frames.alloc(4 * S) || e->last_tok.alloc(4 * S) || e->seg_has_text.alloc(4 * S)) break;
    {
      std::vector<uint32_t> mask((g.vocab + 31) / 32, 0u);
      mask[0] = 3u;                                       // ids 0 ('-', blank) and 1 ('|', silence) render to "" (recognition.py:47-52)
      if (upload(&e->silent_mask, mask.data(), 4 * mask.size())) break;
    }
    if (cudaMemsetAsync(e->kv_cache.p, 0, e->kv_cache.bytes, e->stream) != cudaSuccess) { set_error("memset failed"); break; }
    if (g.split && d == 512 && g.n_heads == 8 && (g.seg_rows == 16 || g.seg_rows == 8) && g.rc_rows == 4) {
      // EXACT: ring rows pre-split [hi 512 | lo 512] bf16 = 16 head-sized pieces per row (attention_exact_kernel)
      e->attn_tma = !make_tmap_bf16_heads(&e->tm_kv, e->kv_cache.p, (uint64_t)S * g.n_layers * 2 * g.ring, 16, (uint32_t)g.seg_rows) &&
                    !make_tmap_bf16_heads(&e->tm_rc, e->rc_kv.p, (uint64_t)B * 2 * g.rc_rows, 16, (uint32_t)g.rc_rows);
      if (!e->attn_tma) fprintf(stderr, "asr_b200: tensor-core EXACT attention disabled: %s\n", asr_last_error());
    }
    if (!g.split && d == 512 && g.n_heads == 8 && (g.seg_rows == 16 || g.seg_rows == 8) && g.rc_rows == 4) {
      // not fatal: without the maps the CTA-per-stream attention kernel serves every batch size
      e->attn_tma = !make_tmap_bf16_heads(&e->tm_kv, e->kv_cache.p, (uint64_t)S * g.n_layers * 2 * g.ring, 8, (uint32_t)g.seg_rows) &&
                    !make_tmap_bf16_heads(&e->tm_rc, e->rc_kv.p, (uint64_t)B * 2 * g.rc_rows, 8, (uint32_t)g.rc_rows);
      if (!e->attn_tma) fprintf(stderr, "asr_b200: streaming attention disabled: %s\n", asr_last_error());
    }
    {
      // TMA-store epilogues (gemm.cuh): per GEMM the tensor maps of its destinations + a kernel-parameter copy of its bias
      auto host_bias = [&](const float* dev, int n, TsMaps* ts) { if (n <= kTsBiasMax) memcpy(ts->bias, weights + (dev - e->w_f32.as<float>()), 4 * (size_t)n); };
      memset(&e->ts_ctc1, 0, sizeof(TsMaps));
      e->ts_ctc1.c0 = e->a_ctc.tm_st;
      host_bias(e->ctc_b1, g.ctc_hidden, &e->ts_ctc1);
      TsMaps qkv;
      memset(&qkv, 0, sizeof(qkv));
      if (!g.split && d % 256 == 0 && 32 % g.seg_rows == 0 && 32 % g.rc_rows == 0) {
        // stream-tiled Q | K | V projection (kPairTileQKV): every destination is a TMA box
        const uint64_t Bm = (uint64_t)B, rowb = (uint64_t)d * 2;
        const uint64_t dims3[3] = {(uint64_t)d, (uint64_t)g.rows, Bm}, str3[2] = {rowb, rowb * g.rows};
        const uint32_t a_seg[3] = {64, (uint32_t)g.seg_rows, (uint32_t)(128 / g.seg_rows)}, a_rc[3] = {64, (uint32_t)g.rc_rows, (uint32_t)(128 / g.rc_rows)};
        const uint32_t q_seg[3] = {64, (uint32_t)g.seg_rows, (uint32_t)(32 / g.seg_rows)}, q_rc[3] = {64, (uint32_t)g.rc_rows, (uint32_t)(32 / g.rc_rows)};
        const uint64_t dims4[4] = {(uint64_t)d, (uint64_t)g.rc_rows, 2, Bm}, str4[3] = {rowb, rowb * g.rc_rows, rowb * g.rc_rows * 2};
        const uint32_t rc4[4] = {64, (uint32_t)g.rc_rows, 1, (uint32_t)(32 / g.rc_rows)};
        e->qkv_ts_ready = !make_tmap_bf16_nd(&e->tm_a_ln_seg, e->a_ln.buf.p, 3, dims3, str3, a_seg) &&
                          !make_tmap_bf16_nd(&qkv.a_rc, e->a_ln.buf.p, 3, dims3, str3, a_rc) &&
                          !make_tmap_bf16_nd(&qkv.c0, e->q.p, 3, dims3, str3, q_seg) && !make_tmap_bf16_nd(&qkv.c1, e->q.p, 3, dims3, str3, q_rc) &&
                          !make_tmap_bf16_2d(&qkv.c2, e->kv_cache.p, (uint64_t)d, (uint64_t)S * g.n_layers * 2 * g.ring, (uint64_t)d, (uint32_t)g.seg_rows) &&
                          !make_tmap_bf16_nd(&qkv.c3, e->rc_kv.p, 4, dims4, str4, rc4);
        if (!e->qkv_ts_ready) fprintf(stderr, "asr_b200: stream-tiled QKV projection disabled: %s\n", asr_last_error());
      }
      for (int l = 0; l < g.n_layers; ++l) {
        LayerW& L = e->layers[l];
        L.ts_qkv = qkv;
        host_bias(L.bqkv, 3 * d, &L.ts_qkv);
        memset(&L.ts_ffn1, 0, sizeof(TsMaps));
        L.ts_ffn1.c0 = e->a_h.tm_st;
        host_bias(L.b1, f, &L.ts_ffn1);
      }
    }
    if (fill_i32(e->past_len.as<int>(), 0, S, e->stream) || fill_i32(e->n_frames.as<int>(), 0, S, e->stream) ||
        fill_i32(e->prev_id.as<int>(), -1, S, e->stream) || fill_i32(e->last_tok.as<int>(), -1, S, e->stream) || fill_i32(e->seg_has_text.as<int>(), 0, S, e->stream)) break;
    e->slot_open.assign(S, 0);
    e->slot_inflight.assign(S, 0);
    e->slot_stamp.assign(S, 0u);
    e->free_slots.resize(S);
    for (size_t i = 0; i < S; ++i) e->free_slots[i] = (int)(S - 1 - i);
    // ---- pinned staging: [pcm (f32 worst case) | slots | outputs]
    const size_t in_bytes = round_up(pcm_bytes(e, B, ASR_PCM_F32), 256) + round_up(4 * (size_t)B, 256);
    const size_t out_bytes = 2 * round_up(4 * (size_t)Mc, 256) + 7 * round_up(4 * (size_t)B, 256) + round_up(4 * (size_t)Mc * g.vocab, 256) +
                             round_up(2 * (size_t)B * BEAM_MAX_LEN, 256) + 4096;
    e->h_out_off = in_bytes; e->h_stage_bytes = in_bytes + out_bytes;
    if (cudaMallocHost(&e->h_buf[0], e->h_stage_bytes) != cudaSuccess || cudaMallocHost(&e->h_buf[1], e->h_stage_bytes) != cudaSuccess) {
      set_error("cudaMallocHost(2 x %zu) failed", e->h_stage_bytes); break;
    }
    e->h_stage = e->h_buf[0];
    if (cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); break; }
    bool ev_ok = true;
    for (int i = 0; i < 2; ++i)
      ev_ok = ev_ok && cudaEventCreateWithFlags(&e->ev_in[i], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&e->ev_done[i], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&e->ev_pre[i], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreate(&e->ev_t0[i]) == cudaSuccess && cudaEventCreate(&e->ev_t1[i]) == cudaSuccess;
    for (int i = 0; i < AsrEngine::kResetRing; ++i) ev_ok = ev_ok && cudaEventCreateWithFlags(&e->ev_reset[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ev_ok) { set_error("cudaEventCreate failed"); break; }
    if (cudaMallocHost((void**)&e->h_reset, 4 * (size_t)cfg->max_sessions * AsrEngine::kResetRing) != cudaSuccess ||
        e->d_reset.alloc(4 * (size_t)cfg->max_sessions * AsrEngine::kResetRing)) {
      set_error("reset staging allocation failed"); break;
    }
    use_buffer(e, 0);
    if (cudaStreamSynchronize(e->stream) != cudaSuccess) { set_error("engine init: %s", cudaGetErrorString(cudaGetLastError())); break; }
    rc = 0;
  } while (0);
  if (rc) { destroy_engine(e); return -1; }
  *out = e;
  return 0;
}

int run_fbank_kind(AsrEngine* e, int kind, int n, int fmt, int n_samples, int subtract_mean, float* d_out, int* n_frames_out) {
  const Geo& g = e->geo;
  if (kind == ASR_FBANK_MELSPEC128) {
    if (n_samples != g.chunk_len) { set_error("melspec128 takes chunk_length = %d samples per stream", g.chunk_len); return -1; }
    *n_frames_out = g.frames;
    use_buffer(e, 0);
    return run_fbank_melspec(e, n, fmt, d_out, false);
  }
  if (kind != ASR_FBANK_KALDI80) { set_error("unknown fbank kind %d", kind); return -1; }
  const FbankPlan& pl = e->kaldi80;
  if (n_samples < 400 || (size_t)n * n_samples * (fmt == ASR_PCM_F32 ? 4 : 2) > e->d_pcm.bytes || ((size_t)n_samples * (fmt == ASR_PCM_F32 ? 4 : 2)) % 16) {
    set_error("kaldi80: n_samples %d unsupported (>= 400, 16-byte multiple, fits the staging buffer)", n_samples); return -1;
  }
  FbankParams P;
  memset(&P, 0, sizeof(P));
  P.pcm = e->d_pcm.p; P.pcm_is_f32 = fmt == ASR_PCM_F32; P.pcm_stride = n_samples; P.n_samples = n_samples;
  P.n_frames = 1 + (n_samples - 400) / 160; P.hop = 160; P.frame_len = 400; P.frame_off = 0; P.nc = pl.nc; P.kaldi = 1;
  P.in_scale = 1.0f; P.preemph = 0.97f; P.log_floor = 1.1920928955078125e-07f;
  P.window = pl.window.as<float>(); P.tw = pl.tw.as<float2>(); P.w2 = pl.w2.as<float2>();
  P.mel_start = pl.mel_start.as<int>(); P.mel_cnt = pl.mel_cnt.as<int>(); P.mel_off = pl.mel_off.as<int>(); P.mel_w = pl.mel_w.as<float>();
  P.n_mels = pl.n_mels; P.out_f32 = d_out; P.out_op = nullptr;
  *n_frames_out = P.n_frames;
  if ((size_t)n * P.n_frames * P.n_mels * 4 > e->fb_f32.bytes) { set_error("kaldi80: output exceeds the feature buffer"); return -1; }
  { ProfScope ps(e, ASR_PROF_FBANK); if (fbank_launch(P, n, e->stream)) return -1; }
  if (subtract_mean) {
    ProfScope ps(e, ASR_PROF_FBANK);
    if (subtract_mean_launch(d_out, n, P.n_frames, P.n_mels, e->stream)) return -1;
  }
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------ hooks of the native scheduler (sched.cu)
namespace asr {

int engine_submit_gather(AsrEngine* e, int n, const int32_t* slots, const int16_t* base, int64_t row_stride, const int32_t* rows, const int64_t* offsets,
                         bool device_gather, bool want_lp, int* ticket) {
  std::lock_guard<std::mutex> lk(e->mu);
  if (device_gather) return submit_rings(e, n, slots, base, row_stride, rows, offsets, want_lp, ticket);
  if (n < 0 || n > e->cfg.max_batch) { set_error("n = %d outside [0, max_batch]", n); return -1; }
  if (e->pend[e->cur_buf].active) { set_error("two steps are already in flight: collect the oldest ticket first"); return -1; }
  int16_t* dst = reinterpret_cast<int16_t*>(e->h_buf[e->cur_buf]);
  const size_t L = e->geo.chunk_len;
  parallel_rows(n, 8, [&](int a, int b) {
    for (int i = a; i < b; ++i) memcpy(dst + (size_t)i * L, base + (size_t)rows[i] * row_stride + offsets[i], L * sizeof(int16_t));
  });
  return submit_step(e, n, slots, dst, ASR_PCM_I16, want_lp, ticket);
}

// Pre-staging: gather `n_rows` chunks into the pinned staging buffer of the NEXT step and start their H2D copy now — before the caller
// knows which of them will run (that depends on the results of the step still in flight: VAD gate, endpoints).  The copy overlaps the
// kernels of the running step; engine_submit_prestaged later launches the chain on a subset through a row-index indirection.
int engine_prestage(AsrEngine* e, int n_rows, const int16_t* base, int64_t row_stride, const int32_t* rows, const int64_t* offsets, bool device_gather) {
  std::lock_guard<std::mutex> lk(e->mu);
  const int b = e->cur_buf;
  if (n_rows < 0 || n_rows > e->cfg.max_batch) { set_error("prestage: %d rows outside [0, max_batch]", n_rows); return -1; }
  if (e->pend[b].active) { set_error("prestage: the next staging buffer still belongs to a step in flight"); return -1; }
  e->prestaged_rows[b] = -1; const int fb_was_staged = e->fb_staged[b]; e->fb_staged[b] = 0; (void)fb_was_staged;
  if (!n_rows) return 0;
  ASR_CUDA_OK(cudaSetDevice(e->device));
  if (device_gather) {                                      // the rings are pinned + mapped: a kernel on the copy stream reads the chunks over PCIe
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, base) != cudaSuccess || pa.type != cudaMemoryTypeHost) {
      cudaGetLastError();
      set_error("prestage: device gather needs the audio rings in pinned host memory (asr_host_alloc)"); return -1;
    }
    long long* h_off = reinterpret_cast<long long*>(e->h_buf[b]);
    for (int i = 0; i < n_rows; ++i) h_off[i] = (long long)rows[i] * row_stride + offsets[i];
    ASR_CUDA_OK(cudaMemcpyAsync(e->d_src_off[b].p, h_off, 8 * (size_t)n_rows, cudaMemcpyHostToDevice, e->copy_stream));
    if (gather_rings_launch(reinterpret_cast<const int16_t*>(pa.devicePointer), e->d_src_off[b].as<long long>(), reinterpret_cast<int16_t*>(dev_pcm(e, b)), n_rows,
                            e->geo.chunk_len, e->copy_stream)) return -1;
  } else {
    int16_t* dst = reinterpret_cast<int16_t*>(e->h_buf[b]);
    const size_t L = e->geo.chunk_len;
    parallel_rows(n_rows, 8, [&](int a, int c) {
      for (int i = a; i < c; ++i) memcpy(dst + (size_t)i * L, base + (size_t)rows[i] * row_stride + offsets[i], L * sizeof(int16_t));
    });
    ASR_CUDA_OK(cudaMemcpyAsync(dev_pcm(e, b), dst, pcm_bytes(e, n_rows, ASR_PCM_I16), cudaMemcpyHostToDevice, e->copy_stream));
  }
  e->prestaged_rows[b] = n_rows;
  if (e->spec_fbank) {
    // the front-end of every staged chunk, queued on the compute stream behind the running step: it executes while the host collects that
    // step and decides which chunks run (nothing in it depends on the decision)
    if (!e->a_fb_stage.buf.p && make_operand(e, &e->a_fb_stage, e->cfg.max_batch * e->geo.frames, e->geo.n_mels)) return -1;
    ASR_CUDA_OK(cudaEventRecord(e->ev_pre[b], e->copy_stream));
    ASR_CUDA_OK(cudaStreamWaitEvent(e->stream, e->ev_pre[b], 0));
    pdl_set_active(false);
    if (run_fbank_melspec(e, n_rows, ASR_PCM_I16, nullptr, true, &e->a_fb_stage, dev_pcm(e, b))) return -1;
    e->fb_staged[b] = 1;
  }
  return 0;
}

// The step over batch elements i < n: session slots[i], audio = pre-staged row staged_index[i].
int engine_submit_prestaged(AsrEngine* e, int n, const int32_t* slots, const int32_t* staged_index, bool want_lp, int* ticket) {
  std::lock_guard<std::mutex> lk(e->mu);
  const int b = e->cur_buf;
  if (e->pend[b].active) { set_error("two steps are already in flight: collect the oldest ticket first"); return -1; }
  if (e->prestaged_rows[b] < 0) { set_error("submit_prestaged: nothing was pre-staged for this step"); return -1; }
  if (check_step_args(e, n, slots)) return -1;
  for (int i = 0; i < n; ++i)
    if (staged_index[i] < 0 || staged_index[i] >= e->prestaged_rows[b]) { set_error("submit_prestaged: staged row %d out of range", staged_index[i]); return -1; }
  const auto t0 = std::chrono::steady_clock::now();
  const int n_staged = e->prestaged_rows[b];
  e->prestaged_rows[b] = -1; const int fb_was_staged = e->fb_staged[b]; e->fb_staged[b] = 0; (void)fb_was_staged;
  if (n) {
    ASR_CUDA_OK(cudaSetDevice(e->device));
    uint8_t* hs = reinterpret_cast<uint8_t*>(e->h_buf[b]);
    int32_t* h_slots = reinterpret_cast<int32_t*>(hs + slots_off(e));
    memcpy(h_slots, slots, 4 * (size_t)n);
    ASR_CUDA_OK(cudaMemcpyAsync(dev_slots(e, b), h_slots, 4 * (size_t)n, cudaMemcpyHostToDevice, e->copy_stream));
    // the index list travels in the (unused) float32 half of the PCM staging area: int16 batches fill at most the first half
    int32_t* h_index = reinterpret_cast<int32_t*>(hs + pcm_bytes(e, e->cfg.max_batch, ASR_PCM_I16));
    memcpy(h_index, staged_index, 4 * (size_t)n);
    ASR_CUDA_OK(cudaMemcpyAsync(e->d_row_index[b].p, h_index, 4 * (size_t)n, cudaMemcpyHostToDevice, e->copy_stream));
    ASR_CUDA_OK(cudaEventRecord(e->ev_in[b], e->copy_stream));
    ASR_CUDA_OK(cudaStreamWaitEvent(e->stream, e->ev_in[b], 0));
    use_buffer(e, b);
    e->act_row_index = e->d_row_index[b].as<int>();
    e->act_fb_stage = fb_was_staged != 0;
    e->act_fb_identity = false;
    if (e->act_fb_stage && n == n_staged) {                   // every staged chunk runs, in staging order: no compaction needed
      bool id = true;
      for (int i = 0; i < n && id; ++i) id = staged_index[i] == i;
      e->act_fb_identity = id;
    }
    if (e->ev_t0[b]) cudaEventRecord(e->ev_t0[b], e->stream);
    const int rc = run_step_chain(e, n, ASR_PCM_I16, want_lp);
    e->act_row_index = nullptr; e->act_fb_stage = false; e->act_fb_identity = false;
    if (rc) return -1;
    if (enqueue_d2h(e, b, n, want_lp)) return -1;
    if (e->ev_t1[b]) cudaEventRecord(e->ev_t1[b], e->stream);
    ASR_CUDA_OK(cudaEventRecord(e->ev_done[b], e->stream));
  }
  e->pend[b].n = n; e->pend[b].want_lp = want_lp; e->pend[b].active = 1; e->pend[b].t0 = t0;
  mark_inflight(e, b, n, slots);
  e->cur_buf ^= 1;
  if (ticket) *ticket = b;
  return 0;
}

int engine_collect_view(AsrEngine* e, int ticket, StepView* v) {
  std::lock_guard<std::mutex> lk(e->mu);
  const int n = (ticket == 0 || ticket == 1) ? e->pend[ticket].n : 0;
  const bool want_lp = (ticket == 0 || ticket == 1) && e->pend[ticket].want_lp;
  const auto t0 = (ticket == 0 || ticket == 1) ? e->pend[ticket].t0 : std::chrono::steady_clock::now();
  if (collect_step(e, ticket, nullptr)) return -1;
  *v = StepView();
  v->n = n;
  const uint8_t* ho = reinterpret_cast<const uint8_t*>(e->h_buf[ticket]) + e->h_out_off;
  for (auto& it : out_layout(e, n, want_lp)) {
    const void* p = ho + it.hoff;
    switch (it.field) {
      case 0: v->argmax_ids = (const int32_t*)p; break; case 1: v->new_tokens = (const int32_t*)p; break; case 2: v->n_new = (const int32_t*)p; break;
      case 3: v->blank_frames = (const int32_t*)p; break; case 4: v->has_token = (const int32_t*)p; break; case 5: v->logprobs = (const float*)p; break;
      case 6: v->beam_tokens = (const int16_t*)p; break; case 7: v->beam_len = (const int32_t*)p; break; case 8: v->beam_score = (const float*)p; break;
      case 9: v->has_text = (const int32_t*)p; break; case 10: v->flags = (const int32_t*)p; break;
    }
  }
  if (n) record_step(e, n, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  return 0;
}

int engine_reset_async(AsrEngine* e, int n, const int32_t* slots) {
  std::lock_guard<std::mutex> lk(e->mu);
  if (n <= 0) return 0;
  ASR_CUDA_OK(cudaSetDevice(e->device));
  return reset_slots_async(e, n, slots);
}

int engine_open_slot(AsrEngine* e, int32_t* slot) { return asr_session_open(e, slot); }
int engine_close_slot(AsrEngine* e, int32_t slot) { return asr_session_close(e, slot); }
int engine_wait_inputs(AsrEngine* e) { return asr_wait_inputs(e); }

}  // namespace asr

// ================================================================================================ C ABI
extern "C" {

const char* asr_last_error(void) { return g_err; }
int asr_abi_version(void) { return ASR_B200_ABI_VERSION; }

int asr_default_config(AsrConfig* c, int low_latency) {
  if (!c) { set_error("null config"); return -1; }
  memset(c, 0, sizeof(*c));
  c->abi_version = ASR_B200_ABI_VERSION;
  c->sample_rate = 16000; c->hop = 160; c->n_fft = 800; c->win = 400; c->n_mels = 128;
  c->segment_size = low_latency ? 32 : 64; c->context_size = 16; c->bias = 4; c->stride = 4;
  c->d_model = 512; c->n_heads = 8; c->ffn_dim = 2048; c->n_layers = 20; c->left_context = 32; c->ctc_hidden = 512; c->vocab = 804;
  c->precision = ASR_PRECISION_FAST; c->max_sessions = 1024; c->max_batch = 256;
  return 0;
}

int asr_weights_count(const AsrConfig* cfg, uint64_t* n) {
  Geo g;
  if (!cfg || !n) { set_error("null argument"); return -1; }
  AsrConfig c = *cfg;
  if (c.max_batch <= 0) c.max_batch = 1;
  if (c.max_sessions <= 0) c.max_sessions = 1;
  if (fill_geo(c, &g)) return -1;
  *n = weights_count(g);
  return 0;
}

int asr_chunk_geometry(const AsrConfig* cfg, int32_t* chunk_length, int32_t* segment_length, int32_t* seg_rows) {
  Geo g;
  if (!cfg) { set_error("null config"); return -1; }
  AsrConfig c = *cfg;
  if (c.max_batch <= 0) c.max_batch = 1;
  if (c.max_sessions <= 0) c.max_sessions = 1;
  if (fill_geo(c, &g)) return -1;
  if (chunk_length) *chunk_length = g.chunk_len;
  if (segment_length) *segment_length = cfg->segment_size * cfg->hop;
  if (seg_rows) *seg_rows = g.seg_rows;
  return 0;
}

int asr_engine_create(const AsrConfig* cfg, const float* weights, uint64_t n_floats, int device, AsrEngine** out) {
  try { return create_engine(cfg, weights, n_floats, device, out); } catch (const std::exception& ex) { set_error("engine create: %s", ex.what()); return -1; }
}

int asr_engine_destroy(AsrEngine* e) { destroy_engine(e); return 0; }

int asr_session_open(AsrEngine* e, int32_t* slot_out) {
  if (!e || !slot_out) { set_error("null argument"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  if (e->free_slots.empty()) { set_error("no free session slot (max_sessions = %d)", e->cfg.max_sessions); return -1; }
  ASR_CUDA_OK(cudaSetDevice(e->device));
  const int32_t s = e->free_slots.back();
  if (reset_slots_async(e, 1, &s)) return -1;
  e->free_slots.pop_back();
  e->slot_open[s] = 1;
  *slot_out = s;
  return 0;
}

int asr_session_reset(AsrEngine* e, int32_t slot) {
  if (!e) { set_error("null engine"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  if (slot < 0 || slot >= e->cfg.max_sessions || !e->slot_open[slot]) { set_error("slot %d is not an open session", slot); return -1; }
  ASR_CUDA_OK(cudaSetDevice(e->device));
  return reset_slots_async(e, 1, &slot);
}

int asr_session_reset_many(AsrEngine* e, int32_t n, const int32_t* slots) {
  if (!e || (n > 0 && !slots)) { set_error("null argument"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  if (n <= 0) return 0;
  if (n > e->cfg.max_sessions) { set_error("reset_many: n = %d > max_sessions", n); return -1; }
  for (int i = 0; i < n; ++i)
    if (slots[i] < 0 || slots[i] >= e->cfg.max_sessions || !e->slot_open[slots[i]]) { set_error("slot %d is not an open session", slots[i]); return -1; }
  ASR_CUDA_OK(cudaSetDevice(e->device));
  return reset_slots_async(e, n, slots);
}

/* Batch assembly in native code: copies, for i < n, chunk_length samples starting at base[rows[i] * row_stride + offsets[i]] into
 * row i of the pinned staging buffer of the next step (multi-threaded memcpy).  Returns that buffer through *pinned_out. */
int asr_gather_pcm(AsrEngine* e, int32_t n, const int16_t* base, int64_t row_stride, const int32_t* rows, const int64_t* offsets, void** pinned_out) {
  if (!e || !base || (n > 0 && (!rows || !offsets))) { set_error("null argument"); return -1; }
  if (n < 0 || n > e->cfg.max_batch) { set_error("n = %d outside [0, max_batch]", n); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  if (e->pend[e->cur_buf].active) { set_error("gather_pcm: the next staging buffer still belongs to a step in flight (asr_collect it first)"); return -1; }
  int16_t* dst = reinterpret_cast<int16_t*>(e->h_buf[e->cur_buf]);
  const size_t L = e->geo.chunk_len;
  parallel_rows(n, 8, [&](int a, int b) {
    for (int i = a; i < b; ++i) memcpy(dst + (size_t)i * L, base + (size_t)rows[i] * row_stride + offsets[i], L * sizeof(int16_t));
  });
  if (pinned_out) *pinned_out = dst;
  return 0;
}

/* Energy gate support (stand-in for the WebRTC gate of stream.py:166-189): peaks[i] = max |x| over
 * base[rows[i]*row_stride + offsets[i] + from .. + to).  Pure host helper, multi-threaded; no device work. */
int asr_pcm_peaks(int32_t n, const int16_t* base, int64_t row_stride, const int32_t* rows, const int64_t* offsets, int32_t from, int32_t to,
                  int32_t* peaks) {
  if (n < 0 || !base || (n > 0 && (!rows || !offsets || !peaks)) || from < 0 || to < from) { set_error("asr_pcm_peaks: bad argument"); return -1; }
  parallel_rows(n, 8, [&](int a, int b) {
    for (int i = a; i < b; ++i) {
      const int16_t* x = base + (size_t)rows[i] * row_stride + offsets[i];
      int lo = 0, hi = 0;
      for (int j = from; j < to; ++j) { const int v = x[j]; lo = v < lo ? v : lo; hi = v > hi ? v : hi; }
      peaks[i] = std::max(hi, -lo);
    }
  });
  return 0;
}

int asr_session_close(AsrEngine* e, int32_t slot) {
  if (!e) { set_error("null engine"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  if (slot < 0 || slot >= e->cfg.max_sessions || !e->slot_open[slot]) { set_error("slot %d is not an open session", slot); return -1; }
  if (e->slot_inflight[slot]) { set_error("slot %d rides a submitted step: asr_collect its ticket before closing the session", slot); return -1; }
  e->slot_open[slot] = 0;
  e->free_slots.push_back(slot);
  return 0;
}

int asr_step(AsrEngine* e, int32_t n, const int32_t* slots, const void* pcm, int32_t fmt, const AsrStepOut* out) {
  if (!e) { set_error("null engine"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  const auto t0 = std::chrono::steady_clock::now();
  if (e->pend[0].active || e->pend[1].active) { set_error("asr_step while an asr_submit ticket is in flight"); return -1; }
  int ticket = 0;
  if (submit_step(e, n, slots, pcm, fmt, out && out->logprobs, &ticket)) return -1;
  if (collect_step(e, ticket, out)) return -1;
  if (n) record_step(e, n, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  return 0;
}

int asr_submit(AsrEngine* e, int32_t n, const int32_t* slots, const void* pcm, int32_t fmt, int32_t want_logprobs, int32_t* ticket) {
  if (!e || !ticket) { set_error("null argument"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  return submit_step(e, n, slots, pcm, fmt, want_logprobs != 0, ticket);
}

int asr_submit_rings(AsrEngine* e, int32_t n, const int32_t* slots, const int16_t* base, int64_t row_stride, const int32_t* rows,
                     const int64_t* offsets, int32_t want_logprobs, int32_t* ticket) {
  if (!e || !ticket) { set_error("null argument"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  return submit_rings(e, n, slots, base, row_stride, rows, offsets, want_logprobs != 0, ticket);
}

/* Blocks until the input transfers (H2D copies / ring gathers) of every submitted step have completed: after it returns the caller may
 * overwrite the host memory those steps read from. */
int asr_wait_inputs(AsrEngine* e) {
  if (!e) { set_error("null engine"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  ASR_CUDA_OK(cudaSetDevice(e->device));
  ASR_CUDA_OK(cudaStreamSynchronize(e->copy_stream));
  return 0;
}

/* Pinned, device-mapped host memory for the scheduler's session audio rings (asr_submit_rings reads them over PCIe). */
void* asr_host_alloc(uint64_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
    set_error("cudaHostAlloc(%llu) failed: %s", (unsigned long long)bytes, cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  return p;
}
int asr_host_free(void* p) {
  if (p && cudaFreeHost(p) != cudaSuccess) { set_error("cudaFreeHost failed: %s", cudaGetErrorString(cudaGetLastError())); return -1; }
  return 0;
}

int asr_collect(AsrEngine* e, int32_t ticket, const AsrStepOut* out) {
  if (!e) { set_error("null engine"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  const int n = (ticket == 0 || ticket == 1) ? e->pend[ticket].n : 0;
  const auto t0 = (ticket == 0 || ticket == 1) ? e->pend[ticket].t0 : std::chrono::steady_clock::now();
  if (collect_step(e, ticket, out)) return -1;
  if (n) record_step(e, n, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  return 0;
}

int asr_stage(AsrEngine* e, int32_t n, const int32_t* slots, const void* pcm, int32_t fmt) {
  if (!e) { set_error("null engine"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  if (ticket_pending(e)) { set_error("asr_stage while an asr_submit ticket is in flight (staging buffer 0 may still be read)"); return -1; }
  e->staged_fmt = fmt;
  use_buffer(e, 0);
  return stage_inputs(e, 0, n, slots, pcm, fmt, e->stream);
}

int asr_run_staged(AsrEngine* e, int32_t n, int32_t want_logprobs) {
  if (!e) { set_error("null engine"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  if (n <= 0 || n > e->cfg.max_batch) { set_error("n = %d outside (0, max_batch]", n); return -1; }
  if (ticket_pending(e)) { set_error("asr_run_staged while an asr_submit ticket is in flight"); return -1; }
  ASR_CUDA_OK(cudaSetDevice(e->device));
  use_buffer(e, 0);
  if (run_step_chain(e, n, e->staged_fmt, want_logprobs != 0)) return -1;
  ++e->steps; e->stream_chunks += n;
  return 0;
}

int asr_pipeline_gpu_time(AsrEngine* e, double* total_ms, uint64_t* n_steps, int32_t reset) {
  if (!e || !total_ms || !n_steps) { set_error("null argument"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  *total_ms = e->pipe_gpu_ms; *n_steps = e->pipe_gpu_n;
  if (reset) { e->pipe_gpu_ms = 0.0; e->pipe_gpu_n = 0; }
  return 0;
}

int asr_fetch(AsrEngine* e, int32_t n, const AsrStepOut* out) {
  if (!e) { set_error("null engine"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  ASR_CUDA_OK(cudaSetDevice(e->device));
  return fetch_outputs(e, n, out, false);
}

int asr_sync(AsrEngine* e) {
  if (!e) { set_error("null engine"); return -1; }
  ASR_CUDA_OK(cudaSetDevice(e->device));
  ASR_CUDA_OK(cudaStreamSynchronize(e->stream));
  return 0;
}

void* asr_stream_handle(AsrEngine* e) { return e ? (void*)e->stream : nullptr; }

void* asr_pinned_pcm(AsrEngine* e, uint64_t* capacity_bytes) {
  if (!e) return nullptr;
  if (capacity_bytes) *capacity_bytes = pcm_bytes(e, e->cfg.max_batch, ASR_PCM_F32);
  return e->h_buf[e->cur_buf];                              // the buffer the NEXT asr_step / asr_submit stages from
}

int asr_stage_raw(AsrEngine* e, const void* pcm, uint64_t bytes) {
  if (!e || !pcm) { set_error("null argument"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  if (ticket_pending(e)) { set_error("asr_stage_raw while an asr_submit ticket is in flight"); return -1; }
  if (bytes > e->d_pcm.bytes) { set_error("asr_stage_raw: %llu bytes > staging capacity %zu", (unsigned long long)bytes, e->d_pcm.bytes); return -1; }
  ASR_CUDA_OK(cudaSetDevice(e->device));
  memcpy(e->h_buf[0], pcm, bytes);
  ASR_CUDA_OK(cudaMemcpyAsync(e->d_pcm.p, e->h_buf[0], bytes, cudaMemcpyHostToDevice, e->stream));
  return 0;
}

static int ensure_fb_buffer(AsrEngine* e, size_t bytes) {
  if (e->fb_f32.bytes >= bytes && e->fb_f32.p) return 0;
  e->fb_f32.free();
  return e->fb_f32.alloc(bytes);
}

int asr_fbank_staged(AsrEngine* e, int32_t kind, int32_t n, int32_t fmt, int32_t n_samples) {
  if (!e) { set_error("null engine"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  ASR_CUDA_OK(cudaSetDevice(e->device));
  const int frames = kind == ASR_FBANK_MELSPEC128 ? e->geo.frames : 1 + (n_samples - 400) / 160;
  const int mels = kind == ASR_FBANK_MELSPEC128 ? e->geo.n_mels : 80;
  if (n <= 0 || frames <= 0) { set_error("asr_fbank_staged: bad sizes"); return -1; }
  if (ensure_fb_buffer(e, (size_t)n * frames * mels * 4)) return -1;
  int nf = 0;
  return run_fbank_kind(e, kind, n, fmt, n_samples, 0, e->fb_f32.as<float>(), &nf);
}

int asr_fbank(AsrEngine* e, int32_t kind, int32_t n, const void* pcm, int32_t fmt, int32_t n_samples, int32_t subtract_mean, float* out) {
  if (!e || !pcm || !out) { set_error("null argument"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  if (n <= 0) return 0;
  if (ticket_pending(e)) { set_error("asr_fbank while an asr_submit ticket is in flight"); return -1; }
  ASR_CUDA_OK(cudaSetDevice(e->device));
  const size_t in_bytes = (size_t)n * n_samples * (fmt == ASR_PCM_F32 ? 4 : 2);
  if (in_bytes > e->d_pcm.bytes) { set_error("asr_fbank: input of %zu bytes exceeds staging capacity %zu (raise max_batch)", in_bytes, e->d_pcm.bytes); return -1; }
  const int frames = kind == ASR_FBANK_MELSPEC128 ? e->geo.frames : 1 + (n_samples - 400) / 160;
  const int mels = kind == ASR_FBANK_MELSPEC128 ? e->geo.n_mels : 80;
  if (frames <= 0) { set_error("asr_fbank: too few samples"); return -1; }
  const size_t out_bytes = (size_t)n * frames * mels * 4;
  if (ensure_fb_buffer(e, out_bytes)) return -1;
  ASR_CUDA_OK(cudaMemcpyAsync(e->d_pcm.p, pcm, in_bytes, cudaMemcpyHostToDevice, e->stream));
  int nf = 0;
  if (run_fbank_kind(e, kind, n, fmt, n_samples, subtract_mean, e->fb_f32.as<float>(), &nf)) return -1;
  ASR_CUDA_OK(cudaMemcpyAsync(out, e->fb_f32.p, out_bytes, cudaMemcpyDeviceToHost, e->stream));
  ASR_CUDA_OK(cudaStreamSynchronize(e->stream));
  return 0;
}

int asr_set_beam(AsrEngine* e, int32_t beam, int32_t cand_k) {
  if (!e) { set_error("null engine"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  if (beam < 0 || beam > BEAM_MAX || (beam > 0 && (cand_k < 1 || cand_k > BEAM_CAND_MAX))) {
    set_error("asr_set_beam: beam must be in [0, %d], cand_k in [1, %d]", BEAM_MAX, BEAM_CAND_MAX); return -1;
  }
  ASR_CUDA_OK(cudaSetDevice(e->device));
  drop_step_graphs(e);                                     // the beam kernel's arguments are part of what the graphs captured
  if (beam > 0 && !e->bm_tokens.p) {
    const size_t S = e->cfg.max_sessions, B = e->cfg.max_batch;
    if (e->bm_n.alloc(4 * S) || e->bm_cur.alloc(4 * S) || e->bm_len.alloc(4 * S * BEAM_MAX) || e->bm_last.alloc(4 * S * BEAM_MAX) ||
        e->bm_pb.alloc(4 * S * BEAM_MAX) || e->bm_pnb.alloc(4 * S * BEAM_MAX) || e->bm_hash.alloc(8 * S * BEAM_MAX) ||
        e->bm_tokens.alloc(2 * S * 2 * BEAM_MAX * BEAM_MAX_LEN) || e->d_beam_tok.alloc(2 * B * BEAM_MAX_LEN) || e->d_beam_len.alloc(4 * B) ||
        e->d_beam_score.alloc(4 * B) || e->bm_cand_tok.alloc(4 * B * e->geo.seg_rows * BEAM_CAND_MAX) ||
        e->bm_cand_lp.alloc(4 * B * e->geo.seg_rows * BEAM_CAND_MAX) || e->bm_row_stat.alloc(8 * B * e->geo.seg_rows)) return -1;
  }
  e->beam = beam; e->cand_k = cand_k;
  if (beam > 0) {
    if (beam_reset_launch(beam_params(e, 0), -1, e->cfg.max_sessions, e->stream)) return -1;
    ASR_CUDA_OK(cudaStreamSynchronize(e->stream));
  }
  return 0;
}

int asr_set_silent_ids(AsrEngine* e, int32_t n, const int32_t* ids) {
  if (!e || (n > 0 && !ids)) { set_error("null argument"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  std::vector<uint32_t> mask((e->geo.vocab + 31) / 32, 0u);
  for (int i = 0; i < n; ++i) {
    if (ids[i] < 0 || ids[i] >= e->geo.vocab) { set_error("asr_set_silent_ids: id %d outside the vocabulary of %d", ids[i], e->geo.vocab); return -1; }
    mask[ids[i] >> 5] |= 1u << (ids[i] & 31);
  }
  ASR_CUDA_OK(cudaSetDevice(e->device));
  ASR_CUDA_OK(cudaStreamSynchronize(e->stream));            // configuration call: not on the serving path
  ASR_CUDA_OK(cudaMemcpy(e->silent_mask.p, mask.data(), 4 * mask.size(), cudaMemcpyHostToDevice));
  return 0;
}

int asr_get_stats(AsrEngine* e, AsrStats* out) {
  if (!e || !out) { set_error("null argument"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  memset(out, 0, sizeof(*out));
  out->steps = e->steps; out->stream_chunks = e->stream_chunks; out->kernel_launches = e->launches;
  if (!e->step_ms.empty()) {
    std::vector<float> v = e->step_ms;
    std::sort(v.begin(), v.end());
    double s = 0; for (float x : v) s += x;
    out->step_ms_mean = s / v.size();
    out->step_ms_p50 = v[v.size() / 2];
    out->step_ms_p99 = v[std::min(v.size() - 1, (size_t)(0.99 * v.size()))];
    out->step_ms_max = v.back();
  }
  return 0;
}

int asr_profile_enable(AsrEngine* e, int32_t on) {
  if (!e) { set_error("null engine"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  e->prof_on = on != 0;
  return 0;
}

int asr_profile_read(AsrEngine* e, double* ms, uint64_t* launches) {
  if (!e) { set_error("null engine"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  ASR_CUDA_OK(cudaSetDevice(e->device));
  ASR_CUDA_OK(cudaStreamSynchronize(e->stream));
  for (auto& r : e->prof_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { e->prof_ms[r.cat] += t; ++e->prof_n[r.cat]; }
    e->prof_pool.push_back(r.a); e->prof_pool.push_back(r.b);
  }
  e->prof_recs.clear();
  for (int i = 0; i < ASR_PROF_COUNT; ++i) {
    if (ms) ms[i] = e->prof_ms[i];
    if (launches) launches[i] = e->prof_n[i];
    e->prof_ms[i] = 0; e->prof_n[i] = 0;
  }
  return 0;
}

int asr_debug_step_partial(AsrEngine* e, int32_t n, const int32_t* slots, const void* pcm, int32_t fmt, int32_t n_layers) {
  if (!e) { set_error("null engine"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  if (n_layers < 0 || n_layers > e->geo.n_layers) { set_error("n_layers out of range"); return -1; }
  if (ticket_pending(e)) { set_error("asr_debug_step_partial while an asr_submit ticket is in flight"); return -1; }
  use_buffer(e, 0);
  if (stage_inputs(e, 0, n, slots, pcm, fmt, e->stream)) return -1;
  if (n == 0) return 0;
  if (run_pipeline(e, n, fmt, n_layers, false, false)) return -1;
  ASR_CUDA_OK(cudaStreamSynchronize(e->stream));
  return 0;
}

int asr_debug_decode_logits(AsrEngine* e, int32_t n, const int32_t* slots, const float* logits, const AsrStepOut* out) {
  if (!e || (n > 0 && !logits)) { set_error("null argument"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  if (ticket_pending(e)) { set_error("asr_debug_decode_logits while an asr_submit ticket is in flight"); return -1; }
  if (check_step_args(e, n, slots)) return -1;
  if (n == 0) return 0;
  ASR_CUDA_OK(cudaSetDevice(e->device));
  use_buffer(e, 0);
  const bool want_lp = out && out->logprobs;
  ASR_CUDA_OK(cudaMemcpyAsync(e->logits.p, logits, 4 * (size_t)n * e->geo.seg_rows * e->geo.vocab, cudaMemcpyHostToDevice, e->stream));
  ASR_CUDA_OK(cudaMemcpyAsync(dev_slots(e, 0), slots, 4 * (size_t)n, cudaMemcpyHostToDevice, e->stream));
  pdl_set_active(false);
  if (run_decode(e, n, want_lp)) return -1;
  if (enqueue_d2h(e, 0, n, want_lp)) return -1;
  ASR_CUDA_OK(cudaStreamSynchronize(e->stream));
  deliver(e, 0, n, want_lp, out);
  return 0;
}

int asr_debug_read(AsrEngine* e, int32_t which, float* out, uint64_t n_floats) {
  if (!e || !out) { set_error("null argument"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  const DevBuf* src = which == 0 ? &e->x : which == 1 ? &e->x1 : which == 2 ? &e->x2 : which == 3 ? &e->q : which == 4 ? &e->logits :
                      which == 5 ? &e->bm_cand_lp : which == 6 ? &e->bm_cand_tok : which == 7 ? &e->bm_row_stat : nullptr;
  if (!src || n_floats * 4 > src->bytes) { set_error("asr_debug_read: bad buffer id %d or size", which); return -1; }
  ASR_CUDA_OK(cudaSetDevice(e->device));
  ASR_CUDA_OK(cudaStreamSynchronize(e->stream));
  ASR_CUDA_OK(cudaMemcpy(out, src->p, n_floats * 4, cudaMemcpyDeviceToHost));
  return 0;
}

int asr_debug_read_state(AsrEngine* e, int32_t slot, int32_t layer, int32_t which, float* out, int32_t* past_length) {
  if (!e || !out) { set_error("null argument"); return -1; }
  std::lock_guard<std::mutex> lk(e->mu);
  const Geo& g = e->geo;
  if (slot < 0 || slot >= e->cfg.max_sessions || layer < 0 || layer >= g.n_layers || (which != 0 && which != 1)) { set_error("asr_debug_read_state: bad index"); return -1; }
  ASR_CUDA_OK(cudaSetDevice(e->device));
  ASR_CUDA_OK(cudaStreamSynchronize(e->stream));
  int pl = 0;
  ASR_CUDA_OK(cudaMemcpy(&pl, e->past_len.as<int>() + slot, 4, cudaMemcpyDeviceToHost));
  if (past_length) *past_length = pl;
  const size_t esz = g.split ? 4 : 2, d = g.d_model;
  std::vector<uint8_t> ring((size_t)g.ring * d * esz);
  const uint8_t* base = reinterpret_cast<const uint8_t*>(e->kv_cache.p) + ((size_t)layer * e->layer_stride + (size_t)slot * e->slot_stride + (size_t)which * g.ring * d) * esz;
  ASR_CUDA_OK(cudaMemcpy(ring.data(), base, ring.size(), cudaMemcpyDeviceToHost));
  const int lv = std::min(pl, g.left);
  memset(out, 0, sizeof(float) * g.left * d);
  for (int i = 0; i < lv; ++i) {                       // reference layout: right-aligned, oldest first (TA:emformer.py:395-396)
    const int rr = ((pl - lv + i) % g.ring + g.ring) % g.ring;
    float* o = out + (size_t)(g.left - lv + i) * d;
    if (g.split) {                                      // pre-split rows [hi d | lo d] bf16: x = hi + lo
      const uint16_t* h = reinterpret_cast<const uint16_t*>(ring.data()) + (size_t)rr * 2 * d;
      for (size_t c = 0; c < d; ++c) {
        uint32_t uh = (uint32_t)h[c] << 16, ul = (uint32_t)h[d + c] << 16;
        float fh, fl;
        memcpy(&fh, &uh, 4); memcpy(&fl, &ul, 4);
        o[c] = fh + fl;
      }
    } else {
      const uint16_t* h = reinterpret_cast<const uint16_t*>(ring.data()) + (size_t)rr * d;
      for (size_t c = 0; c < d; ++c) { uint32_t u = (uint32_t)h[c] << 16; memcpy(o + c, &u, 4); }
    }
  }
  return 0;
}

int asr_debug_gemm(int32_t impl, int32_t M, int32_t N, int32_t K, int32_t split, int32_t bn, const float* A, const float* B, const float* bias,
                   float* C, int device) {
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0 || K % 64) { set_error("asr_debug_gemm: bad arguments"); return -1; }
  ASR_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  ASR_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  const int ld = split ? 2 * K : K, lo = split ? K : 0;
  const int Mp = (int)round_up(M, 128);
  DevBuf dA32, dB32, dA, dB, dC, dbias;
  int rc = -1;
  do {
    if (dA32.alloc(4 * (size_t)M * K) || dB32.alloc(4 * (size_t)N * K) || dA.alloc(2 * (size_t)Mp * ld) || dB.alloc(2 * (size_t)N * ld) ||
        dC.alloc(4 * (size_t)M * N) || dbias.alloc(4 * (size_t)N)) break;
    if (cudaMemcpy(dA32.p, A, 4 * (size_t)M * K, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(dB32.p, B, 4 * (size_t)N * K, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("H2D failed"); break; }
    if (bias && cudaMemcpy(dbias.p, bias, 4 * (size_t)N, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("H2D failed"); break; }
    cudaMemset(dA.p, 0, dA.bytes);
    if (convert_weight(dA32.as<float>(), dA.as<bf16>(), M, K, ld, lo, 0) || convert_weight(dB32.as<float>(), dB.as<bf16>(), N, K, ld, lo, 0)) break;
    const GemmProblem p = make_problem(M, N, K, split);
    EpiF32 epi{dC.as<float>(), bias ? dbias.as<float>() : nullptr, nullptr, N, N};
    if (impl != 0) { set_error("asr_debug_gemm: impl %d does not exist (0 = tcgen05)", impl); break; }
    {
      CUtensorMap ta, tb;
      if (make_tmap_bf16_2d(&ta, dA.p, ld, Mp, ld, 128) || make_tmap_bf16_2d(&tb, dB.p, ld, N, ld, bn == kPairTile ? 128 : bn)) break;
      if (gemm_tc<EpiF32>(ta, tb, p, epi, bn, prop.multiProcessorCount, 0)) break;
    }
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { set_error("asr_debug_gemm: kernel failed: %s", cudaGetErrorString(err)); break; }
    if (cudaMemcpy(C, dC.p, 4 * (size_t)M * N, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("D2H failed"); break; }
    rc = 0;
  } while (0);
  dA32.free(); dB32.free(); dA.free(); dB.free(); dC.free(); dbias.free();
  return rc;
}

/* Diagnostic: act(A B^T + bias) written as a bf16 operand through the EpiOperand epilogue (act: 0 none, 1 GELU, 2 SiLU), returned as fp32.
 * bn = 512: pair kernel, LSU epilogue; 515: pair kernel, TMA-store epilogue; 64 / 128 / 256: one-CTA kernels. */
int asr_debug_gemm_operand(int32_t M, int32_t N, int32_t K, int32_t bn, int32_t act, const float* A, const float* B, const float* bias, float* out, int device) {
  if (!A || !B || !bias || !out || M <= 0 || N <= 0 || K <= 0 || K % 64) { set_error("asr_debug_gemm_operand: bad arguments"); return -1; }
  ASR_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  ASR_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  const int Mp = (int)round_up(M, 256);
  DevBuf dA32, dB32, dA, dB, dO, dbias;
  std::vector<bf16> ho;
  int rc = -1;
  do {
    if (dA32.alloc(4 * (size_t)M * K) || dB32.alloc(4 * (size_t)N * K) || dA.alloc(2 * (size_t)Mp * K) || dB.alloc(2 * (size_t)N * K) || dO.alloc(2 * (size_t)Mp * N) ||
        dbias.alloc(4 * (size_t)N)) break;
    if (cudaMemcpy(dA32.p, A, 4 * (size_t)M * K, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(dB32.p, B, 4 * (size_t)N * K, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(dbias.p, bias, 4 * (size_t)N, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("H2D failed"); break; }
    cudaMemset(dA.p, 0, dA.bytes); cudaMemset(dO.p, 0xff, dO.bytes);
    if (convert_weight(dA32.as<float>(), dA.as<bf16>(), M, K, K, 0, 0) || convert_weight(dB32.as<float>(), dB.as<bf16>(), N, K, K, 0, 0)) break;
    const GemmProblem p = make_problem(M, N, K, 0);
    const bool pair = bn == kPairTile || bn == kPairTileTS;
    CUtensorMap ta, tb;
    TsMaps ts = {};
    if (make_tmap_bf16_2d(&ta, dA.p, K, Mp, K, 128) || make_tmap_bf16_2d(&tb, dB.p, K, N, K, pair ? 128 : bn)) break;
    if (bn == kPairTileTS && (N > kTsBiasMax || make_tmap_bf16_2d(&ts.c0, dO.p, N, Mp, N, 32))) { if (N > kTsBiasMax) set_error("N > %d", kTsBiasMax); break; }
    if (N <= kTsBiasMax) memcpy(ts.bias, bias, 4 * (size_t)N);
    EpiOperand epi{dO.as<bf16>(), dbias.as<float>(), N, 0, act};
    if (gemm_tc<EpiOperand>(ta, tb, p, epi, bn, prop.multiProcessorCount, 0, bn == kPairTileTS ? &ts : nullptr)) break;
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { set_error("asr_debug_gemm_operand: kernel failed: %s", cudaGetErrorString(err)); break; }
    ho.resize((size_t)M * N);
    if (cudaMemcpy(ho.data(), dO.p, 2 * ho.size(), cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("D2H failed"); break; }
    for (size_t i = 0; i < ho.size(); ++i) out[i] = __bfloat162float(ho[i]);
    rc = 0;
  } while (0);
  dA32.free(); dB32.free(); dA.free(); dB.free(); dO.free(); dbias.free();
  return rc;
}

/* Diagnostic: the fused GEMM + residual + LayerNorm kernel on host data.  W: [512, K].  g2/b2 nullable.  out_op_f32: the bf16 operand
 * output converted back to fp32, [M, 512] (or [M / compact_rows * compact_seg, 512] when compact_rows > 0). */
int asr_debug_gemm_ln(int32_t M, int32_t K, int32_t split, const float* A, const float* W, const float* bias, const float* res, const float* g1,
                      const float* b1, const float* g2, const float* b2, int32_t f32_normed, int32_t compact_rows, int32_t compact_seg,
                      float* out_f32, float* out_op_f32, int32_t iters, float* ms_out, int32_t pair, int device) {
  if (!A || !W || !bias || !res || !g1 || !b1 || !out_f32 || !out_op_f32 || M <= 0 || K <= 0 || K % 64) { set_error("asr_debug_gemm_ln: bad arguments"); return -1; }
  ASR_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  ASR_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  const int N = 512, ld = split ? 2 * K : K, lo = split ? K : 0, old_ = split ? 2 * N : N, olo = split ? N : 0;
  const int Mp = (int)round_up(M, 256);
  const int Mo = compact_rows > 0 ? M / compact_rows * compact_seg : M;
  DevBuf dA32, dW32, dA, dW, dv, dres, dout, dop, dyc;
  std::vector<bf16> hop;
  int rc = -1;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  do {
    if (dA32.alloc(4 * (size_t)M * K) || dW32.alloc(4 * (size_t)N * K) || dA.alloc(2 * (size_t)Mp * ld) || dW.alloc(2 * (size_t)N * ld) ||
        dv.alloc(4 * (size_t)N * 5) || dres.alloc(4 * (size_t)M * N) || dout.alloc(4 * (size_t)M * N) || dop.alloc(2 * (size_t)Mp * old_)) break;
    bool ok = cudaMemcpy(dA32.p, A, 4 * (size_t)M * K, cudaMemcpyHostToDevice) == cudaSuccess && cudaMemcpy(dW32.p, W, 4 * (size_t)N * K, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(dres.p, res, 4 * (size_t)M * N, cudaMemcpyHostToDevice) == cudaSuccess;
    const float* vecs[5] = {bias, g1, b1, g2, b2};
    for (int i = 0; i < 5 && ok; ++i)
      if (vecs[i]) ok = cudaMemcpy(dv.as<float>() + i * N, vecs[i], 4 * (size_t)N, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) { set_error("H2D failed"); break; }
    cudaMemset(dA.p, 0, dA.bytes); cudaMemset(dop.p, 0, dop.bytes);
    if (convert_weight(dA32.as<float>(), dA.as<bf16>(), M, K, ld, lo, 0) || convert_weight(dW32.as<float>(), dW.as<bf16>(), N, K, ld, lo, 0)) break;
    const GemmProblem p = make_problem(M, N, K, split);
    CUtensorMap ta, tb, tb128;
    if (make_tmap_bf16_2d(&ta, dA.p, ld, Mp, ld, 128) || make_tmap_bf16_2d(&tb, dW.p, ld, N, ld, 256) || make_tmap_bf16_2d(&tb128, dW.p, ld, N, ld, 128)) break;
    LnEpilogue ep{dv.as<float>(), dres.as<float>(), dv.as<float>() + N, dv.as<float>() + 2 * N, g2 ? dv.as<float>() + 3 * N : nullptr,
                  b2 ? dv.as<float>() + 4 * N : nullptr, dout.as<float>(), dop.as<bf16>(), old_, olo, f32_normed, compact_rows, compact_seg};
    if (pair == 2 && g2) {                                   // pair = 2: cta_group::2 shape with the one-pass second-LN statistics
      std::vector<float> c(N + 3 * (N / 32) + 2);
      double mb = 0.0, vb = 0.0;
      for (int j = 0; j < N; ++j) mb += b1[j];
      mb /= N;
      for (int j = 0; j < N; ++j) vb += (b1[j] - mb) * (b1[j] - mb);
      vb /= N;
      for (int ch = 0; ch < N / 32; ++ch) {
        double s1 = 0.0, s2 = 0.0, s3 = 0.0;
        for (int j = ch * 32; j < ch * 32 + 32; ++j) { c[j] = (float)(g1[j] * (b1[j] - mb)); s1 += g1[j]; s2 += (double)g1[j] * g1[j]; s3 += c[j]; }
        c[N + ch] = (float)s1; c[N + N / 32 + ch] = (float)s2; c[N + 2 * (N / 32) + ch] = (float)s3;
      }
      c[N + 3 * (N / 32)] = (float)mb; c[N + 3 * (N / 32) + 1] = (float)vb;
      if (dyc.alloc(4 * c.size()) || cudaMemcpy(dyc.p, c.data(), 4 * c.size(), cudaMemcpyHostToDevice) != cudaSuccess) { set_error("H2D failed"); break; }
      ep.y_consts = dyc.as<float>();
    }
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int n_it = iters > 0 ? iters : 1;
    for (int it = 0; it < n_it + 1 && ok; ++it) {
      if (it == 1) cudaEventRecord(e0, 0);
      ok = !gemm_ln(ta, tb, tb128, p, ep, pair == 3 ? 2 : (pair != 0), prop.multiProcessorCount, 0);
    }
    if (!ok) break;
    cudaEventRecord(e1, 0);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { set_error("asr_debug_gemm_ln: kernel failed: %s", cudaGetErrorString(err)); break; }
    if (ms_out) { float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); *ms_out = ms / n_it; }
    hop.resize((size_t)Mo * old_);
    if (cudaMemcpy(out_f32, dout.p, 4 * (size_t)M * N, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(hop.data(), dop.p, 2 * hop.size(), cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("D2H failed"); break; }
    for (int r = 0; r < Mo; ++r)
      for (int c = 0; c < N; ++c)
        out_op_f32[(size_t)r * N + c] = __bfloat162float(hop[(size_t)r * old_ + c]) + (olo ? __bfloat162float(hop[(size_t)r * old_ + olo + c]) : 0.f);
    rc = 0;
  } while (0);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  dA32.free(); dW32.free(); dA.free(); dW.free(); dv.free(); dres.free(); dout.free(); dop.free(); dyc.free();
  return rc;
}

/* Times `iters` launches of the tcgen05 GEMM on random bf16 operands already in HBM (diagnostic microbenchmark).
 * epi: 0 = plain fp32 store, 1 = + bias + fp32 residual, 2 = bias + GELU -> bf16 operand.  Returns mean ms per launch. */
int asr_debug_gemm_time(int32_t M, int32_t N, int32_t K, int32_t split, int32_t bn, int32_t epi_kind, int32_t iters, float* ms_out, int device) {
  if (M <= 0 || N <= 0 || K <= 0 || K % 64 || !ms_out || iters <= 0) { set_error("asr_debug_gemm_time: bad arguments"); return -1; }
  ASR_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  ASR_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  const int ld = split ? 2 * K : K;
  const int Mp = (int)round_up(M, 256);
  DevBuf dA, dB, dC, dR, dO, dbias;
  int rc = -1;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  do {
    if (dA.alloc(2 * (size_t)Mp * ld) || dB.alloc(2 * (size_t)N * ld) || dC.alloc(4 * (size_t)M * N) || dR.alloc(4 * (size_t)M * N) ||
        dO.alloc(2 * (size_t)M * N) || dbias.alloc(4 * (size_t)N)) break;
    cudaMemset(dA.p, 0x3c, dA.bytes); cudaMemset(dB.p, 0x3c, dB.bytes); cudaMemset(dR.p, 0, dR.bytes); cudaMemset(dbias.p, 0, dbias.bytes);
    const GemmProblem p = make_problem(M, N, K, split);
    CUtensorMap ta, tb;
    const bool pair = bn == kPairTile || bn == kPairTileTS;
    if (make_tmap_bf16_2d(&ta, dA.p, ld, Mp, ld, 128) || make_tmap_bf16_2d(&tb, dB.p, ld, N, ld, pair ? 128 : bn)) break;
    TsMaps ts = {};                                        // (bias zero, like dbias)
    if (bn == kPairTileTS && make_tmap_bf16_2d(&ts.c0, dO.p, N, M, N, 32)) break;
    EpiF32 e_plain{dC.as<float>(), nullptr, nullptr, N, N};
    EpiF32 e_res{dC.as<float>(), dbias.as<float>(), dR.as<float>(), N, N};
    EpiOperand e_op{dO.as<bf16>(), dbias.as<float>(), N, 0, ACT_GELU};
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bool ok = true;
    for (int it = 0; it < iters + 3 && ok; ++it) {
      if (it == 3) cudaEventRecord(e0, 0);
      if (epi_kind == 0) ok = !gemm_tc<EpiF32>(ta, tb, p, e_plain, bn, prop.multiProcessorCount, 0);
      else if (epi_kind == 1) ok = !gemm_tc<EpiF32>(ta, tb, p, e_res, bn, prop.multiProcessorCount, 0);
      else if (epi_kind == 3) ok = !gemm_tc<EpiNull>(ta, tb, p, EpiNull{}, bn, prop.multiProcessorCount, 0);
      else ok = !gemm_tc<EpiOperand>(ta, tb, p, e_op, bn, prop.multiProcessorCount, 0, bn == kPairTileTS ? &ts : nullptr);
    }
    if (!ok) break;
    cudaEventRecord(e1, 0);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { set_error("asr_debug_gemm_time: %s", cudaGetErrorString(err)); break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_out = ms / iters;
    rc = 0;
  } while (0);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  dA.free(); dB.free(); dC.free(); dR.free(); dO.free(); dbias.free();
  return rc;
}

}  // extern "C"
