// Native session scheduler: the per-connection loop of the reference server for thousands of sessions at once.
//
// Host-side mirror of handle_connection_impl (streaming_decoder/streaming_server.py:367-546) and of Stream
// (streaming_decoder/stream.py: :23-26 initial zero buffer, :78-87 accept_waveform, :110-125 update_stream, :127-163 endpoint_detected,
// :159-160 advance by segment_length, :166-189 VAD skip) and of the v1 cross-stream batcher StreamingE2E.process
// (streaming_decoder_v1/streaming_asr.py:41-119); endpoint rules of online_endpoint.py:42-94.  Session state is struct-of-arrays
// (one row per session) owned here and shown to Python as numpy views; a tick is two calls:
//
//   asr_sched_submit   ready scan (strict LRU under backlog) -> energy gate for sessions without text in their segment -> skip
//                      bookkeeping -> batch assembly straight into the engine's pinned staging buffer (or a device gather out of pinned
//                      rings) -> H2D + kernel chain + D2H enqueued.  No Python between "which sessions are ready" and the launch.
//   asr_sched_collect  wait for the ticket -> update_stream for every served session straight out of the pinned result area ->
//                      endpoint rules -> the fired sessions' state reset in ONE stream-ordered launch.
//
// The bookkeeping halves (plan / commit / update / endpoints) are also exported on their own: they need no GPU, so the CPU tests drive
// them with a scripted engine against a scalar restatement of the reference loop, and a caller with its own VAD or language model
// (relative cost of utils.py:126-139) interposes between them.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <numeric>
#include <stdlib.h>
#include <thread>
#include <vector>

#include "../../include/asr_b200.h"
#include "common.cuh"
#include "sched_hooks.h"

using namespace asr;

namespace {

struct Tick {
  bool active = false;
  int ticket = -1;                 // engine ticket (real engine only)
  std::vector<int32_t> rows;       // sessions run through the model, batch order
  std::vector<int32_t> slots;
  std::vector<int64_t> offsets;
  std::vector<int32_t> skipped;    // VAD-gated sessions of the tick (not run)
  // results (filled by update / endpoints)
  std::vector<int32_t> n_new, new_tok;     // [n], [n * S]
  std::vector<uint8_t> final_, overflow;   // [n]
  std::vector<int32_t> final_rule;         // [n] rule index or -1
  std::vector<int32_t> fin_rows, fin_rule, fin_ntok, fin_tok_off;    // endpoints of the tick (served and skipped sessions), in firing order
  std::vector<double> fin_utt;
  std::vector<int32_t> fin_tok;            // concatenated tokens of the finished segments
  StepView view;                           // pinned result area of the step (real engine)
};

}  // namespace

struct AsrScheduler {
  AsrEngine* eng = nullptr;
  AsrSchedConfig cfg;
  int CAP = 0;
  bool pinned = false;
  std::mutex mu;
  // per-session arrays (capacity rows)
  int16_t* audio = nullptr;
  std::vector<int64_t> rd, wr, n_frames, chunk_processed, chunk_total, segment, last_served;
  std::vector<uint8_t> active, inflight, contain, overflow;
  std::vector<int32_t> slot, tok, ntok;
  std::vector<double> trailing, rel_cost;
  int64_t seq = 0;
  // endpoint rules (online_endpoint.py:4-21), declaration order = evaluation order
  std::vector<uint8_t> must;
  std::vector<double> min_sil, min_utt, max_cost;
  Tick tick[2];
  int next_tick = 0;
  std::vector<int32_t> ready;      // scratch: the last ready scan
  std::vector<int32_t> scratch_rows;
  std::vector<int32_t> peaks;
  // pre-staging (asr_sched_prestage): chunks gathered + copied to the device before the tick that runs them is decided
  std::vector<int64_t> abs_rd;     // samples consumed per session since open (ring compaction moves rd, not this)
  bool pre_valid = false;
  bool ring_reads_pending = false;   // device gather: a gather kernel may still be reading the pinned rings (compaction must wait for it)
  std::vector<int32_t> pre_rows, pre_peaks, pre_index_of;     // staged sessions; gate peaks of their chunks; session row -> staged row (-1)
  std::vector<int64_t> pre_abs;    // abs_rd of every staged chunk: a chunk consumed otherwise in the meantime (VAD skip) is not run from the stage
};

namespace {

double chunk_seconds(const AsrScheduler* s) { return (double)s->cfg.segment_length / (double)s->cfg.sample_rate; }

// numpy's round(x, 2) (round-half-even on x * 100, then / 100): what `round(self.trailing_blank_duration, 2)` (stream.py:140) gives
double round2(double x) { return nearbyint(x * 100.0) / 100.0; }

void clear_segment(AsrScheduler* s, int r) {
  s->ntok[r] = 0; s->n_frames[r] = 0; s->chunk_processed[r] = 0; s->contain[r] = 0; s->trailing[r] = 0.0; s->overflow[r] = 0;
}

void advance(AsrScheduler* s, int r) {
  s->rd[r] += s->cfg.segment_length;                       // stream.py:159-160
  s->abs_rd[r] += s->cfg.segment_length;
  s->last_served[r] = s->seq++;
}

// sessions with a full chunk buffered and none in flight; under backlog the longest-waiting first (nobody starves)
void ready_scan(AsrScheduler* s, int max_rows) {
  const int cap = s->cfg.capacity;
  int lim = s->cfg.max_batch;
  if (max_rows > 0 && max_rows < lim) lim = max_rows;
  s->ready.clear();
  if (s->pre_valid) {                                       // only chunks that are already on their way to the device
    for (size_t i = 0; i < s->pre_rows.size(); ++i) {
      const int r = s->pre_rows[i];
      if (s->active[r] && !s->inflight[r] && s->abs_rd[r] == s->pre_abs[i] && s->wr[r] - s->rd[r] >= s->cfg.chunk_length) s->ready.push_back(r);
    }
  } else {
    for (int r = 0; r < cap; ++r)
      if (s->active[r] && !s->inflight[r] && s->wr[r] - s->rd[r] >= s->cfg.chunk_length) s->ready.push_back(r);
  }
  if ((int)s->ready.size() > lim) {
    std::stable_sort(s->ready.begin(), s->ready.end(), [&](int a, int b) { return s->last_served[a] < s->last_served[b]; });
    s->ready.resize(lim);
  }
}

// detect_endpointing for one session (online_endpoint.py:69-94): index of the first activated rule, -1 if none
int first_rule(const AsrScheduler* s, double utt, double sil, double cost) {
  const bool nonsil = utt > sil;                            // online_endpoint.py:59
  for (size_t k = 0; k < s->must.size(); ++k)
    if ((nonsil || !s->must[k]) && sil >= s->min_sil[k] && cost < s->max_cost[k] && utt >= s->min_utt[k]) return (int)k;
  return -1;
}

// endpoint_detected (stream.py:127-157) for the listed sessions; fired sessions are recorded in the tick and their segment state cleared
void run_endpoints(AsrScheduler* s, Tick& t, const std::vector<int32_t>& rows, bool served) {
  if (s->must.empty()) return;
  for (size_t j = 0; j < rows.size(); ++j) {
    const int r = rows[j];
    const double utt = (double)s->chunk_processed[r] * (double)s->cfg.segment_length / (double)s->cfg.sample_rate;
    s->trailing[r] = round2(s->trailing[r]);
    const int k = first_rule(s, utt, s->trailing[r], s->rel_cost[r]);
    if (k < 0) continue;
    t.fin_rows.push_back(r); t.fin_rule.push_back(k); t.fin_utt.push_back(utt);
    t.fin_ntok.push_back(s->ntok[r]); t.fin_tok_off.push_back((int32_t)t.fin_tok.size());
    t.fin_tok.insert(t.fin_tok.end(), s->tok.begin() + (size_t)r * s->cfg.max_tokens, s->tok.begin() + (size_t)r * s->cfg.max_tokens + s->ntok[r]);
    if (served) { t.final_[j] = 1; t.final_rule[j] = k; }
    clear_segment(s, r);
    s->segment[r] += 1;
  }
}

// plan: ready scan -> gate -> skip bookkeeping (+ endpoint rules on the skipped sessions).  keep: optional caller-computed gate decision
// for s->ready (the caller ran asr_sched_ready first); gate_threshold >= 0: native energy gate on the new samples of the chunk.
int plan_tick(AsrScheduler* s, Tick& t, int max_rows, int gate_threshold, const uint8_t* keep) {
  if (!keep) ready_scan(s, max_rows);
  t = Tick();
  const int n_ready = (int)s->ready.size();
  std::vector<uint8_t> kp(n_ready, 1);
  if (keep) memcpy(kp.data(), keep, n_ready);
  else if (gate_threshold >= 0 && n_ready && s->pre_valid && !s->pre_peaks.empty()) {
    for (int i = 0; i < n_ready; ++i)
      if (!s->contain[s->ready[i]]) kp[i] = s->pre_peaks[s->pre_index_of[s->ready[i]]] >= gate_threshold;
  } else if (gate_threshold >= 0 && n_ready) {
    // the gate is consulted only for sessions without text in the current segment (stream.py:166-189, streaming_server.py:374-379)
    s->scratch_rows.clear();
    std::vector<int64_t> offs;
    for (int i = 0; i < n_ready; ++i) if (!s->contain[s->ready[i]]) { s->scratch_rows.push_back(s->ready[i]); offs.push_back(s->rd[s->ready[i]]); }
    s->peaks.resize(s->scratch_rows.size());
    if (!s->scratch_rows.empty() &&
        asr_pcm_peaks((int32_t)s->scratch_rows.size(), s->audio, s->CAP, s->scratch_rows.data(), offs.data(), s->cfg.buffer_length, s->cfg.chunk_length, s->peaks.data()))
      return -1;
    size_t q = 0;
    for (int i = 0; i < n_ready; ++i) if (!s->contain[s->ready[i]]) kp[i] = s->peaks[q++] >= gate_threshold;
  }
  for (int i = 0; i < n_ready; ++i) {
    const int r = s->ready[i];
    if (kp[i]) { t.rows.push_back(r); t.slots.push_back(s->slot[r]); t.offsets.push_back(s->rd[r]); }
    else t.skipped.push_back(r);
  }
  return 0;
}

// after the step was enqueued successfully: the served sessions are in flight and their buffers advance; the skipped chunks are consumed
// without touching encoder state (stream.py:183-189) and run through the endpoint rules
void commit_tick(AsrScheduler* s, Tick& t) {
  for (int r : t.skipped) {
    s->trailing[r] += chunk_seconds(s);
    s->chunk_processed[r] += 1; s->chunk_total[r] += 1;
    advance(s, r);
  }
  const size_t n = t.rows.size();
  t.final_.assign(n, 0); t.final_rule.assign(n, -1); t.overflow.assign(n, 0);
  run_endpoints(s, t, t.skipped, false);
  for (int r : t.rows) { s->inflight[r] = 1; advance(s, r); }
  t.active = true;
}

// update_stream (stream.py:110-125) for every served session from the step's outputs
void update_tick(AsrScheduler* s, Tick& t, const int32_t* n_new, const int32_t* new_tokens, const int32_t* blank_frames, const int32_t* has_token,
                 const int32_t* has_text, const int32_t* flags) {
  const int S = s->cfg.seg_rows, MT = s->cfg.max_tokens;
  const size_t n = t.rows.size();
  t.n_new.assign(n_new, n_new + n);
  t.new_tok.assign(new_tokens, new_tokens + n * S);
  for (size_t j = 0; j < n; ++j) {
    const int r = t.rows[j];
    s->inflight[r] = 0;
    int32_t* tk = s->tok.data() + (size_t)r * MT;
    for (int i = 0; i < n_new[j]; ++i) {
      if (s->ntok[r] < MT) tk[s->ntok[r]++] = new_tokens[j * S + i];
      else s->overflow[r] = 1;                                 // never silently: the tick reports it (TickResult.overflow)
    }
    if (flags && (flags[j] & ASR_FLAG_BEAM_TRUNCATED)) s->overflow[r] = 1;
    t.overflow[j] = s->overflow[r];
    s->n_frames[r] += S;
    s->chunk_processed[r] += 1; s->chunk_total[r] += 1;
    const bool text = (has_text ? has_text[j] : has_token[j]) != 0;
    if (text) {
      // greedy_search's last_blank (recognition.py:38-43): float32 product when an id > 1 exists (it does when the text is non-empty)
      s->trailing[r] = has_token[j] ? (double)((float)blank_frames[j] * 0.04f) : 0.04 * (double)blank_frames[j];
      s->contain[r] = 1;
    } else {
      s->trailing[r] += chunk_seconds(s);                      // stream.py:124-125
    }
  }
}

int check_row(const AsrScheduler* s, int row) {
  if (row < 0 || row >= s->cfg.capacity || !s->active[row]) { set_error("scheduler row %d is not an open session", row); return -1; }
  return 0;
}

void fill_result(const Tick& t, AsrSchedResult* res) {
  if (!res) return;
  memset(res, 0, sizeof(*res));
  res->n = (int32_t)t.rows.size();
  res->rows = t.rows.data(); res->n_new = t.n_new.data(); res->new_tokens = t.new_tok.data();
  res->final_flags = t.final_.data(); res->final_rule = t.final_rule.data(); res->overflow = t.overflow.data();
  res->n_skipped = (int32_t)t.skipped.size(); res->skipped = t.skipped.data();
  res->n_final = (int32_t)t.fin_rows.size(); res->final_rows = t.fin_rows.data(); res->final_rule_of = t.fin_rule.data();
  res->final_utt = t.fin_utt.data(); res->final_ntok = t.fin_ntok.data(); res->final_tok_off = t.fin_tok_off.data(); res->final_tok = t.fin_tok.data();
  res->argmax_ids = t.view.argmax_ids; res->blank_frames = t.view.blank_frames; res->has_token = t.view.has_token; res->has_text = t.view.has_text;
  res->flags = t.view.flags; res->beam_tokens = t.view.beam_tokens; res->beam_len = t.view.beam_len; res->beam_score = t.view.beam_score;
  res->logprobs = t.view.logprobs;
}

}  // namespace

extern "C" {

int asr_sched_create(const AsrSchedConfig* cfg, AsrEngine* engine, AsrScheduler** out) {
  if (!cfg || !out) { set_error("null argument"); return -1; }
  if (cfg->capacity <= 0 || cfg->chunk_length <= 0 || cfg->segment_length <= 0 || cfg->buffer_length < 0 || cfg->seg_rows <= 0 || cfg->max_batch <= 0 ||
      cfg->backlog_chunks < 0 || cfg->max_tokens <= 0 || cfg->sample_rate <= 0) { set_error("asr_sched_create: bad configuration"); return -1; }
  if (cfg->device_gather && !engine) { set_error("asr_sched_create: device gather needs an engine (pinned rings)"); return -1; }
  AsrScheduler* s = new AsrScheduler();
  s->eng = engine; s->cfg = *cfg;
  s->CAP = cfg->chunk_length + cfg->backlog_chunks * cfg->segment_length;
  const size_t n = (size_t)cfg->capacity;
  const size_t audio_bytes = n * (size_t)s->CAP * sizeof(int16_t);
  if (cfg->device_gather) {
    s->audio = reinterpret_cast<int16_t*>(asr_host_alloc(audio_bytes));
    s->pinned = true;
  } else {
    s->audio = reinterpret_cast<int16_t*>(malloc(audio_bytes ? audio_bytes : 16));
  }
  if (!s->audio) { if (!cfg->device_gather) set_error("asr_sched_create: cannot allocate %zu bytes of audio rings", audio_bytes); delete s; return -1; }
  memset(s->audio, 0, audio_bytes);
  s->rd.assign(n, 0); s->wr.assign(n, 0); s->n_frames.assign(n, 0); s->chunk_processed.assign(n, 0); s->chunk_total.assign(n, 0);
  s->segment.assign(n, 0); s->last_served.assign(n, 0); s->abs_rd.assign(n, 0); s->pre_index_of.assign(n, -1);
  s->active.assign(n, 0); s->inflight.assign(n, 0); s->contain.assign(n, 0); s->overflow.assign(n, 0);
  s->slot.assign(n, -1); s->tok.assign(n * (size_t)cfg->max_tokens, 0); s->ntok.assign(n, 0);
  s->trailing.assign(n, 0.0); s->rel_cost.assign(n, cfg->relative_cost);
  *out = s;
  return 0;
}

int asr_sched_destroy(AsrScheduler* s) {
  if (!s) return 0;
  if (s->audio) { if (s->pinned) asr_host_free(s->audio); else free(s->audio); }
  delete s;
  return 0;
}

int asr_sched_arrays(AsrScheduler* s, AsrSchedArrays* a) {
  if (!s || !a) { set_error("null argument"); return -1; }
  a->audio = s->audio; a->audio_row_samples = s->CAP;
  a->rd = s->rd.data(); a->wr = s->wr.data(); a->active = s->active.data(); a->inflight = s->inflight.data(); a->slot = s->slot.data();
  a->tok = s->tok.data(); a->ntok = s->ntok.data(); a->n_frames = s->n_frames.data(); a->chunk_processed = s->chunk_processed.data();
  a->chunk_processed_total = s->chunk_total.data(); a->trailing = s->trailing.data(); a->contain_token = s->contain.data();
  a->segment = s->segment.data(); a->last_served = s->last_served.data(); a->relative_cost = s->rel_cost.data(); a->overflow = s->overflow.data();
  return 0;
}

int asr_sched_set_rules(AsrScheduler* s, int32_t n, const uint8_t* must_contain_nonsilence, const double* min_trailing_silence,
                        const double* min_utterance_length, const double* max_relative_cost) {
  if (!s || n < 0 || (n > 0 && (!must_contain_nonsilence || !min_trailing_silence || !min_utterance_length || !max_relative_cost))) { set_error("asr_sched_set_rules: bad argument"); return -1; }
  std::lock_guard<std::mutex> lk(s->mu);
  s->must.assign(must_contain_nonsilence, must_contain_nonsilence + n);
  s->min_sil.assign(min_trailing_silence, min_trailing_silence + n);
  s->min_utt.assign(min_utterance_length, min_utterance_length + n);
  s->max_cost.assign(max_relative_cost, max_relative_cost + n);
  return 0;
}

int asr_sched_open(AsrScheduler* s, int32_t row, int32_t slot) {
  if (!s) { set_error("null scheduler"); return -1; }
  std::lock_guard<std::mutex> lk(s->mu);
  if (row < 0 || row >= s->cfg.capacity || s->active[row]) { set_error("asr_sched_open: row %d is taken or out of range", row); return -1; }
  if (s->eng && slot < 0 && engine_open_slot(s->eng, &slot)) return -1;
  s->slot[row] = slot;
  memset(s->audio + (size_t)row * s->CAP, 0, sizeof(int16_t) * (size_t)s->cfg.buffer_length);     // stream.py:23: buffer_length leading zeros
  s->rd[row] = 0; s->wr[row] = s->cfg.buffer_length; s->abs_rd[row] = 0;
  s->active[row] = 1; s->inflight[row] = 0;
  clear_segment(s, row);
  s->chunk_total[row] = 0; s->segment[row] = 0; s->rel_cost[row] = s->cfg.relative_cost;
  s->last_served[row] = s->seq++;
  return 0;
}

int asr_sched_close(AsrScheduler* s, int32_t row) {
  if (!s) { set_error("null scheduler"); return -1; }
  std::lock_guard<std::mutex> lk(s->mu);
  if (check_row(s, row)) return -1;
  if (s->inflight[row]) { set_error("close: the session has a chunk in flight; collect its tick first"); return -1; }
  if (s->eng && engine_close_slot(s->eng, s->slot[row])) return -1;
  s->active[row] = 0; s->slot[row] = -1;
  return 0;
}

/* Endpoint decided by the caller (final-pass decoder, client EOS): emission := [], state := init (streaming_server.py:514-515, :530). */
int asr_sched_reset_rows(AsrScheduler* s, int32_t n, const int32_t* rows) {
  if (!s || (n > 0 && !rows)) { set_error("null argument"); return -1; }
  std::lock_guard<std::mutex> lk(s->mu);
  std::vector<int32_t> slots;
  for (int i = 0; i < n; ++i) {
    if (check_row(s, rows[i])) return -1;
    if (s->inflight[rows[i]]) { set_error("reset: session row %d has a chunk in flight; collect its tick first", rows[i]); return -1; }
    slots.push_back(s->slot[rows[i]]);
  }
  if (s->eng && n > 0 && engine_reset_async(s->eng, n, slots.data())) return -1;
  for (int i = 0; i < n; ++i) { clear_segment(s, rows[i]); s->segment[rows[i]] += 1; }
  return 0;
}

namespace {
// device gather: a gather kernel reads the chunks straight out of the pinned rings; moving ring contents must wait for it
int wait_ring_reads(AsrScheduler* s) {
  if (!s->pinned || !s->ring_reads_pending) return 0;
  if (engine_wait_inputs(s->eng)) return -1;
  s->ring_reads_pending = false;
  return 0;
}

// stream.py:78-87 for one session (lock held): messages of <= 100 samples are dropped; the unread tail moves to the front when the ring is full
int accept_locked(AsrScheduler* s, int32_t row, const int16_t* pcm, int64_t n) {
  if (check_row(s, row)) return -1;
  if (n <= 100) return 0;
  int16_t* a = s->audio + (size_t)row * s->CAP;
  if (s->wr[row] + n > s->CAP) {
    const int64_t live = s->wr[row] - s->rd[row];
    if (live + n > s->CAP) { set_error("session row %d: backlog of %lld samples exceeds the %d-sample buffer", row, (long long)(live + n), s->CAP); return 1; }
    // (device gather: the caller has waited for pending ring reads before any compaction, see wait_ring_reads)
    memmove(a, a + s->rd[row], sizeof(int16_t) * (size_t)live);
    s->rd[row] = 0; s->wr[row] = live;
  }
  memcpy(a + s->wr[row], pcm, sizeof(int16_t) * (size_t)n);
  s->wr[row] += n;
  return 0;
}
}  // namespace

/* stream.py:78-87: append int16 samples (messages of <= 100 samples are dropped); returns 1 when the backlog does not fit. */
int asr_sched_accept(AsrScheduler* s, int32_t row, const int16_t* pcm, int64_t n) {
  if (!s || (n > 0 && !pcm)) { set_error("null argument"); return -1; }
  std::lock_guard<std::mutex> lk(s->mu);
  if (row >= 0 && row < s->cfg.capacity && s->wr[row] + n > s->CAP && wait_ring_reads(s)) return -1;
  return accept_locked(s, row, pcm, n);
}

/* The same for many sessions at once: block[i * samples .. + samples) is appended to session rows[i]. */
int asr_sched_accept_block(AsrScheduler* s, int32_t n, const int32_t* rows, const int16_t* block, int64_t samples) {
  if (!s || (n > 0 && (!rows || !block))) { set_error("null argument"); return -1; }
  std::lock_guard<std::mutex> lk(s->mu);
  for (int i = 0; i < n; ++i) {                             // validate first: the parallel part below cannot fail half way
    if (check_row(s, rows[i])) return -1;
    if (samples > 100 && s->wr[rows[i]] - s->rd[rows[i]] + samples > s->CAP) {
      set_error("session row %d: backlog of %lld samples exceeds the %d-sample buffer", rows[i], (long long)(s->wr[rows[i]] - s->rd[rows[i]] + samples), s->CAP);
      return 1;
    }
  }
  for (int i = 0; i < n; ++i)
    if (s->wr[rows[i]] + samples > s->CAP) { if (wait_ring_reads(s)) return -1; break; }      // some ring will be compacted
  static const int env_cap = [] { const char* v = getenv("ASR_B200_HOST_THREADS"); return v ? std::max(1, atoi(v)) : 8; }();
  const int hw = (int)std::thread::hardware_concurrency();
  const int nt = std::max(1, std::min({8, env_cap, hw > 0 ? hw : 1, n / 64 + 1}));
  std::vector<int> rcs(nt, 0);
  std::vector<std::thread> th;
  auto work = [&](int t) {
    for (int i = (int)((long long)n * t / nt); i < (int)((long long)n * (t + 1) / nt); ++i)
      if (int rc = accept_locked(s, rows[i], block + (size_t)i * samples, samples)) { rcs[t] = rc; return; }
  };
  for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  for (int rc : rcs) if (rc) return rc;
  return 0;
}

int asr_sched_ready(AsrScheduler* s, int32_t max_rows, const int32_t** rows, int32_t* n) {
  if (!s || !rows || !n) { set_error("null argument"); return -1; }
  std::lock_guard<std::mutex> lk(s->mu);
  ready_scan(s, max_rows);
  *rows = s->ready.data(); *n = (int32_t)s->ready.size();
  return 0;
}

int asr_sched_plan(AsrScheduler* s, int32_t max_rows, int32_t gate_threshold, const uint8_t* keep, AsrSchedPlan* plan) {
  if (!s || !plan) { set_error("null argument"); return -1; }
  std::lock_guard<std::mutex> lk(s->mu);
  const int id = s->next_tick;
  Tick& t = s->tick[id];
  if (t.active) { set_error("two ticks are already in flight: collect the oldest first"); return -1; }
  if (plan_tick(s, t, max_rows, gate_threshold, keep)) return -1;
  plan->tick = id; plan->n = (int32_t)t.rows.size(); plan->rows = t.rows.data(); plan->slots = t.slots.data(); plan->offsets = t.offsets.data();
  plan->n_skipped = (int32_t)t.skipped.size(); plan->skipped = t.skipped.data();
  return 0;
}

int asr_sched_commit(AsrScheduler* s, int32_t tick, AsrSchedResult* res) {
  if (!s || tick < 0 || tick > 1 || tick != s->next_tick) { set_error("asr_sched_commit: not the planned tick"); return -1; }
  std::lock_guard<std::mutex> lk(s->mu);
  Tick& t = s->tick[tick];
  commit_tick(s, t);
  if (t.rows.empty()) t.active = false;                      // nothing to collect: the tick is complete (skips and their endpoints only) ...
  else s->next_tick ^= 1;                                    // ... and its slot is reused by the next plan
  fill_result(t, res);
  return 0;
}

int asr_sched_update(AsrScheduler* s, int32_t tick, const AsrStepOut* out) {
  if (!s || tick < 0 || tick > 1 || !out || !out->n_new || !out->new_tokens || !out->blank_frames || !out->has_token) { set_error("asr_sched_update: bad argument"); return -1; }
  std::lock_guard<std::mutex> lk(s->mu);
  Tick& t = s->tick[tick];
  if (!t.active) { set_error("asr_sched_update: tick %d is not in flight", tick); return -1; }
  update_tick(s, t, out->n_new, out->new_tokens, out->blank_frames, out->has_token, out->has_text, out->flags);
  return 0;
}

int asr_sched_endpoints(AsrScheduler* s, int32_t tick, AsrSchedResult* res) {
  if (!s || tick < 0 || tick > 1) { set_error("asr_sched_endpoints: bad tick"); return -1; }
  std::lock_guard<std::mutex> lk(s->mu);
  Tick& t = s->tick[tick];
  if (!t.active) { set_error("asr_sched_endpoints: tick %d is not in flight", tick); return -1; }
  const size_t before = t.fin_rows.size();
  run_endpoints(s, t, t.rows, true);
  t.active = false;
  if (s->eng && t.fin_rows.size() > before) {
    std::vector<int32_t> slots;
    for (size_t i = before; i < t.fin_rows.size(); ++i) slots.push_back(s->slot[t.fin_rows[i]]);
    if (engine_reset_async(s->eng, (int)slots.size(), slots.data())) return -1;
  }
  fill_result(t, res);
  return 0;
}

/* A session dropped out of a failed tick: nothing of it was applied, it becomes eligible again (its chunk is lost). */
int asr_sched_abort(AsrScheduler* s, int32_t tick) {
  if (!s || tick < 0 || tick > 1) { set_error("asr_sched_abort: bad tick"); return -1; }
  std::lock_guard<std::mutex> lk(s->mu);
  Tick& t = s->tick[tick];
  for (int r : t.rows) s->inflight[r] = 0;
  t.active = false;
  return 0;
}

/* Pre-staging for one-tick-per-pass pipelining: every session that has a full chunk buffered — INCLUDING the sessions of the tick still in
 * flight, whose next chunk is already in their ring — is gathered into the next step's pinned staging buffer and its H2D copy starts now,
 * overlapping the kernels of the running tick.  Which of the staged chunks run is decided by the next asr_sched_submit, after the running
 * tick was collected (VAD gate, endpoint resets): it launches on a subset through a row-index indirection.  gate_threshold >= 0 also
 * precomputes the energy-gate peaks of the staged chunks.  Returns the number of staged chunks (0: nothing staged, the next submit
 * assembles its batch the usual way).  With device gather (pinned rings) the gather kernel is what gets pre-issued. */
int asr_sched_prestage(AsrScheduler* s, int32_t gate_threshold, int32_t* n_staged) {
  if (!s) { set_error("null scheduler"); return -1; }
  if (!s->eng) { set_error("asr_sched_prestage needs an engine"); return -1; }
  std::lock_guard<std::mutex> lk(s->mu);
  for (int r : s->pre_rows) s->pre_index_of[r] = -1;
  s->pre_valid = false; s->pre_rows.clear(); s->pre_abs.clear(); s->pre_peaks.clear();
  if (n_staged) *n_staged = 0;
  const int cap = s->cfg.capacity;
  std::vector<int64_t> offs;
  for (int r = 0; r < cap && (int)s->pre_rows.size() < s->cfg.max_batch; ++r)
    if (s->active[r] && s->wr[r] - s->rd[r] >= s->cfg.chunk_length) { s->pre_rows.push_back(r); s->pre_abs.push_back(s->abs_rd[r]); offs.push_back(s->rd[r]); }
  if (s->pre_rows.empty()) return 0;
  const int n = (int)s->pre_rows.size();
  if (s->pinned) s->ring_reads_pending = true;
  if (engine_prestage(s->eng, n, s->audio, s->CAP, s->pre_rows.data(), offs.data(), s->pinned)) { s->pre_rows.clear(); s->pre_abs.clear(); return -1; }
  if (gate_threshold >= 0) {
    s->pre_peaks.resize(n);
    if (asr_pcm_peaks(n, s->audio, s->CAP, s->pre_rows.data(), offs.data(), s->cfg.buffer_length, s->cfg.chunk_length, s->pre_peaks.data())) return -1;
  }
  for (int i = 0; i < n; ++i) s->pre_index_of[s->pre_rows[i]] = i;
  s->pre_valid = true;
  if (n_staged) *n_staged = n;
  return 0;
}

int asr_sched_submit(AsrScheduler* s, int32_t max_rows, int32_t gate_threshold, const uint8_t* keep, int32_t want_logprobs, AsrSchedResult* res, int32_t* tick_out) {
  if (!s || !tick_out) { set_error("null argument"); return -1; }
  if (!s->eng) { set_error("asr_sched_submit needs an engine (use plan / commit / update / endpoints with your own)"); return -1; }
  std::unique_lock<std::mutex> lk(s->mu);
  const int id = s->next_tick;
  Tick& t = s->tick[id];
  if (t.active) { set_error("two ticks are already in flight: collect the oldest first"); return -1; }
  if (plan_tick(s, t, max_rows, gate_threshold, keep)) return -1;
  const bool staged = s->pre_valid;
  s->pre_valid = false;                                      // a stage serves one tick
  if (!t.rows.empty()) {
    // nothing of the tick is applied before the step is enqueued: a failure here leaves every session as it was
    if (staged) {
      std::vector<int32_t> idx(t.rows.size());
      for (size_t i = 0; i < t.rows.size(); ++i) idx[i] = s->pre_index_of[t.rows[i]];
      if (engine_submit_prestaged(s->eng, (int)t.rows.size(), t.slots.data(), idx.data(), want_logprobs != 0, &t.ticket)) return -1;
    } else {
      if (s->pinned) s->ring_reads_pending = true;
      if (engine_submit_gather(s->eng, (int)t.rows.size(), t.slots.data(), s->audio, s->CAP, t.rows.data(), t.offsets.data(), s->pinned, want_logprobs != 0, &t.ticket)) return -1;
    }
  }
  if (staged) for (int r : s->pre_rows) s->pre_index_of[r] = -1;
  commit_tick(s, t);
  if (!t.rows.empty()) s->next_tick ^= 1;                    // a tick without a step has nothing to collect: its slot is reused by the next submit
  if (!t.fin_rows.empty()) {                      // endpoints of skipped sessions: reset their (idle) encoder state
    std::vector<int32_t> slots;
    for (int r : t.fin_rows) slots.push_back(s->slot[r]);
    if (engine_reset_async(s->eng, (int)slots.size(), slots.data())) return -1;
  }
  if (t.rows.empty()) t.active = false;
  fill_result(t, res);
  *tick_out = id;
  return 0;
}

int asr_sched_collect(AsrScheduler* s, int32_t tick, int32_t endpoints, AsrSchedResult* res) {
  if (!s || tick < 0 || tick > 1) { set_error("asr_sched_collect: bad tick"); return -1; }
  if (!s->eng) { set_error("asr_sched_collect needs an engine"); return -1; }
  Tick& t = s->tick[tick];
  {
    std::lock_guard<std::mutex> lk(s->mu);
    if (!t.active) { set_error("asr_sched_collect: tick %d is not in flight", tick); return -1; }
  }
  // the wait for the device happens outside the scheduler lock: the receive path (asr_sched_accept) keeps running; the tick's own
  // sessions are in flight, nobody else touches them
  StepView view;
  const int rc = engine_collect_view(s->eng, t.ticket, &view);
  std::lock_guard<std::mutex> lk(s->mu);
  if (rc) {
    for (int r : t.rows) s->inflight[r] = 0;               // the step is lost; its sessions must not stay locked out of the next ticks
    t.active = false;
    return -1;
  }
  t.view = view;
  update_tick(s, t, t.view.n_new, t.view.new_tokens, t.view.blank_frames, t.view.has_token, t.view.has_text, t.view.flags);
  if (endpoints) {
    const size_t before = t.fin_rows.size();
    run_endpoints(s, t, t.rows, true);
    t.active = false;
    if (t.fin_rows.size() > before) {
      std::vector<int32_t> slots;
      for (size_t i = before; i < t.fin_rows.size(); ++i) slots.push_back(s->slot[t.fin_rows[i]]);
      if (engine_reset_async(s->eng, (int)slots.size(), slots.data())) return -1;
    }
  }
  fill_result(t, res);
  return 0;
}

}  // extern "C"
