// CTC prefix beam search, one warp per stream, beam state carried across chunks (north-star: "CTC greedy/prefix
// beam search as a warp-per-stream kernel"; beam = 10 in BASELINE config #4).  The reference has no CTC prefix beam
// (its final pass is the third-party flashlight lexicon decoder, recognition.py:220-300), so the semantics are those
// of oracle/ctc_beam_oracle.py (Hannun et al. 2014, Alg. 1 without LM; parity unpinned vs the reference):
//   candidates per frame  = [stay(j) for the B beam entries] + [ext(i, k) for parent i x the cand_k best non-blank tokens]
//   stay(j)               : p_b' = p_tot(j) + lp[blank];  p_nb' = p_nb(j) + lp[last(j)]  (+ merged extensions that spell j)
//   ext(i, k)             : p_nb' = (tok_k == last(i) ? p_b(i) : p_tot(i)) + lp[tok_k]
//   next beam             = the `beam` best by logaddexp(p_b', p_nb'), ties to the lower enumeration index.
// What is serial per stream is only the recursion over the chunk's frames; everything that is not was moved out of it:
//   * the cand_k best tokens of every frame come from ctc_greedy_kernel (parallel over rows), together with the row's (max, lse),
//     so a frame costs one gathered logit per beam entry (lp[last(j)]) instead of a scan of the 804-wide row, and the
//     [rows, vocab] log-prob array is neither written nor read;
//   * hypotheses are not copied per frame: during the chunk an entry carries (origin = entry at chunk start, tokens appended
//     since), and the token strings — per-session double buffer in HBM, [2][16][1024] int16 — are rebuilt ONCE per chunk,
//     warp-parallel, 16 bytes per lane.
// Prefix identity is a 64-bit rolling hash + length.
#include "kernels.cuh"

namespace asr {

namespace {

constexpr int BW = 4;            // warps (streams) per CTA
constexpr int BSEG = 32;         // frames per chunk limit (seg_rows is 16, or 8 in the low-latency geometry)

__device__ __forceinline__ float lae(float a, float b) {        // logaddexp with -inf handling
  const float m = fmaxf(a, b), n = fminf(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + log1pf(expf(n - m));
}

__device__ __forceinline__ unsigned long long hext(unsigned long long h, int c) {
  h ^= (unsigned long long)(c + 1) * 0x9E3779B97F4A7C15ull;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 29;
  return h;
}

__device__ __forceinline__ uint32_t bkey(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }

struct __align__(16) WarpState {
  int16_t app[BEAM_MAX][BSEG];                                    // tokens appended to the entry since the chunk began
  int len[BEAM_MAX], last[BEAM_MAX];
  float pb[BEAM_MAX], pnb[BEAM_MAX], ptot[BEAM_MAX];
  unsigned long long hash[BEAM_MAX];
  int org[BEAM_MAX], app_n[BEAM_MAX];                             // entry at chunk start this one descends from; tokens appended since
  float npb[BEAM_MAX], npnb[BEAM_MAX], ntot[BEAM_MAX];
  int dead[BEAM_MAX];
  int sel[BEAM_MAX];
  int ctok[BSEG * BEAM_CAND_MAX];                                 // the chunk's extension candidates, all frames
  float clp[BSEG * BEAM_CAND_MAX];
  float rmax[BSEG], rlse[BSEG], lp_blank[BSEG];
};

__global__ void __launch_bounds__(BW * 32) beam_kernel(BeamParams P) {
  __shared__ WarpState ws_all[BW];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * BW + warp;
  pdl_launch_dependents();
  pdl_wait();
  if (w >= P.n) return;
  WarpState& S = ws_all[warp];
  const int slot = P.slots[w];
  const int B = P.beam, K = P.cand_k, SR = P.seg_rows;
  int nb = P.n_beam[slot];
  const int cur = P.cur[slot];
  const size_t row0 = (size_t)w * SR;
  if (lane < BEAM_MAX) {
    const size_t o = (size_t)slot * BEAM_MAX + lane;
    S.len[lane] = P.len[o]; S.last[lane] = P.last[o]; S.pb[lane] = P.pb[o]; S.pnb[lane] = P.pnb[o]; S.hash[lane] = P.hash[o];
    S.org[lane] = lane; S.app_n[lane] = 0;
  }
  for (int i = lane; i < SR * BEAM_CAND_MAX; i += 32) { S.ctok[i] = P.cand_tok[row0 * BEAM_CAND_MAX + i]; S.clp[i] = P.cand_lp[row0 * BEAM_CAND_MAX + i]; }
  if (lane < SR) {
    const float m = P.row_stat[2 * (row0 + lane)], lse = P.row_stat[2 * (row0 + lane) + 1];
    S.rmax[lane] = m; S.rlse[lane] = lse;
    S.lp_blank[lane] = (P.logits[(row0 + lane) * P.vocab] - m) - lse;
  }
  __syncwarp();

  for (int r = 0; r < SR; ++r) {
    const int* ct = S.ctok + r * BEAM_CAND_MAX;
    const float* cl = S.clp + r * BEAM_CAND_MAX;
    // log-prob of the entry's last token in this frame: the one value of the row that is not among the candidates in general
    float lp_last = -INFINITY;
    if (lane < nb && S.len[lane] > 0) lp_last = (P.logits[(row0 + r) * P.vocab + S.last[lane]] - S.rmax[r]) - S.rlse[r];
    if (lane < BEAM_MAX) { S.dead[lane] = 0; S.ptot[lane] = lane < nb ? lae(S.pb[lane], S.pnb[lane]) : -INFINITY; }
    __syncwarp();
    // ---- stay candidates (lane j), with the extensions that spell the same prefix merged in
    if (lane < nb) {
      const int j = lane;
      float pbn = S.ptot[j] + S.lp_blank[r];
      float pnbn = S.len[j] > 0 ? S.pnb[j] + lp_last : -INFINITY;
      if (S.len[j] > 0) {
        int kk = -1;
        for (int k = 0; k < K; ++k)
          if (ct[k] == S.last[j]) kk = k;
        if (kk >= 0) {
          for (int i = 0; i < nb; ++i) {
            if (S.len[i] + 1 == S.len[j] && S.len[i] < P.max_len && hext(S.hash[i], S.last[j]) == S.hash[j]) {
              const float val = ((S.len[i] > 0 && S.last[i] == S.last[j]) ? S.pb[i] : S.ptot[i]) + cl[kk];
              pnbn = lae(pnbn, val);
              atomicOr(&S.dead[i], 1 << kk);
            }
          }
        }
      }
      S.npb[j] = pbn; S.npnb[j] = pnbn; S.ntot[j] = lae(pbn, pnbn);
    }
    __syncwarp();
    // ---- scores of all candidates; lane owns enumeration indices q = lane + 32 m
    constexpr int QM = (BEAM_MAX + BEAM_MAX * BEAM_CAND_MAX + 31) / 32;
    float sc[QM];
    const int nq = nb + nb * K;
#pragma unroll
    for (int m = 0; m < QM; ++m) {
      const int q = lane + 32 * m;
      float s = -INFINITY;
      if (q < nb) s = S.ntot[q];
      else if (q < nq) {
        const int i = (q - nb) / K, k = (q - nb) - i * K;
        if (!((S.dead[i] >> k) & 1) && S.len[i] < P.max_len)
          s = ((S.len[i] > 0 && S.last[i] == ct[k]) ? S.pb[i] : S.ptot[i]) + cl[k];
      }
      sc[m] = s;
    }
    // ---- select the `beam` best (value desc, enumeration index asc)
    int n_new = 0;
    for (int t = 0; t < B; ++t) {
      float best = -INFINITY; int bq = 0x7fffffff;
#pragma unroll
      for (int m = 0; m < QM; ++m)
        if (sc[m] > best) { best = sc[m]; bq = lane + 32 * m; }
      const uint32_t mx = __reduce_max_sync(0xffffffffu, bkey(best));
      if (mx == bkey(-INFINITY)) break;
      const int wq = (int)__reduce_min_sync(0xffffffffu, (uint32_t)(bkey(best) == mx ? bq : 0x7fffffff));
#pragma unroll
      for (int m = 0; m < QM; ++m)
        if (wq == lane + 32 * m) sc[m] = -INFINITY;
      if (lane == 0) S.sel[t] = wq;
      ++n_new;
    }
    __syncwarp();
    // ---- the new beam, entry t on lane t: read the parent, then (after everyone has read) write in place
    int n_len = 0, n_last = 0, n_org = 0, n_app = 0, app = -1;
    float n_pb = 0.f, n_pnb = 0.f;
    unsigned long long n_hash = 0;
    uint4 a4[BSEG / 8];
    if (lane < n_new) {
      const int q = S.sel[lane];
      int src;
      if (q < nb) { src = q; n_pb = S.npb[q]; n_pnb = S.npnb[q]; }
      else {
        src = (q - nb) / K;
        const int k = (q - nb) - src * K;
        app = ct[k];
        n_pb = -INFINITY;
        n_pnb = ((S.len[src] > 0 && S.last[src] == app) ? S.pb[src] : S.ptot[src]) + cl[k];
      }
      n_len = S.len[src] + (app >= 0); n_last = app >= 0 ? app : S.last[src];
      n_hash = app >= 0 ? hext(S.hash[src], app) : S.hash[src];
      n_org = S.org[src]; n_app = S.app_n[src];
      const uint4* s4 = reinterpret_cast<const uint4*>(S.app[src]);
#pragma unroll
      for (int i = 0; i < BSEG / 8; ++i) a4[i] = s4[i];
    }
    __syncwarp();
    if (lane < n_new) {
      S.len[lane] = n_len; S.last[lane] = n_last; S.pb[lane] = n_pb; S.pnb[lane] = n_pnb; S.hash[lane] = n_hash; S.org[lane] = n_org;
      uint4* d4 = reinterpret_cast<uint4*>(S.app[lane]);
#pragma unroll
      for (int i = 0; i < BSEG / 8; ++i) d4[i] = a4[i];
      if (app >= 0) S.app[lane][n_app++] = (int16_t)app;
      S.app_n[lane] = n_app;
    }
    nb = n_new;
    __syncwarp();
  }
  // ---- rebuild the token strings once: new entry t = string of its origin (old buffer) + the tokens appended during the chunk
  const int16_t* src_buf = P.tokens + ((size_t)slot * 2 + cur) * BEAM_MAX * BEAM_MAX_LEN;
  int16_t* dst_buf = P.tokens + ((size_t)slot * 2 + (cur ^ 1)) * BEAM_MAX * BEAM_MAX_LEN;
  int16_t* out0 = P.out_tokens + (size_t)w * BEAM_MAX_LEN;
  for (int t = 0; t < nb; ++t) {
    const int an = S.app_n[t], L0 = S.len[t] - an;
    const uint4* s4 = reinterpret_cast<const uint4*>(src_buf + S.org[t] * BEAM_MAX_LEN);   // rows are 16-byte aligned: whole uint4 pieces
    uint4* d4 = reinterpret_cast<uint4*>(dst_buf + t * BEAM_MAX_LEN);
    uint4* o4 = reinterpret_cast<uint4*>(out0);
    for (int i = lane; i * 8 < L0; i += 32) { const uint4 x = s4[i]; d4[i] = x; if (t == 0) o4[i] = x; }
    __syncwarp();
    if (lane < an) { const int16_t x = S.app[t][lane]; dst_buf[t * BEAM_MAX_LEN + L0 + lane] = x; if (t == 0) out0[L0 + lane] = x; }
  }
  // ---- write back the state and the best hypothesis
  if (lane < BEAM_MAX) {
    const size_t o = (size_t)slot * BEAM_MAX + lane;
    P.len[o] = S.len[lane]; P.last[o] = S.last[lane]; P.pb[o] = S.pb[lane]; P.pnb[o] = S.pnb[lane]; P.hash[o] = S.hash[lane];
  }
  if (lane == 0) { P.n_beam[slot] = nb; P.cur[slot] = cur ^ 1; }
  int full = 0;                                                   // a hypothesis that can no longer be extended: the caller must know
  if (lane < nb) full = S.len[lane] >= P.max_len;
  full = __any_sync(0xffffffffu, full);
  if (lane == 0) { P.out_len[w] = S.len[0]; P.out_score[w] = lae(S.pb[0], S.pnb[0]); if (full) P.out_flags[w] |= 1; }
}

__global__ void beam_reset_kernel(BeamParams P, int slot) {
  const int i = threadIdx.x;
  if (i < BEAM_MAX) {
    const size_t o = (size_t)slot * BEAM_MAX + i;
    P.len[o] = 0; P.last[o] = -1; P.pb[o] = i == 0 ? 0.f : -INFINITY; P.pnb[o] = -INFINITY; P.hash[o] = 0x1234567ull;
  }
  if (i == 0) { P.n_beam[slot] = 1; P.cur[slot] = 0; }
}

__global__ void beam_reset_many_kernel(BeamParams P, const int* __restrict__ slots, int n) {     // one warp per listed slot
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, i = threadIdx.x & 31;
  if (w >= n) return;
  const int slot = slots[w];
  if (i < BEAM_MAX) {
    const size_t o = (size_t)slot * BEAM_MAX + i;
    P.len[o] = 0; P.last[o] = -1; P.pb[o] = i == 0 ? 0.f : -INFINITY; P.pnb[o] = -INFINITY; P.hash[o] = 0x1234567ull;
  }
  if (i == 0) { P.n_beam[slot] = 1; P.cur[slot] = 0; }
}

__global__ void beam_reset_all_kernel(BeamParams P, int n_slots) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += gridDim.x * blockDim.x) {
    for (int i = 0; i < BEAM_MAX; ++i) {
      const size_t o = (size_t)s * BEAM_MAX + i;
      P.len[o] = 0; P.last[o] = -1; P.pb[o] = i == 0 ? 0.f : -INFINITY; P.pnb[o] = -INFINITY; P.hash[o] = 0x1234567ull;
    }
    P.n_beam[s] = 1; P.cur[s] = 0;
  }
}

}  // namespace

int beam_launch(const BeamParams& P, cudaStream_t st) {
  if (P.n <= 0) return 0;
  if (P.beam < 1 || P.beam > BEAM_MAX || P.cand_k < 1 || P.cand_k > BEAM_CAND_MAX || P.seg_rows > BSEG || P.max_len > BEAM_MAX_LEN - 1) {
    set_error("beam: unsupported parameters (beam %d <= %d, cand_k %d <= %d, frames per chunk %d <= %d)", P.beam, BEAM_MAX, P.cand_k, BEAM_CAND_MAX, P.seg_rows, BSEG);
    return -1;
  }
  ASR_CUDA_OK(launch_pdl(beam_kernel, dim3((P.n + BW - 1) / BW), dim3(BW * 32), 0, st, P));
  return 0;
}

int beam_reset_many_launch(const BeamParams& P, const int* d_slots, int n, cudaStream_t st) {
  if (n <= 0) return 0;
  beam_reset_many_kernel<<<(n * 32 + 127) / 128, 128, 0, st>>>(P, d_slots, n);
  ASR_CUDA_OK(cudaGetLastError());
  return 0;
}

int beam_reset_launch(const BeamParams& P, int slot, int n_slots_all, cudaStream_t st) {
  if (slot >= 0) beam_reset_kernel<<<1, 32, 0, st>>>(P, slot);
  else beam_reset_all_kernel<<<64, 128, 0, st>>>(P, n_slots_all);
  ASR_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace asr
