// CTC prefix beam search, one warp per stream, beam state carried across chunks (north-star: "CTC greedy/prefix
// beam search as a warp-per-stream kernel"; beam = 10 in BASELINE config #4).  The reference has no CTC prefix beam
// (its final pass is the third-party flashlight lexicon decoder, recognition.py:220-300), so the semantics are those
// of oracle/ctc_beam_oracle.py (Hannun et al. 2014, Alg. 1 without LM; parity unpinned vs the reference):
//   candidates per frame  = [stay(j) for the B beam entries] + [ext(i, k) for parent i x the cand_k best non-blank tokens]
//   stay(j)               : p_b' = p_tot(j) + lp[blank];  p_nb' = p_nb(j) + lp[last(j)]  (+ merged extensions that spell j)
//   ext(i, k)             : p_nb' = (tok_k == last(i) ? p_b(i) : p_tot(i)) + lp[tok_k]
//   next beam             = the `beam` best by logaddexp(p_b', p_nb'), ties to the lower enumeration index.
// Prefix identity is a 64-bit rolling hash + length; token strings live in a per-session double buffer in HBM
// ([2][16][1024] int16) and are copied warp-parallel, 16 bytes per lane, when a beam entry is (re)built.
#include "kernels.cuh"

namespace asr {

namespace {

constexpr int BW = 4;          // warps (streams) per CTA
constexpr int BV = 32;         // vocab <= 1024

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

struct WarpState {
  int len[BEAM_MAX], last[BEAM_MAX];
  float pb[BEAM_MAX], pnb[BEAM_MAX], ptot[BEAM_MAX];
  unsigned long long hash[BEAM_MAX];
  // scratch
  float npb[BEAM_MAX], npnb[BEAM_MAX], ntot[BEAM_MAX];
  int cand_tok[BEAM_CAND_MAX];
  float cand_lp[BEAM_CAND_MAX];
  int dead[BEAM_MAX];
  int sel[BEAM_MAX];
  int o_len[BEAM_MAX], o_last[BEAM_MAX];
  float o_pb[BEAM_MAX], o_pnb[BEAM_MAX];
  unsigned long long o_hash[BEAM_MAX];
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
  const int B = P.beam, K = P.cand_k;
  int nb = P.n_beam[slot], cur = P.cur[slot];
  if (lane < BEAM_MAX) {
    const size_t o = (size_t)slot * BEAM_MAX + lane;
    S.len[lane] = P.len[o]; S.last[lane] = P.last[o]; S.pb[lane] = P.pb[o]; S.pnb[lane] = P.pnb[o]; S.hash[lane] = P.hash[o];
  }
  __syncwarp();
  int16_t* tok_base = P.tokens + (size_t)slot * 2 * BEAM_MAX * BEAM_MAX_LEN;

  for (int r = 0; r < P.seg_rows; ++r) {
    const float* row = P.logprobs + ((size_t)w * P.seg_rows + r) * P.vocab;
    float v[BV];
#pragma unroll
    for (int i = 0; i < BV; ++i) {
      const int c = lane + 32 * i;
      v[i] = (c < P.vocab && c != 0) ? row[c] : -INFINITY;       // blank (id 0) is never an extension candidate
    }
    // ---- the cand_k best non-blank tokens (ties: lower id)
    for (int k = 0; k < K; ++k) {
      float best = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
      for (int i = 0; i < BV; ++i)
        if (v[i] > best) { best = v[i]; bi = lane + 32 * i; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
#pragma unroll
      for (int i = 0; i < BV; ++i)
        if (bi == lane + 32 * i) v[i] = -INFINITY;
      if (lane == 0) { S.cand_tok[k] = bi; S.cand_lp[k] = best; }
    }
    if (lane < BEAM_MAX) { S.dead[lane] = 0; S.ptot[lane] = lane < nb ? lae(S.pb[lane], S.pnb[lane]) : -INFINITY; }
    __syncwarp();
    // ---- stay candidates (lane j), with the extensions that spell the same prefix merged in
    if (lane < nb) {
      const int j = lane;
      const float lp_blank = row[0];
      float pbn = S.ptot[j] + lp_blank;
      float pnbn = S.len[j] > 0 ? S.pnb[j] + row[S.last[j]] : -INFINITY;
      if (S.len[j] > 0) {
        int kk = -1;
        for (int k = 0; k < K; ++k)
          if (S.cand_tok[k] == S.last[j]) kk = k;
        if (kk >= 0) {
          for (int i = 0; i < nb; ++i) {
            if (S.len[i] + 1 == S.len[j] && S.len[i] < P.max_len && hext(S.hash[i], S.last[j]) == S.hash[j]) {
              const float val = ((S.len[i] > 0 && S.last[i] == S.last[j]) ? S.pb[i] : S.ptot[i]) + S.cand_lp[kk];
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
          s = ((S.len[i] > 0 && S.last[i] == S.cand_tok[k]) ? S.pb[i] : S.ptot[i]) + S.cand_lp[k];
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
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oq = __shfl_xor_sync(0xffffffffu, bq, o);
        if (ob > best || (ob == best && oq < bq)) { best = ob; bq = oq; }
      }
      if (best == -INFINITY) break;
#pragma unroll
      for (int m = 0; m < QM; ++m)
        if (bq == lane + 32 * m) sc[m] = -INFINITY;
      if (lane == 0) S.sel[t] = bq;
      ++n_new;
    }
    __syncwarp();
    // ---- rebuild the beam into the other token buffer
    const int16_t* src_buf = tok_base + (size_t)cur * BEAM_MAX * BEAM_MAX_LEN;
    int16_t* dst_buf = tok_base + (size_t)(cur ^ 1) * BEAM_MAX * BEAM_MAX_LEN;
    for (int t = 0; t < n_new; ++t) {
      const int q = S.sel[t];
      int src, app = -1;
      float pbn, pnbn;
      if (q < nb) { src = q; pbn = S.npb[q]; pnbn = S.npnb[q]; }
      else {
        src = (q - nb) / K;
        const int k = (q - nb) - src * K;
        app = S.cand_tok[k];
        pbn = -INFINITY;
        pnbn = ((S.len[src] > 0 && S.last[src] == app) ? S.pb[src] : S.ptot[src]) + S.cand_lp[k];
      }
      const int L = S.len[src];
      {                                                            // rows are BEAM_MAX_LEN * 2 bytes apart and 16-byte aligned: whole uint4 pieces
        const uint4* s4 = reinterpret_cast<const uint4*>(src_buf + src * BEAM_MAX_LEN);
        uint4* d4 = reinterpret_cast<uint4*>(dst_buf + t * BEAM_MAX_LEN);
        for (int i = lane; i * 8 < L; i += 32) d4[i] = s4[i];
      }
      __syncwarp();
      if (lane == 0) {
        if (app >= 0) dst_buf[t * BEAM_MAX_LEN + L] = (int16_t)app;
        S.o_len[t] = L + (app >= 0); S.o_last[t] = app >= 0 ? app : S.last[src];
        S.o_pb[t] = pbn; S.o_pnb[t] = pnbn; S.o_hash[t] = app >= 0 ? hext(S.hash[src], app) : S.hash[src];
      }
    }
    __syncwarp();
    if (lane < BEAM_MAX && lane < n_new) {
      S.len[lane] = S.o_len[lane]; S.last[lane] = S.o_last[lane]; S.pb[lane] = S.o_pb[lane]; S.pnb[lane] = S.o_pnb[lane]; S.hash[lane] = S.o_hash[lane];
    }
    nb = n_new; cur ^= 1;
    __syncwarp();
  }
  // ---- write back the state and the best hypothesis
  if (lane < BEAM_MAX) {
    const size_t o = (size_t)slot * BEAM_MAX + lane;
    P.len[o] = S.len[lane]; P.last[o] = S.last[lane]; P.pb[o] = S.pb[lane]; P.pnb[o] = S.pnb[lane]; P.hash[o] = S.hash[lane];
  }
  if (lane == 0) { P.n_beam[slot] = nb; P.cur[slot] = cur; }
  const int L0 = S.len[0];
  {
    const uint4* s4 = reinterpret_cast<const uint4*>(tok_base + (size_t)cur * BEAM_MAX * BEAM_MAX_LEN);
    uint4* d4 = reinterpret_cast<uint4*>(P.out_tokens + (size_t)w * BEAM_MAX_LEN);
    for (int i = lane; i * 8 < L0; i += 32) d4[i] = s4[i];
  }
  int full = 0;                                                   // a hypothesis that can no longer be extended: the caller must know
  if (lane < nb) full = S.len[lane] >= P.max_len;
  full = __any_sync(0xffffffffu, full);
  if (lane == 0) { P.out_len[w] = L0; P.out_score[w] = lae(S.pb[0], S.pnb[0]); if (full) P.out_flags[w] |= 1; }
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
  if (P.beam < 1 || P.beam > BEAM_MAX || P.cand_k < 1 || P.cand_k > BEAM_CAND_MAX || P.vocab > 32 * BV || P.max_len > BEAM_MAX_LEN - 1) {
    set_error("beam: unsupported parameters (beam %d <= %d, cand_k %d <= %d, vocab %d)", P.beam, BEAM_MAX, P.cand_k, BEAM_CAND_MAX, P.vocab);
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
