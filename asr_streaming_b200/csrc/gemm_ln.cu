// GEMM (N = 512 = d_model) with the residual add and the following LayerNorm(s) fused into the epilogue.
//
// Replaces, per Emformer layer (TA:emformer.py:416-440):
//   out_proj  : x1 = x + attn W_o^T + b_o                      -> fp32 x1  and  bf16 LN_ff(x1)        (A operand of FFN1)
//   FFN2      : y  = LN_out(x1 + h W_2^T + b_2)                -> fp32 y   and  bf16 LN_in'(y)         (A operand of the next QKV)
//   last FFN2 : y  = LN_out(...)                               -> fp32 y   and  bf16 y of the segment rows (A operand of the CTC head)
// i.e. the separate LayerNorm passes (ln_to_operand / ln_out_fused, 41 launches and ~3.2 MB of HBM traffic per stream-chunk) and
// the fp32 x2 round trip disappear.
//
// LayerNorm needs whole rows, a CTA's TMEM holds 128 x 512 fp32 only once.  To keep the accumulators double-buffered the row
// is split over a cluster of two CTAs: both compute the same 128 rows, CTA r the columns [256 r, 256 r + 256).  The mainloop is
// the one of gemm_tc_kernel<256> (TMA producer warp, single-thread tcgen05.mma issuer, 3-stage smem ring).  The 16 epilogue
// warps work in up to three passes over their accumulator chunks (thread = row for the statistics, TMEM as the scratch):
//   R1   tcgen05.ld -> + bias + residual (residual read with full-line accesses, transposed through smem into row layout)
//        -> per-chunk (sum, M2) around the chunk mean (Chan et al. combination: two-pass accuracy) -> tcgen05.st v back
//   X    row statistics exchange: 4 warps x 2 CTAs own pieces of a row; inside the CTA through shared memory + a named barrier,
//        across the pair with st.async into the peer's shared memory signalling the peer's mbarrier (no cluster-scope fences)
//   R1b  (two-LN form only) tcgen05.ld v -> y = LN_a(v) in row layout -> statistics of y -> tcgen05.st y -> exchange
//   R2   tcgen05.ld -> 16-byte transpose through smem -> fp32 row + bf16 LN_b(row) stores, 4 rows x 128 B per instruction
#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace asr {

namespace {

using namespace tc;

constexpr int LN_N = 512;
#ifndef ASR_LN_KNOCK
// diagnostic builds only (python -m asr_streaming_b200.build --out=... -DASR_LN_KNOCK=n, tools/gemm_ln_knock.py): bit 0 = no residual loads, 1 = no
// global stores, 2 = no statistics exchange, 3 = stores go to a 4096-row (L2-resident) window, 4 = residual read from such a window
#define ASR_LN_KNOCK 0
#endif
#ifdef ASR_EPI_TIMING
__device__ unsigned long long g_ln_clk[8];                  // [0] tiles, [1] wait tfull, [2] R1, [3] exchange(s), [4] R2
#define LN_T(var) const long long var = clock64()
#else
#define LN_T(var)
#endif
constexpr int L_STATS4_BYTES = 4 * BM * 16 + 2 * BM * 16;      // float4 payload of the derived second-LN statistics (PAIR shape only)
// Three shapes of the same kernel (SHAPE):
//   0  cluster of 2: CTA r = columns [256 r, +256) of the same 128 rows; 1-CTA MMA 128 x 256; stage = A 16 KB + B 32 KB, 3 stages
//   1  cluster of 4 (PAIR): ranks (2n + m) -> column half n, row half m of a 256-row tile; the two CTAs with equal n form a
//      cta_group::2 pair (MMA 256 x 256, each CTA stages its own 128 A rows and HALF of the pair's B rows: 32 KB per stage, 64
//      instead of 96 B/clk/SM of L2 ingest at full tensor rate).  Statistics are exchanged between the CTAs with equal m (rank ^ 2).
//      148 SMs hold 34 such clusters (136 SMs).  For long K and large M (FFN2 from ~440 streams per step on).
//   2  cluster of 4 (QUAD): CTA r = columns [128 r, +128) of the same 128 rows, one 32-column chunk per epilogue warp.  For small M:
//      a 256-stream step has only 40 row tiles, i.e. 80 CTAs in shape 0 — this shape puts 160 CTAs on the 148 SMs and halves the
//      epilogue's critical path per tile.
template <int SHAPE> struct LnCfg {
  static constexpr bool PAIR = SHAPE == 1;
  static constexpr int NSPLIT = SHAPE == 2 ? 4 : 2;           // CTAs that share a row
  static constexpr int BN = LN_N / NSPLIT;                    // columns per CTA: 256 / 256 / 128
  static constexpr int CHUNKS = BN / 128;                     // 32-column chunks per epilogue warp: 2 / 2 / 1
  static constexpr int CLUSTER = PAIR ? 4 : NSPLIT;
  static constexpr int STAGES = SHAPE == 0 ? 3 : 4;
  static constexpr int STAGE_BYTES = PAIR ? (BM + BN / 2) * BK * 2 : (BM + BN) * BK * 2;       // 48 KB / 32 KB / 32 KB
  static constexpr int STATS2_BYTES = 4 * BM * 8 + 2 * (NSPLIT - 1) * BM * 8;                  // local[warp in quarter][row] + remote[set][peer][row], float2
  static constexpr int STATS_BYTES = STATS2_BYTES + (PAIR ? L_STATS4_BYTES : 0);
};
constexpr int L_EW = 16;                                   // epilogue warps: 4 per TMEM lane quarter, 2 chunks of 32 columns each
constexpr int L_THREADS = 64 + 32 * L_EW;
constexpr int L_LD = 36;                                   // fp32 row stride of the per-warp transpose tile (conflict-free both ways)
constexpr int L_XPOSE_FLOATS = 32 * L_LD;
constexpr int L_XPOSE_BYTES = L_EW * L_XPOSE_FLOATS * 4;
template <int SHAPE> constexpr int ln_smem_bytes() { return LnCfg<SHAPE>::STAGES * LnCfg<SHAPE>::STAGE_BYTES + L_XPOSE_BYTES + LnCfg<SHAPE>::STATS_BYTES + 1024 + 512; }
static_assert(ln_smem_bytes<0>() <= 232448 && ln_smem_bytes<1>() <= 232448 && ln_smem_bytes<2>() <= 232448, "gemm_ln: shared memory budget");

__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// (sum, M2 around the own mean) of 32 values held by one thread
__device__ __forceinline__ void chunk_stats(const float (&v)[32], float& sum, float& m2) {
  float s[4] = {0.f, 0.f, 0.f, 0.f};                        // four independent chains: the epilogue is latency-, not issue-bound
#pragma unroll
  for (int j = 0; j < 32; j += 4) { s[0] += v[j]; s[1] += v[j + 1]; s[2] += v[j + 2]; s[3] += v[j + 3]; }
  const float t = (s[0] + s[1]) + (s[2] + s[3]);
  const float m = t * (1.0f / 32.0f);
  float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float d = v[j + k] - m; q[k] = fmaf(d, d, q[k]); }
  }
  sum = t; m2 = (q[0] + q[1]) + (q[2] + q[3]);
}

struct RowStats { float mean, rstd; };

// Everything one epilogue warp needs for the row-statistics exchange of its TMEM lane quarter.
struct XchgCtx {
  uint32_t local_part;       // shared::cta address of local[wq][row] (this thread's slot); the quarter's 4 slots are BM*8 bytes apart
  uint32_t local_row;        // shared::cta address of local[0][row]
  uint32_t remote_set0;      // shared::cta address of remote[set 0][slot 0][row] (what the peers wrote for us); slot stride BM*8, set stride slots*BM*8
  uint32_t peer_dst0[3];     // shared::cluster address, in peer i's memory, of remote[set 0][my slot there][row]
  uint32_t bar0;             // shared::cta address of xbar[0][quarter]; set 1 is 4*8 bytes further
  uint32_t peer_bar0[3];     // shared::cluster address of peer i's xbar[0][quarter]
  uint32_t local4_row, local4_part, remote4_set0, peer_remote4_0;   // the float4 exchange (PAIR shape only, one peer)
  int named_bar;             // 1 + quarter
  int wq;
  uint32_t rank;             // 0 .. NSPLIT-1: which column slice this CTA owns (partials are summed in slice order in every CTA)
};

// All 128 threads of a lane quarter (4 warps) in all NSPLIT CTAs that share the rows call this once per round with their partial:
// sum over their N_T values and M2 around that partial's own mean.  Returns mean and 1/sqrt(var + eps) of the whole 512-wide row
// (biased variance, torch.nn.LayerNorm).  Two levels, Chan's pairwise combination at both (two-pass accuracy, fixed order =>
// bit-reproducible):
//   inside the CTA   partials through shared memory, two named-barrier syncs of the quarter's 128 threads
//   across the CTAs  warp wq == 0 pushes the CTA's combined (sum, M2) into every peer's shared memory with st.async, which signals
//                    the peer's mbarrier (complete_tx) — no cluster-scope fence, no L1 invalidation on the consumer side.
// Two mbarriers / receive slots per quarter alternate by round, so a packet of round r+2 can never be counted in round r.
template <int NSPLIT, int N_T>
__device__ __forceinline__ RowStats exchange_row_stats(float sum, float m2, const XchgCtx& X, int round) {
  constexpr float N_CTA = 4.0f * N_T;
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(X.local_part), "f"(sum), "f"(m2) : "memory");
  asm volatile("bar.sync %0, 128;" ::"r"(X.named_bar) : "memory");
  float2 p[4];
#pragma unroll
  for (int w = 0; w < 4; ++w)
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(p[w].x), "=f"(p[w].y) : "r"(X.local_row + (uint32_t)(w * BM) * 8u) : "memory");
  asm volatile("bar.sync %0, 128;" ::"r"(X.named_bar) : "memory");          // everyone has read: the slots may be rewritten next round
  const float s_cta = (p[0].x + p[1].x) + (p[2].x + p[3].x);
  const float mean_cta = s_cta * (1.0f / N_CTA);
  float q_cta = 0.f;
#pragma unroll
  for (int w = 0; w < 4; ++w) { const float d = p[w].x * (1.0f / N_T) - mean_cta; q_cta += p[w].y + (float)N_T * d * d; }
  const uint32_t set = (uint32_t)(round & 1), parity = (uint32_t)((round >> 1) & 1);
  const uint32_t bar = X.bar0 + set * 32u;
  constexpr uint32_t SET_STRIDE = (uint32_t)((NSPLIT - 1) * BM * 8);
  if (X.wq == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(8 * (NSPLIT - 1))) : "memory");
#pragma unroll
    for (int i = 0; i < NSPLIT - 1; ++i)
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
                   ::"r"(X.peer_dst0[i] + set * SET_STRIDE), "f"(s_cta), "f"(q_cta), "r"(X.peer_bar0[i] + set * 32u) : "memory");
  }
  mbar_wait(bar, parity);
  float sv[NSPLIT], qv[NSPLIT];                                   // per column slice, in slice order
#pragma unroll
  for (int r = 0; r < NSPLIT; ++r) {
    if (r == (int)X.rank) { sv[r] = s_cta; qv[r] = q_cta; }
    else {
      const int slot = r < (int)X.rank ? r : r - 1;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(sv[r]), "=f"(qv[r]) : "r"(X.remote_set0 + set * SET_STRIDE + (uint32_t)(slot * BM * 8)) : "memory");
    }
  }
  float tot = 0.f;
#pragma unroll
  for (int r = 0; r < NSPLIT; ++r) tot += sv[r];
  const float mean = tot * (1.0f / LN_N);
  float m2_all = 0.f;
#pragma unroll
  for (int r = 0; r < NSPLIT; ++r) { const float d = sv[r] * (1.0f / N_CTA) - mean; m2_all += qv[r] + N_CTA * d * d; }
  RowStats r;
  r.mean = mean;
  r.rstd = 1.0f / sqrtf(m2_all * (1.0f / LN_N) + 1e-5f);
  return r;
}

// Plain sums of three per-thread partials over the 8 owners of a row (same protocol and barriers as exchange_row_stats: it is just
// another round).  Used for the statistics of y = LN_a(v) derived from v without a second pass over the data.
__device__ __forceinline__ float3 exchange_sum3(float a, float b, float c, const XchgCtx& X, int round) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(X.local4_part), "f"(a), "f"(b), "f"(c), "f"(0.f) : "memory");
  asm volatile("bar.sync %0, 128;" ::"r"(X.named_bar) : "memory");
  float4 p[4];
#pragma unroll
  for (int w = 0; w < 4; ++w)
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(p[w].x), "=f"(p[w].y), "=f"(p[w].z), "=f"(p[w].w)
                 : "r"(X.local4_row + (uint32_t)(w * BM) * 16u) : "memory");
  asm volatile("bar.sync %0, 128;" ::"r"(X.named_bar) : "memory");
  const float a_cta = (p[0].x + p[1].x) + (p[2].x + p[3].x), b_cta = (p[0].y + p[1].y) + (p[2].y + p[3].y), c_cta = (p[0].z + p[1].z) + (p[2].z + p[3].z);
  const uint32_t set = (uint32_t)(round & 1), parity = (uint32_t)((round >> 1) & 1);
  const uint32_t bar = X.bar0 + set * 32u;
  if (X.wq == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 16;" ::"r"(bar) : "memory");
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(X.peer_remote4_0 + set * (uint32_t)(BM * 16)), "f"(a_cta), "f"(b_cta), "f"(c_cta), "f"(0.f), "r"(X.peer_bar0[0] + set * 32u) : "memory");
  }
  mbar_wait(bar, parity);
  float4 o;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w) : "r"(X.remote4_set0 + set * (uint32_t)(BM * 16)) : "memory");
  // CTA 0's sums first in both CTAs: identical results
  return X.rank == 0 ? make_float3(a_cta + o.x, b_cta + o.y, c_cta + o.z) : make_float3(o.x + a_cta, o.y + b_cta, o.z + c_cta);
}

// Per 32-value chunk of one row: besides (sum, M2) also the chunk-centred weighted moments that give the statistics of
// y = LN(v; g, b) once the row mean is known.  With w = v - m_c (m_c = chunk mean), gc = g (b - mean b):
//   p1 = sum g w,  p2 = sum (g w)^2,  p3 = sum g^2 w,  p4 = sum gc w.
struct ChunkMoments { float sum, m2, p1, p2, p3, p4; };
__device__ __forceinline__ ChunkMoments chunk_moments(const float (&v)[32], const float* __restrict__ g, const float* __restrict__ gc) {
  float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 32; j += 4) { s[0] += v[j]; s[1] += v[j + 1]; s[2] += v[j + 2]; s[3] += v[j + 3]; }
  const float t = (s[0] + s[1]) + (s[2] + s[3]);
  const float m = t * (1.0f / 32.0f);
  float q[4] = {0.f, 0.f, 0.f, 0.f}, p1[2] = {0.f, 0.f}, p2[2] = {0.f, 0.f}, p3[2] = {0.f, 0.f}, p4[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(g + j)), c4 = __ldg(reinterpret_cast<const float4*>(gc + j));
    const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = v[j + k] - m, gw = gg[k] * d;
      q[k] = fmaf(d, d, q[k]);
      p1[k & 1] += gw; p2[k & 1] = fmaf(gw, gw, p2[k & 1]); p3[k & 1] = fmaf(gg[k], gw, p3[k & 1]); p4[k & 1] = fmaf(cc[k], d, p4[k & 1]);
    }
  }
  ChunkMoments r;
  r.sum = t; r.m2 = (q[0] + q[1]) + (q[2] + q[3]); r.p1 = p1[0] + p1[1]; r.p2 = p2[0] + p2[1]; r.p3 = p3[0] + p3[1]; r.p4 = p4[0] + p4[1];
  return r;
}

template <int SHAPE>
__global__ void __launch_bounds__(L_THREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmProblem p, LnEpilogue ep) {
  using Cfg = LnCfg<SHAPE>;
  constexpr bool PAIR = Cfg::PAIR;
  constexpr int L_STAGES = Cfg::STAGES, L_STAGE_BYTES = Cfg::STAGE_BYTES, CLUSTER = Cfg::CLUSTER, LBN = Cfg::BN, NSPLIT = Cfg::NSPLIT;
  constexpr int CHUNKS = Cfg::CHUNKS, N_T = 32 * CHUNKS;
  constexpr uint32_t TMEM_COLS = 2 * LBN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t xpose_base = smem_base + L_STAGES * L_STAGE_BYTES;
  const uint32_t stats_base = xpose_base + L_XPOSE_BYTES;
  const uint32_t bar_base = stats_base + Cfg::STATS_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (L_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * L_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * L_STAGES + 2 + s); };
  auto xq_bar = [&](int set, int q) { return bar_base + 8u * (2 * L_STAGES + 4 + 4 * set + q); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * L_STAGES + 12);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t nrank = PAIR ? rank >> 1 : rank;              // column slice
  const uint32_t mrank = PAIR ? rank & 1u : 0u;                // row half inside the pair's 256-row tile
  constexpr int TILE_M = PAIR ? 2 * BM : BM;
  const int cluster_id = blockIdx.x / CLUSTER, n_clusters = gridDim.x / CLUSTER;
  const int m_tiles = (p.M + TILE_M - 1) / TILE_M;
  const int kb_per_pass = p.K / BK;
  const int total_kb = kb_per_pass * p.passes;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < L_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), PAIR ? 2 * L_EW : L_EW); }
    for (int q = 0; q < 8; ++q) mbar_init(xq_bar(q >> 2, q & 3), 32);       // the 32 lanes of the quarter's warp 0 arm 8 bytes each
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  cluster_sync_all();                                       // the peer's barriers exist before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int m_blk = cluster_id; m_blk < m_tiles; m_blk += n_clusters) {
        const int row_a = m_blk * TILE_M + (int)mrank * BM;
        const int row_b = (int)nrank * LBN + (PAIR ? (int)mrank * (LBN / 2) : 0);
        for (int ps = 0; ps < p.passes; ++ps) {
          for (int kb = 0; kb < kb_per_pass; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = smem_base + stage * L_STAGE_BYTES;
            const uint32_t sb = sa + BM * BK * 2;
            if (PAIR) {
              const uint32_t lbar = full_bar(stage) & kPeerBitMask;            // the pair leader's barrier collects both CTAs' bytes
              if (mrank == 0) mbar_expect_tx(full_bar(stage), 2u * (uint32_t)L_STAGE_BYTES);
              tma_load_2d_pair(sa, &tmA, p.a_koff[ps] + kb * BK, row_a, lbar);
              tma_load_2d_pair(sb, &tmB, p.b_koff[ps] + kb * BK, row_b, lbar);
            } else {
              mbar_expect_tx(full_bar(stage), (uint32_t)L_STAGE_BYTES);
              tma_load_2d(sa, &tmA, p.a_koff[ps] + kb * BK, row_a, full_bar(stage));
              tma_load_2d(sb, &tmB, p.b_koff[ps] + kb * BK, row_b, full_bar(stage));
            }
            if (++stage == L_STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (PAIR: the leader CTA of each pair only) =====================
    constexpr uint32_t idesc = PAIR ? ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(LBN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24)) : make_idesc(LBN);
    const uint16_t pair_mask = (uint16_t)(3u << (2 * nrank));
    if (!PAIR || mrank == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
#ifdef ASR_MMA_TIMING
      long long w_empty = 0, w_full = 0, n_t = 0;
      const long long t_begin = clock64();
#endif
      for (int m_blk = cluster_id; m_blk < m_tiles; m_blk += n_clusters) {
        const int n_kb = total_kb;
#ifdef ASR_MMA_TIMING
        const long long ta = clock64();
#endif
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
#ifdef ASR_MMA_TIMING
        w_empty += clock64() - ta; ++n_t;
#endif
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * LBN);
        for (int kb = 0; kb < n_kb; ++kb) {
#ifdef ASR_MMA_TIMING
          const long long tb = clock64();
#endif
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
#ifdef ASR_MMA_TIMING
          w_full += clock64() - tb;
#endif
          if (lane == 0) {
            const uint32_t sa = smem_base + stage * L_STAGE_BYTES;
            const uint64_t adesc = make_smem_desc(sa);
            const uint64_t bdesc = make_smem_desc(sa + BM * BK * 2);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if (PAIR) umma_bf16_pair(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            }
            if (PAIR) {
              umma_commit_pair(empty_bar(stage), pair_mask);
              if (kb == n_kb - 1) umma_commit_pair(tfull_bar(acc), pair_mask);
            } else {
              umma_commit(empty_bar(stage));
              if (kb == n_kb - 1) umma_commit(tfull_bar(acc));
            }
          }
          __syncwarp();
          if (++stage == L_STAGES) { stage = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
#ifdef ASR_MMA_TIMING
      if (blockIdx.x == 0 && lane == 0)
        printf("gemm_ln<%d> MMA thread of CTA 0 (K %d): %lld tiles, per tile %lld clk total, waiting for a free accumulator %lld, for operand stages %lld\n", SHAPE, p.K, n_t,
               (clock64() - t_begin) / (n_t ? n_t : 1), w_empty / (n_t ? n_t : 1), w_full / (n_t ? n_t : 1));
#endif
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;                            // TMEM lane quarter this warp may access
    const int wq = ew >> 2;                                  // which of the 4 warps sharing the quarter: chunks wq and wq + 4
    const int rsub = lane >> 3, csub = 4 * (lane & 7);
    float* xpose = reinterpret_cast<float*>(smem_raw + (xpose_base - smem_u32(smem_raw))) + ew * L_XPOSE_FLOATS;
    XchgCtx X;
    {
      const uint32_t row_off = (uint32_t)(quarter * 32 + lane) * 8u;
      X.local_row = stats_base + row_off;
      X.local_part = X.local_row + (uint32_t)(wq * BM) * 8u;
      X.remote_set0 = stats_base + 4u * BM * 8u + row_off;
      X.bar0 = xq_bar(0, quarter);
#pragma unroll
      for (int i = 0; i < 3; ++i) { X.peer_dst0[i] = 0; X.peer_bar0[i] = 0; }
#pragma unroll
      for (int i = 0; i < NSPLIT - 1; ++i) {
        const uint32_t pn = (uint32_t)i < nrank ? (uint32_t)i : (uint32_t)i + 1u;       // the peer's column slice
        const uint32_t pr = PAIR ? (pn << 1) | mrank : pn;                              // its cluster rank
        const uint32_t my_slot = nrank < pn ? nrank : nrank - 1u;                       // where my partial lives in its receive area
        X.peer_dst0[i] = map_to_cta(X.remote_set0 + my_slot * (uint32_t)(BM * 8), pr);
        X.peer_bar0[i] = map_to_cta(X.bar0, pr);
      }
      if (PAIR) {
        const uint32_t base4 = stats_base + (uint32_t)Cfg::STATS2_BYTES, row4 = (uint32_t)(quarter * 32 + lane) * 16u;
        X.local4_row = base4 + row4;
        X.local4_part = X.local4_row + (uint32_t)(wq * BM) * 16u;
        X.remote4_set0 = base4 + 4u * BM * 16u + row4;
        X.peer_remote4_0 = map_to_cta(X.remote4_set0, rank ^ 2u);
      } else {
        X.local4_row = X.local4_part = X.remote4_set0 = X.peer_remote4_0 = 0;
      }
      X.named_bar = 1 + quarter; X.wq = wq; X.rank = nrank;
    }
    const int col_cta = (int)nrank * LBN;
    int acc = 0; uint32_t acc_phase = 0;
    int round = 0;
    for (int m_blk = cluster_id; m_blk < m_tiles; m_blk += n_clusters) {
      const int row0 = m_blk * TILE_M + (int)mrank * BM + quarter * 32;
      const int my_row = row0 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * LBN);
      if (my_row < p.M) {                                    // residual lines of my row -> L2 while the MMAs of the tile run
        const float* r = ep.res + (size_t)my_row * LN_N + col_cta;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r + wq * 32));
        if (CHUNKS == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(r + (wq + 4) * 32));
      }
      LN_T(lt0);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      LN_T(lt1);

      float raw[32];
      // PAIR shape with two LayerNorms: the statistics of y = LN_a(v) are derived from chunk-centred weighted moments of v
      // gathered in R1 (ep.y_consts), so the pass that materialised y in TMEM (R1b) disappears.
      const bool fuse2 = PAIR && ep.g2 != nullptr && ep.y_consts != nullptr;
      float mom[2][4];                                       // [chunk][p1..p4]
      // ---------------- R1: v = acc + bias + residual, statistics of v
      float s_a = 0.f, q_a = 0.f, s_b = 0.f, q_b = 0.f;
#pragma unroll 1
      for (int h = 0; h < CHUNKS; ++h) {
        const int cc = wq + 4 * h;
        const int col0 = col_cta + cc * 32;
        tmem_ld32(taddr + (uint32_t)(cc * 32), raw);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = row0 + 4 * i + rsub;
          const float4 r4 = (row < p.M && !(ASR_LN_KNOCK & 1)) ? *reinterpret_cast<const float4*>(ep.res + (size_t)((ASR_LN_KNOCK & 16) ? (row & 4095) : row) * LN_N + col0 + csub)
                                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(xpose + (4 * i + rsub) * L_LD + csub) = r4;
        }
        __syncwarp();
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 rr = *reinterpret_cast<const float4*>(xpose + lane * L_LD + j);
          const float4 bb = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
          raw[j] += rr.x + bb.x; raw[j + 1] += rr.y + bb.y; raw[j + 2] += rr.z + bb.z; raw[j + 3] += rr.w + bb.w;
        }
        __syncwarp();
        if (fuse2) {
          const ChunkMoments cm = chunk_moments(raw, ep.g1 + col0, ep.y_consts + col0);
          if (h == 0) { s_a = cm.sum; q_a = cm.m2; } else { s_b = cm.sum; q_b = cm.m2; }
          mom[h][0] = cm.p1; mom[h][1] = cm.p2; mom[h][2] = cm.p3; mom[h][3] = cm.p4;
        } else {
          if (h == 0) chunk_stats(raw, s_a, q_a); else chunk_stats(raw, s_b, q_b);
        }
        tmem_st32(taddr + (uint32_t)(cc * 32), raw);
      }
      tmem_st_wait();
      const float mc_a = s_a * (1.0f / 32.0f), mc_b = s_b * (1.0f / 32.0f);      // chunk means (before the sums are merged)
      if (CHUNKS == 2) {
        const float d = mc_a - mc_b;
        q_a = q_a + q_b + 16.0f * d * d;                     // Chan: n_a n_b / (n_a + n_b) = 16
        s_a += s_b;
      }
      LN_T(lt2);
#if ASR_LN_KNOCK & 4
      RowStats st; st.mean = s_a * (1.0f / N_T); st.rstd = 1.0f / sqrtf(q_a * (1.0f / N_T) + 1e-5f);
#else
      RowStats st = exchange_row_stats<NSPLIT, N_T>(s_a, q_a, X, round++);
#endif
      RowStats st2 = st;                                     // statistics of y (two-LN forms)

      const float* g_fin = ep.g1;
      const float* b_fin = ep.b1;
      if (fuse2) {
        // with u = v - mean = w + delta (delta = chunk mean - row mean) and the per-chunk constants G1 = sum g, G2 = sum g^2, G3 = sum gc:
        //   A1 = sum g u = p1 + delta G1;   A2 = sum (g u)^2 = p2 + 2 delta p3 + delta^2 G2;   A3 = sum gc u = p4 + delta G3
        const float* yc = ep.y_consts + LN_N;
        float a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int h = 0; h < CHUNKS; ++h) {
          const int chunk = (int)nrank * (LBN / 32) + wq + 4 * h;     // 32-column chunk index within the 512-wide row
          const float dl = (h == 0 ? mc_a : mc_b) - st.mean;
          a1 += mom[h][0] + dl * __ldg(yc + chunk);
          a2 += mom[h][1] + 2.0f * dl * mom[h][2] + dl * dl * __ldg(yc + 16 + chunk);
          a3 += mom[h][3] + dl * __ldg(yc + 32 + chunk);
        }
        const float3 A = exchange_sum3(a1, a2, a3, X, round++);
        // y_j = r g_j u_j + b_j:  mean_y = r A1/N + mean(b);  var_y = r^2 (A2/N - (A1/N)^2) + 2 r A3/N + var(b)
        const float mg = A.x * (1.0f / LN_N);
        const float var_y = st.rstd * st.rstd * (A.y * (1.0f / LN_N) - mg * mg) + 2.0f * st.rstd * A.z * (1.0f / LN_N) + __ldg(yc + 49);
        st2.mean = st.rstd * mg + __ldg(yc + 48);
        st2.rstd = 1.0f / sqrtf(fmaxf(var_y, 0.f) + 1e-5f);
      } else if (ep.g2) {
        // ---------------- R1b: y = LN_a(v) in row layout, statistics of y
#pragma unroll 1
        for (int h = 0; h < CHUNKS; ++h) {
          const int cc = wq + 4 * h;
          const int col0 = col_cta + cc * 32;
          tmem_ld32(taddr + (uint32_t)(cc * 32), raw);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 gg = __ldg(reinterpret_cast<const float4*>(ep.g1 + col0 + j));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(ep.b1 + col0 + j));
            raw[j] = (raw[j] - st.mean) * st.rstd * gg.x + bb.x;
            raw[j + 1] = (raw[j + 1] - st.mean) * st.rstd * gg.y + bb.y;
            raw[j + 2] = (raw[j + 2] - st.mean) * st.rstd * gg.z + bb.z;
            raw[j + 3] = (raw[j + 3] - st.mean) * st.rstd * gg.w + bb.w;
          }
          if (h == 0) chunk_stats(raw, s_a, q_a); else chunk_stats(raw, s_b, q_b);
          tmem_st32(taddr + (uint32_t)(cc * 32), raw);
        }
        tmem_st_wait();
        if (CHUNKS == 2) {
          const float d = (s_a - s_b) * (1.0f / 32.0f);
          q_a = q_a + q_b + 16.0f * d * d;
          s_a += s_b;
        }
        st = exchange_row_stats<NSPLIT, N_T>(s_a, q_a, X, round++);
        g_fin = ep.g2; b_fin = ep.b2;
      }

      // ---------------- R2: fp32 row + bf16 LayerNorm(row) with full-line stores
      LN_T(lt3);
      float mu[8], rs[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        mu[i] = __shfl_sync(0xffffffffu, st.mean, 4 * i + rsub);
        rs[i] = __shfl_sync(0xffffffffu, st.rstd, 4 * i + rsub);
      }
      if (fuse2) { g_fin = ep.g2; b_fin = ep.b2; }
#pragma unroll 1
      for (int h = 0; h < CHUNKS; ++h) {
        const int cc = wq + 4 * h;
        const int col = col_cta + cc * 32 + csub;
        tmem_ld32(taddr + (uint32_t)(cc * 32), raw);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(xpose + lane * L_LD + j) = make_float4(raw[j], raw[j + 1], raw[j + 2], raw[j + 3]);
        __syncwarp();
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g_fin + col));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(b_fin + col));
        float4 g1v = gg, b1v = bb;
        if (fuse2) { g1v = __ldg(reinterpret_cast<const float4*>(ep.g1 + col)); b1v = __ldg(reinterpret_cast<const float4*>(ep.b1 + col)); }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = row0 + 4 * i + rsub;
          float4 t = *reinterpret_cast<const float4*>(xpose + (4 * i + rsub) * L_LD + csub);
          float n[4];
          if (fuse2) {                                         // t = v: first LN here (its statistics are mu / rs), second with st2
            t.x = (t.x - mu[i]) * rs[i] * g1v.x + b1v.x; t.y = (t.y - mu[i]) * rs[i] * g1v.y + b1v.y;
            t.z = (t.z - mu[i]) * rs[i] * g1v.z + b1v.z; t.w = (t.w - mu[i]) * rs[i] * g1v.w + b1v.w;
            const float m2 = __shfl_sync(0xffffffffu, st2.mean, 4 * i + rsub), r2 = __shfl_sync(0xffffffffu, st2.rstd, 4 * i + rsub);
            n[0] = (t.x - m2) * r2 * gg.x + bb.x; n[1] = (t.y - m2) * r2 * gg.y + bb.y;
            n[2] = (t.z - m2) * r2 * gg.z + bb.z; n[3] = (t.w - m2) * r2 * gg.w + bb.w;
          } else {
            n[0] = (t.x - mu[i]) * rs[i] * gg.x + bb.x;
            n[1] = (t.y - mu[i]) * rs[i] * gg.y + bb.y;
            n[2] = (t.z - mu[i]) * rs[i] * gg.z + bb.z;
            n[3] = (t.w - mu[i]) * rs[i] * gg.w + bb.w;
          }
          if (row < p.M && !(ASR_LN_KNOCK & 2)) {
            *reinterpret_cast<float4*>(ep.out_f32 + (size_t)((ASR_LN_KNOCK & 8) ? (row & 4095) : row) * LN_N + col) = ep.f32_normed ? make_float4(n[0], n[1], n[2], n[3]) : t;
            int orow = (ASR_LN_KNOCK & 8) ? (row & 4095) : row;
            if (ep.compact_rows) {                           // last layer: only the segment rows feed the CTC head (TA:emformer.py:803)
              const int b = row / ep.compact_rows, tt = row - b * ep.compact_rows;
              orow = tt < ep.compact_seg ? b * ep.compact_seg + tt : -1;
            }
            if (orow >= 0) {
              bf16* o = ep.out_op + (size_t)orow * ep.op_ld + col;
              const __nv_bfloat162 h01 = __floats2bfloat162_rn(n[0], n[1]), h23 = __floats2bfloat162_rn(n[2], n[3]);
              *reinterpret_cast<uint2*>(o) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
              if (ep.op_lo_off) {
                const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
                *reinterpret_cast<uint2*>(o + ep.op_lo_off) = make_uint2(pack_bf16x2(n[0] - f01.x, n[1] - f01.y), pack_bf16x2(n[2] - f23.x, n[3] - f23.y));
              }
            }
          }
        }
        __syncwarp();
      }
#ifdef ASR_EPI_TIMING
      if (blockIdx.x == 0 && threadIdx.x == 64) {
        const long long lt4 = clock64();
        atomicAdd(&g_ln_clk[0], 1ull); atomicAdd(&g_ln_clk[1], (unsigned long long)(lt1 - lt0)); atomicAdd(&g_ln_clk[2], (unsigned long long)(lt2 - lt1));
        atomicAdd(&g_ln_clk[3], (unsigned long long)(lt3 - lt2)); atomicAdd(&g_ln_clk[4], (unsigned long long)(lt4 - lt3));
      }
#endif
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (!PAIR || mrank == 0) mbar_arrive(tempty_bar(acc));
        else mbar_arrive_cta(tempty_bar(acc), rank & ~1u);       // the pair leader's MMA warp waits for both CTAs' epilogues
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

#ifdef ASR_EPI_TIMING
  if (blockIdx.x == 0 && threadIdx.x == 64 && g_ln_clk[0]) {
    printf("gemm_ln<%d> K %d epilogue warp 0 of CTA 0: %llu tiles, per tile: wait tfull %llu clk, R1 %llu, exchange %llu, R2 %llu\n", SHAPE, p.K, g_ln_clk[0],
           g_ln_clk[1] / g_ln_clk[0], g_ln_clk[2] / g_ln_clk[0], g_ln_clk[3] / g_ln_clk[0], g_ln_clk[4] / g_ln_clk[0]);
    g_ln_clk[0] = g_ln_clk[1] = g_ln_clk[2] = g_ln_clk[3] = g_ln_clk[4] = 0;
  }
#endif
  tc_fence_before();
  cluster_sync_all();                                       // the peer may still read this CTA's statistics through DSMEM
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace

namespace {
template <int SHAPE>
int ln_max_clusters(int num_sms, int* out) {
  constexpr int CL = LnCfg<SHAPE>::CLUSTER;
  static int max_clusters_dev[kMaxDevices] = {0};
  int& max_clusters = max_clusters_dev[current_device_index()];
  if (!max_clusters) {
    ASR_CUDA_OK(cudaFuncSetAttribute(gemm_ln_kernel<SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, ln_smem_bytes<SHAPE>()));
    // how many clusters of this shape the device can hold at once (GPC boundaries: 148 SMs hold 74 pairs but only ~34 quads)
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(num_sms / CL * CL); q.blockDim = dim3(L_THREADS); q.dynamicSmemBytes = ln_smem_bytes<SHAPE>();
    cudaLaunchAttribute a[1];
    a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = CL; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    q.attrs = a; q.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_ln_kernel<SHAPE>, &q) != cudaSuccess || n <= 0) { cudaGetLastError(); n = num_sms / CL; }
    max_clusters = n < num_sms / CL ? n : num_sms / CL;
  }
  *out = max_clusters;
  return 0;
}

template <int SHAPE>
int launch_ln(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmProblem& p, const LnEpilogue& ep, int num_sms, cudaStream_t st) {
  constexpr int CL = LnCfg<SHAPE>::CLUSTER, TILE_M = LnCfg<SHAPE>::PAIR ? 2 * BM : BM;
  int max_clusters = 0;
  if (ln_max_clusters<SHAPE>(num_sms, &max_clusters)) return -1;
  const int m_tiles = (p.M + TILE_M - 1) / TILE_M;
  const int clusters = m_tiles < max_clusters ? m_tiles : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL * clusters); cfg.blockDim = dim3(L_THREADS); cfg.dynamicSmemBytes = ln_smem_bytes<SHAPE>(); cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  ASR_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_ln_kernel<SHAPE>, tmA, tmB, p, ep));
  return 0;
}
}  // namespace

// tmB256 / tmB128: tensor maps of the [512, ld] weight with 256- and 128-row boxes.  shape: see LnCfg (0 pair of column halves,
// 1 cta_group::2 quads, 2 four column quarters).
int gemm_ln(const CUtensorMap& tmA, const CUtensorMap& tmB256, const CUtensorMap& tmB128, const GemmProblem& p, const LnEpilogue& ep, int shape,
            int num_sms, cudaStream_t st) {
  if (p.M <= 0) return 0;
  if (p.N != LN_N) { set_error("gemm_ln: N = %d, the fused LayerNorm epilogue is built for d_model = %d", p.N, LN_N); return -1; }
  if (p.K % BK != 0) { set_error("gemm_ln: K=%d not a multiple of %d", p.K, BK); return -1; }
  if (shape == 1) return launch_ln<1>(tmA, tmB128, p, ep, num_sms, st);
  if (shape == 2) return launch_ln<2>(tmA, tmB128, p, ep, num_sms, st);
  return launch_ln<0>(tmA, tmB256, p, ep, num_sms, st);
}

}  // namespace asr
