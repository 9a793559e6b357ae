// tcgen05 / TMEM / TMA GEMM for the encoder's dense projections (Q/KV proj, out_proj,
// FFN1, FFN2, input_linear, CTC head).  Replaces the ATen GEMMs behind torch.nn.Linear in
// TA:emformer.py:161,:164,:207,:366-373, lightspeech/modules/encoder.py:142, decoder.py:66-70.
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0      TMA producer: cp.async.bulk.tensor 2D loads of A[128 x 64] and B[BN x 64] bf16 tiles
//               (128B swizzle) into a kStages-deep shared-memory ring, signalled by mbarrier tx-count.
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (cta_group::1, kind::f16, M=128, N=BN,
//               K=16 per instruction), fp32 accumulators in TMEM, double-buffered (2 x BN columns);
//               tcgen05.commit releases smem stages and publishes finished accumulators.
//   warps 2..9  epilogue, two warps per TMEM lane quarter alternating 32-column blocks: tcgen05.ld 32x32b.x32 (one
//               accumulator row per thread, one chunk ahead), 16-byte transpose through shared memory, then fused bias /
//               residual / GELU / SiLU / Q-scaling / KV-ring scatter with full-line (4 rows x 128 B) global accesses.
#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace asr {

namespace {

using namespace tc;
// 16 epilogue warps (4 per TMEM lane quarter): with 8 the epilogue is latency-bound (tcgen05.ld -> smem transpose -> global),
// measured 242 -> 173 us on the FFN1 shape (profiles/r01_gemm_sweep_*.txt).  96 registers per thread at 576 threads.
constexpr int kEpiWarps = 16;
constexpr int kThreads = 64 + 32 * kEpiWarps;
template <class Epi> struct EpiWarps { static constexpr int value = kEpiWarps; };

constexpr int kStageLd = 36;                          // fp32 row stride of the per-warp transpose tile: STS.128 by row and LDS.128 by
constexpr int kXposeFloats = 32 * kStageLd;           // (4 rows x 8 lanes) are both bank-conflict-free
constexpr int kStagingBytes = kEpiWarps * kXposeFloats * 4;   // pair kernel

// One accumulator tile (128 rows x BN columns of this CTA) from TMEM to global memory.
//   * before the accumulator is complete (overlapping the MMAs): RowCtx of the thread's row, L2 prefetch of its residual rows;
//   * tcgen05.ld (thread = row) runs one 32-column chunk ahead of the math (two register sets);
//   * each chunk is transposed through shared memory so that global accesses are full 128-byte lines (see common.cuh).
#ifdef ASR_EPI_TIMING
__device__ unsigned long long g_epi_clk[8];                 // [0] tiles, [1] wait tfull, [2] TMEM loads, [3] math + staging + stores
#define EPI_T(var) const long long var = clock64()
#else
#define EPI_T(var)
#endif
template <int BN, int kStride, class Epi, bool kBf16Only = false>
__device__ __forceinline__ void epilogue_tile(const Epi& epi, const GemmProblem& p, int row0, int tile_col0, int half, int lane,
                                              uint32_t taddr, uint32_t tfull, uint32_t tfull_phase, float* xpose) {
  constexpr int kChunks = BN / 32;
  if constexpr (IsNullEpi<Epi>::value) { mbar_wait(tfull, tfull_phase); tc_fence_after(); return; }
  const int my_row = row0 + lane;
  const typename Epi::RowCtx my_ctx = epi.row_ctx(my_row < p.M ? my_row : p.M - 1, tile_col0);
  if (my_row < p.M) epi.prefetch_tile(my_row, tile_col0, (p.N - tile_col0) < BN ? (p.N - tile_col0) : BN, my_ctx);
  typename Epi::RowCtx ctx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) ctx[i] = shfl_ctx<typename Epi::RowCtx>(my_ctx, 4 * i + (lane >> 3));
  const float* bias = epi.bias_ptr();
  if constexpr (Epi::kBf16Rows) {
    if (kBf16Only || (epi.bf16_rows() && tile_col0 + BN <= p.N)) {
      // bf16 outputs: bias / activation in the thread = row layout, convert, transpose 16-byte pieces (XOR-swizzled, conflict-free
      // both ways) and store 8 rows x 64 B per instruction — a third of the instructions of the fp32 transpose path below, which is
      // what the K = 512 GEMMs (8 k-blocks of MMA per 128 x 256 accumulator) are bound by.
      typename Epi::RowCtx c4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) c4[i] = shfl_ctx<typename Epi::RowCtx>(my_ctx, 8 * i + (lane >> 2));
      const int sec_col0 = epi.section_col0(tile_col0);
      uint32_t* xw = reinterpret_cast<uint32_t*>(xpose);
      EPI_T(t0);
      mbar_wait(tfull, tfull_phase);
      tc_fence_after();
      auto finish16 = [&](float (&v)[32], int c) {
        const int col0 = tile_col0 + c * 32;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          {
            const float* bp = epi.bias_ptr() + col0 + 8 * q;
            epi.apply8(v + 8 * q, __ldg(reinterpret_cast<const float4*>(bp)), __ldg(reinterpret_cast<const float4*>(bp + 4)), tile_col0);
          }
          uint4 o;
          o.x = pack_bf16x2(v[8 * q], v[8 * q + 1]); o.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
          o.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]); o.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
          *reinterpret_cast<uint4*>(xw + lane * 16 + 4 * (q ^ ((lane >> 1) & 3))) = o;
        }
        __syncwarp();
#ifdef ASR_EPI_TIMING
        const long long ta = clock64();
#endif
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = 8 * i + (lane >> 2), pc = lane & 3;
          const uint4 o = *reinterpret_cast<const uint4*>(xw + r * 16 + 4 * (pc ^ ((r >> 1) & 3)));
          if (row0 + r < p.M) *reinterpret_cast<uint4*>(epi.row_ptr(c4[i], row0 + r, col0 + 8 * pc, sec_col0)) = o;
        }
        __syncwarp();
#ifdef ASR_EPI_TIMING
        if (blockIdx.x == 0 && threadIdx.x == 64) atomicAdd(&g_epi_clk[4], (unsigned long long)(clock64() - ta));
#endif
      };
      float v[32];
      if constexpr (kChunks == 2 * kStride) {
        // two chunks per warp: both TMEM loads in flight before the first chunk's math
        float w[32];
        EPI_T(t1);
        tmem_ld32(taddr + (uint32_t)(half * 32), v);
        tmem_ld32(taddr + (uint32_t)((half + kStride) * 32), w);
        tmem_ld_wait();
        EPI_T(t2);
        finish16(v, half);
        finish16(w, half + kStride);
#ifdef ASR_EPI_TIMING
        if (blockIdx.x == 0 && threadIdx.x == 64) {
          const long long t3 = clock64();
          atomicAdd(&g_epi_clk[0], 1ull); atomicAdd(&g_epi_clk[1], (unsigned long long)(t1 - t0));
          atomicAdd(&g_epi_clk[2], (unsigned long long)(t2 - t1)); atomicAdd(&g_epi_clk[3], (unsigned long long)(t3 - t2));
        }
#endif
      } else {
#pragma unroll 1
        for (int c = half; c < kChunks; c += kStride) {
          tmem_ld32(taddr + (uint32_t)(c * 32), v);
          tmem_ld_wait();
          finish16(v, c);
        }
      }
      return;
    }
  }
  if constexpr (kBf16Only) return;                       // (unreachable: the launcher checks the epilogue type and N)
  mbar_wait(tfull, tfull_phase);
  tc_fence_after();
  if (half >= kChunks) return;
  float va[32];
  auto finish = [&](float (&raw)[32], int c) {
    const int col = tile_col0 + c * 32 + 4 * (lane & 7);
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias && col < p.N) b4 = *reinterpret_cast<const float4*>(bias + col);      // N is a multiple of 4 for every matrix here
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(xpose + lane * kStageLd + j) = make_float4(raw[j], raw[j + 1], raw[j + 2], raw[j + 3]);
    __syncwarp();
    float v[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(xpose + (4 * i + (lane >> 3)) * kStageLd + 4 * (lane & 7));
      v[i][0] = t.x; v[i][1] = t.y; v[i][2] = t.z; v[i][3] = t.w;
    }
    __syncwarp();
    if (row0 < p.M && col < p.N) epi.store(row0, col, lane, p.M, v, ctx, b4);
  };
  if constexpr (kStride >= 4) {          // 16 epilogue warps: thread-level parallelism hides the tcgen05.ld latency, keep registers low
#pragma unroll 1
    for (int c = half; c < kChunks; c += kStride) {
      tmem_ld32(taddr + (uint32_t)(c * 32), va);
      tmem_ld_wait();
      finish(va, c);
    }
    return;
  }
  float vb[32];
  tmem_ld32(taddr + (uint32_t)(half * 32), va);
  tmem_ld_wait();
#pragma unroll 1
  for (int i = 0; half + kStride * i < kChunks; i += 2) {
    const int c0 = half + kStride * i, c1 = c0 + kStride;
    if (c1 < kChunks) tmem_ld32(taddr + (uint32_t)(c1 * 32), vb);
    finish(va, c0);
    tmem_ld_wait();
    if (c1 >= kChunks) break;
    const int c2 = c1 + kStride;
    if (c2 < kChunks) tmem_ld32(taddr + (uint32_t)(c2 * 32), va);
    finish(vb, c1);
    tmem_ld_wait();
  }
}

template <int BN, int EW> struct TileCfg {
  static constexpr int kStageBytes = (BM + BN) * BK * 2;
  static constexpr int kXposeBytes = EW * kXposeFloats * 4;
  static constexpr int kStages = (BN == 256) ? 3 : ((BN == 128) ? 4 : 6);
  static constexpr int kTmemCols = 2 * BN;   // power of two for BN in {64,128,256}
  static constexpr int kSmemBytes = kStages * kStageBytes + kXposeBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kThreadsCta = 64 + 32 * EW;
};

template <int BN, class Epi>
__global__ void __launch_bounds__(64 + 32 * EpiWarps<Epi>::value, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmProblem p, Epi epi) {
  constexpr int EW = EpiWarps<Epi>::value;
  using Cfg = TileCfg<BN, EW>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzle-128B tiles need 1024B alignment
  const uint32_t staging_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  const uint32_t bar_base = staging_base + Cfg::kXposeBytes;
  // barrier layout (8 bytes each): full[kStages] | empty[kStages] | tmem_full[2] | tmem_empty[2] | tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * Cfg::kStages + 4);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BM - 1) / BM, n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int kb_per_pass = p.K / BK;
  const int total_kb = kb_per_pass * p.passes;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), EW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"((uint32_t)Cfg::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  pdl_launch_dependents();
  pdl_wait();                                              // A operand / residual / slot tables come from predecessor kernels

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
        for (int ps = 0; ps < p.passes; ++ps) {
          for (int kb = 0; kb < kb_per_pass; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
            const uint32_t sb = sa + BM * BK * 2;
            mbar_expect_tx(full_bar(stage), (uint32_t)Cfg::kStageBytes);
            tma_load_2d(sa, &tmA, p.a_koff[ps] + kb * BK, m_blk * BM, full_bar(stage));
            tma_load_2d(sb, &tmB, p.b_koff[ps] + kb * BK, n_blk * BN, full_bar(stage));
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(BN);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint64_t adesc = make_smem_desc(sa);
          const uint64_t bdesc = make_smem_desc(sa + BM * BK * 2);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128B swizzle atom: +2 in the (addr >> 4) field
            umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));                    // smem stage reusable once these MMAs retire
          if (kb == total_kb - 1) umma_commit(tfull_bar(acc));   // accumulator complete
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;                           // TMEM lane quarter this warp may read
    const int half = ew >> 2;                               // which of the EW/4 warps sharing the quarter
    float* xpose = reinterpret_cast<float*>(smem_raw + (staging_base - smem_u32(smem_raw))) + ew * kXposeFloats;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
      const int row0 = m_blk * BM + quarter * 32;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
      epilogue_tile<BN, EW / 4, Epi>(epi, p, row0, n_blk * BN, half, lane, taddr, tfull_bar(acc), acc_phase, xpose);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols) : "memory");
  }
}


// ==========================================================================================================
// CTA-pair variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile.  Each CTA loads its own
// 128 rows of A and HALF of the B tile (128 of the 256 weight rows); the leader's single thread issues
// tcgen05.mma.cta_group::2 (M = 256, N = 256), which reads A/B from both CTAs' shared memory and writes each CTA's 128
// accumulator rows into that CTA's TMEM.  Per SM and k-block this moves 32 KB instead of 48 KB through L2 -> shared memory -> tensor
// core (TMA writes + UMMA operand reads of the 1-CTA 128 x 256 tile want 192 B/clk of a 128 B/clk shared-memory pipe at full tensor
// rate, the pair 128), and the smaller stage allows a deeper ring.  Alone (null epilogue) this mainloop runs at 1.75 PFLOP/s even at
// K = 512.
//   full[s]   : leader's barrier only; both CTAs' TMA loads complete_tx on it (cta_group::2 TMA), leader arms 64 KB
//   empty[s]  : per CTA, arrived by the leader's tcgen05.commit multicast (mask 0b11)
//   tfull[a]  : per CTA, same multicast commit after the last k-block
//   tempty[a] : leader's barrier only; 2 x 16 epilogue warps arrive (the peer's through mapa / shared::cluster)
//
// Three epilogue modes (MODE):
//   0  generic: tcgen05.ld -> registers -> shared-memory transpose -> LSU stores (any functor; fp32 outputs, EXACT precision)
//   1  TMA-store, row tiling: bf16 outputs of the thread = row layout are staged ONCE in shared memory in the 128B-swizzled box layout
//      (32 rows x 64 columns per warp, conflict-free 16-byte stores) and leave with one cp.async.bulk.tensor store per warp and tile
//      — no LDS, no STG, no address arithmetic per row; the TMA engine writes whole 128-byte lines.  (FFN1, CTC1)
//   2  TMA-store, stream tiling (QKV, FAST precision): the M tiles are cut along STREAMS instead of rows through 3D maps of the A
//      operand viewed as [stream][row in chunk][k] — "segment tiles" (128 / seg_rows streams x the seg_rows segment rows) and
//      "right-context tiles" (128 / rc_rows streams x the rc_rows look-ahead rows).  A warp's 32 accumulator rows are then whole
//      streams, and every destination of the fused Q | K | V projection is a TMA box: q rows through a 3D map of q, the K / V rows of a
//      stream's segment are one seg_rows x 64 box in that session's ring (the ring advances by seg_rows, so the block never wraps), the
//      right-context K / V rows one 4D box of the scratch.  Replaces the per-row pointer table + 64-byte LSU scatter.
// ==========================================================================================================
constexpr int P_BN = 256;
constexpr int P_STAGES = 4;
constexpr int P_STAGE_BYTES = (BM + P_BN / 2) * BK * 2;             // 32 KB per CTA
constexpr int TS_WARP_BYTES = 32 * 128;                             // one 32-row x 64-column bf16 box per epilogue warp
template <int MODE> struct PairSmem {
  static constexpr int STAGING = MODE == 0 ? kStagingBytes : kEpiWarps * TS_WARP_BYTES;
  static constexpr int BYTES = P_STAGES * P_STAGE_BYTES + STAGING + 1024 + 256;
};

// MODE 1 / 2: one accumulator tile of this warp (32 rows x 64 columns: chunks 2 * half, 2 * half + 1) -> bias / activation in the
// thread = row layout -> packed bf16 -> the warp's staging box (row r at r * 128 B, 16-byte piece j at (j ^ (r & 7)) * 16: the
// CU_TENSOR_MAP_SWIZZLE_128B pattern, bank-conflict-free for the 8 lanes of a store phase) -> fence.proxy.async.  The caller issues
// the TMA store(s) after the __syncwarp() that ends this function.
template <class Epi>
__device__ __forceinline__ void ts_fill_box(const Epi& epi, const TsMaps& ts, int tile_col0, int half, int lane, uint32_t taddr, uint32_t tfull,
                                            uint32_t tfull_phase, uint32_t stg) {
  // the previous tile's store(s) of this warp must have finished READING the staging box (they had a whole tile's time)
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
  mbar_wait(tfull, tfull_phase);
  tc_fence_after();
  float v[32], w[32];
  {
    tmem_ld32(taddr + (uint32_t)(half * 64), v);
    tmem_ld32(taddr + (uint32_t)(half * 64 + 32), w);
    tmem_ld_wait();
  }
  const int col0 = tile_col0 + half * 64;
  const uint32_t row_addr = stg + (uint32_t)lane * 128u;
  const uint32_t sw = (uint32_t)(lane & 7);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    {
      const float* cb = ts.bias + col0 + 8 * q;              // kernel parameter: constant-bank loads, warp-uniform address
      epi.apply8(v + 8 * q, make_float4(cb[0], cb[1], cb[2], cb[3]), make_float4(cb[4], cb[5], cb[6], cb[7]), tile_col0);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + (((uint32_t)q ^ sw) << 4)), "r"(pack_bf16x2(v[8 * q], v[8 * q + 1])),
                 "r"(pack_bf16x2(v[8 * q + 2], v[8 * q + 3])), "r"(pack_bf16x2(v[8 * q + 4], v[8 * q + 5])), "r"(pack_bf16x2(v[8 * q + 6], v[8 * q + 7])) : "memory");
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    {
      const float* cb = ts.bias + col0 + 32 + 8 * q;
      epi.apply8(w + 8 * q, make_float4(cb[0], cb[1], cb[2], cb[3]), make_float4(cb[4], cb[5], cb[6], cb[7]), tile_col0);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + (((uint32_t)(4 + q) ^ sw) << 4)), "r"(pack_bf16x2(w[8 * q], w[8 * q + 1])),
                 "r"(pack_bf16x2(w[8 * q + 2], w[8 * q + 3])), "r"(pack_bf16x2(w[8 * q + 4], w[8 * q + 5])), "r"(pack_bf16x2(w[8 * q + 6], w[8 * q + 7])) : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // my generic-proxy writes before the async-proxy (TMA) read of the box
  __syncwarp();
}

template <class Epi, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ TsMaps ts, GemmProblem p, Epi epi,
                StreamTiling stl) {
  constexpr int NACC = 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t staging_base = smem_base + P_STAGES * P_STAGE_BYTES;
  const uint32_t bar_base = staging_base + PairSmem<MODE>::STAGING;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (P_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * P_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * P_STAGES + NACC + s); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * P_STAGES + 2 * NACC);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Warp roles: TMA producer = warp 0, MMA issuer = warp 1, epilogue = warps 2 .. 17 (TMEM lane quarter = warp % 4).  Placing the two
  // single-thread roles at the highest warp ids instead changes nothing (profiles/r02_gemm_warp_roles.log).
  constexpr int kProducerWarp = 0, kMmaWarp = 1, kEpiWarp0 = 2;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int n_tiles = (p.N + P_BN - 1) / P_BN;
  // MODE 2: pair tiles [0, seg_pairs) are segment tiles, the rest right-context tiles
  const int m_pairs = MODE == 2 ? stl.seg_pairs + stl.rc_pairs : (p.M + 2 * BM - 1) / (2 * BM);
  const int num_tiles = m_pairs * n_tiles;
  const int kb_per_pass = p.K / BK;
  const int total_kb = kb_per_pass * p.passes;

  if (warp == kProducerWarp && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (MODE >= 1) asm volatile("prefetch.tensormap [%0];" ::"l"(&ts.c0) : "memory");
    if (MODE == 2) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&ts.a_rc) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&ts.c1) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&ts.c2) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&ts.c3) : "memory");
    }
    for (int s = 0; s < P_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < NACC; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 2 * kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                                      // peer's barriers are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == kProducerWarp) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        const int m_pair = tile / n_tiles, n_blk = tile - m_pair * n_tiles;
        const int row_a = m_pair * 2 * BM + (int)rank * BM;
        const int row_b = n_blk * P_BN + (int)rank * (P_BN / 2);
        const bool seg_tile = m_pair < stl.seg_pairs;
        const int stream0 = seg_tile ? (2 * m_pair + (int)rank) * stl.spt_seg : (2 * (m_pair - stl.seg_pairs) + (int)rank) * stl.spt_rc;
        for (int ps = 0; ps < p.passes; ++ps) {
          for (int kb = 0; kb < kb_per_pass; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);       // own slot free: the leader's MMAs that read it (in both CTAs) retired
            const uint32_t sa = smem_base + stage * P_STAGE_BYTES;
            const uint32_t sb = sa + BM * BK * 2;
            const uint32_t lbar = full_bar(stage) & kPeerBitMask;
            if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * (uint32_t)P_STAGE_BYTES);
            if constexpr (MODE == 2) {
              if (seg_tile) tma_load_3d_pair(sa, &tmA, p.a_koff[ps] + kb * BK, 0, stream0, lbar);
              else tma_load_3d_pair(sa, &ts.a_rc, p.a_koff[ps] + kb * BK, stl.seg_rows, stream0, lbar);
            } else {
              tma_load_2d_pair(sa, &tmA, p.a_koff[ps] + kb * BK, row_a, lbar);
            }
            tma_load_2d_pair(sb, &tmB, p.b_koff[ps] + kb * BK, row_b, lbar);
            if (++stage == P_STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P_BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
#ifdef ASR_MMA_TIMING
      long long w_empty = 0, w_full = 0, n_t = 0;
      const long long t_begin = clock64();
#endif
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
#ifdef ASR_MMA_TIMING
        const long long ta = clock64();
#endif
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);        // both CTAs' epilogues drained this accumulator stage
        tc_fence_after();
#ifdef ASR_MMA_TIMING
        w_empty += clock64() - ta; ++n_t;
#endif
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * P_BN);
        for (int kb = 0; kb < total_kb; ++kb) {
#ifdef ASR_MMA_TIMING
          const long long tb = clock64();
#endif
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
#ifdef ASR_MMA_TIMING
          w_full += clock64() - tb;
#endif
          if (lane == 0) {
            const uint32_t sa = smem_base + stage * P_STAGE_BYTES;
            const uint64_t adesc = make_smem_desc(sa);
            const uint64_t bdesc = make_smem_desc(sa + BM * BK * 2);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16_pair(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_pair(empty_bar(stage), 3);
            if (kb == total_kb - 1) umma_commit_pair(tfull_bar(acc), 3);
          }
          __syncwarp();
          if (++stage == P_STAGES) { stage = 0; phase ^= 1u; }
        }
        if (++acc == NACC) { acc = 0; acc_phase ^= 1u; }
      }
#ifdef ASR_MMA_TIMING
      if (blockIdx.x == 0 && lane == 0)
        printf("MMA thread of CTA 0 (MODE %d, N %d K %d): %lld tiles, per tile %lld clk total, waiting for a free accumulator %lld, for operand stages %lld\n", MODE, p.N, p.K, n_t,
               (clock64() - t_begin) / (n_t ? n_t : 1), w_empty / (n_t ? n_t : 1), w_full / (n_t ? n_t : 1));
#endif
    }
  } else {
    // ===================== epilogue (16 warps, both CTAs) =====================
    const int ew = warp - kEpiWarp0;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    float* xpose = reinterpret_cast<float*>(smem_raw + (staging_base - smem_u32(smem_raw))) + ew * kXposeFloats;     // MODE 0
    const uint32_t stg = staging_base + (uint32_t)(ew * TS_WARP_BYTES);                                               // MODE 1 / 2
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
      const int m_pair = tile / n_tiles, n_blk = tile - m_pair * n_tiles;
      const int row0 = m_pair * 2 * BM + (int)rank * BM + quarter * 32;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * P_BN);
      if constexpr (MODE == 0) {
        epilogue_tile<P_BN, kEpiWarps / 4, Epi>(epi, p, row0, n_blk * P_BN, half, lane, taddr, tfull_bar(acc), acc_phase, xpose);
      } else if constexpr (MODE == 1) {
        ts_fill_box(epi, ts, n_blk * P_BN, half, lane, taddr, tfull_bar(acc), acc_phase, stg);
        if (lane == 0 && row0 < p.M) {
          tma_store_2d(&ts.c0, stg, n_blk * P_BN + half * 64, row0);
          tma_store_commit();
        }
      } else {
        // ---- MODE 2: Q | K | V with stream tiling
        const bool seg_tile = m_pair < stl.seg_pairs;
        const int spw = seg_tile ? stl.spw_seg : stl.spw_rc;                      // streams per warp (32 rows)
        const int stream_w = (seg_tile ? (2 * m_pair + (int)rank) * stl.spt_seg : (2 * (m_pair - stl.seg_pairs) + (int)rank) * stl.spt_rc) + quarter * spw;
        const int tile_col0 = n_blk * P_BN;
        const int sec = tile_col0 / epi.d;                                        // 0: q, 1: k, 2: v (a tile never straddles the sections)
        const int col_sec = tile_col0 - sec * epi.d + half * 64;
        int kv_row = -1;                                                          // lane i < spw: cache row of stream (stream_w + i)'s segment block
        if (sec > 0 && seg_tile && lane < spw && stream_w + lane < stl.n_streams) {
          const int slot = __ldg(epi.slots + stream_w + lane);
          kv_row = (int)(epi.kv_row0 + (long long)slot * epi.slot_rows + (long long)(sec - 1) * epi.ring + __ldg(epi.past_len + slot) % epi.ring);
        }
        ts_fill_box(epi, ts, tile_col0, half, lane, taddr, tfull_bar(acc), acc_phase, stg);
        if (sec == 0) {
          if (lane == 0 && stream_w < stl.n_streams) {
            if (seg_tile) tma_store_3d(&ts.c0, stg, col_sec, 0, stream_w); else tma_store_3d(&ts.c1, stg, col_sec, stl.seg_rows, stream_w);
            tma_store_commit();
          }
        } else if (seg_tile) {
          if (kv_row >= 0) {
            tma_store_2d(&ts.c2, stg + (uint32_t)(lane * stl.seg_rows * 128), col_sec, kv_row);
            tma_store_commit();
          }
        } else if (lane == 0 && stream_w < stl.n_streams) {
          tma_store_4d(&ts.c3, stg, col_sec, 0, sec - 1, stream_w);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(tempty_bar(acc));
        else mbar_arrive_cta(tempty_bar(acc), 0);
      }
      if (++acc == NACC) { acc = 0; acc_phase ^= 1u; }
    }
    if (MODE >= 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");       // the staging box outlives this CTA's last store
  }

#ifdef ASR_EPI_TIMING
  if (blockIdx.x == 0 && threadIdx.x == 32 * kEpiWarp0 && g_epi_clk[0]) {
    printf("epilogue warp 0 of CTA 0: %llu tiles, per tile: wait tfull %llu clk, TMEM loads %llu clk, math + staging + stores %llu clk (of which LDS + STG %llu)\n", g_epi_clk[0],
           g_epi_clk[1] / g_epi_clk[0], g_epi_clk[2] / g_epi_clk[0], g_epi_clk[3] / g_epi_clk[0], g_epi_clk[4] / g_epi_clk[0]);
    g_epi_clk[0] = g_epi_clk[1] = g_epi_clk[2] = g_epi_clk[3] = g_epi_clk[4] = 0;
  }
#endif
  tc_fence_before();
  cluster_sync_all();                                      // nobody frees TMEM / exits while the peer may still signal or read
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <class Epi, int MODE>
int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB128, const TsMaps& ts, const GemmProblem& p, const Epi& epi, const StreamTiling& stl,
                int num_sms, cudaStream_t st) {
  constexpr int SMEM = PairSmem<MODE>::BYTES;
  static size_t attr_done[kMaxDevices] = {0};
  ASR_CUDA_OK(ensure_dyn_smem(gemm_tc2_kernel<Epi, MODE>, (size_t)SMEM, attr_done));
  const int m_pairs = MODE == 2 ? stl.seg_pairs + stl.rc_pairs : (p.M + 2 * BM - 1) / (2 * BM);
  const int tiles = m_pairs * ((p.N + P_BN - 1) / P_BN);
  const int pairs = tiles < num_sms / 2 ? tiles : num_sms / 2;
  ASR_CUDA_OK(launch_pdl(gemm_tc2_kernel<Epi, MODE>, dim3(2 * pairs), dim3(kThreads), SMEM, st, tmA, tmB128, ts, p, epi, stl));
  return 0;
}

template <int BN, class Epi>
int launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmProblem& p, const Epi& epi, int num_sms, cudaStream_t st) {
  using Cfg = TileCfg<BN, EpiWarps<Epi>::value>;
  static size_t attr_done[kMaxDevices] = {0};
  ASR_CUDA_OK(ensure_dyn_smem(gemm_tc_kernel<BN, Epi>, (size_t)Cfg::kSmemBytes, attr_done));
  const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  ASR_CUDA_OK(launch_pdl(gemm_tc_kernel<BN, Epi>, dim3(grid), dim3(Cfg::kThreadsCta), Cfg::kSmemBytes, st, tmA, tmB, p, epi));
  return 0;
}

}  // namespace

template <class Epi>
int gemm_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmProblem& p, const Epi& epi, int bn, int num_sms, cudaStream_t st,
            const TsMaps* ts, const StreamTiling* stl) {
  if (p.M <= 0) return 0;
  if (p.K % BK != 0) { set_error("gemm_tc: K=%d not a multiple of %d", p.K, BK); return -1; }
  static const TsMaps no_maps = {};
  switch (bn) {
    case kPairTile: return launch_pair<Epi, 0>(tmA, tmB, no_maps, p, epi, StreamTiling{}, num_sms, st);      // tmB must be the 128-row-box map
    case kPairTileTS:
      if constexpr (Epi::kBf16Rows && !Epi::kStreamTiles) {
        if (!ts || !epi.bf16_rows() || p.N % P_BN != 0 || p.N > kTsBiasMax) { set_error("gemm_tc: the TMA-store epilogue needs bf16 row outputs, their tensor map and N %% 256 == 0 (N %d)", p.N); return -1; }
        return launch_pair<Epi, 1>(tmA, tmB, *ts, p, epi, StreamTiling{}, num_sms, st);
      } else break;
    case kPairTileQKV:
      if constexpr (Epi::kStreamTiles) {
        if (!ts || !stl || p.N != 3 * epi.d || p.N > kTsBiasMax || epi.d % P_BN != 0 || 128 % stl->seg_rows || 32 % stl->seg_rows || 32 % stl->rc_rows || stl->seg_rows + stl->rc_rows != epi.rows) {
          set_error("gemm_tc: stream tiling needs the tensor maps, N = 3 d, d %% 256 == 0 and row counts that divide 32"); return -1;
        }
        return launch_pair<Epi, 2>(tmA, tmB, *ts, p, epi, *stl, num_sms, st);
      } else break;
    case 64: return launch_bn<64, Epi>(tmA, tmB, p, epi, num_sms, st);
    case 128: return launch_bn<128, Epi>(tmA, tmB, p, epi, num_sms, st);
    case 256: return launch_bn<256, Epi>(tmA, tmB, p, epi, num_sms, st);
  }
  set_error("gemm_tc: unsupported BN=%d for this epilogue", bn);
  return -1;
}

#define ASR_INST_GEMM(E) template int gemm_tc<E>(const CUtensorMap&, const CUtensorMap&, const GemmProblem&, const E&, int, int, cudaStream_t, const TsMaps*, const StreamTiling*)
ASR_INST_GEMM(EpiF32);
ASR_INST_GEMM(EpiNull);
ASR_INST_GEMM(EpiOperand);
ASR_INST_GEMM(EpiQKV<float>);
ASR_INST_GEMM(EpiQKV<bf16>);
#undef ASR_INST_GEMM

// ------------------------------------------------------------------------------------------
// Host: TMA descriptors.  cuTensorMapEncodeTiled is fetched through the runtime so the library has no
// link-time dependency on libcuda (it must dlopen on a GPU-less box for the symbol-export test).
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_elems, uint32_t box_rows) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
      set_error("cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(e));
      return -1;
    }
    fn = (EncodeTiledFn)p;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (cols %llu rows %llu ld %llu box_rows %u)", (int)r,
              (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)ld_elems, box_rows);
    return -1;
  }
  return 0;
}

// 3D view of a [rows, n_heads * 64] bf16 matrix as (64 dims, rows, heads) with a {64, box_rows, n_heads} box and 128B swizzle: one
// TMA op lands box_rows full rows in shared memory head-major ([head][row][128 B]), the layout the attention warps want.
int make_tmap_bf16_heads(CUtensorMap* out, const void* base, uint64_t rows, uint32_t n_heads, uint32_t box_rows) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) { set_error("cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(e)); return -1; }
    fn = (EncodeTiledFn)p;
  }
  cuuint64_t gdim[3] = {64, rows, n_heads};
  cuuint64_t gstr[2] = {(cuuint64_t)n_heads * 128, 128};            // bytes: row stride, head stride
  cuuint32_t box[3] = {64, box_rows, n_heads};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(3D) failed: CUresult %d (rows %llu heads %u box_rows %u)", (int)r, (unsigned long long)rows, n_heads, box_rows); return -1; }
  return 0;
}

int make_tmap_bf16_nd(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) { set_error("cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(e)); return -1; }
    fn = (EncodeTiledFn)p;
  }
  if (rank < 2 || rank > 4) { set_error("make_tmap_bf16_nd: rank %d", rank); return -1; }
  cuuint64_t gdim[4]; cuuint64_t gstr[3]; cuuint32_t bx[4]; cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(rank %d) failed: CUresult %d", rank, (int)r); return -1; }
  return 0;
}

}  // namespace asr
