// GEMM entry points (tcgen05 / TMEM / TMA).
#pragma once
#include "common.cuh"

namespace asr {

constexpr int kPairTile = 512;      // `bn` value selecting the cta_group::2 kernel (256 x 256 tile per CTA pair; tmB = box-128 map), LSU epilogue
constexpr int kPairTileTS = 515;    // the same kernel with the TMA-store epilogue (bf16-row functors; needs TsMaps::c0 = map of the output, box {64, 32})
constexpr int kPairTileQKV = 516;   // the same kernel with stream tiling + TMA-store epilogue for the fused Q | K | V projection (EpiQKV<bf16>)

// Tensor maps of the TMA-store epilogues (one kernel parameter).  All bf16, 128B swizzle, 64 elements (128 B) wide.
//   kPairTileTS : c0 = the output [rows, N], box {64, 32}
//   kPairTileQKV: tmA = A operand as [stream][row in chunk][K], box {64, seg_rows, 128 / seg_rows};  a_rc = same tensor, box {64, rc_rows,
//                 128 / rc_rows};  c0 / c1 = q as [stream][row][d], boxes {64, seg_rows, 32 / seg_rows} and {64, rc_rows, 32 / rc_rows};
//                 c2 = the K/V cache as [ring rows of all layers and sessions][d], box {64, seg_rows};  c3 = the right-context K/V scratch
//                 [stream][K|V][rc row][d] as 4D, box {64, rc_rows, 1, 32 / rc_rows}
// bias: the GEMM's bias vector, copied into the kernel parameters (constant bank).  In the thread = row layout of the TMEM loads every
// lane needs the SAME 64 bias values per tile: as global loads those are 16 uniform LDG.128 per thread and tile — more L1 / shared-memory
// pipe wavefronts than the epilogue's own stores, on the pipe the mainloop's TMA writes + UMMA operand reads already saturate (measured:
// +19..25 us per launch at M = 81,920); from the constant bank they cost no load/store-pipe traffic at all.
constexpr int kTsBiasMax = 2048;
struct TsMaps { CUtensorMap c0, c1, c2, c3, a_rc; float bias[kTsBiasMax]; };
struct StreamTiling {      // kPairTileQKV: M tiles cut along streams
  int n_streams = 0, seg_rows = 0, rc_rows = 0;
  int spt_seg = 0, spt_rc = 0;       // streams per 128-row CTA tile: 128 / seg_rows, 128 / rc_rows
  int spw_seg = 0, spw_rc = 0;       // streams per 32-row warp slice
  int seg_pairs = 0, rc_pairs = 0;   // 256-row pair tiles of each kind
};
inline StreamTiling make_stream_tiling(int n_streams, int seg_rows, int rc_rows) {
  StreamTiling t;
  t.n_streams = n_streams; t.seg_rows = seg_rows; t.rc_rows = rc_rows;
  t.spt_seg = 128 / seg_rows; t.spt_rc = 128 / rc_rows; t.spw_seg = 32 / seg_rows; t.spw_rc = 32 / rc_rows;
  t.seg_pairs = (n_streams + 2 * t.spt_seg - 1) / (2 * t.spt_seg); t.rc_pairs = (n_streams + 2 * t.spt_rc - 1) / (2 * t.spt_rc);
  return t;
}

// C = A * B^T with fused epilogue.  tmA / tmB: 2D bf16 tensor maps, box {64, 128} and {64, bn}, 128B swizzle.
template <class Epi>
int gemm_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmProblem& p, const Epi& epi, int bn, int num_sms, cudaStream_t st,
            const TsMaps* ts = nullptr, const StreamTiling* stl = nullptr);

// N = 512 GEMM with residual add + LayerNorm(s) fused into the epilogue (gemm_ln.cu).  With v = A W^T + bias + res:
//   g2 == nullptr, f32_normed = 0 :  out_f32 = v,            out_op = bf16 LN(v; g1, b1)                       (out_proj -> FFN1 operand)
//   g2 != nullptr                 :  out_f32 = LN(v; g1,b1), out_op = bf16 LN(out_f32; g2, b2)                 (FFN2 -> next layer's QKV operand)
//   g2 == nullptr, f32_normed = 1 :  out_f32 = LN(v; g1,b1), out_op = bf16 out_f32                              (last FFN2 -> CTC operand)
// compact_rows > 0: out_op keeps only rows t < compact_seg of every block of compact_rows rows, packed (segment rows without right context).
struct LnEpilogue {
  const float* bias;     // [512]
  const float* res;      // [M, 512] fp32
  const float* g1; const float* b1;
  const float* g2; const float* b2;
  float* out_f32;        // [M, 512]
  bf16* out_op;          // [rows, op_ld] bf16 hi (+ lo at op_lo_off)
  int op_ld, op_lo_off;
  int f32_normed;
  int compact_rows, compact_seg;
  // optional (two-LN form, cta_group::2 shape): constants of the first LayerNorm that let its output statistics be derived from
  // v in one pass: [0,512) gc = g1 (b1 - mean b1) | [512,528) sum g1 per 32-column chunk | [528,544) sum g1^2 | [544,560) sum gc |
  // [560] mean b1 | [561] var b1 (biased).  nullptr: the second LayerNorm's statistics take their own pass over TMEM.
  const float* y_consts = nullptr;
};
// tmB256 / tmB128: tensor maps of the [512, ld] weight with 256- / 128-row boxes.  shape 0: cluster of 2 column halves;
// 1: cta_group::2 pairs x 2 column halves (large M, long K); 2: cluster of 4 column quarters (small M).
int gemm_ln(const CUtensorMap& tmA, const CUtensorMap& tmB256, const CUtensorMap& tmB128, const GemmProblem& p, const LnEpilogue& ep, int shape,
            int num_sms, cudaStream_t st);

int make_tmap_bf16_heads(CUtensorMap* out, const void* base, uint64_t rows, uint32_t n_heads, uint32_t box_rows);
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_elems, uint32_t box_rows);
// general bf16 map, rank <= 4, 128B swizzle, innermost box 64 elements: dims / strides (bytes, for dims 1 ..) / box given innermost first
int make_tmap_bf16_nd(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);

}  // namespace asr
