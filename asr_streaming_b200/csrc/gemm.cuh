// GEMM entry points (tcgen05 product path + SIMT cross-check kernel used by tests/bring-up only).
#pragma once
#include "common.cuh"

namespace asr {

constexpr int kPairTileA = 514;     // cta_group::2 kernel with the A tile resident in shared memory (K <= 512, bf16-row epilogues; tmB = box-128 map)
constexpr int kPairTile128 = 513;   // cta_group::2 kernel with a 256 x 128 tile and four accumulator stages (tmB = box-64 map)
constexpr int kPairTile = 512;   // `bn` value selecting the cta_group::2 kernel (256 x 256 tile per CTA pair; tmB = box-128 map)

// C = A * B^T with fused epilogue.  tmA / tmB: 2D bf16 tensor maps, box {64, 128} and {64, bn}, 128B swizzle.
template <class Epi>
int gemm_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmProblem& p, const Epi& epi, int bn, int num_sms, cudaStream_t st);

// Same contract on CUDA cores (fp32 FMA over the same bf16 operands).  Diagnostic cross-check for the
// tcgen05 kernel; selected only by ASR_B200_DEBUG_SIMT_GEMM=1 or asr_debug_gemm().
template <class Epi>
int gemm_simt(const bf16* A, int lda, const bf16* B, int ldb, const GemmProblem& p, const Epi& epi, cudaStream_t st);

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

// Fused feed-forward block: FFN1 (+ bias + GELU) -> L2-resident scratch -> FFN2 + residual + LayerNorm(s) in one persistent kernel.
struct MlpFuse {
  bf16* h = nullptr;          // scratch [clusters * 256, ffn] bf16 (mlp_ln_clusters() slabs)
  const float* b1 = nullptr;  // [ffn] FFN1 bias
  int n_htiles = 0;           // ffn / 256
  int kb1 = 0;                // d_model / 64
  int ffn = 0;
};
int mlp_ln_clusters(int num_sms);
int mlp_ln(const CUtensorMap& tmH, const CUtensorMap& tmW2_128, const CUtensorMap& tmX, const CUtensorMap& tmW1_128, const GemmProblem& p,
           const LnEpilogue& ep, const MlpFuse& mf, int num_sms, cudaStream_t st);

int make_tmap_bf16_heads(CUtensorMap* out, const void* base, uint64_t rows, uint32_t n_heads, uint32_t box_rows);
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_elems, uint32_t box_rows);

}  // namespace asr
