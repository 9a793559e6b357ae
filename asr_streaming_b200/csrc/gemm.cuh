// GEMM entry points (tcgen05 product path + SIMT cross-check kernel used by tests/bring-up only).
#pragma once
#include "common.cuh"

namespace asr {

constexpr int kPairTile = 512;   // `bn` value selecting the cta_group::2 kernel (256 x 256 tile per CTA pair; tmB = box-128 map)

// C = A * B^T with fused epilogue.  tmA / tmB: 2D bf16 tensor maps, box {64, 128} and {64, bn}, 128B swizzle.
template <class Epi>
int gemm_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmProblem& p, const Epi& epi, int bn, int num_sms, cudaStream_t st);

// Same contract on CUDA cores (fp32 FMA over the same bf16 operands).  Diagnostic cross-check for the
// tcgen05 kernel; selected only by ASR_B200_DEBUG_SIMT_GEMM=1 or asr_debug_gemm().
template <class Epi>
int gemm_simt(const bf16* A, int lda, const bf16* B, int ldb, const GemmProblem& p, const Epi& epi, cudaStream_t st);

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_elems, uint32_t box_rows);

}  // namespace asr
