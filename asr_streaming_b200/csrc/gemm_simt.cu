// CUDA-core cross-check GEMM (diagnostic only; see gemm.cuh).  One thread per output row, 32 columns per
// block: the accumulator fragment has the same shape as the tcgen05 epilogue's, so the epilogue functors
// are shared verbatim and a disagreement isolates the TMA/UMMA descriptors.
#include "gemm.cuh"

namespace asr {

namespace {

template <class Epi>
__global__ void __launch_bounds__(128) gemm_simt_kernel(const bf16* __restrict__ A, int lda, const bf16* __restrict__ B, int ldb,
                                                        GemmProblem p, Epi epi) {
  __shared__ float bs[32][65];
  const int row = blockIdx.y * 128 + threadIdx.x;
  const int col0 = blockIdx.x * 32;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  for (int ps = 0; ps < p.passes; ++ps) {
    for (int k0 = 0; k0 < p.K; k0 += 64) {
      __syncthreads();
      for (int i = threadIdx.x; i < 32 * 64; i += 128) {
        const int j = i >> 6, k = i & 63;
        bs[j][k] = (col0 + j < p.N) ? __bfloat162float(B[(size_t)(col0 + j) * ldb + p.b_koff[ps] + k0 + k]) : 0.f;
      }
      __syncthreads();
      if (row < p.M) {
        const bf16* a = A + (size_t)row * lda + p.a_koff[ps] + k0;
        for (int k = 0; k < 64; ++k) {
          const float av = __bfloat162float(a[k]);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = fmaf(av, bs[j][k], acc[j]);
        }
      }
    }
  }
  if (row < p.M) epi.store(row, col0, acc, epi.row_ctx(row, col0));
}

}  // namespace

template <class Epi>
int gemm_simt(const bf16* A, int lda, const bf16* B, int ldb, const GemmProblem& p, const Epi& epi, cudaStream_t st) {
  if (p.M <= 0) return 0;
  dim3 grid((p.N + 31) / 32, (p.M + 127) / 128);
  gemm_simt_kernel<Epi><<<grid, 128, 0, st>>>(A, lda, B, ldb, p, epi);
  ASR_CUDA_OK(cudaGetLastError());
  return 0;
}

template int gemm_simt<EpiF32>(const bf16*, int, const bf16*, int, const GemmProblem&, const EpiF32&, cudaStream_t);
template int gemm_simt<EpiOperand>(const bf16*, int, const bf16*, int, const GemmProblem&, const EpiOperand&, cudaStream_t);
template int gemm_simt<EpiQKV<float>>(const bf16*, int, const bf16*, int, const GemmProblem&, const EpiQKV<float>&, cudaStream_t);
template int gemm_simt<EpiQKV<bf16>>(const bf16*, int, const bf16*, int, const GemmProblem&, const EpiQKV<bf16>&, cudaStream_t);

}  // namespace asr
