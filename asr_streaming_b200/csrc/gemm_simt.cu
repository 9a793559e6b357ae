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
  // hand the 128 x 32 block to the epilogue in the same fragment shape the tcgen05 kernel uses
  __shared__ __align__(16) float tile[4][32][36];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.y * 128 + warp * 32;
  const typename Epi::RowCtx my_ctx = epi.row_ctx(row < p.M ? row : p.M - 1, col0);
#pragma unroll
  for (int j = 0; j < 32; ++j) tile[warp][lane][j] = acc[j];
  __syncwarp();
  typename Epi::RowCtx ctx[8];
  float v[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    ctx[i] = shfl_ctx<typename Epi::RowCtx>(my_ctx, 4 * i + (lane >> 3));
#pragma unroll
    for (int j = 0; j < 4; ++j) v[i][j] = tile[warp][4 * i + (lane >> 3)][4 * (lane & 7) + j];
  }
  const int col = col0 + 4 * (lane & 7);
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (epi.bias_ptr()) {
    const float* bp = epi.bias_ptr();
    b4 = make_float4(col < p.N ? bp[col] : 0.f, col + 1 < p.N ? bp[col + 1] : 0.f, col + 2 < p.N ? bp[col + 2] : 0.f, col + 3 < p.N ? bp[col + 3] : 0.f);
  }
  if (col < p.N) epi.store(row0, col, lane, p.M, v, ctx, b4);
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
