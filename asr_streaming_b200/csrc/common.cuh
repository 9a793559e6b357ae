// Shared device/host helpers for the B200 (sm_100a) lightspeech per-chunk path.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace asr {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------
// Geometry shared by every kernel (mirrors AudioConfig, reference streaming_decoder/utils.py:9-23
// and the model hyper-parameters of SURVEY.md §8a).
// ------------------------------------------------------------------------------------------
struct Geo {
  int chunk_len;     // samples per stream-chunk (13440)
  int hop;           // 160
  int n_fft;         // 800
  int win;           // 400
  int n_mels;        // 128
  int frames;        // 80 fbank frames per chunk
  int stride;        // 4
  int d_model;       // 512
  int n_heads;       // 8
  int ffn;           // 2048
  int n_layers;      // 20
  int seg_rows;      // 16
  int rc_rows;       // 4
  int rows;          // 20 = seg_rows + rc_rows
  int left;          // 32
  int ring;          // left + seg_rows = 48 rows of K/V kept per layer per session
  int ctc_hidden;    // 512
  int vocab;         // 804
  int split;         // 0: bf16 operands (FAST), 1: split-bf16 hi/lo, 3 products (EXACT)
};

#define ASR_CUDA_OK(expr)                                                                       \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      asr::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));    \
      return -1;                                                                                \
    }                                                                                           \
  } while (0)

void set_error(const char* fmt, ...);

// ------------------------------------------------------------------------------------------
// Small device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// Write 8 consecutive fp32 values as a GEMM A-operand row fragment: bf16 "hi" at out[col..col+8) and,
// when lo_off != 0 (EXACT mode), the bf16 residual "lo" at out[lo_off+col ..).  16-byte stores.
__device__ __forceinline__ void store_operand8(bf16* __restrict__ row_ptr, int col, int lo_off, const float* v) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bf16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
    h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[i] = pack_bf16x2(v[2 * i] - __bfloat162float(h0), v[2 * i + 1] - __bfloat162float(h1));
  }
  *reinterpret_cast<uint4*>(row_ptr + col) = make_uint4(h[0], h[1], h[2], h[3]);
  if (lo_off) *reinterpret_cast<uint4*>(row_ptr + lo_off + col) = make_uint4(l[0], l[1], l[2], l[3]);
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// erf(x) = sign(x) * (1 - 2^q(min(|x|, 4))), q = degree-6 minimax fit of log2(erfc) with zero constant term (weights =
// d erf / d q); max abs error 3.1e-7 over the whole line in fp32 (fit + check: DESIGN.md §2).  One MUFU.EX2 + 7 FMA,
// branch-free.  The GELU epilogue of FFN1 runs once per output of a K = 512 GEMM, i.e. it has ~16 issue slots per element
// before it, not the tensor pipe, bounds the kernel; libdevice erff costs ~25 instructions, a rational form 2 MUFU ops.
__device__ __forceinline__ float erf_fast(float x) {
  const float a = fminf(fabsf(x), 4.0f);
  float q = 1.420474000e-04f;
  q = fmaf(q, a, -3.664300360e-03f);
  q = fmaf(q, a, 3.089622360e-02f);
  q = fmaf(q, a, -1.496994580e-01f);
  q = fmaf(q, a, -9.181654620e-01f);
  q = fmaf(q, a, -1.627925070e+00f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q * a));       // q*a in [-27, 0]: no denormal / overflow handling needed
  return copysignf(1.0f - e, x);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// ------------------------------------------------------------------------------------------
// GEMM epilogues.  One thread owns one accumulator row (the shape tcgen05.ld 32x32b delivers) and gets it in
// fragments of 32 consecutive columns; every global access is a 16-byte vector.  Anything that depends only on
// the row (destination of a K/V row in the session ring) is computed once per tile (RowCtx).
// (Measured on B200: a shared-memory transpose to lane = column with 4-byte accesses doubles the epilogue time;
//  profiles/r01_notes.md.)
// ------------------------------------------------------------------------------------------
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_SILU = 2 };

// out[row, col] = acc (+bias) (+res)        fp32 out; n_valid masks a ragged N (CTC vocab = 804)
struct EpiF32 {
  float* out;
  const float* bias;   // nullable
  const float* res;    // nullable, same ld as out
  int ld;
  int n_valid;
  typedef int RowCtx;
  __device__ __forceinline__ RowCtx row_ctx(int, int) const { return 0; }
  __device__ __forceinline__ void store(int row, int col0, float (&v)[32], RowCtx) const {
    float* o = out + (size_t)row * ld + col0;
    const float* r = res ? res + (size_t)row * ld + col0 : nullptr;
    if (col0 + 32 <= n_valid) {
      float4 rr[8];
      if (r) {
#pragma unroll
        for (int j = 0; j < 8; ++j) rr[j] = *reinterpret_cast<const float4*>(r + 4 * j);      // all loads in flight first
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 t = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        if (bias) { const float4 b = *reinterpret_cast<const float4*>(bias + col0 + 4 * j); t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w; }
        if (r) { t.x += rr[j].x; t.y += rr[j].y; t.z += rr[j].z; t.w += rr[j].w; }
        *reinterpret_cast<float4*>(o + 4 * j) = t;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < n_valid) o[j] = v[j] + (bias ? bias[col0 + j] : 0.f) + (r ? r[j] : 0.f);
    }
  }
};

// out = act(acc + bias) written as the next GEMM's A operand (bf16 hi [+ lo])
struct EpiOperand {
  bf16* out;
  const float* bias;
  int ld;       // elements per row (N, or 2N when split)
  int lo_off;   // 0 or N
  int act;
  typedef int RowCtx;
  __device__ __forceinline__ RowCtx row_ctx(int, int) const { return 0; }
  __device__ __forceinline__ void store(int row, int col0, float (&v)[32], RowCtx) const {
    bf16* o = out + (size_t)row * ld;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(bias + col0 + j);
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
    if (act == ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
    } else if (act == ACT_SILU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = silu(v[j]);
    }
#pragma unroll
    for (int j = 0; j < 32; j += 8) store_operand8(o, col0 + j, lo_off, &v[j]);
  }
};

// Fused Q | K | V projection (TA:emformer.py:161,:164): q scaled by d_h^-0.5 (:188);
// K/V of the segment rows -> the session's ring slot (replaces _pack_state cat+slice, :400-414);
// K/V of right-context rows -> per-step scratch (never cached, :314-315).  A tile never straddles the q|k|v
// sections (d % BN == 0), so the destination row pointer is a per-tile, per-row constant.
template <typename T>
struct EpiQKV {
  T* q;                  // [M, d]  (same element type as the K/V cache: bf16 in FAST, fp32 in EXACT)
  T* cache_layer;        // cache + layer * (2 * ring * d)
  size_t slot_stride;    // elements per session slot
  T* rc;                 // [B, 2, rc_rows, d]
  const float* bias;     // [3d]
  const int* slots;      // [B]
  const int* past_len;   // [n_slots]
  int rows, seg_rows, rc_rows, ring, d;
  float qscale;
  typedef T* RowCtx;                          // column 0 of this tile's section (q | k | v) in the destination row
  __device__ __forceinline__ RowCtx row_ctx(int row, int tile_col0) const {
    const int sec = tile_col0 / d;            // 0: q, 1: k, 2: v
    if (sec == 0) return q + (size_t)row * d;
    const int which = sec - 1;
    const int b = row / rows, t = row - b * rows;
    if (t < seg_rows) {
      const int slot = slots[b];
      const int rr = (past_len[slot] + t) % ring;
      return cache_layer + (size_t)slot * slot_stride + ((size_t)which * ring + rr) * d;
    }
    return rc + (((size_t)b * 2 + which) * rc_rows + (t - seg_rows)) * d;
  }
  __device__ __forceinline__ void store(int, int col0, float (&v)[32], RowCtx dst_row) const {
    const int sec = col0 / d;
    const float sc = sec == 0 ? qscale : 1.0f;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(bias + col0 + j);
      v[j] = (v[j] + b.x) * sc; v[j + 1] = (v[j + 1] + b.y) * sc; v[j + 2] = (v[j + 2] + b.z) * sc; v[j + 3] = (v[j + 3] + b.w) * sc;
    }
    T* dst = dst_row + (col0 - sec * d);
    if (sizeof(T) == 4) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 8)
        *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(dst) + j) =
            make_uint4(pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]), pack_bf16x2(v[j + 4], v[j + 5]),
                       pack_bf16x2(v[j + 6], v[j + 7]));
    }
  }
};

// ------------------------------------------------------------------------------------------
// GEMM problem: C[M,N] = sum over passes of A[:, a_koff[p] : +K] * B[:, b_koff[p] : +K]^T
// A: [M, lda] bf16 row-major (K-major), B: [N, ldb] bf16 row-major (nn.Linear weight layout).
// passes = 1 (FAST) or 3 (EXACT: hi*hi + lo*hi + hi*lo with hi|lo stored side by side along K).
// ------------------------------------------------------------------------------------------
struct GemmProblem {
  int M, N, K;
  int passes;
  int a_koff[3];
  int b_koff[3];
};

inline GemmProblem make_problem(int M, int N, int K, int split) {
  GemmProblem p;
  p.M = M; p.N = N; p.K = K;
  p.passes = split ? 3 : 1;
  p.a_koff[0] = 0; p.a_koff[1] = K; p.a_koff[2] = 0;
  p.b_koff[0] = 0; p.b_koff[1] = 0; p.b_koff[2] = K;
  return p;
}

}  // namespace asr
