// Shared device/host helpers for the B200 (sm_100a) lightspeech per-chunk path.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <utility>

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

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: a process that drives several GPUs (GpuRouter: one Engine per
// device) must set it once on each.  `done` is a per-call-site table indexed by device ordinal.
constexpr int kMaxDevices = 64;
inline int current_device_index() { int d = 0; cudaGetDevice(&d); return (d < 0 || d >= kMaxDevices) ? 0 : d; }
template <typename K>
inline cudaError_t ensure_dyn_smem(K kernel, size_t bytes, size_t (&done)[kMaxDevices]) {
  const int d = current_device_index();
  if (done[d] >= bytes) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) done[d] = bytes;
  return e;
}

// Programmatic dependent launch: every kernel of the per-step chain is launched with programmaticStreamSerialization, calls
// pdl_launch_dependents() early (the next kernel's CTAs may be scheduled as soon as SM resources free up and run their
// prologue: barrier init, TMEM allocation, tensor-map prefetch, table loads) and pdl_wait() before its first access to memory
// produced by a predecessor (griddepcontrol.wait returns once all prerequisite grids have completed and flushed).
// Measured on B200: +15 % at 256 streams (2.35 -> 2.05 ms/step), -3.5 % at 4096 streams (launch latency is already hidden
// behind 100+ us kernels and early-resident dependents only take SM slots), so the engine enables it per step for small
// batches only.  ASR_B200_NO_PDL=1 falls back to plain stream order everywhere.
bool pdl_enabled();
void pdl_set_active(bool on);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

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

// Packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: two fp32 lanes per instruction at the scalar issue rate).  The epilogues
// of the K = 512 GEMMs have ~16 issue slots per output element before they, not the tensor pipe, bound the kernel.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2 f2_bcast(float v) { return f2_pack(v, v); }
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// erf(x) = sign(x) * (1 - 2^q(min(|x|, 4))), q = degree-6 minimax fit of log2(erfc) with zero constant term (weights =
// d erf / d q); max abs error 3.1e-7 over the whole line in fp32 (fit + check: DESIGN.md §2).  One MUFU.EX2 + 7 FMA,
// branch-free; libdevice erff costs ~25 instructions, a rational form 2 MUFU ops.
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

// Exact-erf GELU (torch.nn.GELU() default, TA:emformer.py:42-43) on two values at once.  With u = |x| and
// e = erfc(u / sqrt 2) = 2^(u' * p(u')), u' = min(u, 4 sqrt 2) (the 1/sqrt 2 is folded into the coefficients of erf_fast's fit):
//     gelu(x) = 0.5 x (1 + sign(x)(1 - e)) = (0.5 x + 0.5 u) - 0.5 u e
// which needs no sign handling: 9 packed FMA-pipe instructions, 2 MUFU.EX2 and 4 ALU-pipe ops per PAIR (the scalar form
// is ~19 instructions per element).  Max abs error vs float64 5.8e-7 on [-12, 12] (checked in numpy with fp32 rounding).
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  const float u0 = fabsf(x0), u1 = fabsf(x1);
  const f32x2 u = f2_pack(u0, u1);
  const f32x2 a = f2_pack(fminf(u0, 5.65685425f), fminf(u1, 5.65685425f));
  f32x2 q = f2_fma(f2_bcast(1.775592500e-05f), a, f2_bcast(-6.477629067e-04f));
  q = f2_fma(q, a, f2_bcast(7.724056020e-03f));
  q = f2_fma(q, a, f2_bcast(-5.292675272e-02f));
  q = f2_fma(q, a, f2_bcast(-4.590827227e-01f));
  q = f2_fma(q, a, f2_bcast(-1.151116848e+00f));
  float t0, t1, e0, e1;
  f2_unpack(f2_mul(q, a), t0, t1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));        // exponent in [-27, 0]
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
  const f32x2 s = f2_fma(f2_bcast(0.5f), f2_pack(x0, x1), f2_mul(u, f2_bcast(0.5f)));
  f2_unpack(f2_fma(f2_mul(u, f2_bcast(-0.5f)), f2_pack(e0, e1), s), x0, x1);
}
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

template <typename C> __device__ __forceinline__ C shfl_ctx(C v, int src);
template <> __device__ __forceinline__ int shfl_ctx<int>(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <> __device__ __forceinline__ unsigned long long shfl_ctx<unsigned long long>(unsigned long long v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// ------------------------------------------------------------------------------------------
// GEMM epilogues.  tcgen05.ld hands every thread one accumulator ROW (32 columns at a time).  Storing in that shape makes
// each 16-byte store instruction touch 32 different 128-byte lines, and the K = 512 GEMMs of this model become store-bound
// at ~2 TB/s (profiles/r01_gemm_sweep_rowstore_epilogue.txt: time tracks output bytes while the same mainloop reaches
// 1.7 PFLOP/s at K = 4096).  So the kernel transposes each 32x32 block through a padded shared-memory tile with 16-byte
// accesses; afterwards lane l owns 4 consecutive columns 4*(l&7).. of the 8 rows 4*i + (l>>3): every global load/store
// instruction covers 4 rows x one full 128-byte line.  The functors get that fragment:
//     v[i][0..3]  = rows row0 + 4*i + (lane>>3),  columns col .. col+3        (col = col0 + 4*(lane&7))
//     ctx[i]      = RowCtx of those rows (computed once per tile by lane = row, gathered by shuffle)
//     b4          = bias[col .. col+3]
// ------------------------------------------------------------------------------------------
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_SILU = 2 };

// Diagnostic only (asr_debug_gemm_time epi 3): waits for the accumulator and drops it — the mainloop's own speed.
struct EpiNull {
  static constexpr bool kBf16Rows = false;
  static constexpr bool kStreamTiles = false;
  typedef int RowCtx;
  __device__ __forceinline__ RowCtx row_ctx(int, int) const { return 0; }
  __device__ __forceinline__ const float* bias_ptr() const { return nullptr; }
  __device__ __forceinline__ void prefetch_tile(int, int, int, RowCtx) const {}
  __device__ __forceinline__ void store(int, int, int, int, float (&)[8][4], const RowCtx (&)[8], const float4&) const {}
};
template <class E> struct IsNullEpi { static constexpr bool value = false; };
template <> struct IsNullEpi<EpiNull> { static constexpr bool value = true; };

// out[row, col] = acc (+bias) (+res)        fp32 out; n_valid masks a ragged N (CTC vocab = 804)
struct EpiF32 {
  static constexpr bool kBf16Rows = false;
  static constexpr bool kStreamTiles = false;
  float* out;
  const float* bias;   // nullable
  const float* res;    // nullable, same ld as out
  int ld;
  int n_valid;
  typedef int RowCtx;
  __device__ __forceinline__ RowCtx row_ctx(int, int) const { return 0; }
  __device__ __forceinline__ const float* bias_ptr() const { return bias; }
  // residual rows of this thread for the whole tile -> L2, issued while the MMAs of the tile are still running
  __device__ __forceinline__ void prefetch_tile(int row, int col0, int n_cols, RowCtx) const {
    if (res) {
      const float* r = res + (size_t)row * ld + col0;
      for (int c = 0; c < n_cols; c += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(r + c));
    }
  }
  __device__ __forceinline__ void store(int row0, int col, int lane, int M, float (&v)[8][4], const RowCtx (&)[8], const float4& b4) const {
    const int rsub = lane >> 3;
    if (col + 4 <= n_valid) {
      float4 rr[8];
      if (res) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = row0 + 4 * i + rsub;
          rr[i] = row < M ? *reinterpret_cast<const float4*>(res + (size_t)row * ld + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = row0 + 4 * i + rsub;
        float4 t = make_float4(v[i][0] + b4.x, v[i][1] + b4.y, v[i][2] + b4.z, v[i][3] + b4.w);
        if (res) { t.x += rr[i].x; t.y += rr[i].y; t.z += rr[i].z; t.w += rr[i].w; }
        if (row < M) *reinterpret_cast<float4*>(out + (size_t)row * ld + col) = t;
      }
    } else {
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = row0 + 4 * i + rsub;
        if (row < M)
          for (int j = 0; j < 4; ++j)
            if (col + j < n_valid) {
              const size_t o = (size_t)row * ld + col + j;
              out[o] = v[i][j] + bb[j] + (res ? res[o] : 0.f);
            }
      }
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
  __device__ __forceinline__ const float* bias_ptr() const { return bias; }
  __device__ __forceinline__ void prefetch_tile(int, int, int, RowCtx) const {}
  // Packed-bf16 row path (see epilogue_tile): bias + activation on the 32 values a thread holds of ITS row (columns col0 ..), then
  // the rows leave as 16-byte pieces.  Same operations in the same order as store() below: bit-identical results.
  static constexpr bool kBf16Rows = true;
  static constexpr bool kStreamTiles = false;
  __host__ __device__ __forceinline__ bool bf16_rows() const { return lo_off == 0; }
  // v: 8 values of the thread's row; ba / bb: the bias of their columns
  __device__ __forceinline__ void apply8(float* v, const float4& ba, const float4& bb, int) const {
    if (act == ACT_GELU) {
      f2_unpack(f2_add(f2_pack(v[0], v[1]), f2_pack(ba.x, ba.y)), v[0], v[1]);
      f2_unpack(f2_add(f2_pack(v[2], v[3]), f2_pack(ba.z, ba.w)), v[2], v[3]);
      f2_unpack(f2_add(f2_pack(v[4], v[5]), f2_pack(bb.x, bb.y)), v[4], v[5]);
      f2_unpack(f2_add(f2_pack(v[6], v[7]), f2_pack(bb.z, bb.w)), v[6], v[7]);
      gelu_erf2(v[0], v[1]); gelu_erf2(v[2], v[3]); gelu_erf2(v[4], v[5]); gelu_erf2(v[6], v[7]);
    } else {
      const float b8[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float x = v[k] + b8[k]; v[k] = act == ACT_SILU ? silu(x) : x; }
    }
  }
  __device__ __forceinline__ int section_col0(int) const { return 0; }
  __device__ __forceinline__ bf16* row_ptr(RowCtx, int row, int col, int) const { return out + (size_t)row * ld + col; }
  __device__ __forceinline__ void store(int row0, int col, int lane, int M, float (&v)[8][4], const RowCtx (&)[8], const float4& b4) const {
    const int rsub = lane >> 3;
    if (act == ACT_GELU) {
      const f32x2 b01 = f2_pack(b4.x, b4.y), b23 = f2_pack(b4.z, b4.w);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f2_unpack(f2_add(f2_pack(v[i][0], v[i][1]), b01), v[i][0], v[i][1]);
        f2_unpack(f2_add(f2_pack(v[i][2], v[i][3]), b23), v[i][2], v[i][3]);
        gelu_erf2(v[i][0], v[i][1]);
        gelu_erf2(v[i][2], v[i][3]);
      }
    } else {
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float x = v[i][j] + bb[j];
          v[i][j] = act == ACT_SILU ? silu(x) : x;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row0 + 4 * i + rsub;
      if (row < M) {
        bf16* o = out + (size_t)row * ld + col;
        // packed conversions (F2FP.BF16.PACK_AB); single-value cvt goes through the XU pipe (ncu: 48 % XU in this epilogue)
        const __nv_bfloat162 h01 = __floats2bfloat162_rn(v[i][0], v[i][1]), h23 = __floats2bfloat162_rn(v[i][2], v[i][3]);
        uint2 h;
        h.x = *reinterpret_cast<const uint32_t*>(&h01);
        h.y = *reinterpret_cast<const uint32_t*>(&h23);
        *reinterpret_cast<uint2*>(o) = h;
        if (lo_off) {
          const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
          uint2 l;
          l.x = pack_bf16x2(v[i][0] - f01.x, v[i][1] - f01.y);
          l.y = pack_bf16x2(v[i][2] - f23.x, v[i][3] - f23.y);
          *reinterpret_cast<uint2*>(o + lo_off) = l;
        }
      }
    }
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
  long long kv_row0 = 0;     // stream tiling: first row of this layer's slab in the [rows, d] view of the whole cache
  int slot_rows = 0;         //                rows per session inside the slab (2 * ring)
  typedef unsigned long long RowCtx;          // T* of column 0 of this tile's section (q | k | v) in the destination row
  __device__ __forceinline__ RowCtx row_ctx(int row, int tile_col0) const {
    const int sec = tile_col0 / d;            // 0: q, 1: k, 2: v
    if (sec == 0) return (RowCtx)(q + (size_t)row * d);
    const int which = sec - 1;
    const int b = row / rows, t = row - b * rows;
    if (t < seg_rows) {
      const int slot = slots[b];
      const int rr = (past_len[slot] + t) % ring;
      return (RowCtx)(cache_layer + (size_t)slot * slot_stride + ((size_t)which * ring + rr) * d);
    }
    return (RowCtx)(rc + (((size_t)b * 2 + which) * rc_rows + (t - seg_rows)) * d);
  }
  __device__ __forceinline__ const float* bias_ptr() const { return bias; }
  __device__ __forceinline__ void prefetch_tile(int, int, int, RowCtx) const {}
  // Packed-bf16 row path (bf16 K/V cache and q only): same arithmetic as store(), (acc + bias) * scale.
  static constexpr bool kBf16Rows = sizeof(T) == 2;
  static constexpr bool kStreamTiles = sizeof(T) == 2;       // gemm.cuh kPairTileQKV: M tiles cut along streams, every destination a TMA box
  __host__ __device__ __forceinline__ bool bf16_rows() const { return true; }
  __device__ __forceinline__ void apply8(float* v, const float4& ba, const float4& bb, int tile_col0) const {
    const float sc = tile_col0 < d ? qscale : 1.0f;          // a tile never straddles the q | k | v sections
    v[0] = (v[0] + ba.x) * sc; v[1] = (v[1] + ba.y) * sc; v[2] = (v[2] + ba.z) * sc; v[3] = (v[3] + ba.w) * sc;
    v[4] = (v[4] + bb.x) * sc; v[5] = (v[5] + bb.y) * sc; v[6] = (v[6] + bb.z) * sc; v[7] = (v[7] + bb.w) * sc;
  }
  __device__ __forceinline__ int section_col0(int tile_col0) const { return tile_col0 / d * d; }      // first column of the tile's q | k | v section
  __device__ __forceinline__ bf16* row_ptr(RowCtx ctx, int, int col, int sec_col0) const { return reinterpret_cast<bf16*>(ctx) + (col - sec_col0); }
  __device__ __forceinline__ void store(int row0, int col, int lane, int M, float (&v)[8][4], const RowCtx (&ctx)[8], const float4& b4) const {
    const int rsub = lane >> 3;
    const int sec = col / d;
    const float sc = sec == 0 ? qscale : 1.0f;
    const int lc = col - sec * d;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row0 + 4 * i + rsub;
      const float x0 = (v[i][0] + b4.x) * sc, x1 = (v[i][1] + b4.y) * sc, x2 = (v[i][2] + b4.z) * sc, x3 = (v[i][3] + b4.w) * sc;
      if (row < M) {
        T* dst = reinterpret_cast<T*>(ctx[i]) + lc;
        if (sizeof(T) == 4) {
          if (sec == 0) {
            *reinterpret_cast<float4*>(dst) = make_float4(x0, x1, x2, x3);
          } else {
            // EXACT K/V rows are stored pre-split, [hi d bf16 | lo d bf16] in the 4d bytes of the row: x = hi + lo to 2^-17, which is
            // what the three-MMA attention (layers.cu: attention_exact_kernel) consumes without any conversion
            const __nv_bfloat162 h01 = __floats2bfloat162_rn(x0, x1), h23 = __floats2bfloat162_rn(x2, x3);
            const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
            bf16* sp = reinterpret_cast<bf16*>(ctx[i]) + lc;
            *reinterpret_cast<uint2*>(sp) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
            *reinterpret_cast<uint2*>(sp + d) = make_uint2(pack_bf16x2(x0 - f01.x, x1 - f01.y), pack_bf16x2(x2 - f23.x, x3 - f23.y));
          }
        } else *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(x0, x1), pack_bf16x2(x2, x3));
      }
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
