// Memory-bound fused kernels of the Emformer layer and the CTC decode.
#include <stdlib.h>

#include "kernels.cuh"

namespace asr {

namespace {

// ------------------------------------------------------------------------------------------
// LayerNorm(512): one warp per row, 16 values per lane held in registers, two-pass variance (matches
// torch.nn.LayerNorm: biased variance, eps = 1e-5), fp32 statistics.
// ------------------------------------------------------------------------------------------
constexpr int LN_D = 512;
constexpr int LN_V = LN_D / 128;   // float4 per lane

__device__ __forceinline__ void ln_load(const float* __restrict__ x, int lane, float (&v)[LN_V * 4]) {
#pragma unroll
  for (int i = 0; i < LN_V; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(x + i * 128 + lane * 4);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}

__device__ __forceinline__ void ln_apply(float (&v)[LN_V * 4], const float* __restrict__ g, const float* __restrict__ b, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_V * 4; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / LN_D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_V * 4; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
  const float var = warp_sum(q) * (1.0f / LN_D);
  const float rstd = 1.0f / sqrtf(var + 1e-5f);
#pragma unroll
  for (int i = 0; i < LN_V; ++i) {
    const float4 gg = *reinterpret_cast<const float4*>(g + i * 128 + lane * 4);
    const float4 bb = *reinterpret_cast<const float4*>(b + i * 128 + lane * 4);
    v[4 * i] = (v[4 * i] - mean) * rstd * gg.x + bb.x;
    v[4 * i + 1] = (v[4 * i + 1] - mean) * rstd * gg.y + bb.y;
    v[4 * i + 2] = (v[4 * i + 2] - mean) * rstd * gg.z + bb.z;
    v[4 * i + 3] = (v[4 * i + 3] - mean) * rstd * gg.w + bb.w;
  }
}

__device__ __forceinline__ void ln_store_operand(bf16* __restrict__ o, int lo_off, int lane, const float (&v)[LN_V * 4]) {
#pragma unroll
  for (int i = 0; i < LN_V; ++i) {
    const int c = i * 128 + lane * 4;
    const bf16 h0 = __float2bfloat16_rn(v[4 * i]), h1 = __float2bfloat16_rn(v[4 * i + 1]);
    const bf16 h2 = __float2bfloat16_rn(v[4 * i + 2]), h3 = __float2bfloat16_rn(v[4 * i + 3]);
    uint2 h;
    h.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    h.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
    *reinterpret_cast<uint2*>(o + c) = h;
    if (lo_off) {
      uint2 l;
      l.x = pack_bf16x2(v[4 * i] - __bfloat162float(h0), v[4 * i + 1] - __bfloat162float(h1));
      l.y = pack_bf16x2(v[4 * i + 2] - __bfloat162float(h2), v[4 * i + 3] - __bfloat162float(h3));
      *reinterpret_cast<uint2*>(o + lo_off + c) = l;
    }
  }
}

__global__ void __launch_bounds__(256) ln_to_operand_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                                            bf16* __restrict__ out, int ld, int lo_off, int M) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();
  if (row >= M) return;
  float v[LN_V * 4];
  ln_load(x + (size_t)row * LN_D, lane, v);
  ln_apply(v, g, b, lane);
  ln_store_operand(out + (size_t)row * ld, lo_off, lane, v);
}

__global__ void __launch_bounds__(256) ln_out_fused_kernel(const float* __restrict__ x2, const float* __restrict__ g1, const float* __restrict__ b1,
                                                           float* __restrict__ y, const float* __restrict__ g2, const float* __restrict__ b2,
                                                           bf16* __restrict__ out, int ld, int lo_off, int M, int rows, int seg_rows) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();
  if (row >= M) return;
  float v[LN_V * 4];
  ln_load(x2 + (size_t)row * LN_D, lane, v);
  ln_apply(v, g1, b1, lane);
#pragma unroll
  for (int i = 0; i < LN_V; ++i)
    *reinterpret_cast<float4*>(y + (size_t)row * LN_D + i * 128 + lane * 4) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  if (g2) {
    ln_apply(v, g2, b2, lane);
    ln_store_operand(out + (size_t)row * ld, lo_off, lane, v);
  } else {
    const int b = row / rows, t = row - b * rows;
    if (t < seg_rows) ln_store_operand(out + ((size_t)b * seg_rows + t) * ld, lo_off, lane, v);
  }
}

// ------------------------------------------------------------------------------------------
// Chunk attention (TA:emformer.py:184-204, :127-144) for one (stream, head) per warp.
// Keys = [valid left context from the ring (oldest first), this chunk's segment rows (already in the ring,
// written by the QKV epilogue), right-context rows (scratch)]; softmax in fp32 over exactly
// lv + seg + rc keys — the reference slices the cache, it never attends to zero padding (:391-398).
// ------------------------------------------------------------------------------------------
constexpr int AT_DH = 64;
constexpr int AT_MAXK = 64;
constexpr int AT_KST = 68;     // padded fp32 row stride of the staged K / V tile (conflict-free LDS.128)
constexpr int AT_WARPS = 4;

template <typename T>
__device__ __forceinline__ void stage_rows(const AttnParams<T>& P, const T* cache_slot, const T* rc_b, int which, int head, int lv, int pl,
                                           int n_keys, float* __restrict__ dst, int lane) {
  constexpr int VEC = 16 / (int)sizeof(T);              // elements per 16-byte load
  constexpr int PER_ROW = AT_DH / VEC;
  for (int i = lane; i < n_keys * PER_ROW; i += 32) {
    const int j = i / PER_ROW, c = (i - j * PER_ROW) * VEC;
    const T* src;
    if (j < lv + P.seg_rows) {
      const int rr = (pl - lv + j + P.ring) % P.ring;     // left rows: pl-lv+j ; segment rows: pl + (j-lv)
      src = cache_slot + ((size_t)which * P.ring + rr) * P.d;
    } else {
      src = rc_b + ((size_t)which * P.rc_rows + (j - lv - P.seg_rows)) * P.d;
    }
    float* o = dst + j * AT_KST + c;
    if (sizeof(T) == 4) {
      // EXACT ring rows are pre-split [hi d bf16 | lo d bf16] (see EpiQKV::store): x = hi + lo
      const bf16* sp = reinterpret_cast<const bf16*>(src) + head * AT_DH + c;
      const uint2 hi = *reinterpret_cast<const uint2*>(sp), lo = *reinterpret_cast<const uint2*>(sp + P.d);
      const float2 h0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi.x)), h1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi.y));
      const float2 l0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lo.x)), l1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lo.y));
      *reinterpret_cast<float4*>(o) = make_float4(h0.x + l0.x, h0.y + l0.y, h1.x + l1.x, h1.y + l1.y);
    } else {
      const int4 raw = *reinterpret_cast<const int4*>(src + head * AT_DH + c);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h[e]); o[2 * e] = f.x; o[2 * e + 1] = f.y; }
    }
  }
}

template <typename T, int ROWS>
__global__ void __launch_bounds__(AT_WARPS * 32) attention_kernel(AttnParams<T> P) {
  extern __shared__ __align__(16) float at_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x, head = blockIdx.y * AT_WARPS + warp;
  float* s_q = at_smem + warp * (ROWS * AT_DH + AT_MAXK * AT_KST + AT_MAXK * ROWS);
  float* s_kv = s_q + ROWS * AT_DH;
  float* s_p = s_kv + AT_MAXK * AT_KST;                   // pT[key][ROWS]
  pdl_launch_dependents();
  pdl_wait();

  const int slot = P.slots[b];
  const int pl = P.past_len[slot];
  const int lv = pl < P.left ? pl : P.left;
  const int n_keys = lv + P.seg_rows + P.rc_rows;
  const T* cache_slot = P.cache_layer + (size_t)slot * P.slot_stride;
  const T* rc_b = P.rc + (size_t)b * 2 * P.rc_rows * P.d;

  for (int i = lane; i < ROWS * (AT_DH / 4); i += 32) {
    const int r = i / (AT_DH / 4), c = (i - r * (AT_DH / 4)) * 4;
    const T* qs = P.q + ((size_t)b * ROWS + r) * P.d + head * AT_DH + c;
    *reinterpret_cast<float4*>(s_q + r * AT_DH + c) = make_float4(to_f32<T>(qs[0]), to_f32<T>(qs[1]), to_f32<T>(qs[2]), to_f32<T>(qs[3]));
  }
  stage_rows<T>(P, cache_slot, rc_b, 0, head, lv, pl, n_keys, s_kv, lane);
  __syncwarp();

  // ---- scores: lane owns keys (lane, lane + 32), all ROWS queries
  float s0[ROWS], s1[ROWS];
#pragma unroll
  for (int i = 0; i < ROWS; ++i) { s0[i] = 0.f; s1[i] = 0.f; }
  const float* k0 = s_kv + lane * AT_KST;
  const float* k1 = s_kv + (lane + 32) * AT_KST;
  const bool v0 = lane < n_keys, v1 = lane + 32 < n_keys;
#pragma unroll 2
  for (int c = 0; c < AT_DH; c += 4) {
    const float4 ka = v0 ? *reinterpret_cast<const float4*>(k0 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 kb = v1 ? *reinterpret_cast<const float4*>(k1 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      const float4 qq = *reinterpret_cast<const float4*>(s_q + i * AT_DH + c);
      s0[i] = fmaf(qq.x, ka.x, fmaf(qq.y, ka.y, fmaf(qq.z, ka.z, fmaf(qq.w, ka.w, s0[i]))));
      s1[i] = fmaf(qq.x, kb.x, fmaf(qq.y, kb.y, fmaf(qq.z, kb.z, fmaf(qq.w, kb.w, s1[i]))));
    }
  }
  // ---- softmax over keys (fp32), probabilities to smem transposed
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    const float a = v0 ? s0[i] : -INFINITY, c = v1 ? s1[i] : -INFINITY;
    const float m = warp_max(fmaxf(a, c));
    const float e0 = v0 ? expf(a - m) : 0.f, e1 = v1 ? expf(c - m) : 0.f;
    const float inv = 1.0f / warp_sum(e0 + e1);
    s_p[lane * ROWS + i] = e0 * inv;
    s_p[(lane + 32) * ROWS + i] = e1 * inv;
  }
  __syncwarp();
  stage_rows<T>(P, cache_slot, rc_b, 1, head, lv, pl, n_keys, s_kv, lane);   // V over the K tile
  __syncwarp();
  // ---- out[i][d] = sum_j p[i][j] * v[j][d]; lane owns d = lane, lane + 32
  float o0[ROWS], o1[ROWS];
#pragma unroll
  for (int i = 0; i < ROWS; ++i) { o0[i] = 0.f; o1[i] = 0.f; }
  for (int j = 0; j < n_keys; ++j) {
    const float va = s_kv[j * AT_KST + lane], vb = s_kv[j * AT_KST + lane + 32];
#pragma unroll
    for (int i = 0; i < ROWS; i += 4) {
      const float4 pp = *reinterpret_cast<const float4*>(s_p + j * ROWS + i);
      o0[i] = fmaf(pp.x, va, o0[i]); o1[i] = fmaf(pp.x, vb, o1[i]);
      o0[i + 1] = fmaf(pp.y, va, o0[i + 1]); o1[i + 1] = fmaf(pp.y, vb, o1[i + 1]);
      o0[i + 2] = fmaf(pp.z, va, o0[i + 2]); o1[i + 2] = fmaf(pp.z, vb, o1[i + 2]);
      o0[i + 3] = fmaf(pp.w, va, o0[i + 3]); o1[i + 3] = fmaf(pp.w, vb, o1[i + 3]);
    }
  }
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    bf16* o = P.out + ((size_t)b * ROWS + i) * P.ld + head * AT_DH;
    const bf16 h0 = __float2bfloat16_rn(o0[i]), h1 = __float2bfloat16_rn(o1[i]);
    o[lane] = h0; o[lane + 32] = h1;
    if (P.lo_off) {
      o[P.lo_off + lane] = __float2bfloat16_rn(o0[i] - __bfloat162float(h0));
      o[P.lo_off + lane + 32] = __float2bfloat16_rn(o1[i] - __bfloat162float(h1));
    }
  }
}


// ------------------------------------------------------------------------------------------
// FAST-mode chunk attention on the tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate).
// One CTA per stream, one warp per head.  The whole K and V row set of the stream (valid left context +
// segment rows from the ring, right-context rows from scratch; full 512-wide rows = 1 KB each) is brought into
// shared memory with 16-byte cp.async in one go, so the ~105 KB per stream is in flight at once and each byte of
// the K/V ring is read from HBM exactly once per layer.  Rows are padded to 1040 B: bank = (4*key + word) % 32,
// conflict-free for the B-fragment loads of QK^T and for ldmatrix.trans of V.  S = QK^T lives in registers
// (2 m-tiles x 8 n-tiles), softmax in fp32 on the fragments, P re-used as the A fragments of P*V.
// ------------------------------------------------------------------------------------------
constexpr int AM_ROWB = 1040;                 // bytes per staged K/V row (512 bf16 + 16 B pad)
constexpr int AM_WARPS = 8;

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int ROWS>
__global__ void __launch_bounds__(AM_WARPS * 32, 2) attention_mma_kernel(AttnParams<bf16> P) {
  extern __shared__ __align__(16) uint8_t am_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int b = blockIdx.x, head = warp;
  pdl_launch_dependents();
  pdl_wait();
  const int slot = P.slots[b];
  const int pl = P.past_len[slot];
  const int lv = pl < P.left ? pl : P.left;
  const int n_keys = lv + P.seg_rows + P.rc_rows;            // <= 64
  const int kmax = P.left + P.seg_rows + P.rc_rows;
  uint8_t* s_k = am_smem;                                     // [kmax][1040]
  uint8_t* s_v = s_k + (size_t)kmax * AM_ROWB;                // [kmax][1040]
  uint8_t* s_zero = s_v + (size_t)kmax * AM_ROWB;             // one zero row for keys >= n_keys
  const bf16* cache_slot = P.cache_layer + (size_t)slot * P.slot_stride;
  const bf16* rc_b = P.rc + (size_t)b * 2 * P.rc_rows * P.d;

  // ---- stage K and V rows: n_keys rows x 64 chunks of 16 B each, for K then V
  const int chunks = n_keys * 64;
  for (int i = tid; i < 2 * chunks; i += AM_WARPS * 32) {
    const int which = i >= chunks;
    const int r = (i - which * chunks) >> 6, c = (i & 63);
    const bf16* src;
    if (r < lv + P.seg_rows) {
      const int rr = (pl - lv + r + P.ring) % P.ring;
      src = cache_slot + ((size_t)which * P.ring + rr) * P.d;
    } else {
      src = rc_b + ((size_t)which * P.rc_rows + (r - lv - P.seg_rows)) * P.d;
    }
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared((which ? s_v : s_k) + (size_t)r * AM_ROWB + c * 16);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + c * 8) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = tid; i < AM_ROWB / 16; i += AM_WARPS * 32) reinterpret_cast<uint4*>(s_zero)[i] = make_uint4(0, 0, 0, 0);

  // ---- Q fragments straight from global (bf16, already scaled): rows >= ROWS are zero padding
  uint32_t qa[2][4][4];                                       // [m-tile][k-step][reg]
  const bf16* qb = P.q + (size_t)b * ROWS * P.d + head * AT_DH;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int r0 = mt * 16 + g, r1 = r0 + 8, c0 = ks * 16 + 2 * tig;
      qa[mt][ks][0] = r0 < ROWS ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r0 * P.d + c0) : 0u;
      qa[mt][ks][1] = r1 < ROWS ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r1 * P.d + c0) : 0u;
      qa[mt][ks][2] = r0 < ROWS ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r0 * P.d + c0 + 8) : 0u;
      qa[mt][ks][3] = r1 < ROWS ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r1 * P.d + c0 + 8) : 0u;
    }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- S = Q K^T
  float sc[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) sc[mt][nt][e] = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (nt * 8 < n_keys) {                                     // warp-uniform
      int key = nt * 8 + g;
      key = key < n_keys ? key : n_keys - 1;                    // clamp: garbage columns are masked below
      const uint8_t* kr = s_k + (size_t)key * AM_ROWB + head * (AT_DH * 2) + tig * 4;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr + ks * 32);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kr + ks * 32 + 16);
        mma_bf16_16816(sc[0][nt], qa[0][ks], b0, b1);
        mma_bf16_16816(sc[1][nt], qa[1][ks], b0, b1);
      }
    }
  }
  // ---- softmax over keys, fp32, per query row (rows g and g+8 of each m-tile); 4 lanes share a row
  uint32_t pa[2][4][4];                                        // P as A fragments: [m-tile][k-step of 16 keys][reg]
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {                           // hr = 0: row g (regs 0,1), hr = 1: row g+8 (regs 2,3)
      float m = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const bool valid = nt * 8 + 2 * tig + e < n_keys;
          float& x = sc[mt][nt][2 * hr + e];
          x = valid ? x : -INFINITY;
          m = fmaxf(m, x);
        }
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      float sum = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float& x = sc[mt][nt][2 * hr + e];
          x = __expf(x - m);                                   // exp(-inf) = 0 for masked keys
          sum += x;
        }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = 1.0f / sum;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        sc[mt][nt][2 * hr] *= inv;
        sc[mt][nt][2 * hr + 1] *= inv;
      }
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      pa[mt][kk][0] = pack_bf16x2(sc[mt][2 * kk][0], sc[mt][2 * kk][1]);
      pa[mt][kk][1] = pack_bf16x2(sc[mt][2 * kk][2], sc[mt][2 * kk][3]);
      pa[mt][kk][2] = pack_bf16x2(sc[mt][2 * kk + 1][0], sc[mt][2 * kk + 1][1]);
      pa[mt][kk][3] = pack_bf16x2(sc[mt][2 * kk + 1][2], sc[mt][2 * kk + 1][3]);
    }
  }
  // ---- O = P V : B fragments of V via ldmatrix.x4.trans (two 8-dim n-tiles per instruction)
  float oc[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int dn = 0; dn < 8; ++dn)
#pragma unroll
      for (int e = 0; e < 4; ++e) oc[mt][dn][e] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    if (kk * 16 < n_keys) {                                     // warp-uniform
      // lane -> (matrix id = lane/8, row in matrix = lane%8): matrices 0,1 = keys +0..7, +8..15 at dims dn; 2,3 = same keys at dims dn+1
      const int mi = lane >> 3, ri = lane & 7;
      const int key = kk * 16 + (mi & 1) * 8 + ri;
      const uint8_t* row = key < n_keys ? s_v + (size_t)key * AM_ROWB : s_zero - head * (AT_DH * 2);   // zero row for padded keys
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {                          // pairs of 8-dim tiles
        const uint8_t* src = (key < n_keys ? row + head * (AT_DH * 2) : s_zero) + (key < n_keys ? (dp * 2 + (mi >> 1)) * 16 : 0);
        const uint32_t addr = (uint32_t)__cvta_generic_to_shared(src);
        uint32_t v0, v1, v2, v3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(addr));
        mma_bf16_16816(oc[0][2 * dp], pa[0][kk], v0, v1);
        mma_bf16_16816(oc[1][2 * dp], pa[1][kk], v0, v1);
        mma_bf16_16816(oc[0][2 * dp + 1], pa[0][kk], v2, v3);
        mma_bf16_16816(oc[1][2 * dp + 1], pa[1][kk], v2, v3);
      }
    }
  }
  // ---- write the A operand of out_proj (bf16; FAST mode has no lo half)
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int r = mt * 16 + hr * 8 + g;
      if (r < ROWS) {
        bf16* o = P.out + ((size_t)b * ROWS + r) * P.ld + head * AT_DH + 2 * tig;
#pragma unroll
        for (int dn = 0; dn < 8; ++dn)
          *reinterpret_cast<uint32_t*>(o + dn * 8) = pack_bf16x2(oc[mt][dn][2 * hr], oc[mt][dn][2 * hr + 1]);
      }
    }
}

// ------------------------------------------------------------------------------------------
// Streaming variant (FAST precision, from 148 streams per step on): one persistent CTA per SM walks over its streams with the
// K/V staging double-buffered and fed by TMA.  The ring advances by seg_rows = 16 rows per chunk, so a stream's cached keys are
// 1..3 whole 16-row blocks (16 KB contiguous each) plus the 4 right-context rows: at most 4 TMA ops per K (and per V) through
// a 3D (dim, row, head) tensor map whose box lands in shared memory head-major with the 128B swizzle — the B-fragment loads of
// Q K^T and the ldmatrix.trans of V are bank-conflict-free without padding.  (A first version issued one 1 KB cp.async.bulk
// per row: the load pipe alone then ran at 3.0 TB/s, 2880 copies per SM per launch at ~95 clk each.)  K and V halves have
// their own full / empty mbarriers: K is released right after Q K^T so the K rows of the stream after next are already in
// flight during softmax and P V.  16 head warps = (head, 16-row query tile): per-warp instruction latency, not issue rate,
// bounds the math (ncu: 34 % issue-active with 8 warps).  Same fragments and summation order as attention_mma_kernel.
// ------------------------------------------------------------------------------------------
template <int ROWS> struct AsCfg {
  static constexpr int MT = (ROWS + 15) / 16;                 // 16-row query tiles per stream (20 rows -> 2, 12 rows -> 1)
  static constexpr int HEAD_WARPS = AM_WARPS * MT;
  static constexpr int THREADS = (HEAD_WARPS + 1) * 32;
};
constexpr int AS_RC_ROWS = 4;                                  // right-context rows per chunk (context_size 16 / stride 4)
constexpr int AS_RING_ROWS = 48;                               // upper bound of left context + segment rows staged per stream
constexpr int AS_RC_OFF = AS_RING_ROWS * 1024;                 // right-context block [head][4 rows][128 B] behind the ring blocks
constexpr int AS_HALF_BYTES = AS_RC_OFF + AS_RC_ROWS * 1024;   // K (or V) half of a buffer: 52 KB

__device__ __forceinline__ void as_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void as_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void as_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void as_mbar_wait(uint32_t bar, uint32_t parity) {      // bounded: a byte-count bug traps instead of hanging the box
  uint32_t done = 0;
  long long t0 = 0;
  for (int it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (it == 64) t0 = clock64();
    if (it > 64 && (it & 1023) == 0 && clock64() - t0 > 4000000000LL) {
      printf("asr attention: mbarrier wait timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void as_tma_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// byte offset, inside a K (or V) half buffer, of the 128-byte row of (key, head); the first n_ring keys live in SEG-row blocks
// ([block][head][SEG rows][128 B], one TMA box each), the right-context keys in the block behind them
template <int SEG>
__device__ __forceinline__ uint32_t as_row_off(int key, int n_ring, int head) {
  return key < n_ring ? (uint32_t)((key / SEG) * (SEG * 1024) + head * (SEG * 128) + (key % SEG) * 128)
                      : (uint32_t)(AS_RC_OFF + head * (AS_RC_ROWS * 128) + (key - n_ring) * 128);
}
// 128B swizzle: the 16-byte chunk index is XORed with bits [7,10) of the (1024-aligned) shared-memory offset
__device__ __forceinline__ uint32_t as_swz(uint32_t row_off, int chunk) { return row_off + ((uint32_t)(chunk ^ ((row_off >> 7) & 7)) << 4); }

// A fragments of one 16-row query tile (rows row_base + g, + 8), zero for rows >= ROWS
template <int ROWS>
__device__ __forceinline__ void as_load_q(uint32_t (&qa)[4][4], const bf16* qb, int d, int row_base, int g, int tig) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int r0 = row_base + g, r1 = r0 + 8, c0 = ks * 16 + 2 * tig;
    qa[ks][0] = r0 < ROWS ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r0 * d + c0) : 0u;
    qa[ks][1] = r1 < ROWS ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r1 * d + c0) : 0u;
    qa[ks][2] = r0 < ROWS ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r0 * d + c0 + 8) : 0u;
    qa[ks][3] = r1 < ROWS ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r1 * d + c0 + 8) : 0u;
  }
}

template <int ROWS>
__global__ void __launch_bounds__(AsCfg<ROWS>::THREADS, 1)
attention_stream_kernel(const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmRC, AttnParams<bf16> P, int n_streams) {
  constexpr int HEAD_WARPS = AsCfg<ROWS>::HEAD_WARPS;
  constexpr int SEG = ROWS - AS_RC_ROWS;                        // segment rows per chunk = rows per ring block (16, or 8 in low-latency mode)
  constexpr int BLOCK_BYTES = SEG * 1024;
  extern __shared__ uint8_t as_smem_raw[];
  __shared__ __align__(8) unsigned long long as_bars[8];       // kfull[2] | vfull[2] | kempty[2] | vempty[2]
  __shared__ int as_meta[2];                                   // per buffer: valid left-context rows of its stream
  __shared__ __align__(16) uint8_t s_zero[128];                // a zero row for keys >= n_keys in P V
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tig = lane & 3;
  const uint32_t smem0 = ((uint32_t)__cvta_generic_to_shared(as_smem_raw) + 1023u) & ~1023u;    // swizzled TMA boxes need 1024 B alignment
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(as_bars);
  auto kfull = [&](int b) { return bar0 + 8u * b; };
  auto vfull = [&](int b) { return bar0 + 16u + 8u * b; };
  auto kempty = [&](int b) { return bar0 + 32u + 8u * b; };
  auto vempty = [&](int b) { return bar0 + 48u + 8u * b; };
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmKV) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmRC) : "memory");
    for (int b = 0; b < 2; ++b) {
      as_mbar_init(kfull(b), 1); as_mbar_init(vfull(b), 1); as_mbar_init(kempty(b), HEAD_WARPS); as_mbar_init(vempty(b), HEAD_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) reinterpret_cast<uint32_t*>(s_zero)[tid] = 0u;
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();

  if (warp == HEAD_WARPS) {
    // ===================== producer: one lane, <= 8 TMA ops per stream =====================
    if (lane == 0) {
      int it = 0;
      for (int b = blockIdx.x; b < n_streams; b += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t par = (uint32_t)((it >> 1) & 1);
        const int slot = P.slots[b];
        const int pl = P.past_len[slot];
        const int lv = pl < P.left ? pl : P.left;
        const int nb = lv / SEG + 1;                              // ring blocks holding [valid left context | segment]
        const int first_blk = (pl - lv) / SEG;
        const int ring_blocks = P.ring / SEG;
        const uint32_t bytes = (uint32_t)(nb * BLOCK_BYTES + AS_RC_ROWS * 1024);
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {                 // K first: the head warps start on it while V is in flight
          as_mbar_wait(which ? vempty(buf) : kempty(buf), par ^ 1u);   // the head warps are done with this half (first pass: free)
          if (!which) as_meta[buf] = lv;
          const uint32_t full = which ? vfull(buf) : kfull(buf);
          as_mbar_expect_tx(full, bytes);
          const uint32_t dst = smem0 + (uint32_t)((buf * 2 + which) * AS_HALF_BYTES);
          const long long row0 = P.cache_row0 + (long long)slot * P.slot_rows + (long long)which * P.ring;
          for (int j = 0; j < nb; ++j)
            as_tma_3d(dst + (uint32_t)(j * BLOCK_BYTES), &tmKV, 0, (int)(row0 + ((first_blk + j) % ring_blocks) * SEG), 0, full);
          as_tma_3d(dst + (uint32_t)AS_RC_OFF, &tmRC, 0, (b * 2 + which) * AS_RC_ROWS, 0, full);
        }
      }
    }
    return;
  }

  // ===================== head warps: warp = (query tile, head) =====================
  const int head = warp % AM_WARPS;
  const int row_base = (warp / AM_WARPS) * 16;
  const bool second_half = row_base + 8 < ROWS;                 // rows row_base + 8 + g exist at all (warp-uniform)
  uint32_t qa[4][4];
  if (blockIdx.x < n_streams) as_load_q<ROWS>(qa, P.q + (size_t)blockIdx.x * ROWS * P.d + head * AT_DH, P.d, row_base, g, tig);
  const uint32_t zero_addr = (uint32_t)__cvta_generic_to_shared(s_zero);
  int it = 0;
  for (int b = blockIdx.x; b < n_streams; b += gridDim.x, ++it) {
    const int buf = it & 1;
    const uint32_t par = (uint32_t)((it >> 1) & 1);
    const uint32_t s_k = smem0 + (uint32_t)(buf * 2 * AS_HALF_BYTES);
    const uint32_t s_v = s_k + (uint32_t)AS_HALF_BYTES;
    as_mbar_wait(kfull(buf), par);
    const int lv = as_meta[buf];
    const int n_ring = lv + SEG;
    const int n_keys = n_ring + AS_RC_ROWS;

    // ---- S = Q K^T
    float sc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) sc[nt][e] = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt * 8 < n_keys) {                                     // warp-uniform
        // K fragments by ldmatrix.x4: lane -> (matrix lane / 8, row lane % 8); the four matrices of one instruction are the b0 | b1 pairs of two
        // 16-dim steps (dims [32 hf, 32 hf + 32) of key nt * 8 + row) — 2 instructions and 2 address computations per 8 keys instead of 8 LDS.32
        int key = nt * 8 + (lane & 7);
        key = key < n_keys ? key : n_keys - 1;                    // clamp: garbage columns are masked below
        const uint32_t ro = as_row_off<SEG>(key, n_ring, head);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t k0, k1, k2, k3;
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(k0), "=r"(k1), "=r"(k2), "=r"(k3) : "r"(s_k + as_swz(ro, 4 * hf + (lane >> 3))));
          mma_bf16_16816(sc[nt], qa[2 * hf], k0, k1);
          mma_bf16_16816(sc[nt], qa[2 * hf + 1], k2, k3);
        }
      }
    }
    __syncwarp();
    if (lane == 0) as_mbar_arrive(kempty(buf));                 // this warp no longer reads the K half
    // the Q fragments are dead now: fetch the next stream's while softmax and P V run
    if (b + (int)gridDim.x < n_streams) as_load_q<ROWS>(qa, P.q + (size_t)(b + gridDim.x) * ROWS * P.d + head * AT_DH, P.d, row_base, g, tig);

    // ---- softmax over keys, fp32, per query row (rows g and g+8 of the tile); 4 lanes share a row
    uint32_t pa[4][4];
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      if (hr == 1 && !second_half) {                             // padding rows only: their probabilities are never used
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { sc[nt][2] = 0.f; sc[nt][3] = 0.f; }
        continue;
      }
      float m = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const bool valid = nt * 8 + 2 * tig + e < n_keys;
          float& x = sc[nt][2 * hr + e];
          x = valid ? x : -INFINITY;
          m = fmaxf(m, x);
        }
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      float sum = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float& x = sc[nt][2 * hr + e];
          x = __expf(x - m);                                     // exp(-inf) = 0 for masked keys
          sum += x;
        }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = 1.0f / sum;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        sc[nt][2 * hr] *= inv;
        sc[nt][2 * hr + 1] *= inv;
      }
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      pa[kk][0] = pack_bf16x2(sc[2 * kk][0], sc[2 * kk][1]);
      pa[kk][1] = pack_bf16x2(sc[2 * kk][2], sc[2 * kk][3]);
      pa[kk][2] = pack_bf16x2(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
      pa[kk][3] = pack_bf16x2(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
    }
    // ---- O = P V : B fragments of V via ldmatrix.x4.trans (two 8-dim n-tiles per instruction)
    as_mbar_wait(vfull(buf), par);
    float oc[8][4];
#pragma unroll
    for (int dn = 0; dn < 8; ++dn)
#pragma unroll
      for (int e = 0; e < 4; ++e) oc[dn][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (kk * 16 < n_keys) {
        // lane -> (matrix id = lane/8, row in matrix = lane%8): matrices 0,1 = keys +0..7, +8..15 at dims dn; 2,3 = same keys at dims dn+1
        const int mi = lane >> 3, ri = lane & 7;
        const int key = kk * 16 + (mi & 1) * 8 + ri;
        const uint32_t ro = as_row_off<SEG>(key < n_keys ? key : 0, n_ring, head);
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          const uint32_t addr = key < n_keys ? s_v + as_swz(ro, dp * 2 + (mi >> 1)) : zero_addr;
          uint32_t v0, v1, v2, v3;
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(addr));
          mma_bf16_16816(oc[2 * dp], pa[kk], v0, v1);
          mma_bf16_16816(oc[2 * dp + 1], pa[kk], v2, v3);
        }
      }
    }
    __syncwarp();
    if (lane == 0) as_mbar_arrive(vempty(buf));                 // this warp no longer reads the V half
    // ---- write the A operand of out_proj
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int r = row_base + hr * 8 + g;
      if (r < ROWS) {
        bf16* o = P.out + ((size_t)b * ROWS + r) * P.ld + head * AT_DH + 2 * tig;
#pragma unroll
        for (int dn = 0; dn < 8; ++dn)
          *reinterpret_cast<uint32_t*>(o + dn * 8) = pack_bf16x2(oc[dn][2 * hr], oc[dn][2 * hr + 1]);
      }
    }
  }
}

template <int ROWS>
int attention_stream_launch(const AttnParams<bf16>& P, int n_streams, int num_sms, cudaStream_t st) {
  const size_t smem = (size_t)4 * AS_HALF_BYTES + 1024;
  static size_t attr_done[kMaxDevices] = {0};
  ASR_CUDA_OK(ensure_dyn_smem(attention_stream_kernel<ROWS>, smem, attr_done));
  const int grid = n_streams < num_sms ? n_streams : num_sms;
  ASR_CUDA_OK(launch_pdl(attention_stream_kernel<ROWS>, dim3(grid), dim3(AsCfg<ROWS>::THREADS), smem, st, *P.h_tm_cache, *P.h_tm_rc, P, n_streams));
  return 0;
}

// ------------------------------------------------------------------------------------------
// EXACT-precision chunk attention on the tensor cores.  In EXACT mode the K/V ring rows are stored PRE-SPLIT by the QKV epilogue:
// [hi 512 bf16 | lo 512 bf16] with x = hi + lo (hi = bf16(x), lo = bf16(x - hi)) — the same 4 bytes per element as fp32, relative
// error 2^-17.  Seen through the head-major tensor map a row is 16 pieces of 128 B (pieces 0-7: hi of head p, 8-15: lo of head
// p - 8), so this kernel is attention_stream_kernel with three MMAs per fragment pair (hi*hi + lo*hi + hi*lo, fp32 accumulation —
// the EXACT GEMMs' scheme), fragments read with the same LDS.32 / ldmatrix.trans, no conversion of K or V at all; q (fp32) and
// the probabilities are split in registers.  ONE staging buffer (rows are twice as wide: K 104 KB + V 104 KB): K and V keep their
// own full / empty barriers, so the next stream's K arrives during softmax and P V of this one and its V during the next Q K^T.
// Replaces the CUDA-core kernel (1.47 ms per launch at 4096 streams: half of an EXACT step) from 148 streams per step on.
// ------------------------------------------------------------------------------------------
constexpr int AX_RC_OFF = AS_RING_ROWS * 2048;                 // split rows: 2 KB each
constexpr int AX_HALF_BYTES = AX_RC_OFF + AS_RC_ROWS * 2048;   // K (or V): 104 KB

// byte offset of piece p (hi: head, lo: 8 + head) of key's row inside the K (or V) buffer
template <int SEG>
__device__ __forceinline__ uint32_t ax_row_off(int key, int n_ring, int piece) {
  return key < n_ring ? (uint32_t)((key / SEG) * (SEG * 2048) + piece * (SEG * 128) + (key % SEG) * 128)
                      : (uint32_t)(AX_RC_OFF + piece * (AS_RC_ROWS * 128) + (key - n_ring) * 128);
}
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  const float2 hf = __bfloat1622float2(h);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = pack_bf16x2(x0 - hf.x, x1 - hf.y);
}

template <int ROWS>
__device__ __forceinline__ void ax_load_q(uint32_t (&qh)[4][4], uint32_t (&ql)[4][4], const float* qb, int d, int row_base, int g, int tig) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int r0 = row_base + g, r1 = r0 + 8, c0 = ks * 16 + 2 * tig;
    const float2 z = make_float2(0.f, 0.f);
    const float2 a0 = r0 < ROWS ? *reinterpret_cast<const float2*>(qb + (size_t)r0 * d + c0) : z;
    const float2 a1 = r1 < ROWS ? *reinterpret_cast<const float2*>(qb + (size_t)r1 * d + c0) : z;
    const float2 a2 = r0 < ROWS ? *reinterpret_cast<const float2*>(qb + (size_t)r0 * d + c0 + 8) : z;
    const float2 a3 = r1 < ROWS ? *reinterpret_cast<const float2*>(qb + (size_t)r1 * d + c0 + 8) : z;
    split2(a0.x, a0.y, qh[ks][0], ql[ks][0]); split2(a1.x, a1.y, qh[ks][1], ql[ks][1]);
    split2(a2.x, a2.y, qh[ks][2], ql[ks][2]); split2(a3.x, a3.y, qh[ks][3], ql[ks][3]);
  }
}

template <int ROWS>
__global__ void __launch_bounds__(AsCfg<ROWS>::THREADS, 1)
attention_exact_kernel(const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmRC, AttnParams<float> P, int n_streams) {
  constexpr int HEAD_WARPS = AsCfg<ROWS>::HEAD_WARPS;
  constexpr int SEG = ROWS - AS_RC_ROWS;
  constexpr int BLOCK_BYTES = SEG * 2048;
  extern __shared__ uint8_t as_smem_raw[];
  __shared__ __align__(8) unsigned long long ax_bars[4];       // kfull | vfull | kempty | vempty
  __shared__ int ax_meta;
  __shared__ __align__(16) uint8_t ax_zero[16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tig = lane & 3;
  const uint32_t smem0 = ((uint32_t)__cvta_generic_to_shared(as_smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(ax_bars);
  const uint32_t kfull = bar0, vfull = bar0 + 8u, kempty = bar0 + 16u, vempty = bar0 + 24u;
  const uint32_t zero_addr = (uint32_t)__cvta_generic_to_shared(ax_zero);
  if (tid < 4) reinterpret_cast<uint32_t*>(ax_zero)[tid] = 0u;
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmKV) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmRC) : "memory");
    as_mbar_init(kfull, 1); as_mbar_init(vfull, 1); as_mbar_init(kempty, HEAD_WARPS); as_mbar_init(vempty, HEAD_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();

  if (warp == HEAD_WARPS) {
    // ===================== producer: one lane, <= 8 TMA ops per stream =====================
    if (lane == 0) {
      int it = 0;
      for (int b = blockIdx.x; b < n_streams; b += gridDim.x, ++it) {
        const uint32_t par = (uint32_t)(it & 1);
        const int slot = P.slots[b];
        const int pl = P.past_len[slot];
        const int lv = pl < P.left ? pl : P.left;
        const int nb = lv / SEG + 1;
        const int first_blk = (pl - lv) / SEG;
        const int ring_blocks = P.ring / SEG;
        const uint32_t bytes = (uint32_t)(nb * BLOCK_BYTES + AS_RC_ROWS * 2048);
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
          as_mbar_wait(which ? vempty : kempty, par ^ 1u);
          if (!which) ax_meta = lv;
          const uint32_t full = which ? vfull : kfull;
          as_mbar_expect_tx(full, bytes);
          const uint32_t dst = smem0 + (uint32_t)(which * AX_HALF_BYTES);
          const long long row0 = P.cache_row0 + (long long)slot * P.slot_rows + (long long)which * P.ring;
          for (int j = 0; j < nb; ++j)
            as_tma_3d(dst + (uint32_t)(j * BLOCK_BYTES), &tmKV, 0, (int)(row0 + ((first_blk + j) % ring_blocks) * SEG), 0, full);
          as_tma_3d(dst + (uint32_t)AX_RC_OFF, &tmRC, 0, (b * 2 + which) * AS_RC_ROWS, 0, full);
        }
      }
    }
    return;
  }

  // ===================== head warps: warp = (query tile, head) =====================
  const int head = warp % AM_WARPS;
  const int row_base = (warp / AM_WARPS) * 16;
  const bool second_half = row_base + 8 < ROWS;
  const uint32_t s_k = smem0, s_v = smem0 + (uint32_t)AX_HALF_BYTES;
  int it = 0;
  for (int b = blockIdx.x; b < n_streams; b += gridDim.x, ++it) {
    const uint32_t par = (uint32_t)(it & 1);
    // q split into 32 fragment registers: loaded per stream (not prefetched across P V — the 544-thread CTA has 96 registers a thread)
    uint32_t qh[4][4], ql[4][4];
    ax_load_q<ROWS>(qh, ql, P.q + (size_t)b * ROWS * P.d + head * AT_DH, P.d, row_base, g, tig);
    as_mbar_wait(kfull, par);
    const int lv = ax_meta;
    const int n_ring = lv + SEG;
    const int n_keys = n_ring + AS_RC_ROWS;

    // ---- S = Q K^T, three MMAs per fragment pair
    float sc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) sc[nt][e] = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt * 8 < n_keys) {
        int key = nt * 8 + (lane & 7);                            // ldmatrix.x4: lane -> (matrix lane / 8, row lane % 8), as in attention_stream_kernel
        key = key < n_keys ? key : n_keys - 1;                    // clamp: garbage columns are masked below
        const uint32_t rh = ax_row_off<SEG>(key, n_ring, head), rl = ax_row_off<SEG>(key, n_ring, head + 8);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {                          // dims [32 hf, 32 hf + 32): (b0, b1) of steps ks = 2 hf and 2 hf + 1, hi and lo pieces
          uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(h0), "=r"(h1), "=r"(h2), "=r"(h3) : "r"(s_k + as_swz(rh, 4 * hf + (lane >> 3))));
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(l0), "=r"(l1), "=r"(l2), "=r"(l3) : "r"(s_k + as_swz(rl, 4 * hf + (lane >> 3))));
          mma_bf16_16816(sc[nt], ql[2 * hf], h0, h1);             // small terms first
          mma_bf16_16816(sc[nt], qh[2 * hf], l0, l1);
          mma_bf16_16816(sc[nt], qh[2 * hf], h0, h1);
          mma_bf16_16816(sc[nt], ql[2 * hf + 1], h2, h3);
          mma_bf16_16816(sc[nt], qh[2 * hf + 1], l2, l3);
          mma_bf16_16816(sc[nt], qh[2 * hf + 1], h2, h3);
        }
      }
    }
    __syncwarp();
    if (lane == 0) as_mbar_arrive(kempty);

    // ---- softmax over keys, fp32
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      if (hr == 1 && !second_half) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { sc[nt][2] = 0.f; sc[nt][3] = 0.f; }
        continue;
      }
      float m = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const bool valid = nt * 8 + 2 * tig + e < n_keys;
          float& x = sc[nt][2 * hr + e];
          x = valid ? x : -INFINITY;
          m = fmaxf(m, x);
        }
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      float sum = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float& x = sc[nt][2 * hr + e];
          x = expf(x - m);
          sum += x;
        }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = 1.0f / sum;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        sc[nt][2 * hr] *= inv;
        sc[nt][2 * hr + 1] *= inv;
      }
    }
    // ---- O = P V, three MMAs per fragment pair; B fragments of V (hi and lo) via ldmatrix.x4.trans
    as_mbar_wait(vfull, par);
    float oc[8][4];
#pragma unroll
    for (int dn = 0; dn < 8; ++dn)
#pragma unroll
      for (int e = 0; e < 4; ++e) oc[dn][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (kk * 16 < n_keys) {
        uint32_t ph[4], pl_[4];
        split2(sc[2 * kk][0], sc[2 * kk][1], ph[0], pl_[0]);
        split2(sc[2 * kk][2], sc[2 * kk][3], ph[1], pl_[1]);
        split2(sc[2 * kk + 1][0], sc[2 * kk + 1][1], ph[2], pl_[2]);
        split2(sc[2 * kk + 1][2], sc[2 * kk + 1][3], ph[3], pl_[3]);
        const int mi = lane >> 3, ri = lane & 7;
        const int key = kk * 16 + (mi & 1) * 8 + ri;
        const bool kv = key < n_keys;
        const uint32_t rh = ax_row_off<SEG>(kv ? key : 0, n_ring, head), rl = ax_row_off<SEG>(kv ? key : 0, n_ring, head + 8);
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          const uint32_t ah = kv ? s_v + as_swz(rh, dp * 2 + (mi >> 1)) : zero_addr;
          const uint32_t al = kv ? s_v + as_swz(rl, dp * 2 + (mi >> 1)) : zero_addr;
          uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(h0), "=r"(h1), "=r"(h2), "=r"(h3) : "r"(ah));
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(l0), "=r"(l1), "=r"(l2), "=r"(l3) : "r"(al));
          mma_bf16_16816(oc[2 * dp], pl_, h0, h1);
          mma_bf16_16816(oc[2 * dp], ph, l0, l1);
          mma_bf16_16816(oc[2 * dp], ph, h0, h1);
          mma_bf16_16816(oc[2 * dp + 1], pl_, h2, h3);
          mma_bf16_16816(oc[2 * dp + 1], ph, l2, l3);
          mma_bf16_16816(oc[2 * dp + 1], ph, h2, h3);
        }
      }
    }
    __syncwarp();
    if (lane == 0) as_mbar_arrive(vempty);
    // ---- write the A operand of out_proj (hi | lo halves)
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int r = row_base + hr * 8 + g;
      if (r < ROWS) {
        bf16* o = P.out + ((size_t)b * ROWS + r) * P.ld + head * AT_DH + 2 * tig;
#pragma unroll
        for (int dn = 0; dn < 8; ++dn) {
          uint32_t hi, lo;
          split2(oc[dn][2 * hr], oc[dn][2 * hr + 1], hi, lo);
          *reinterpret_cast<uint32_t*>(o + dn * 8) = hi;
          if (P.lo_off) *reinterpret_cast<uint32_t*>(o + P.lo_off + dn * 8) = lo;
        }
      }
    }
  }
}

template <int ROWS>
int attention_exact_launch(const AttnParams<float>& P, int n_streams, int num_sms, cudaStream_t st) {
  const size_t smem = (size_t)2 * AX_HALF_BYTES + 1024;
  static size_t attr_done[kMaxDevices] = {0};
  ASR_CUDA_OK(ensure_dyn_smem(attention_exact_kernel<ROWS>, smem, attr_done));
  const int grid = n_streams < num_sms ? n_streams : num_sms;
  ASR_CUDA_OK(launch_pdl(attention_exact_kernel<ROWS>, dim3(grid), dim3(AsCfg<ROWS>::THREADS), smem, st, *P.h_tm_cache, *P.h_tm_rc, P, n_streams));
  return 0;
}

template <int ROWS>
int attention_mma_launch(const AttnParams<bf16>& P, int n_streams, cudaStream_t st) {
  const int kmax = P.left + P.seg_rows + P.rc_rows;
  const size_t smem = (size_t)(2 * kmax + 1) * AM_ROWB;
  static size_t attr_done[kMaxDevices] = {0};
  ASR_CUDA_OK(ensure_dyn_smem(attention_mma_kernel<ROWS>, smem, attr_done));
  ASR_CUDA_OK(launch_pdl(attention_mma_kernel<ROWS>, dim3(n_streams), dim3(AM_WARPS * 32), smem, st, P));
  return 0;
}

template <typename T, int ROWS>
int attention_launch_rows(const AttnParams<T>& P, int n_streams, cudaStream_t st) {
  static_assert(ROWS % 4 == 0, "ROWS must be a multiple of 4");
  const size_t smem = sizeof(float) * AT_WARPS * (ROWS * AT_DH + AT_MAXK * AT_KST + AT_MAXK * ROWS);
  static size_t attr_done[kMaxDevices] = {0};
  ASR_CUDA_OK(ensure_dyn_smem(attention_kernel<T, ROWS>, smem, attr_done));
  dim3 grid(n_streams, P.n_heads / AT_WARPS);
  ASR_CUDA_OK(launch_pdl(attention_kernel<T, ROWS>, grid, dim3(AT_WARPS * 32), smem, st, P));
  return 0;
}

// ------------------------------------------------------------------------------------------
// CTC log-softmax + argmax + incremental greedy (decoder.py:69; recognition.py:33-57).  One CTA per stream.
// The reference re-scans the whole accumulated emission every chunk; carrying (prev_id, n_frames,
// last_tok_frame) per session gives the identical token sequence and last_blank incrementally.
// ------------------------------------------------------------------------------------------
constexpr int CTC_MAXV = 32;   // vocab <= 1024

// order-preserving float <-> uint32 (so redux.sync max works on log-probs); -inf maps below every finite value
__device__ __forceinline__ uint32_t f2key(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float key2f(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// NV = ceil(vocab / 32) columns per lane (26 for the 804-entry vocabulary: the loops below are fully unrolled over it)
template <int NV>
__global__ void __launch_bounds__(256, 4) ctc_greedy_kernel(CtcParams P) {
  __shared__ int s_ids[64];
  __shared__ float s_clp[8][32];                                 // per warp: the row's elements that can be among its cand_k best
  __shared__ int s_ctok[8][32];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();
  for (int r = warp; r < P.seg_rows; r += 8) {
    const size_t row = (size_t)b * P.seg_rows + r;
    const float* z = P.logits + row * P.vocab;
    float v[NV];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < P.vocab ? z[c] : -INFINITY;
      m = fmaxf(m, v[i]);
    }
    m = warp_max(m);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (lane + 32 * i < P.vocab) ? __expf(v[i] - m) : 0.f;    // ex2.approx: 2e-7 relative per term, 1e-6 on the log-probs (bound 2e-4)
    const float lse = logf(warp_sum(s));
    float best = -INFINITY; int bi = 0x7fffffff;
    float lm = -INFINITY;                                          // lane maximum over the non-blank ids
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < P.vocab) {
        const float lp = (v[i] - m) - lse;                       // log_softmax(dim=2)
        v[i] = lp;
        if (P.logprobs) P.logprobs[row * P.vocab + c] = lp;
        if (lp > best) { best = lp; bi = c; }                     // first max wins inside a lane (c increasing)
        if (c != 0) lm = fmaxf(lm, lp);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {                            // torch.argmax: lowest index among equal maxima
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (P.cand_k > 0) {
      // ---- the cand_k best non-blank ids of the row (value desc, id asc) for the prefix beam search, found in parallel over rows here
      // instead of serially per frame inside the beam kernel.  Threshold T = the cand_k-th largest of the 32 lane maxima: every
      // member of the row's top cand_k is >= T, and usually only ~cand_k elements are; those are compacted into shared memory and ranked
      // one per lane.  More than 32 elements >= T (ties on a flat row) take the plain selection over the registers.
      if (lane == 0) { P.row_stat[2 * row] = m; P.row_stat[2 * row + 1] = lse; }
      uint32_t w = f2key(lm);
      for (int k = 0; k + 1 < P.cand_k; ++k) {
        const uint32_t mx = __reduce_max_sync(0xffffffffu, w);
        const uint32_t bal = __ballot_sync(0xffffffffu, w == mx);
        if (lane == __ffs(bal) - 1) w = 0u;
      }
      float T = key2f(__reduce_max_sync(0xffffffffu, w));
      if (!(T == T)) T = -INFINITY;                                // fewer than cand_k lanes hold a finite value (tiny vocabularies): everything finite qualifies
      int cnt = 0;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        const bool f = c < P.vocab && c != 0 && v[i] >= T && v[i] > -INFINITY;
        const uint32_t bal = __ballot_sync(0xffffffffu, f);
        if (bal) {                                                 // warp-uniform
          const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
          if (f && pos < 32) { s_clp[warp][pos] = v[i]; s_ctok[warp][pos] = c; }
          cnt += __popc(bal);
        }
      }
      __syncwarp();
      float my_lp = -INFINITY; int my_tok = 0;
      if (cnt <= 32) {
        float x = lane < cnt ? s_clp[warp][lane] : -INFINITY;
        const int xc = lane < cnt ? s_ctok[warp][lane] : 0x7fffffff;
        for (int k = 0; k < P.cand_k; ++k) {
          const uint32_t mx = __reduce_max_sync(0xffffffffu, f2key(x));
          const int wc = (int)__reduce_min_sync(0xffffffffu, (uint32_t)(f2key(x) == mx ? xc : 0x7fffffff));
          if (lane == k) { my_lp = key2f(mx); my_tok = mx == f2key(-INFINITY) ? 0x7fffffff : wc; }
          if (xc == wc) x = -INFINITY;
        }
      } else {
        for (int k = 0; k < P.cand_k; ++k) {
          float cb = -INFINITY; int ci = 0x7fffffff;
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            if (c < P.vocab && c != 0 && v[i] > cb) { cb = v[i]; ci = c; }
          }
          const uint32_t mx = __reduce_max_sync(0xffffffffu, f2key(cb));
          const int wc = (int)__reduce_min_sync(0xffffffffu, (uint32_t)(f2key(cb) == mx ? ci : 0x7fffffff));
          if (lane == k) { my_lp = key2f(mx); my_tok = mx == f2key(-INFINITY) ? 0x7fffffff : wc; }
#pragma unroll
          for (int i = 0; i < NV; ++i)
            if (wc == lane + 32 * i) v[i] = -INFINITY;
        }
      }
      __syncwarp();
      if (lane < P.cand_k) {                                       // an exhausted row (fewer finite ids than cand_k) yields lp = -inf entries: never selected
        P.cand_tok[row * BEAM_CAND_MAX + lane] = my_tok;
        P.cand_lp[row * BEAM_CAND_MAX + lane] = my_lp;
      }
    }
    if (lane == 0) { s_ids[r] = bi; P.argmax_ids[row] = bi; }
  }
  __syncthreads();
  if (warp == 0 && P.seg_rows <= 32) {
    // the carry update of greedy_search, one lane per frame of the chunk (ballots instead of a serial scan by one thread)
    const int slot = P.slots[b];
    const int prev0 = P.prev_id[slot], nf = P.n_frames[slot];
    int lt = P.last_tok_frame[slot], ht = P.seg_has_text[slot];
    const bool in = lane < P.seg_rows;
    const int id = in ? s_ids[lane] : 0;
    const int prev = lane == 0 ? prev0 : (in ? s_ids[lane - 1] : 0);
    const uint32_t is_new = __ballot_sync(0xffffffffu, in && id != prev && id != 0);        // unique_consecutive, then drop blank
    const uint32_t is_tok = __ballot_sync(0xffffffffu, in && id > 1);                       // tokens_idx = indices > 1
    const uint32_t is_txt = __ballot_sync(0xffffffffu, in && !((P.silent_mask[id >> 5] >> (id & 31)) & 1u));   // `if text:` (stream.py:121)
    if ((is_new >> lane) & 1u) P.new_tokens[(size_t)b * P.seg_rows + __popc(is_new & ((1u << lane) - 1u))] = id;
    if (lane == 0) {
      if (is_tok) lt = nf + 31 - __clz(is_tok);
      ht |= is_txt != 0u;
      const int nf1 = nf + P.seg_rows;
      P.prev_id[slot] = s_ids[P.seg_rows - 1]; P.n_frames[slot] = nf1; P.last_tok_frame[slot] = lt; P.seg_has_text[slot] = ht;
      P.n_new[b] = __popc(is_new);
      P.has_token[b] = lt >= 0;
      P.has_text[b] = ht;
      P.flags[b] = 0;
      P.blank_frames[b] = lt >= 0 ? nf1 - 1 - lt : nf1;
      P.past_len[slot] += P.seg_rows;                                                     // TA:emformer.py:413 (state[3] + update_length)
    }
  } else if (threadIdx.x == 0 && P.seg_rows > 32) {
    const int slot = P.slots[b];
    int prev = P.prev_id[slot], nf = P.n_frames[slot], lt = P.last_tok_frame[slot], ht = P.seg_has_text[slot], n_new = 0;
    for (int r = 0; r < P.seg_rows; ++r) {
      const int id = s_ids[r];
      if (id != prev && id != 0) P.new_tokens[(size_t)b * P.seg_rows + n_new++] = id;   // unique_consecutive, then drop blank
      if (id > 1) lt = nf;                                                              // tokens_idx = indices > 1
      ht |= !((P.silent_mask[id >> 5] >> (id & 31)) & 1u);                              // `if text:` (stream.py:121): an id that renders to something
      prev = id; ++nf;
    }
    P.prev_id[slot] = prev; P.n_frames[slot] = nf; P.last_tok_frame[slot] = lt; P.seg_has_text[slot] = ht;
    P.n_new[b] = n_new;
    P.has_token[b] = lt >= 0;
    P.has_text[b] = ht;
    P.flags[b] = 0;
    P.blank_frames[b] = lt >= 0 ? nf - 1 - lt : nf;
    P.past_len[slot] += P.seg_rows;                                                     // TA:emformer.py:413 (state[3] + update_length)
  }
}

__global__ void convert_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int rows, int cols, int ld, int lo_off) {
  const size_t n = (size_t)rows * cols;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (size_t)r * cols);
    const float x = src[i];
    const bf16 h = __float2bfloat16_rn(x);
    dst[(size_t)r * ld + c] = h;
    if (lo_off) dst[(size_t)r * ld + lo_off + c] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}

__global__ void subtract_mean_kernel(float* __restrict__ x, int n_frames, int n_mels) {
  // one CTA per stream, one thread per mel bin
  float* p = x + (size_t)blockIdx.x * n_frames * n_mels;
  for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < n_frames; ++t) s += p[(size_t)t * n_mels + m];
    const float mean = s / (float)n_frames;
    for (int t = 0; t < n_frames; ++t) p[(size_t)t * n_mels + m] -= mean;
  }
}

__global__ void reset_slots_kernel(const int* __restrict__ slots, int n, int* past_len, int* n_frames, int* prev_id, int* last_tok, int* has_text) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = slots[i];
    past_len[s] = 0; n_frames[s] = 0; prev_id[s] = -1; last_tok[s] = -1; has_text[s] = 0;
  }
}

// Batch assembly on the device: chunk i = chunk_len int16 samples at base + src_off[i] (pinned, mapped host memory: the sessions'
// audio rings, read over PCIe) -> row i of the step's PCM buffer.  The kernel runs on the copy stream UNDER the compute stream's step, whose
// persistent GEMM CTAs take ~55 k of an SM's 64 k registers: a gather CTA must be small enough to sit next to one (128 threads, few
// registers, one CTA per SM, persistent over the chunks) or the next GEMM's CTAs wait for gather CTAs that are themselves waiting on PCIe
// round trips (a 256-thread CTA per chunk cost the step 2 ms of its 14 at 4096 streams; 74 persistent 128-thread CTAs cost it 0.5 ms).  Four
// independent 16-byte loads per thread keep ~150 KB in flight on the link.
constexpr int GR_THREADS = 128;
__global__ void __launch_bounds__(GR_THREADS) gather_rings_kernel(const int16_t* __restrict__ base, const long long* __restrict__ src_off,
                                                                  int16_t* __restrict__ dst, int n, int chunk_len) {
  for (int c = blockIdx.x; c < n; c += gridDim.x) {
    const int16_t* s = base + src_off[c];
    int16_t* d = dst + (size_t)c * chunk_len;
    if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0 && (chunk_len & 7) == 0) {
      const uint4* s4 = reinterpret_cast<const uint4*>(s);
      uint4* d4 = reinterpret_cast<uint4*>(d);
      const int nv = chunk_len / 8;
      int i = threadIdx.x;
      for (; i + 3 * GR_THREADS < nv; i += 4 * GR_THREADS) {
        const uint4 a0 = s4[i], a1 = s4[i + GR_THREADS], a2 = s4[i + 2 * GR_THREADS], a3 = s4[i + 3 * GR_THREADS];
        d4[i] = a0; d4[i + GR_THREADS] = a1; d4[i + 2 * GR_THREADS] = a2; d4[i + 3 * GR_THREADS] = a3;
      }
      for (; i < nv; i += GR_THREADS) d4[i] = s4[i];
    } else {
      for (int i = threadIdx.x; i < chunk_len; i += GR_THREADS) d[i] = s[i];
    }
  }
}

// dst block i = src block idx[i] (blocks of vec_per_block 16-byte vectors): the pre-computed fbank operands of the chunks that run
__global__ void __launch_bounds__(256) gather_blocks_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, const int* __restrict__ idx, int vec_per_block) {
  const uint4* s = src + (size_t)idx[blockIdx.x] * vec_per_block;
  uint4* d = dst + (size_t)blockIdx.x * vec_per_block;
  for (int i = threadIdx.x; i < vec_per_block; i += blockDim.x) d[i] = s[i];
}

__global__ void fill_i32_kernel(int* p, int v, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace

int ln_to_operand(const float* x, const float* g, const float* b, bf16* out, int ld, int lo_off, int M, int d, cudaStream_t st) {
  if (M <= 0) return 0;
  if (d != LN_D) { set_error("layer norm kernels are built for d_model = %d (got %d)", LN_D, d); return -1; }
  ASR_CUDA_OK(launch_pdl(ln_to_operand_kernel, dim3((M + 7) / 8), dim3(256), 0, st, x, g, b, out, ld, lo_off, M));
  return 0;
}

int ln_out_fused(const float* x2, const float* g1, const float* b1, float* y, const float* g2, const float* b2, bf16* out, int ld,
                 int lo_off, int M, int d, int rows, int seg_rows, cudaStream_t st) {
  if (M <= 0) return 0;
  if (d != LN_D) { set_error("layer norm kernels are built for d_model = %d (got %d)", LN_D, d); return -1; }
  ASR_CUDA_OK(launch_pdl(ln_out_fused_kernel, dim3((M + 7) / 8), dim3(256), 0, st, x2, g1, b1, y, g2, b2, out, ld, lo_off, M, rows, seg_rows));
  return 0;
}

template <typename T>
int attention_launch(const AttnParams<T>& P, int n_streams, cudaStream_t st) {
  if (n_streams <= 0) return 0;
  if (P.d != P.n_heads * AT_DH || P.n_heads % AT_WARPS != 0 || P.left + P.seg_rows + P.rc_rows > AT_MAXK) {
    set_error("attention: unsupported geometry (d %d heads %d keys %d)", P.d, P.n_heads, P.left + P.seg_rows + P.rc_rows);
    return -1;
  }
  if constexpr (sizeof(T) == 2) {
    if (P.n_heads == AM_WARPS && P.d == 512 && P.lo_off == 0) {
      const char* sm_env = getenv("ASR_B200_ATTN_STREAM_MIN");       // read per launch so a test can flip it inside one process
      const int stream_min = sm_env ? atoi(sm_env) : 148;       // measured on B200: wins from 256 streams per step on (1.98 vs 2.00 ms)
      int num_sms = 148;
      cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, current_device_index());
      const bool tma_ok = P.h_tm_cache && P.h_tm_rc && P.rc_rows == AS_RC_ROWS && P.rows == P.seg_rows + P.rc_rows && P.ring <= AS_RING_ROWS &&
                          P.ring % P.seg_rows == 0 && P.left % P.seg_rows == 0 && (P.seg_rows == 16 || P.seg_rows == 8);
      if (n_streams >= stream_min && tma_ok) {         // persistent, double-buffered streaming kernel for large batches
        if (P.rows == 20) return attention_stream_launch<20>(P, n_streams, num_sms, st);
        if (P.rows == 12) return attention_stream_launch<12>(P, n_streams, num_sms, st);
      }
      if (P.rows == 20) return attention_mma_launch<20>(P, n_streams, st);
      if (P.rows == 12) return attention_mma_launch<12>(P, n_streams, st);
    }
  }
  if constexpr (sizeof(T) == 4) {
    const char* sm_env = getenv("ASR_B200_ATTN_STREAM_MIN");
    const int stream_min = sm_env ? atoi(sm_env) : 148;
    const bool tma_ok = P.h_tm_cache && P.h_tm_rc && P.n_heads == AM_WARPS && P.d == 512 && P.rc_rows == AS_RC_ROWS && P.rows == P.seg_rows + P.rc_rows &&
                        P.ring <= AS_RING_ROWS && P.ring % P.seg_rows == 0 && P.left % P.seg_rows == 0 && (P.seg_rows == 16 || P.seg_rows == 8);
    if (n_streams >= stream_min && tma_ok) {
      int num_sms = 148;
      cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, current_device_index());
      if (P.rows == 20) return attention_exact_launch<20>(P, n_streams, num_sms, st);
      if (P.rows == 12) return attention_exact_launch<12>(P, n_streams, num_sms, st);
    }
  }
  if (P.rows == 20) return attention_launch_rows<T, 20>(P, n_streams, st);
  if (P.rows == 12) return attention_launch_rows<T, 12>(P, n_streams, st);
  set_error("attention: unsupported rows per chunk %d (built for 20 and 12)", P.rows);
  return -1;
}
template int attention_launch<float>(const AttnParams<float>&, int, cudaStream_t);
template int attention_launch<bf16>(const AttnParams<bf16>&, int, cudaStream_t);

int ctc_greedy_launch(const CtcParams& P, int n_streams, cudaStream_t st) {
  if (n_streams <= 0) return 0;
  if (P.vocab > 32 * CTC_MAXV || P.seg_rows > 64) { set_error("ctc: vocab %d / seg_rows %d too large", P.vocab, P.seg_rows); return -1; }
  if (P.vocab <= 26 * 32) { ASR_CUDA_OK(launch_pdl(ctc_greedy_kernel<26>, dim3(n_streams), dim3(256), 0, st, P)); }
  else { ASR_CUDA_OK(launch_pdl(ctc_greedy_kernel<CTC_MAXV>, dim3(n_streams), dim3(256), 0, st, P)); }
  return 0;
}

int convert_weight(const float* src, bf16* dst, int rows, int cols, int ld, int lo_off, cudaStream_t st) {
  convert_weight_kernel<<<296, 256, 0, st>>>(src, dst, rows, cols, ld, lo_off);
  ASR_CUDA_OK(cudaGetLastError());
  return 0;
}

int subtract_mean_launch(float* x, int n_streams, int n_frames, int n_mels, cudaStream_t st) {
  subtract_mean_kernel<<<n_streams, 128, 0, st>>>(x, n_frames, n_mels);
  ASR_CUDA_OK(cudaGetLastError());
  return 0;
}

int reset_slots_launch(const int* slots, int n, int* past_len, int* n_frames, int* prev_id, int* last_tok, int* has_text, cudaStream_t st) {
  if (n <= 0) return 0;
  reset_slots_kernel<<<(n + 255) / 256, 256, 0, st>>>(slots, n, past_len, n_frames, prev_id, last_tok, has_text);
  ASR_CUDA_OK(cudaGetLastError());
  return 0;
}

int gather_rings_launch(const int16_t* base_dev, const long long* src_off, int16_t* dst, int n, int chunk_len, cudaStream_t st) {
  if (n <= 0) return 0;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ctas = sms / 2 > 0 ? sms / 2 : 1;       // measured at 4096 streams: 37 or 74 CTAs cost the concurrent step 0.9 ms less than 148 (PCIe-bound either way)
  gather_rings_kernel<<<n < ctas ? n : ctas, GR_THREADS, 0, st>>>(base_dev, src_off, dst, n, chunk_len);
  ASR_CUDA_OK(cudaGetLastError());
  return 0;
}

int gather_blocks_launch(const void* src, void* dst, const int* idx, int n, size_t block_bytes, cudaStream_t st) {
  if (n <= 0) return 0;
  if (block_bytes % 16) { set_error("gather_blocks: block of %zu bytes is not a multiple of 16", block_bytes); return -1; }
  gather_blocks_kernel<<<n, 256, 0, st>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), idx, (int)(block_bytes / 16));
  ASR_CUDA_OK(cudaGetLastError());
  return 0;
}

int fill_i32(int* p, int v, size_t n, cudaStream_t st) {
  if (!n) return 0;
  fill_i32_kernel<<<148, 256, 0, st>>>(p, v, n);
  ASR_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace asr
