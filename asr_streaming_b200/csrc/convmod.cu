// Streaming convolution module with per-session cache (north-star "depthwise causal conv with cache,
// Swish ... as fused memory-bound kernels").  Streaming form of the reference's ConvolutionBlock
// (lightspeech/layers/block.py:129-171):
//     pre_norm (ScaleBiasNorm, normalization.py:9-19) -> pointwise_conv1 -> SiLU -> depthwise_conv(k, padding (k-1)/2, groups = d)
//     -> BatchNorm1d (eval) -> SiLU -> pointwise_conv2
// The reference pads (k-1)/2 frames on both sides of the whole utterance.  Chunk by chunk the same numbers come out of a causal
// window over [k-1 cached frames | new frames] with (k-1)/2 frames of latency: z[q] = sum_i w[i] h[q-(k-1)+i] = y[q-(k-1)/2], so the
// streamed output IS the reference block's output on the full sequence, delayed (tests pin it against the unmodified module).
//   scale_bias_operand_kernel  pre_norm fused into the bf16 (hi|lo) A-operand conversion of pointwise_conv1
//   gemm_tc (tcgen05)          the two pointwise convolutions (1x1 conv == Linear over channels)
//   dwconv_cache_kernel        SiLU on load, k-tap depthwise window held in registers (thread = channel, coalesced over channels),
//                              conv bias + folded BatchNorm affine + SiLU, bf16 operand store, cache update — one pass over HBM:
//                              reads T*d + (k-1)*d fp32, writes T*d bf16 (x2 in EXACT) + (k-1)*d fp32 per stream-chunk.
// The Emformer path of the lightspeech model has no convolution module, so nothing in the engine calls this; it is exposed through
// its own C entry points (asr_convmod_*) for an encoder that has one (Squeezeformer / Conformer blocks, layers/block.py:9-75).
#include <math.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "../../include/asr_b200.h"
#include "gemm.cuh"
#include "kernels.cuh"

namespace asr {
namespace {

constexpr int CM_MAX_K = 32;       // depthwise taps

__global__ void __launch_bounds__(256) scale_bias_operand_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                                  const float* __restrict__ bias, bf16* __restrict__ out, int ld,
                                                                  int lo_off, size_t n8, int d) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = (i * 8) / d;
    const int col = (int)((i * 8) - row * d);
    const float4 a = *reinterpret_cast<const float4*>(x + i * 8), b = *reinterpret_cast<const float4*>(x + i * 8 + 4);
    const float4 s0 = *reinterpret_cast<const float4*>(scale + col), s1 = *reinterpret_cast<const float4*>(scale + col + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(bias + col), b1 = *reinterpret_cast<const float4*>(bias + col + 4);
    const float v[8] = {s0.x * a.x + b0.x, s0.y * a.y + b0.y, s0.z * a.z + b0.z, s0.w * a.w + b0.w,
                        s1.x * b.x + b1.x, s1.y * b.y + b1.y, s1.z * b.z + b1.z, s1.w * b.w + b1.w};
    store_operand8(out + row * ld, col, lo_off, v);
  }
}

struct DwParams {
  const float* h;          // [n*T, d]  pointwise_conv1 output (pre-activation)
  const int* slots;        // [n]
  float* cache;            // [max_sessions, k-1, d]  SiLU(h) of the last k-1 frames of every session
  const float* w;          // [d, k]   depthwise_conv.weight (PyTorch layout [d, 1, k])
  const float* cb;         // [d]      depthwise_conv.bias
  const float* bn_a;       // [d]      gamma / sqrt(var + eps)
  const float* bn_b;       // [d]      beta - mean * bn_a
  bf16* out;               // A operand of pointwise_conv2 [n*T, ld]
  int ld, lo_off, d, k, T;
};

// thread = (stream, channel): the window [k-1 cached | T new] lives in registers, every global access is coalesced over channels.
// Window and taps are right-aligned in CM_MAX_K-sized register arrays (zero taps in front) so that every register index is static.
template <int T>
__global__ void __launch_bounds__(128) dwconv_cache_kernel(DwParams P) {
  const int c = blockIdx.x * 128 + threadIdx.x, b = blockIdx.y;
  if (c >= P.d) return;
  const int kc = P.k - 1, pad = CM_MAX_K - P.k;          // pad = unused leading taps / window entries
  float win[CM_MAX_K - 1 + T];
  float* cache = P.cache + (size_t)P.slots[b] * kc * P.d + c;
#pragma unroll
  for (int i = 0; i < CM_MAX_K - 1; ++i) win[i] = i >= pad ? cache[(size_t)(i - pad) * P.d] : 0.f;
  const float* h = P.h + (size_t)b * T * P.d + c;
#pragma unroll
  for (int t = 0; t < T; ++t) win[CM_MAX_K - 1 + t] = silu(h[(size_t)t * P.d]);                 // block.py:155
  float w[CM_MAX_K];
#pragma unroll
  for (int j = 0; j < CM_MAX_K; ++j) w[j] = j >= pad ? P.w[(size_t)c * P.k + (j - pad)] : 0.f;
  const float cb = P.cb[c], a = P.bn_a[c], bb = P.bn_b[c];
  bf16* o = P.out + (size_t)b * T * P.ld + c;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    float acc = cb;                                                                           // block.py:160: z[t] = sum_i w[i] h[t-(k-1)+i]
#pragma unroll
    for (int j = 0; j < CM_MAX_K; ++j) acc = fmaf(w[j], win[t + j], acc);
    const float y = silu(fmaf(acc, a, bb));                                                   // BatchNorm (eval) + SiLU, block.py:161-162
    const bf16 hi = __float2bfloat16_rn(y);
    o[(size_t)t * P.ld] = hi;
    if (P.lo_off) o[(size_t)t * P.ld + P.lo_off] = __float2bfloat16_rn(y - __bfloat162float(hi));
  }
  // new cache = the last k-1 frames of [old cache | new frames] = the last k-1 entries of the window
#pragma unroll
  for (int i = 0; i < CM_MAX_K - 1; ++i)
    if (i >= pad) cache[(size_t)(i - pad) * P.d] = win[T + i];
}

struct DevMem {
  void* p = nullptr;
  size_t bytes = 0;
  int alloc(size_t n) { bytes = n; ASR_CUDA_OK(cudaMalloc(&p, n ? n : 16)); return 0; }
  void free() { if (p) cudaFree(p); p = nullptr; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

size_t up(size_t v, size_t m) { return (v + m - 1) / m * m; }

}  // namespace
}  // namespace asr

using namespace asr;

struct AsrConvModule {
  int device = 0, num_sms = 148, d = 0, k = 0, T = 0, max_sessions = 0, max_batch = 0, split = 0;
  cudaStream_t stream = nullptr;
  std::mutex mu;
  DevMem w_f32, w_bf16, x, h, y, cache, slots, bn_a, bn_b, a_in, a_mid;
  const float *scale = nullptr, *bias = nullptr, *b1 = nullptr, *dw_w = nullptr, *dw_b = nullptr, *b2 = nullptr;
  bf16 *W1 = nullptr, *W2 = nullptr;
  int ld = 0, lo_off = 0;
  CUtensorMap tm_in, tm_mid, tm_w1[2], tm_w2[2];     // weight maps with 128- and 256-row boxes
};

namespace {

uint64_t convmod_count(int d, int k) { return (uint64_t)2 * d + (uint64_t)d * d + d + (uint64_t)d * k + d + 4 * (uint64_t)d + (uint64_t)d * d + d; }

void convmod_free(AsrConvModule* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  for (DevMem* b : {&m->w_f32, &m->w_bf16, &m->x, &m->h, &m->y, &m->cache, &m->slots, &m->bn_a, &m->bn_b, &m->a_in, &m->a_mid}) b->free();
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

int convmod_gemm(AsrConvModule* m, const CUtensorMap& ta, const CUtensorMap (&tw)[2], int M, const float* bias, float* out) {
  const GemmProblem p = make_problem(M, m->d, m->d, m->split);
  const int tiles256 = ((M + 127) / 128) * ((m->d + 255) / 256);
  const int bn = tiles256 >= m->num_sms ? 256 : 128;
  EpiF32 epi{out, bias, nullptr, m->d, m->d};
  return gemm_tc<EpiF32>(ta, tw[bn == 256], p, epi, bn, m->num_sms, m->stream);
}

}  // namespace

extern "C" {

int asr_convmod_weights_count(int32_t d_model, int32_t kernel, uint64_t* n_floats) {
  if (!n_floats || d_model <= 0 || kernel <= 0) { set_error("asr_convmod_weights_count: bad arguments"); return -1; }
  *n_floats = convmod_count(d_model, kernel);
  return 0;
}

/* weights (fp32, in this order): pre_norm.scale[d], pre_norm.bias[d], pointwise_conv1.weight[d,d], .bias[d], depthwise_conv.weight[d,k],
 * .bias[d], norm.weight[d], norm.bias[d], norm.running_mean[d], norm.running_var[d], pointwise_conv2.weight[d,d], .bias[d] */
int asr_convmod_create(int32_t d_model, int32_t kernel, int32_t rows_per_chunk, int32_t max_sessions, int32_t max_batch, int32_t precision,
                       const float* weights, uint64_t n_floats, int32_t device, AsrConvModule** out) {
  if (!weights || !out) { set_error("null argument"); return -1; }
  if (d_model % 128 || kernel < 1 || kernel > CM_MAX_K || !(kernel & 1) || (rows_per_chunk != 8 && rows_per_chunk != 16 && rows_per_chunk != 32) ||
      max_sessions <= 0 || max_batch <= 0) {
    set_error("convmod: unsupported geometry (d_model %% 128 == 0, odd kernel <= %d, rows per chunk in {8,16,32})", CM_MAX_K); return -1;
  }
  if (n_floats != convmod_count(d_model, kernel)) { set_error("convmod: weights blob has %llu floats, needs %llu", (unsigned long long)n_floats, (unsigned long long)convmod_count(d_model, kernel)); return -1; }
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { set_error("no CUDA device: the B200 path has no CPU fallback"); return -1; }
  if (device < 0 || device >= n_dev) { set_error("device %d out of range", device); return -1; }
  ASR_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  ASR_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) { set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor); return -1; }
  AsrConvModule* m = new AsrConvModule();
  m->device = device; m->num_sms = prop.multiProcessorCount; m->d = d_model; m->k = kernel; m->T = rows_per_chunk;
  m->max_sessions = max_sessions; m->max_batch = max_batch; m->split = precision == ASR_PRECISION_EXACT;
  const int d = d_model, k = kernel;
  m->ld = m->split ? 2 * d : d; m->lo_off = m->split ? d : 0;
  int rc = -1;
  do {
    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); break; }
    if (m->w_f32.alloc(4 * n_floats) || cudaMemcpy(m->w_f32.p, weights, 4 * n_floats, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("weights upload failed"); break; }
    const float* w = m->w_f32.as<float>();
    m->scale = w; m->bias = w + d;
    const float* W1 = w + 2 * d; m->b1 = W1 + (size_t)d * d;
    m->dw_w = m->b1 + d; m->dw_b = m->dw_w + (size_t)d * k;
    const size_t bn_off = (size_t)(m->dw_b + d - w);
    const float* W2 = w + bn_off + 4 * d; m->b2 = W2 + (size_t)d * d;
    // BatchNorm1d (eval) folded into one affine per channel (block.py:163): a = gamma / sqrt(var + eps), b = beta - mean * a
    std::vector<float> a(d), b(d);
    const float *g = weights + bn_off, *be = g + d, *mu = be + d, *var = mu + d;
    for (int c = 0; c < d; ++c) { a[c] = g[c] / sqrtf(var[c] + 1e-5f); b[c] = be[c] - mu[c] * a[c]; }
    if (m->bn_a.alloc(4 * d) || m->bn_b.alloc(4 * d) || cudaMemcpy(m->bn_a.p, a.data(), 4 * d, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(m->bn_b.p, b.data(), 4 * d, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("BatchNorm upload failed"); break; }
    if (m->w_bf16.alloc(2 * 2 * (size_t)d * m->ld)) break;
    m->W1 = m->w_bf16.as<bf16>(); m->W2 = m->W1 + (size_t)d * m->ld;
    if (convert_weight(W1, m->W1, d, d, m->ld, m->lo_off, m->stream) || convert_weight(W2, m->W2, d, d, m->ld, m->lo_off, m->stream)) break;
    const size_t Mp = up((size_t)max_batch * m->T, 128);
    if (m->x.alloc(4 * Mp * d) || m->h.alloc(4 * Mp * d) || m->y.alloc(4 * Mp * d) || m->slots.alloc(4 * (size_t)max_batch) ||
        m->a_in.alloc(2 * Mp * m->ld) || m->a_mid.alloc(2 * Mp * m->ld) || m->cache.alloc(4 * (size_t)max_sessions * (k - 1) * d + 16)) break;
    cudaMemsetAsync(m->a_in.p, 0, m->a_in.bytes, m->stream); cudaMemsetAsync(m->a_mid.p, 0, m->a_mid.bytes, m->stream);
    cudaMemsetAsync(m->cache.p, 0, m->cache.bytes, m->stream);
    bool ok = !make_tmap_bf16_2d(&m->tm_in, m->a_in.p, m->ld, Mp, m->ld, 128) && !make_tmap_bf16_2d(&m->tm_mid, m->a_mid.p, m->ld, Mp, m->ld, 128);
    for (int i = 0; i < 2 && ok; ++i)
      ok = !make_tmap_bf16_2d(&m->tm_w1[i], m->W1, m->ld, d, m->ld, i ? 256 : 128) && !make_tmap_bf16_2d(&m->tm_w2[i], m->W2, m->ld, d, m->ld, i ? 256 : 128);
    if (!ok) break;
    if (cudaStreamSynchronize(m->stream) != cudaSuccess) { set_error("convmod init: %s", cudaGetErrorString(cudaGetLastError())); break; }
    rc = 0;
  } while (0);
  if (rc) { convmod_free(m); return -1; }
  *out = m;
  return 0;
}

int asr_convmod_destroy(AsrConvModule* m) { convmod_free(m); return 0; }

/* Start of an utterance for the listed sessions: their cached frames become zeros (the reference's left zero padding). */
int asr_convmod_reset(AsrConvModule* m, int32_t n, const int32_t* slots) {
  if (!m || (n > 0 && !slots)) { set_error("null argument"); return -1; }
  std::lock_guard<std::mutex> lk(m->mu);
  ASR_CUDA_OK(cudaSetDevice(m->device));
  const size_t per = 4 * (size_t)(m->k - 1) * m->d;
  for (int i = 0; i < n; ++i) {
    if (slots[i] < 0 || slots[i] >= m->max_sessions) { set_error("convmod: slot %d out of range", slots[i]); return -1; }
    ASR_CUDA_OK(cudaMemsetAsync(m->cache.as<uint8_t>() + (size_t)slots[i] * per, 0, per, m->stream));
  }
  ASR_CUDA_OK(cudaStreamSynchronize(m->stream));
  return 0;
}

/* One chunk (rows_per_chunk frames) for each listed session.  x, y: host fp32 [n, rows_per_chunk, d_model]; y is the reference
 * ConvolutionBlock's output delayed by (kernel-1)/2 frames. */
int asr_convmod_step(AsrConvModule* m, int32_t n, const int32_t* slots, const float* x, float* y) {
  if (!m || (n > 0 && (!slots || !x || !y))) { set_error("null argument"); return -1; }
  if (n < 0 || n > m->max_batch) { set_error("convmod: n = %d outside [0, max_batch = %d]", n, m->max_batch); return -1; }
  if (!n) return 0;
  for (int i = 0; i < n; ++i) {
    if (slots[i] < 0 || slots[i] >= m->max_sessions) { set_error("convmod: slot %d out of range", slots[i]); return -1; }
    for (int j = 0; j < i; ++j) if (slots[j] == slots[i]) { set_error("convmod: slot %d appears twice in one step", slots[i]); return -1; }
  }
  std::lock_guard<std::mutex> lk(m->mu);
  ASR_CUDA_OK(cudaSetDevice(m->device));
  pdl_set_active(false);
  const int d = m->d, M = n * m->T;
  const size_t bytes = 4 * (size_t)M * d;
  ASR_CUDA_OK(cudaMemcpyAsync(m->x.p, x, bytes, cudaMemcpyHostToDevice, m->stream));
  ASR_CUDA_OK(cudaMemcpyAsync(m->slots.p, slots, 4 * (size_t)n, cudaMemcpyHostToDevice, m->stream));
  const size_t n8 = (size_t)M * d / 8;
  scale_bias_operand_kernel<<<(unsigned)std::min<size_t>((n8 + 255) / 256, 148 * 8), 256, 0, m->stream>>>(m->x.as<float>(), m->scale, m->bias, m->a_in.as<bf16>(),
                                                                                                        m->ld, m->lo_off, n8, d);
  ASR_CUDA_OK(cudaGetLastError());
  if (convmod_gemm(m, m->tm_in, m->tm_w1, M, m->b1, m->h.as<float>())) return -1;
  DwParams P{m->h.as<float>(), m->slots.as<int>(), m->cache.as<float>(), m->dw_w, m->dw_b, m->bn_a.as<float>(), m->bn_b.as<float>(),
             m->a_mid.as<bf16>(), m->ld, m->lo_off, d, m->k, m->T};
  const dim3 grid(d / 128, n);
  if (m->T == 8) dwconv_cache_kernel<8><<<grid, 128, 0, m->stream>>>(P);
  else if (m->T == 16) dwconv_cache_kernel<16><<<grid, 128, 0, m->stream>>>(P);
  else dwconv_cache_kernel<32><<<grid, 128, 0, m->stream>>>(P);
  ASR_CUDA_OK(cudaGetLastError());
  if (convmod_gemm(m, m->tm_mid, m->tm_w2, M, m->b2, m->y.as<float>())) return -1;
  ASR_CUDA_OK(cudaMemcpyAsync(y, m->y.p, bytes, cudaMemcpyDeviceToHost, m->stream));
  cudaError_t e = cudaStreamSynchronize(m->stream);
  if (e != cudaSuccess) { set_error("convmod step failed: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

}  // extern "C"
