// Log-mel front-ends as fused sm_100a kernels: PCM load -> window -> shared-memory Stockham FFT -> power ->
// sparse-triangle mel projection -> log -> fp32 features and/or the bf16 GEMM A-operand of input_linear.
//
//   melspec128 : the reference's extract_filterbank (lightspeech/datas/audio.py:9-30 ->
//                torchaudio MelSpectrogram(n_fft=800, win=400 Hann periodic zero-padded centred, hop=160,
//                power=2, 128 HTK mels 0..8000 Hz, norm=None) -> clamp(1e-5).log()).
//   kaldi80    : north-star front-end / BASELINE config #2 (TA:compliance/kaldi.py:514-607: snip_edges,
//                remove-DC, pre-emphasis 0.97, povey window, 512-pt FFT, 80 Kaldi-mel bins from 20 Hz, log(max(eps,.))).
//
// One CTA per stream-chunk; the chunk's PCM is staged once in shared memory with 16-byte loads (frames overlap
// 60 %, so each sample is read from HBM once); one warp per frame.  A real 2*NC-point FFT of a 400-sample frame is
// computed as an NC-point complex FFT (NC = 400: radix 4,4,5,5; NC = 256: radix 4,4,4,4) of z[m] = y[2m] + i*y[2m+1]
// followed by the split post-process.  The centred zero padding of torch.stft only changes the phase, not |X|^2.
#include "kernels.cuh"

namespace asr {

namespace {

constexpr int kWarps = 8;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

template <int R> __device__ __forceinline__ void butterfly(float2 (&v)[R]);

template <> __device__ __forceinline__ void butterfly<4>(float2 (&v)[4]) {
  const float2 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]), a2 = cadd(v[1], v[3]);
  const float2 d = csub(v[1], v[3]);
  const float2 a3 = make_float2(d.y, -d.x);          // -i * (v1 - v3)
  v[0] = cadd(a0, a2); v[1] = cadd(a1, a3); v[2] = csub(a0, a2); v[3] = csub(a1, a3);
}

template <> __device__ __forceinline__ void butterfly<5>(float2 (&v)[5]) {
  const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;   // cos(2pi/5), cos(4pi/5)
  const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;    // sin(2pi/5), sin(4pi/5)
  const float2 p1 = cadd(v[1], v[4]), p2 = cadd(v[2], v[3]), d1 = csub(v[1], v[4]), d2 = csub(v[2], v[3]);
  const float2 A1 = make_float2(v[0].x + c1 * p1.x + c2 * p2.x, v[0].y + c1 * p1.y + c2 * p2.y);
  const float2 A2 = make_float2(v[0].x + c2 * p1.x + c1 * p2.x, v[0].y + c2 * p1.y + c1 * p2.y);
  const float2 B1 = make_float2(s1 * d1.x + s2 * d2.x, s1 * d1.y + s2 * d2.y);
  const float2 B2 = make_float2(s2 * d1.x - s1 * d2.x, s2 * d1.y - s1 * d2.y);
  v[0] = make_float2(v[0].x + p1.x + p2.x, v[0].y + p1.y + p2.y);
  v[1] = make_float2(A1.x + B1.y, A1.y - B1.x);      // A1 - i*B1
  v[4] = make_float2(A1.x - B1.y, A1.y + B1.x);
  v[2] = make_float2(A2.x + B2.y, A2.y - B2.x);
  v[3] = make_float2(A2.x - B2.y, A2.y + B2.x);
}

// One Stockham autosort pass of radix R over NC points held in shared memory (one warp).
template <int NC, int R>
__device__ __forceinline__ void fft_pass(const float2* __restrict__ in, float2* __restrict__ out, int Ns, const float2* __restrict__ tw, int lane) {
  constexpr int Q = NC / R;
  for (int j = lane; j < Q; j += 32) {
    const int k = j % Ns;
    const int tstep = k * (NC / (Ns * R));
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = in[j + r * Q];
#pragma unroll
    for (int r = 1; r < R; ++r) v[r] = cmul(v[r], tw[tstep * r]);
    butterfly<R>(v);
    const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) out[j0 + r * Ns] = v[r];
  }
  __syncwarp();
}

template <int NC> __device__ __forceinline__ void fft_warp(float2* a, float2* b, const float2* tw, int lane);
template <> __device__ __forceinline__ void fft_warp<400>(float2* a, float2* b, const float2* tw, int lane) {
  fft_pass<400, 4>(a, b, 1, tw, lane);
  fft_pass<400, 4>(b, a, 4, tw, lane);
  fft_pass<400, 5>(a, b, 16, tw, lane);
  fft_pass<400, 5>(b, a, 80, tw, lane);
}
template <> __device__ __forceinline__ void fft_warp<256>(float2* a, float2* b, const float2* tw, int lane) {
  fft_pass<256, 4>(a, b, 1, tw, lane);
  fft_pass<256, 4>(b, a, 4, tw, lane);
  fft_pass<256, 4>(a, b, 16, tw, lane);
  fft_pass<256, 4>(b, a, 64, tw, lane);
}

template <typename PcmT> __device__ __forceinline__ float pcm_to_float(PcmT v);
template <> __device__ __forceinline__ float pcm_to_float<int16_t>(int16_t v) { return (float)v; }
template <> __device__ __forceinline__ float pcm_to_float<float>(float v) { return v; }

template <int NC, typename PcmT, bool KALDI>
__global__ void __launch_bounds__(kWarps * 32) fbank_kernel(FbankParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  // layout: tw[NC] float2 | w2[NC+1] float2 (padded to even) | window[frame_len] | per-warp 2*NC float2 | pcm
  float2* s_tw = reinterpret_cast<float2*>(smem);
  float2* s_w2 = s_tw + NC;
  float* s_win = reinterpret_cast<float*>(s_w2 + NC + 2);
  float2* s_fft = reinterpret_cast<float2*>(s_win + ((P.frame_len + 3) & ~3));
  PcmT* s_pcm = reinterpret_cast<PcmT*>(s_fft + kWarps * 2 * NC);

  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < NC; i += blockDim.x) s_tw[i] = P.tw[i];
  for (int i = tid; i <= NC; i += blockDim.x) s_w2[i] = P.w2[i];
  for (int i = tid; i < P.frame_len; i += blockDim.x) s_win[i] = P.window[i];
  pdl_launch_dependents();
  pdl_wait();                                              // the tables above are constants; PCM / outputs are not
  {
    // 16-byte vectorised, coalesced PCM stage-in (stream base is 16B aligned: pcm_stride * sizeof(PcmT) % 16 == 0)
    const int n_vec = (P.n_samples * (int)sizeof(PcmT)) / 16;
    const int4* src = reinterpret_cast<const int4*>(reinterpret_cast<const PcmT*>(P.pcm) + (size_t)b * P.pcm_stride);
    int4* dst = reinterpret_cast<int4*>(s_pcm);
    for (int i = tid; i < n_vec; i += blockDim.x) dst[i] = __ldg(src + i);
    for (int i = n_vec * (16 / (int)sizeof(PcmT)) + tid; i < P.n_samples; i += blockDim.x)
      s_pcm[i] = reinterpret_cast<const PcmT*>(P.pcm)[(size_t)b * P.pcm_stride + i];
  }
  __syncthreads();

  float2* fa = s_fft + warp * 2 * NC;
  float2* fb = fa + NC;
  const int half = P.frame_len / 2;                                  // 200 complex inputs
  const float in_scale = P.in_scale;

  for (int f = warp; f < P.n_frames; f += kWarps) {
    const PcmT* x = s_pcm + f * P.hop + P.frame_off;
    float mean = 0.f;
    if (KALDI) {                                                     // remove_dc_offset (kaldi.py:218-221)
      float s = 0.f;
      for (int n = lane; n < P.frame_len; n += 32) s += pcm_to_float<PcmT>(x[n]);
      mean = warp_sum(s) / (float)P.frame_len;
    }
    for (int m = lane; m < NC; m += 32) {
      float2 z = make_float2(0.f, 0.f);
      if (m < half) {
        float x0 = pcm_to_float<PcmT>(x[2 * m]) * in_scale, x1 = pcm_to_float<PcmT>(x[2 * m + 1]) * in_scale;
        if (KALDI) {                                                 // pre-emphasis with replicate pad (kaldi.py:228-233)
          const float xm1 = pcm_to_float<PcmT>(x[m == 0 ? 0 : 2 * m - 1]) - mean;
          x0 -= mean; x1 -= mean;
          const float y0 = x0 - P.preemph * xm1, y1 = x1 - P.preemph * x0;
          x0 = y0; x1 = y1;
        }
        z = make_float2(x0 * s_win[2 * m], x1 * s_win[2 * m + 1]);
      }
      fa[m] = z;
    }
    __syncwarp();
    fft_warp<NC>(fa, fb, s_tw, lane);                                // result in fa
    // split post-process: Y[k] = E[k] + W[k] * O[k], k = 0..NC;  power -> fb (as floats)
    float* pw = reinterpret_cast<float*>(fb);
    for (int k = lane; k <= NC; k += 32) {
      const float2 zk = fa[k == NC ? 0 : k];
      const float2 zr = fa[k == 0 ? 0 : NC - k];
      const float2 E = make_float2(0.5f * (zk.x + zr.x), 0.5f * (zk.y - zr.y));
      const float2 O = make_float2(0.5f * (zk.y + zr.y), -0.5f * (zk.x - zr.x));
      const float2 Y = cadd(E, cmul(s_w2[k], O));
      pw[k] = Y.x * Y.x + Y.y * Y.y;
    }
    __syncwarp();
    // sparse triangle mel projection (<= 2 mels per FFT bin) + log
    const size_t orow = (size_t)b * P.n_frames + f;
    for (int m = lane; m < P.n_mels; m += 32) {
      const int s0 = __ldg(P.mel_start + m), cnt = __ldg(P.mel_cnt + m);
      const float* w = P.mel_w + __ldg(P.mel_off + m);
      float acc = 0.f;
      for (int i = 0; i < cnt; ++i) acc = fmaf(pw[s0 + i], __ldg(w + i), acc);
      const float v = logf(fmaxf(acc, P.log_floor));
      if (P.out_f32) P.out_f32[orow * P.n_mels + m] = v;
      if (P.out_op) {
        bf16* o = P.out_op + orow * P.op_ld + m;
        const bf16 h = __float2bfloat16_rn(v);
        *o = h;
        if (P.op_lo_off) o[P.op_lo_off] = __float2bfloat16_rn(v - __bfloat162float(h));
      }
    }
    __syncwarp();
  }
}

template <int NC, typename PcmT, bool KALDI>
int launch(const FbankParams& P, int n_streams, cudaStream_t st) {
  const size_t smem = sizeof(float2) * (NC + NC + 2) + sizeof(float) * ((P.frame_len + 3) & ~3) + sizeof(float2) * kWarps * 2 * NC +
                      ((P.n_samples * sizeof(PcmT) + 15) & ~(size_t)15);
  static size_t attr_done[kMaxDevices] = {0};
  ASR_CUDA_OK(ensure_dyn_smem(fbank_kernel<NC, PcmT, KALDI>, smem, attr_done));
  ASR_CUDA_OK(launch_pdl(fbank_kernel<NC, PcmT, KALDI>, dim3(n_streams), dim3(kWarps * 32), smem, st, P));
  return 0;
}

}  // namespace

int fbank_launch(const FbankParams& P, int n_streams, cudaStream_t st) {
  if (n_streams <= 0) return 0;
  const int need = (P.n_frames - 1) * P.hop + P.frame_off + P.frame_len;
  if (need > P.n_samples || P.frame_len > 2 * P.nc || (P.frame_len & 1)) { set_error("fbank: bad geometry"); return -1; }
  if (P.nc == 400 && !P.kaldi) return P.pcm_is_f32 ? launch<400, float, false>(P, n_streams, st) : launch<400, int16_t, false>(P, n_streams, st);
  if (P.nc == 256 && P.kaldi) return P.pcm_is_f32 ? launch<256, float, true>(P, n_streams, st) : launch<256, int16_t, true>(P, n_streams, st);
  set_error("fbank: unsupported plan nc=%d kaldi=%d", P.nc, P.kaldi);
  return -1;
}

}  // namespace asr
