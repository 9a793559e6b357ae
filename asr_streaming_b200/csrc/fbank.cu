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
// 60 %, so each sample is read from HBM once).  A real 2*NC-point FFT of a 400-sample frame is computed as an NC-point complex
// FFT of z[m] = y[2m] + i*y[2m+1] followed by the split post-process; the complex FFT is a TWO-STEP REGISTER FFT (fft_regs.cuh):
// NC = 16 x N2 (256 = 16 x 16, 400 = 16 x 25), a warp transforms two frames at once — step 1: DFT-16 per lane in registers
// straight from the staged PCM (window / pre-emphasis applied on load; for NC = 400 half of the inputs are the zero padding, so
// 8-input DFT-16s), twiddles in registers, one conflict-free exchange through shared memory, step 2: DFT-N2 per lane in registers.
// Shared-memory traffic per frame: ~13 KB instead of ~35 KB for four radix passes through shared memory (the first version, which
// ran at 78 % l1tex throughput with 4-way conflicts on its scattered writes).  The centred zero padding of torch.stft only changes
// the phase, not |X|^2.
#include "kernels.cuh"
#include "fft_regs.cuh"

namespace asr {

namespace {

constexpr int kWarps = 8;

template <typename PcmT> __device__ __forceinline__ float pcm_to_float(PcmT v);
template <> __device__ __forceinline__ float pcm_to_float<int16_t>(int16_t v) { return (float)v; }
template <> __device__ __forceinline__ float pcm_to_float<float>(float v) { return v; }

// x[i], x[i + 1] for even i (frames start on even samples: hop and frame_off are even, checked by the launcher)
template <typename PcmT> __device__ __forceinline__ float2 pcm_pair(const PcmT* x, int i);
template <> __device__ __forceinline__ float2 pcm_pair<int16_t>(const int16_t* x, int i) {
  const uint32_t u = *reinterpret_cast<const uint32_t*>(x + i);
  return make_float2((float)(int16_t)(u & 0xffffu), (float)(int16_t)(u >> 16));
}
template <> __device__ __forceinline__ float2 pcm_pair<float>(const float* x, int i) { return *reinterpret_cast<const float2*>(x + i); }

template <int NC> struct FbSmem {
  typedef fftr::TwoStep<NC> TS;
  static constexpr int PW = (NC + 1 + 3) & ~3;                        // power spectrum floats per frame
  static constexpr int WARP_BYTES = 2 * TS::EX * 8 + 2 * PW * 4;
};

template <int NC, typename PcmT, bool KALDI>
__global__ void __launch_bounds__(kWarps * 32, KALDI ? 3 : 2) fbank_kernel(FbankParams P) {
  typedef fftr::TwoStep<NC> TS;
  extern __shared__ __align__(16) uint8_t smem[];
  // layout: w2[NC+2] float2 | window[frame_len] | per-warp { exchange / spectrum 2 x EX float2, power 2 x PW floats } | pcm
  float2* s_w2 = reinterpret_cast<float2*>(smem);
  float* s_win = reinterpret_cast<float*>(s_w2 + NC + 2);
  uint8_t* s_warp = reinterpret_cast<uint8_t*>(s_win + ((P.frame_len + 3) & ~3));
  PcmT* s_pcm = reinterpret_cast<PcmT*>(s_warp + kWarps * FbSmem<NC>::WARP_BYTES);
  float* s_blk = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_pcm) + ((P.n_samples * sizeof(PcmT) + 15) & ~(size_t)15));   // KALDI: block sums

  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i <= NC; i += blockDim.x) s_w2[i] = P.w2[i];
  for (int i = tid; i < P.frame_len; i += blockDim.x) s_win[i] = P.window[i] * P.in_scale;   // int16 -> [-1, 1) folded into the window
  // step-1 twiddles of this lane's column n2: W_NC^(n2 * k1), kept in registers for every frame of the chunk
  float2 tw[16];
  {
    const int n2 = TS::N2 == 16 ? (lane & 15) : (lane < TS::N2 ? lane : 0);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) tw[k1] = __ldg(P.tw + n2 * k1);
  }
  pdl_launch_dependents();
  pdl_wait();                                              // the tables above are constants; PCM / outputs are not
  {
    // 16-byte vectorised, coalesced PCM stage-in (stream base is 16B aligned: pcm_stride * sizeof(PcmT) % 16 == 0)
    const int n_vec = (P.n_samples * (int)sizeof(PcmT)) / 16;
    const size_t srow = P.row_index ? (size_t)P.row_index[b] : (size_t)b;
    const int4* src = reinterpret_cast<const int4*>(reinterpret_cast<const PcmT*>(P.pcm) + srow * P.pcm_stride);
    int4* dst = reinterpret_cast<int4*>(s_pcm);
    for (int i = tid; i < n_vec; i += blockDim.x) dst[i] = __ldg(src + i);
    for (int i = n_vec * (16 / (int)sizeof(PcmT)) + tid; i < P.n_samples; i += blockDim.x)
      s_pcm[i] = reinterpret_cast<const PcmT*>(P.pcm)[srow * P.pcm_stride + i];
  }
  __syncthreads();
  if (KALDI) {
    // remove_dc_offset needs every frame's mean: frames overlap, so sum blocks of gcd(hop, frame_len) samples once per chunk and
    // add frame_len / blk of them per frame (for int16 input every partial sum is an integer below 2^24: exact in any order)
    const int n_blk = ((P.n_frames - 1) * P.hop + P.frame_len) / P.blk;
    for (int t = tid; t < n_blk; t += blockDim.x) {
      const PcmT* x = s_pcm + P.frame_off + t * P.blk;
      float s = 0.f;
      for (int n = 0; n < P.blk; n += 2) { const float2 u = pcm_pair<PcmT>(x, n); s += u.x + u.y; }
      s_blk[t] = s;
    }
    __syncthreads();
  }

  float2* ex = reinterpret_cast<float2*>(s_warp + warp * FbSmem<NC>::WARP_BYTES);
  float* pw = reinterpret_cast<float*>(ex + 2 * TS::EX);
  const int half = P.frame_len / 2;                                  // 200 complex inputs
  const int n_pairs = (P.n_frames + 1) >> 1;

  for (int p = warp; p < n_pairs; p += kWarps) {
    const int f0 = 2 * p;
    const int f1 = f0 + 1 < P.n_frames ? f0 + 1 : f0;                // odd frame count: the second slot recomputes the first
    float mean0 = 0.f, mean1 = 0.f;
    if (KALDI) {                                                     // remove_dc_offset (kaldi.py:218-221)
      const int per = P.frame_len / P.blk, b0 = f0 * P.hop / P.blk, b1 = f1 * P.hop / P.blk;
      float s0 = 0.f, s1 = 0.f;
      for (int j = 0; j < per; ++j) { s0 += s_blk[b0 + j]; s1 += s_blk[b1 + j]; }
      mean0 = s0 / (float)P.frame_len;
      mean1 = s1 / (float)P.frame_len;
    }
    // ---- step 1: DFT-16 over n1 of z[n1 * N2 + n2] (this lane's column n2), twiddle, -> exchange buffer
#pragma unroll
    for (int r = 0; r < TS::ROUNDS; ++r) {
      int h, n2;
      bool active;
      fftr::step1_slot<NC>(lane, r, h, n2, active);
      if (active) {
        const PcmT* x = s_pcm + (h ? f1 : f0) * P.hop + P.frame_off;
        const float mean = h ? mean1 : mean0;
        float2 a[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
          const int m = n1 * TS::N2 + n2;
          float2 z = make_float2(0.f, 0.f);
          if ((!TS::NNZ8 || n1 < 8) && m < half) {
            const float2 xp = pcm_pair<PcmT>(x, 2 * m);
            float x0 = xp.x, x1 = xp.y;
            if (KALDI) {                                             // pre-emphasis with replicate pad (kaldi.py:228-233)
              const float xm1 = pcm_to_float<PcmT>(x[m == 0 ? 0 : 2 * m - 1]) - mean;
              x0 -= mean; x1 -= mean;
              const float y0 = x0 - P.preemph * xm1, y1 = x1 - P.preemph * x0;
              x0 = y0; x1 = y1;
            }
            const float2 w = *reinterpret_cast<const float2*>(s_win + 2 * m);
            z = make_float2(x0 * w.x, x1 * w.y);
          }
          a[n1] = z;
        }
        fftr::step1<NC>(a, tw, ex + h * TS::EX, n2);
      }
    }
    __syncwarp();
    // ---- step 2: DFT-N2 over n2 for (frame lane / 16, k1 = lane % 16); spectrum back into the buffer in natural order
    {
      float2 X[TS::N2];
      float2* exf = ex + (lane >> 4) * TS::EX;
      fftr::step2_compute<NC>(exf, lane & 15, X);
      __syncwarp();
      fftr::step2_store<NC>(exf, lane & 15, X);
    }
    __syncwarp();
    // ---- both frames, whole warp: split post-process Y[k] = E[k] + W[k] * O[k], k = 0..NC; power; sparse mel projection; log.
    // Bins k and NC - k share E, O and the product T = W[k] * O:  Y[k] = E + T,  Y[NC - k] = conj(E - T).
    float* pw1 = pw + FbSmem<NC>::PW;
    for (int k = lane + 1; k <= NC / 2; k += 32) {
      const float2 w = s_w2[k];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float2* fa = ex + h * TS::EX;
        const float2 zk = fa[k], zr = fa[NC - k];
        const float2 E = make_float2(0.5f * (zk.x + zr.x), 0.5f * (zk.y - zr.y));
        const float2 O = make_float2(0.5f * (zk.y + zr.y), -0.5f * (zk.x - zr.x));
        const float2 T = fftr::cmul(w, O);
        const float2 Yp = fftr::cadd(E, T), Ym = fftr::csub(E, T);
        float* o = h ? pw1 : pw;
        o[k] = Yp.x * Yp.x + Yp.y * Yp.y;
        o[NC - k] = Ym.x * Ym.x + Ym.y * Ym.y;
      }
    }
    if (lane < 2) {                                                  // k = 0 and k = NC (Nyquist) are real: z0.x +- z0.y
      const float2 z0 = ex[lane * TS::EX];
      float* o = lane ? pw1 : pw;
      o[0] = (z0.x + z0.y) * (z0.x + z0.y);
      o[NC] = (z0.x - z0.y) * (z0.x - z0.y);
    }
    __syncwarp();
    const bool two = f0 + 1 < P.n_frames;
    const size_t orow = (size_t)b * P.n_frames + f0;
    for (int m = lane; m < P.n_mels; m += 32) {
      const int s0 = __ldg(P.mel_start + m), cnt = __ldg(P.mel_cnt + m);
      const float* w = P.mel_w + __ldg(P.mel_off + m);
      float acc0 = 0.f, acc1 = 0.f;
      for (int i = 0; i < cnt; ++i) {
        const float wi = __ldg(w + i);
        acc0 = fmaf(pw[s0 + i], wi, acc0);
        acc1 = fmaf(pw1[s0 + i], wi, acc1);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && !two) break;
        const float v = __logf(fmaxf(h ? acc1 : acc0, P.log_floor));     // MUFU.LG2 * ln 2: abs error ~1e-6 over the log-mel range, 16x fewer instructions
        const size_t row = orow + h;
        if (P.out_f32) P.out_f32[row * P.n_mels + m] = v;
        if (P.out_op) {
          bf16* o = P.out_op + row * P.op_ld + m;
          const bf16 hv = __float2bfloat16_rn(v);
          *o = hv;
          if (P.op_lo_off) o[P.op_lo_off] = __float2bfloat16_rn(v - __bfloat162float(hv));
        }
      }
    }
    __syncwarp();
  }
}

template <int NC, typename PcmT, bool KALDI>
int launch(const FbankParams& P, int n_streams, cudaStream_t st) {
  const size_t smem = sizeof(float2) * (NC + 2) + sizeof(float) * ((P.frame_len + 3) & ~3) + (size_t)kWarps * FbSmem<NC>::WARP_BYTES +
                      ((P.n_samples * sizeof(PcmT) + 15) & ~(size_t)15) + (KALDI ? sizeof(float) * (size_t)(P.n_samples / (P.blk > 0 ? P.blk : 1) + 1) : 0);
  static size_t attr_done[kMaxDevices] = {0};
  ASR_CUDA_OK(ensure_dyn_smem(fbank_kernel<NC, PcmT, KALDI>, smem, attr_done));
  ASR_CUDA_OK(launch_pdl(fbank_kernel<NC, PcmT, KALDI>, dim3(n_streams), dim3(kWarps * 32), smem, st, P));
  return 0;
}

}  // namespace

static int gcd_int(int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; }

int fbank_launch(const FbankParams& P_in, int n_streams, cudaStream_t st) {
  if (n_streams <= 0) return 0;
  FbankParams P = P_in;
  P.blk = gcd_int(P.hop, P.frame_len);
  if (P.kaldi && (P.blk & 1)) { set_error("fbank: hop and frame length must share an even divisor"); return -1; }
  const int need = (P.n_frames - 1) * P.hop + P.frame_off + P.frame_len;
  if (need > P.n_samples || P.frame_len > 2 * P.nc || (P.frame_len & 1) || (P.hop & 1) || (P.frame_off & 1)) { set_error("fbank: bad geometry"); return -1; }
  if (P.nc == 400 && !P.kaldi) return P.pcm_is_f32 ? launch<400, float, false>(P, n_streams, st) : launch<400, int16_t, false>(P, n_streams, st);
  if (P.nc == 256 && P.kaldi) return P.pcm_is_f32 ? launch<256, float, true>(P, n_streams, st) : launch<256, int16_t, true>(P, n_streams, st);
  set_error("fbank: unsupported plan nc=%d kaldi=%d", P.nc, P.kaldi);
  return -1;
}

}  // namespace asr
