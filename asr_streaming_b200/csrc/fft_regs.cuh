// Register-resident DFT-16 and DFT-25 (forward, exp(-2*pi*i*n*k/N)) built from radix-4 / radix-5 butterflies, for the fbank
// kernels' two-step FFTs (256 = 16 x 16, 400 = 16 x 25).  Host-compilable (tests/test_fft_regs.py runs them with g++ against
// numpy) — every index below is a compile-time constant after unrolling, so the arrays live in registers on the device.
#pragma once
#ifndef ASR_HD
#define ASR_HD __host__ __device__ __forceinline__
#endif

namespace asr {
namespace fftr {

ASR_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
ASR_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
ASR_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// in-place forward DFT-4, natural order
ASR_HD void dft4(float2& v0, float2& v1, float2& v2, float2& v3) {
  const float2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3);
  const float2 d = csub(v1, v3);
  const float2 a3 = make_float2(d.y, -d.x);          // -i * (v1 - v3)
  v0 = cadd(a0, a2); v1 = cadd(a1, a3); v2 = csub(a0, a2); v3 = csub(a1, a3);
}

// in-place forward DFT-5, natural order
ASR_HD void dft5(float2& v0, float2& v1, float2& v2, float2& v3, float2& v4) {
  const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;   // cos(2pi/5), cos(4pi/5)
  const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;    // sin(2pi/5), sin(4pi/5)
  const float2 p1 = cadd(v1, v4), p2 = cadd(v2, v3), d1 = csub(v1, v4), d2 = csub(v2, v3);
  const float2 A1 = make_float2(v0.x + c1 * p1.x + c2 * p2.x, v0.y + c1 * p1.y + c2 * p2.y);
  const float2 A2 = make_float2(v0.x + c2 * p1.x + c1 * p2.x, v0.y + c2 * p1.y + c1 * p2.y);
  const float2 B1 = make_float2(s1 * d1.x + s2 * d2.x, s1 * d1.y + s2 * d2.y);
  const float2 B2 = make_float2(s2 * d1.x - s1 * d2.x, s2 * d1.y - s1 * d2.y);
  v0 = make_float2(v0.x + p1.x + p2.x, v0.y + p1.y + p2.y);
  v1 = make_float2(A1.x + B1.y, A1.y - B1.x);        // A1 - i*B1
  v4 = make_float2(A1.x - B1.y, A1.y + B1.x);
  v2 = make_float2(A2.x + B2.y, A2.y - B2.x);
  v3 = make_float2(A2.x - B2.y, A2.y + B2.x);
}

template <int E> ASR_HD float2 w16() {
  if constexpr (E == 1) return make_float2(9.238795325e-01f, -3.826834324e-01f);
  if constexpr (E == 2) return make_float2(7.071067812e-01f, -7.071067812e-01f);
  if constexpr (E == 3) return make_float2(3.826834324e-01f, -9.238795325e-01f);
  if constexpr (E == 4) return make_float2(6.123233996e-17f, -1.000000000e+00f);
  if constexpr (E == 6) return make_float2(-7.071067812e-01f, -7.071067812e-01f);
  if constexpr (E == 9) return make_float2(-9.238795325e-01f, 3.826834324e-01f);
  return make_float2(1.f, 0.f);
}
template <int E> ASR_HD float2 w25() {
  if constexpr (E == 1) return make_float2(9.685831611e-01f, -2.486898872e-01f);
  if constexpr (E == 2) return make_float2(8.763066800e-01f, -4.817536741e-01f);
  if constexpr (E == 3) return make_float2(7.289686274e-01f, -6.845471059e-01f);
  if constexpr (E == 4) return make_float2(5.358267950e-01f, -8.443279255e-01f);
  if constexpr (E == 6) return make_float2(6.279051953e-02f, -9.980267284e-01f);
  if constexpr (E == 8) return make_float2(-4.257792916e-01f, -9.048270525e-01f);
  if constexpr (E == 9) return make_float2(-6.374239897e-01f, -7.705132428e-01f);
  if constexpr (E == 12) return make_float2(-9.921147013e-01f, -1.253332336e-01f);
  if constexpr (E == 16) return make_float2(-6.374239897e-01f, 7.705132428e-01f);
  return make_float2(1.f, 0.f);
}
// v * W16^E with the trivial cases folded
template <int E> ASR_HD float2 mul_w16(float2 v) {
  if constexpr (E == 0) return v;
  else if constexpr (E == 4) return make_float2(v.y, -v.x);
  else if constexpr (E == 2) return make_float2(0.70710678118654752f * (v.x + v.y), 0.70710678118654752f * (v.y - v.x));
  else if constexpr (E == 6) return make_float2(0.70710678118654752f * (v.y - v.x), -0.70710678118654752f * (v.x + v.y));
  else return cmul(v, w16<E>());
}
template <int E> ASR_HD float2 mul_w25(float2 v) {
  if constexpr (E == 0) return v;
  else return cmul(v, w25<E>());
}

// Forward DFT-16 of a[0..16) (a[n] = 0 for n >= NNZ may be promised with NNZ = 8: the first radix-4 stage then has two inputs).
// n = 4 na + nb, k = ka + 4 kb:  W16^(nk) = W4^(na ka) * W16^(nb ka) * W4^(nb kb).  Result: out[k], natural order.
template <int NNZ>
ASR_HD void dft16(const float2 (&a)[16], float2 (&out)[16]) {
  float2 t[16];                                       // t[4 ka + nb]
#define ASR_DFT16_S1(NB)                                                                                   \
  {                                                                                                        \
    float2 x0 = a[NB], x1 = a[4 + NB], x2, x3;                                                             \
    if constexpr (NNZ > 8) { x2 = a[8 + NB]; x3 = a[12 + NB]; dft4(x0, x1, x2, x3); }                      \
    else { const float2 s = cadd(x0, x1), d = csub(x0, x1);                                                \
           x3 = make_float2(x0.x - x1.y, x0.y + x1.x); x1 = make_float2(x0.x + x1.y, x0.y - x1.x); x0 = s; x2 = d; } \
    t[NB] = x0; t[4 + NB] = mul_w16<NB * 1>(x1); t[8 + NB] = mul_w16<NB * 2>(x2); t[12 + NB] = mul_w16<NB * 3>(x3); \
  }
  ASR_DFT16_S1(0) ASR_DFT16_S1(1) ASR_DFT16_S1(2) ASR_DFT16_S1(3)
#undef ASR_DFT16_S1
#pragma unroll
  for (int ka = 0; ka < 4; ++ka) {
    dft4(t[4 * ka], t[4 * ka + 1], t[4 * ka + 2], t[4 * ka + 3]);          // over nb -> kb
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) out[ka + 4 * kb] = t[4 * ka + kb];
  }
}

// Forward DFT-25: n = 5 na + nb, k = ka + 5 kb:  W25^(nk) = W5^(na ka) * W25^(nb ka) * W5^(nb kb).
ASR_HD void dft25(const float2 (&a)[25], float2 (&out)[25]) {
  float2 t[25];                                       // t[5 ka + nb]
#define ASR_DFT25_S1(NB)                                                                                   \
  {                                                                                                        \
    float2 x0 = a[NB], x1 = a[5 + NB], x2 = a[10 + NB], x3 = a[15 + NB], x4 = a[20 + NB];                  \
    dft5(x0, x1, x2, x3, x4);                                                                              \
    t[NB] = x0; t[5 + NB] = mul_w25<NB * 1>(x1); t[10 + NB] = mul_w25<NB * 2>(x2); t[15 + NB] = mul_w25<NB * 3>(x3); \
    t[20 + NB] = mul_w25<NB * 4>(x4);                                                                      \
  }
  ASR_DFT25_S1(0) ASR_DFT25_S1(1) ASR_DFT25_S1(2) ASR_DFT25_S1(3) ASR_DFT25_S1(4)
#undef ASR_DFT25_S1
#pragma unroll
  for (int ka = 0; ka < 5; ++ka) {
    dft5(t[5 * ka], t[5 * ka + 1], t[5 * ka + 2], t[5 * ka + 3], t[5 * ka + 4]);
#pragma unroll
    for (int kb = 0; kb < 5; ++kb) out[ka + 5 * kb] = t[5 * ka + kb];
  }
}

// ------------------------------------------------------------------------------------------
// Two-step FFT of NC = 16 x N2 complex points (N2 = 16 or 25), TWO frames per warp, one exchange through shared memory:
//   X[k1 + 16 k2] = sum_n2 W_N2^(n2 k2) * [ W_NC^(n2 k1) * sum_n1 z[n1 N2 + n2] W_16^(n1 k1) ]
// step 1: a lane owns one n2 of one frame: DFT-16 over n1 in registers, twiddle, -> ex[h][k1 * ST + n2]
// step 2: a lane owns (frame h = lane / 16, k1 = lane % 16): DFT-N2 over n2 in registers -> ex[h][k1 + 16 k2] = X in natural order
// ST = 17 for N2 = 16 (padding: both the consecutive-n2 writes and the stride-ST reads are bank-conflict free), 25 for N2 = 25
// (25 float2 = 50 words: 16 lanes land on 16 different even banks).  NNZ8: only z[m], m < NC / 2, are non-zero (the reference's
// 400-sample window zero-padded to n_fft 800), so the DFT-16s of step 1 see 8 inputs.
// ------------------------------------------------------------------------------------------
template <int NC> struct TwoStep {
  static constexpr int N2 = NC / 16;
  static constexpr int ST = N2 == 16 ? 17 : N2;
  static constexpr int EX = 16 * ST;                  // float2 per frame in the exchange buffer (>= NC)
  static constexpr int ROUNDS = N2 == 16 ? 1 : 2;     // N2 = 16: lanes = (frame, n2); N2 = 25: one frame per round, lanes < 25
  static constexpr bool NNZ8 = NC == 400;
};

// lane's slot in step 1 of round r: frame h, column n2 (active == false: nothing to do)
template <int NC> ASR_HD void step1_slot(int lane, int r, int& h, int& n2, bool& active) {
  if (TwoStep<NC>::N2 == 16) { h = lane >> 4; n2 = lane & 15; active = true; }
  else { h = r; n2 = lane; active = lane < TwoStep<NC>::N2; }
}

// a[n1] = z[n1 * N2 + n2] already loaded (zeros beyond the frame); tw[k1] = W_NC^(n2 * k1)
template <int NC> ASR_HD void step1(const float2 (&a)[16], const float2 (&tw)[16], float2* ex_frame, int n2) {
  float2 A[16];
  dft16<TwoStep<NC>::NNZ8 ? 8 : 16>(a, A);
  ex_frame[n2] = A[0];
#pragma unroll
  for (int k1 = 1; k1 < 16; ++k1) ex_frame[k1 * TwoStep<NC>::ST + n2] = cmul(A[k1], tw[k1]);
}

template <int NC> ASR_HD void step2_compute(const float2* ex_frame, int k1, float2 (&X)[TwoStep<NC>::N2]) {
  constexpr int N2 = TwoStep<NC>::N2;
  float2 a[N2];
#pragma unroll
  for (int n2 = 0; n2 < N2; ++n2) a[n2] = ex_frame[k1 * TwoStep<NC>::ST + n2];
  if constexpr (N2 == 16) dft16<16>(a, X);
  else dft25(a, X);
}
template <int NC> ASR_HD void step2_store(float2* ex_frame, int k1, const float2 (&X)[TwoStep<NC>::N2]) {
#pragma unroll
  for (int k2 = 0; k2 < TwoStep<NC>::N2; ++k2) ex_frame[k1 + 16 * k2] = X[k2];
}

}  // namespace fftr
}  // namespace asr
