// Host harness for asr_streaming_b200/csrc/fft_regs.cuh (tests/test_fft_regs.py): emulates the 32 lanes of a warp running the
// two-step register FFT on two frames and prints the spectra; also the bare DFT-16 / DFT-25.  Built with g++ (no GPU involved).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <vector_types.h>
#include <vector_functions.h>
#define ASR_HD static inline
#include "fft_regs.cuh"
using namespace asr::fftr;

template <int NC>
static void run(const std::vector<float2>& z /* [2][NC] */, std::vector<float2>& out) {
  typedef TwoStep<NC> TS;
  std::vector<float2> ex(2 * TS::EX, make_float2(0.f, 0.f));
  for (int r = 0; r < TS::ROUNDS; ++r)
    for (int lane = 0; lane < 32; ++lane) {
      int h, n2; bool act;
      step1_slot<NC>(lane, r, h, n2, act);
      if (!act) continue;
      float2 a[16], tw[16];
      for (int n1 = 0; n1 < 16; ++n1) a[n1] = z[h * NC + n1 * TS::N2 + n2];
      for (int k1 = 0; k1 < 16; ++k1) { const double ang = -2.0 * M_PI * (n2 * k1) / NC; tw[k1] = make_float2((float)cos(ang), (float)sin(ang)); }
      step1<NC>(a, tw, ex.data() + h * TS::EX, n2);
    }
  float2 X[32][TS::N2];
  for (int lane = 0; lane < 32; ++lane) step2_compute<NC>(ex.data() + (lane >> 4) * TS::EX, lane & 15, X[lane]);
  for (int lane = 0; lane < 32; ++lane) step2_store<NC>(ex.data() + (lane >> 4) * TS::EX, lane & 15, X[lane]);
  out.resize(2 * NC);
  for (int h = 0; h < 2; ++h) for (int k = 0; k < NC; ++k) out[h * NC + k] = ex[h * TS::EX + k];
}

int main(int argc, char** argv) {
  const int nc = argc > 1 ? atoi(argv[1]) : 256;
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 32768.0f - 1.0f; };
  if (nc == 16 || nc == 25) {
    float2 a16[16], o16[16], a25[25], o25[25];
    for (int i = 0; i < 25; ++i) { a25[i] = make_float2(rnd(), rnd()); if (i < 16) a16[i] = a25[i]; }
    if (nc == 16) { dft16<16>(a16, o16); for (int i = 0; i < 16; ++i) printf("%.9g %.9g %.9g %.9g\n", a16[i].x, a16[i].y, o16[i].x, o16[i].y); }
    else { dft25(a25, o25); for (int i = 0; i < 25; ++i) printf("%.9g %.9g %.9g %.9g\n", a25[i].x, a25[i].y, o25[i].x, o25[i].y); }
    return 0;
  }
  std::vector<float2> z(2 * nc), out;
  for (int h = 0; h < 2; ++h)
    for (int m = 0; m < nc; ++m) z[h * nc + m] = (nc == 400 && m >= 200) ? make_float2(0.f, 0.f) : make_float2(rnd(), rnd());
  if (nc == 256) run<256>(z, out); else run<400>(z, out);
  for (int i = 0; i < 2 * nc; ++i) printf("%.9g %.9g %.9g %.9g\n", z[i].x, z[i].y, out[i].x, out[i].y);
  return 0;
}
