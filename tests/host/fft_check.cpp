// Host emulation of the warp-level 512-point two-for-one FFT used by fbank_fast.cuh:
// runs the SAME templated register FFTs (csrc/fft_regs.cuh) lane by lane with the same
// exchange / partner index arithmetic and prints the power spectra of four real frames.
// tests/test_host_fft.py compares them with numpy.  Build: g++ -O2 -std=c++17.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../dl_sound_classification_b200/csrc/fft_regs.cuh"
using namespace b200;

int main(int argc, char** argv) {
  // four real frames of 512 samples from stdin (binary float32), output 4 x 256 powers
  std::vector<float> in(4 * 512);
  if (fread(in.data(), 4, in.size(), stdin) != in.size()) return 1;
  static float2 E[512];
  float2 Y[2][32][16];   // [fft][lane][slot]  stage-1 results (slot = register index)
  float2 U[32][32];      // [lane][n2]
  // stage 1: lane l holds n = l + 32 j
  for (int f = 0; f < 2; ++f)
    for (int l = 0; l < 32; ++l) {
      float2 z[16];
      for (int j = 0; j < 16; ++j) z[j] = make_float2(in[(2 * f) * 512 + l + 32 * j], in[(2 * f + 1) * 512 + l + 32 * j]);
      fft_dif<16>(z);
      for (int k1 = 0; k1 < 16; ++k1) {
        int slot = bitrev_n(k1, 4);
        double ang = -2.0 * M_PI * (double)(l * k1) / 512.0;
        float2 w = make_float2((float)cos(ang), (float)sin(ang));
        float2 a = z[slot];
        z[slot] = make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
      }
      for (int s = 0; s < 16; ++s) Y[f][l][s] = z[s];
    }
  // exchange through E (one FFT at a time), stage-2 gather: lane = k1 + 16 * fft
  for (int f = 0; f < 2; ++f) {
    for (int l = 0; l < 32; ++l)
      for (int k1 = 0; k1 < 16; ++k1) E[k1 * 32 + (l ^ k1)] = Y[f][l][bitrev_n(k1, 4)];
    for (int k1 = 0; k1 < 16; ++k1)
      for (int n2 = 0; n2 < 32; ++n2) U[k1 + 16 * f][n2] = E[k1 * 32 + (n2 ^ k1)];
  }
  for (int lane = 0; lane < 32; ++lane) {
    float2 u[32];
    for (int i = 0; i < 32; ++i) u[i] = U[lane][i];
    fft_dif<32>(u);
    for (int i = 0; i < 32; ++i) U[lane][i] = u[i];     // X[k1 + 16 k2] = u[bitrev5(k2)]
  }
  std::vector<float> P(4 * 256);
  for (int lane = 0; lane < 32; ++lane) {
    int k1 = lane & 15, f = lane >> 4;
    int src = ((16 - k1) & 15) + 16 * f;
    for (int k2 = 0; k2 < 16; ++k2) {
      float2 zk = U[lane][bitrev_n(k2, 5)];
      float2 zp = (k1 == 0) ? U[lane][bitrev_n((32 - k2) & 31, 5)] : U[src][bitrev_n(31 - k2, 5)];
      float ar = zk.x + zp.x, ai = zk.y - zp.y, br = zk.y + zp.y, bi = zp.x - zk.x;
      int k = k1 + 16 * k2;
      P[(2 * f) * 256 + k] = 0.25f * (ar * ar + ai * ai);
      P[(2 * f + 1) * 256 + k] = 0.25f * (br * br + bi * bi);
    }
  }
  fwrite(P.data(), 4, P.size(), stdout);
  return 0;
}
