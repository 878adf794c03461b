// In-register radix-2 DIF FFTs (N = 2..32) with compile-time twiddles.
//
// Everything is a fully unrolled template over a float2 array held in registers; twiddles
// are float literals rounded from float64 (cos/sin of multiples of 2*pi/32), and rotations by
// multiples of pi/2 and pi/4 are special-cased so no multiply by 0 or 1 is ever issued.
// The same code compiles for the host (tests/host_fft_check.cpp) so the index arithmetic of
// the warp-level 512-point transform is pinned on the CPU test-suite.
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define B200_HD __host__ __device__ __forceinline__
#define B200_CHD __host__ __device__
#else
#include <cmath>
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
#define B200_HD inline
#define B200_CHD
#endif

namespace b200 {

// cos(2*pi*k/32), k = 0..8, rounded from float64
template <int K> struct Cos32;
template <> struct Cos32<0> { static constexpr float v = 1.0f; };
template <> struct Cos32<1> { static constexpr float v = 0.98078528040323043f; };
template <> struct Cos32<2> { static constexpr float v = 0.92387953251128674f; };
template <> struct Cos32<3> { static constexpr float v = 0.83146961230254524f; };
template <> struct Cos32<4> { static constexpr float v = 0.70710678118654752f; };
template <> struct Cos32<5> { static constexpr float v = 0.55557023301960218f; };
template <> struct Cos32<6> { static constexpr float v = 0.38268343236508977f; };
template <> struct Cos32<7> { static constexpr float v = 0.19509032201612825f; };
template <> struct Cos32<8> { static constexpr float v = 0.0f; };

// cos / sin of 2*pi*K/32 for K in [0, 32) via quadrant symmetry
template <int K> struct CS32 {
  static constexpr int k = ((K % 32) + 32) % 32;
  static constexpr int q = k / 8, r = k % 8;
  // angle = q*90deg + r*11.25deg
  static constexpr float c0 = Cos32<r>::v, s0 = Cos32<8 - r>::v;   // cos, sin of the residual
  static constexpr float c = (q == 0) ? c0 : (q == 1) ? -s0 : (q == 2) ? -c0 : s0;
  static constexpr float s = (q == 0) ? s0 : (q == 1) ? c0 : (q == 2) ? -s0 : -c0;
};

#ifdef __CUDACC__
// Packed FP32 pairs on sm_100 (add/sub/mul/fma .f32x2 -> FADD2 / FMUL2 / FFMA2).  Both halves round to nearest like
// the scalar instructions, so every packed form below is bit-identical to the scalar expression next to it; ptxas
// folds the half swaps (.LO_HI), per-half sign patterns (.NP / .PN) and scalar broadcasts (.F32) of the pack / unpack
// moves into operand modifiers, so a complex multiply by a constant is TWO instructions instead of four.
typedef unsigned long long fx2;
__device__ __forceinline__ fx2 fx_pk(float lo, float hi) { fx2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float2 fx_upk(fx2 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ fx2 fx_add(fx2 a, fx2 b) { fx2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ fx2 fx_sub(fx2 a, fx2 b) { fx2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ fx2 fx_mul(fx2 a, fx2 b) { fx2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ fx2 fx_fma(fx2 a, fx2 b, fx2 c) { fx2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
// (a.x c + a.y s, a.y c - a.x s) = fma((a.y, a.x), (s, -s), (a.x, a.y) * (c, c)): the roundings of the scalar form
__device__ __forceinline__ float2 fx_cmul_conj(float2 a, float c, float s) {
  return fx_upk(fx_fma(fx_pk(a.y, a.x), fx_pk(s, -s), fx_mul(fx_pk(a.x, a.y), fx_pk(c, c))));
}
#endif

// a * exp(-2*pi*i * K / 32)
template <int K> B200_HD float2 mul_w32(float2 a) {
  constexpr int k = ((K % 32) + 32) % 32;
  if constexpr (k == 0) return a;
  else if constexpr (k == 8) return make_float2(a.y, -a.x);
  else if constexpr (k == 16) return make_float2(-a.x, -a.y);
  else if constexpr (k == 24) return make_float2(-a.y, a.x);
#ifdef __CUDA_ARCH__
  // rotations by odd multiples of pi/4: ((a.x + a.y) h, (a.y - a.x) h) etc. as one FADD2 (swap + sign pattern) + one FMUL2
  else if constexpr (k == 4) { constexpr float h = Cos32<4>::v; return fx_upk(fx_mul(fx_add(fx_pk(a.x, a.y), fx_pk(a.y, -a.x)), fx_pk(h, h))); }
  else if constexpr (k == 12) { constexpr float h = Cos32<4>::v; return fx_upk(fx_mul(fx_sub(fx_pk(a.y, -a.x), fx_pk(a.x, a.y)), fx_pk(h, h))); }
  else if constexpr (k == 20) { constexpr float h = Cos32<4>::v; return fx_upk(fx_mul(fx_add(fx_pk(a.x, a.y), fx_pk(a.y, -a.x)), fx_pk(-h, -h))); }
  else if constexpr (k == 28) { constexpr float h = Cos32<4>::v; return fx_upk(fx_mul(fx_sub(fx_pk(a.x, a.y), fx_pk(a.y, -a.x)), fx_pk(h, h))); }
  else { return fx_cmul_conj(a, CS32<k>::c, CS32<k>::s); }
#else
  else if constexpr (k == 4) { constexpr float h = Cos32<4>::v; return make_float2((a.x + a.y) * h, (a.y - a.x) * h); }
  else if constexpr (k == 12) { constexpr float h = Cos32<4>::v; return make_float2((a.y - a.x) * h, -(a.x + a.y) * h); }
  else if constexpr (k == 20) { constexpr float h = Cos32<4>::v; return make_float2(-(a.x + a.y) * h, (a.x - a.y) * h); }
  else if constexpr (k == 28) { constexpr float h = Cos32<4>::v; return make_float2((a.x - a.y) * h, (a.x + a.y) * h); }
  else {
    constexpr float c = CS32<k>::c, s = CS32<k>::s;
    // (x + iy)(c - is) = (xc + ys) + i(yc - xs)
    return make_float2(std::fmaf(a.y, s, a.x * c), std::fmaf(-a.x, s, a.y * c));
  }
#endif
}

template <int N> struct BitRev;   // bit reversal of I over log2(N) bits
template <int N, int I> struct BitRevI {
  static constexpr int bits = (N == 2) ? 1 : (N == 4) ? 2 : (N == 8) ? 3 : (N == 16) ? 4 : 5;
  static constexpr int rev(int i) { int r = 0; for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b); return r; }
  static constexpr int v = rev(I);
};
B200_CHD constexpr int bitrev_n(int i, int bits) { int r = 0; for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b); return r; }
B200_CHD constexpr int log2_n(int n) { int b = 0; while ((1 << b) < n) ++b; return b; }

// Complex add / subtract.  On sm_100 each is ONE packed instruction (add.rn.f32x2 -> FADD2): Blackwell's
// dual-FP32 path halves the issue slots of the butterfly adds.
B200_HD float2 cadd(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
B200_HD float2 csub(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
#else
  return make_float2(a.x - b.x, a.y - b.y);
#endif
}

// One DIF stage: blocks of size 2*H starting at BASE; twiddle step = 32 / (2*H).
template <int N, int H, int BASE, int J> struct DifButterfly {
  static B200_HD void run(float2 (&v)[N]) {
    float2 a = v[BASE + J], b = v[BASE + J + H];
    v[BASE + J] = cadd(a, b);
    v[BASE + J + H] = mul_w32<J * (16 / H)>(csub(a, b));
    if constexpr (J + 1 < H) DifButterfly<N, H, BASE, J + 1>::run(v);
  }
};
template <int N, int H, int BASE> struct DifBlocks {
  static B200_HD void run(float2 (&v)[N]) {
    DifButterfly<N, H, BASE, 0>::run(v);
    if constexpr (BASE + 2 * H < N) DifBlocks<N, H, BASE + 2 * H>::run(v);
  }
};
template <int N, int H> struct DifStages {
  static B200_HD void run(float2 (&v)[N]) {
    DifBlocks<N, H, 0>::run(v);
    if constexpr (H > 1) DifStages<N, H / 2>::run(v);
  }
};

// In-place forward DFT of N points (N in {2,4,8,16,32}); X[k] ends up in v[bitrev(k)].
template <int N> B200_HD void fft_dif(float2 (&v)[N]) { DifStages<N, N / 2>::run(v); }

}  // namespace b200
