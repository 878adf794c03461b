// Micro-benchmark: variants of the 441->160 polyphase resampler phase (lanes = hops).
//   V0  taps broadcast from shared memory as float4, scalar FFMA      (fbank_fast.cuh round-1 v2)
//   V1  taps as immediates in the instruction stream, scalar FFMA      (per-group code, switch)
//   V2  taps from shared memory as float2 pairs, packed fma.rn.f32x2
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o exp_resample exp_resample.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "taps_441_160.inc"   // python tools/gen_taps_inc.py

constexpr int RP = 5, LT = 36, WIN = 46, NG = 32;
__constant__ float c_taps[NG * 180];
constexpr int XF = 32 * 441 + 480, RING = 34 * 161 + 2;

template <int G> __device__ __forceinline__ void dispatch_imm(int g, const float* xs, float* yo) {
  if constexpr (G < NG) {
    if (g == G) resample_group_imm<G>(xs, yo);
    else dispatch_imm<G + 1>(g, xs, yo);
  }
}

__device__ __forceinline__ void group_smem(const float4* __restrict__ T4, const float* __restrict__ xs, float* __restrict__ yo) {
  float acc[RP] = {0.f, 0.f, 0.f, 0.f, 0.f};
  float4 cur = make_float4(0, 0, 0, 0);
  int e = 0;
#pragma unroll
  for (int u = 0; u < WIN; ++u) {
    const float xv = xs[u];
#pragma unroll
    for (int r = 0; r < RP; ++r) {
      const int j = u - (r == 0 ? 0 : r == 1 ? 2 : r == 2 ? 4 : r == 3 ? 6 : 10);
      if (j >= 0 && j < LT) {
        if ((e & 3) == 0) cur = T4[e >> 2];
        const float tap = (e & 3) == 0 ? cur.x : (e & 3) == 1 ? cur.y : (e & 3) == 2 ? cur.z : cur.w;
        acc[r] = fmaf(tap, xv, acc[r]);
        ++e;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RP; ++r) yo[r] = acc[r];
}

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  return (unsigned long long)__float_as_uint(a) | ((unsigned long long)__float_as_uint(b) << 32);
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// taps per group stored [r][j] (180 floats); pairs (j, j+1) with x pairs (u, u+1), u even
__device__ __forceinline__ void group_f2(const float2* __restrict__ T2, const float* __restrict__ xs, float* __restrict__ yo) {
  unsigned long long acc[RP] = {0, 0, 0, 0, 0};
#pragma unroll
  for (int u = 0; u < WIN; u += 2) {
    const unsigned long long xp = pack2(xs[u], xs[u + 1]);
#pragma unroll
    for (int r = 0; r < RP; ++r) {
      const int j = u - (r == 0 ? 0 : r == 1 ? 2 : r == 2 ? 4 : r == 3 ? 6 : 10);
      if (j >= 0 && j < LT) {
        const float2 t = T2[(r * LT + j) >> 1];
        acc[r] = fma2(xp, pack2(t.x, t.y), acc[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RP; ++r) yo[r] = __uint_as_float((unsigned)acc[r]) + __uint_as_float((unsigned)(acc[r] >> 32));
}

// V3: taps through the constant bank (LDC, a path separate from the shared-memory pipe)
__device__ __forceinline__ void group_const(int g, const float* __restrict__ xs, float* __restrict__ yo) {
  float acc[RP] = {0.f, 0.f, 0.f, 0.f, 0.f};
  const float* T = c_taps + g * 180;
#pragma unroll
  for (int u = 0; u < WIN; ++u) {
    const float xv = xs[u];
#pragma unroll
    for (int r = 0; r < RP; ++r) {
      const int j = u - (r == 0 ? 0 : r == 1 ? 2 : r == 2 ? 4 : r == 3 ? 6 : 10);
      if (j >= 0 && j < LT) acc[r] = fmaf(T[r * LT + j], xv, acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < RP; ++r) yo[r] = acc[r];
}

template <int V>
__global__ void __launch_bounds__(256, 2) kern(const float* __restrict__ x, const float* __restrict__ taps, float* __restrict__ out, int nchunks) {
  extern __shared__ __align__(16) float smem[];
  float* A = smem;
  float* ring = A + XF;
  float* st = ring + RING;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < XF; i += 256) A[i] = x[(blockIdx.x * 977 + i) % 100000];
  if (V == 0 || V == 2) for (int i = tid; i < NG * 180; i += 256) st[i] = taps[i];
  __syncthreads();
  float chk = 0.f;
  for (int c = 0; c < nchunks; ++c) {
    const float* xl = A + lane * 441;
    float* yl = ring + (2 + lane) * 161;
#pragma unroll 1
    for (int gi = 0; gi < 4; ++gi) {
      const int g = warp * 4 + gi;
      if (V == 0) group_smem(reinterpret_cast<const float4*>(st + g * 180), xl + 1 + g * 13, yl + RP * g);
      else if (V == 1) dispatch_imm<0>(g, xl, yl);
      else if (V == 3) group_const(g, xl + 1 + g * 13, yl + RP * g);
      else group_f2(reinterpret_cast<const float2*>(st + g * 180), xl + 1 + g * 13, yl + RP * g);
    }
    __syncthreads();
    chk += ring[(tid * 7 + c) % (34 * 161)];
    A[(tid * 13 + c) % XF] += chk * 1e-9f;
    __syncthreads();
  }
  out[blockIdx.x * 256 + tid] = chk;
}

template <int V> float run(const float* x, const float* taps, float* out, int grid, int nchunks) {
  size_t smem = (size_t)(XF + RING + ((V == 0 || V == 2) ? NG * 180 : 0)) * 4;
  cudaFuncSetAttribute(kern<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<V><<<grid, 256, smem>>>(x, taps, out, nchunks);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) kern<V><<<grid, 256, smem>>>(x, taps, out, nchunks);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
  return ms / 5;
}

int main() {
  float *x, *taps, *out;
  cudaMalloc(&x, 100000 * 4); cudaMalloc(&taps, NG * 180 * 4); cudaMalloc(&out, 1 << 22);
  std::vector<float> h(100000); for (auto& v : h) v = rand() / (float)RAND_MAX - 0.5f;
  cudaMemcpy(x, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  std::vector<float> t(NG * 180);
  for (int g = 0; g < NG; ++g) for (int r = 0; r < RP; ++r) for (int j = 0; j < LT; ++j) t[g * 180 + r * LT + j] = kImmTaps[RP * g + r][j];
  cudaMemcpy(taps, t.data(), t.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpyToSymbol(c_taps, t.data(), t.size() * 4);
  const int grid = 296, nch = 64;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float m0 = run<0>(x, taps, out, grid, nch), m1 = run<1>(x, taps, out, grid, nch), m2 = run<2>(x, taps, out, grid, nch);
  float m3 = run<3>(x, taps, out, grid, nch);
  printf("V3 const-bank   %.3f ms  %.1f SM-cycles/frame\n", m3, m3 * 1e-3 * clk * 1e3 / (2.0 * nch * 32));
  // each SM runs 2 CTAs x nch chunks of 32 frames per launch
  auto cyc = [&](float ms) { return ms * 1e-3 * clk * 1e3 / (2.0 * nch * 32); };
  printf("clock %d kHz\nV0 smem-float4  %.3f ms  %.1f SM-cycles/frame\nV1 immediates   %.3f ms  %.1f SM-cycles/frame\nV2 smem f32x2   %.3f ms  %.1f SM-cycles/frame\n",
         clk, m0, cyc(m0), m1, cyc(m1), m2, cyc(m2));
  return 0;
}
