// Micro-benchmark: issue / pipe throughput of scalar vs packed FP32 on sm_100a.
//   FFMA      fma.rn.f32            1 FMA / lane / instruction
//   FFMA2     fma.rn.f32x2          2 FMA / lane / instruction
//   FADD2     add.rn.f32x2
//   MIX       alternating FFMA2 and scalar FFMA (does the scalar op use a second pipe?)
//   MIXLDS    FFMA2 with one conflict-free LDS.32 per 2 FFMA2 (the resampler's ratio)
// Prints warp-instructions / clk / SM and FMA lane-ops / clk / SM for 4..16 warps per SM.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o exp_pipes exp_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

constexpr int ILP = 8, ITERS = 2048;

template <int KIND>
__global__ void k(float* out, float seed) {
  __shared__ float sm[1024];
  sm[threadIdx.x & 1023] = seed;
  __syncthreads();
  unsigned long long a2[ILP];
  float a1[ILP];
  const unsigned long long b2 = ((unsigned long long)__float_as_uint(seed) << 32) | __float_as_uint(seed + 1.f);
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a2[i] = b2 + i; a1[i] = seed + i; }
  const float* sp = sm + (threadIdx.x & 31);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (KIND == 0) a1[i] = fma1(a1[i], seed, a1[i]);
      if (KIND == 1) a2[i] = fma2(a2[i], b2, a2[i]);
      if (KIND == 2) a2[i] = add2(a2[i], b2);
      if (KIND == 3) { a2[i] = fma2(a2[i], b2, a2[i]); a1[i] = fma1(a1[i], seed, a1[i]); }
      if (KIND == 4) {
        a2[i] = fma2(a2[i], b2, a2[i]);
        if (i & 1) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(sp + 32 * (i + (it & 7))))); a1[i] += v; }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a1[i] + __uint_as_float((unsigned)a2[i]) + __uint_as_float((unsigned)(a2[i] >> 32));
  if (s == 12345.678f) out[0] = s;
}

template <int KIND>
void run(const char* name, int instr_per_iter_slot, int fma_per_iter_slot) {
  float* d;
  cudaMalloc(&d, 4);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int warps = 4; warps <= 16; warps += 4) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND><<<148, warps * 32>>>(d, 1.0f);
    cudaEventRecord(e0);
    for (int r = 0; r < 10; ++r) k<KIND><<<148, warps * 32>>>(d, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms / 10 * 1e-3 * clk_khz * 1e3;
    const double winst = (double)warps * ITERS * ILP * instr_per_iter_slot;
    printf("%-7s warps/SM %2d: %.3f warp-instr/clk/SM, %.1f FMA-lanes/clk/SM (%.3f ms)\n", name, warps, winst / cycles,
           (double)warps * ITERS * ILP * fma_per_iter_slot * 32 / cycles, ms / 10);
  }
}

int main() {
  run<0>("FFMA", 1, 1);
  run<1>("FFMA2", 1, 2);
  run<2>("FADD2", 1, 2);
  run<3>("MIX", 2, 3);
  run<4>("MIXLDS", 1, 2);
  return 0;
}
