// Micro-benchmark: how long does ONE cp.async.bulk of a resampler input chunk take per SM, and what does an L2 prefetch buy?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/build/exp_tma tools/exp_tma.cu && tools/build/exp_tma
// Every CTA (one per SM) copies chunks of `bytes` into shared memory, `gap` cycles of busy-waiting between copies (the
// resampling of the chunk).  mode 0: plain; 1: bulk L2 prefetch of the NEXT chunk issued at the start of the gap;
// 2: per-line prefetch.global.L2 of the next chunk by all threads; 3: warm (the same chunk every time: L2 hits).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}

__global__ void fill(unsigned* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[n + i] = (unsigned)(i * 2654435761u) ^ (unsigned)(i >> 7);
}

__global__ void __launch_bounds__(128, 1) k(const float* __restrict__ src, size_t clip_floats, int chunks, int bytes, int gap, int mode,
                                            unsigned long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar), dst = (uint32_t)__cvta_generic_to_shared(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned long long wait = 0;
  unsigned parity = 0;
  const int step = bytes / 4 - 36;                               // consecutive chunks overlap by ~34 samples (kept 16-byte aligned)
  for (int c = 0; c < chunks; ++c) {
    // clip_floats == 220500: the real kernel's walk (CTA x takes clips x, x + 148, ..., 15 chunks of 32 hops per clip)
    const bool real = clip_floats == 220500;
    const float* clip = real ? src + (size_t)(blockIdx.x + gridDim.x * (c / 15)) * clip_floats : src + (size_t)blockIdx.x * clip_floats;
    const int cc = real ? c % 15 : c, cn = real ? (c + 1) % 15 : c + 1;
    const float* clipn = real ? src + (size_t)(blockIdx.x + gridDim.x * ((c + 1) / 15)) * clip_floats : clip;
    const float* g = clip + (mode == 3 ? 0 : (size_t)cc * (real ? 14112 : step));
    const float* gn = clipn + (size_t)cn * (real ? 14112 : step);
    long long t0 = 0;
    if (threadIdx.x == 0) {
      t0 = clock64();
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(g), "r"(bytes), "r"(bar_a) : "memory");
    }
    mbar_wait(bar_a, parity);
    parity ^= 1u;
    if (threadIdx.x == 0) wait += (unsigned long long)(clock64() - t0);
    if (mode == 1 && threadIdx.x == 32 && c + 1 < chunks)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gn), "r"(bytes) : "memory");
    if (mode == 2 && c + 1 < chunks)
      for (int l = threadIdx.x; l < bytes / 128; l += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(gn + l * 32));
    const long long t1 = clock64();
    while (clock64() - t1 < gap) {}
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out, wait);
}

int main() {
  const int sms = 148, chunks = 64;
  const size_t clip_floats = (size_t)chunks * 14200 + 64;
  float* src;
  unsigned long long* out;
  cudaMalloc(&src, sms * clip_floats * 4 * 8);                   // 8 disjoint sets: every run reads cold data
  cudaMemset(src, 0, sms * clip_floats * 4 * 8);
  fill<<<4096, 256>>>(reinterpret_cast<unsigned*>(src), sms * clip_floats * 4);   // second half of the sets: pseudo-random bits instead of zeros
  cudaMalloc(&out, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  int set = 0;
  for (int bytes : {14144, 28288, 56576})
    for (int gap : {0, 6000})
      for (int mode = 0; mode < 4; ++mode) {
        if (gap == 0 && (mode == 1 || mode == 2)) continue;
        cudaMemset(out, 0, 8);
        k<<<sms, 128, 60 * 1024>>>(src + (size_t)(set++ % 8) * sms * clip_floats, clip_floats, chunks, bytes, gap, mode, out);
        unsigned long long w = 0;
        cudaMemcpy(&w, out, 8, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        printf("bytes %6d gap %5d mode %d (%s): %7.0f cycles per copy  %s\n", bytes, gap, mode,
               mode == 0 ? "plain" : mode == 1 ? "bulk L2 prefetch" : mode == 2 ? "line L2 prefetch" : "warm/L2-resident",
               (double)w / (sms * chunks), e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  // random data (sets 4..7) and the real kernel's clip walk (1036 clips of 220500 floats, 105 chunks per CTA)
  for (int rnd = 0; rnd < 2; ++rnd)
    for (int mode : {0, 1}) {
      cudaMemset(out, 0, 8);
      k<<<sms, 128, 60 * 1024>>>(src + (size_t)(rnd ? 4 : 0) * sms * clip_floats, clip_floats, chunks, 56576, 6000, mode, out);
      unsigned long long w = 0;
      cudaMemcpy(&w, out, 8, cudaMemcpyDeviceToHost);
      printf("%s data, 56576 B, gap 6000, mode %d: %7.0f cycles per copy\n", rnd ? "random" : "zero", mode, (double)w / (sms * chunks));
    }
  float* real;
  cudaMalloc(&real, (size_t)1036 * 220500 * 4);
  fill<<<4096, 256>>>(reinterpret_cast<unsigned*>(real) - (size_t)1036 * 220500 / 2 * 0, 0);
  cudaMemset(real, 0x3c, (size_t)1036 * 220500 * 4);
  for (int gap : {3500, 6000})
    for (int mode : {0, 1}) {
      cudaMemset(out, 0, 8);
      k<<<sms, 128, 60 * 1024>>>(real, 220500, 105, 56576, gap, mode, out);
      unsigned long long w = 0;
      cudaMemcpy(&w, out, 8, cudaMemcpyDeviceToHost);
      printf("real clip walk (1036 clips x 882000 B), 56576 B, gap %d, mode %d: %7.0f cycles per copy  %s\n", gap, mode, (double)w / (sms * 105),
             cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
