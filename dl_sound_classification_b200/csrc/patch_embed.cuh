// AST patch embedding fed straight from the frontend's features:
//   Conv2d(1, D, kernel 16, stride s) on (B, 1, F, T)  ->  flatten(2).transpose(1, 2)  ->  (B, Fp*Tp, D)
// Reference: PatchEmbed.forward  src/models/ast_mini.py:7-15 (ast_small.py:7-15), ASTModel.forward src/models/ast.py:30,50-56.
//
// It is an im2col GEMM  C[M = B*Fp*Tp, N = D] = A[M, K = 256] * W^T + bias  with A never materialised in HBM: the CTA gathers
// the 16 x 16 patches of 128 consecutive output rows from the (L2-resident) features, rounds them to fp16 -- the reference's
// recommended AST setting is precision "16-mixed" (configs/base_training.yaml:32,48), i.e. this convolution runs in fp16
// autocast there too -- and lays them out in shared memory in the K-major 128-byte-swizzled canonical form the 5th-gen tensor
// core reads.  One elected thread issues tcgen05.mma (M = 128, N = 192, K = 16 per instruction, fp32 accumulation in TMEM);
// the epilogue reads the accumulator back with tcgen05.ld, adds the bias and stores fp16 or fp32 rows.
// A CTA keeps its 192 x 256 weight tile resident in shared memory and walks over the row tiles (persistent grid).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr int PE_THREADS = 256;
constexpr int PE_M = 128, PE_N = 192, PE_K = 256;
constexpr int PE_KB = 64;                              // fp16 elements per 128-byte swizzle row = one k-block
constexpr int PE_NKB = PE_K / PE_KB;                   // 4
constexpr int PE_A_KB_BYTES = PE_M * 128;              // 16 KB per k-block of A
constexpr int PE_B_KB_BYTES = PE_N * 128;              // 24 KB per k-block of W
constexpr int PE_TMEM_COLS = 256;                      // power of two >= PE_N
constexpr size_t PE_SMEM = 1024 + (size_t)PE_NKB * (PE_A_KB_BYTES + PE_B_KB_BYTES) + 64 + PE_N * 4 + 8 * 32 * 208;

struct PatchEmbedParams {
  const float* feat;       // (B, 1, F, T)
  const __half* w;         // (D, 256) fp16: weight[d][0][i][j] at 16 i + j
  const float* bias;       // (D) or nullptr
  void* out;               // (B, Fp * Tp, D) fp16 or fp32
  int B, F, T, D, stride, Fp, Tp, out_f16;
  int64_t M;               // B * Fp * Tp
};

// byte offset of 16-byte chunk c (0..7) of row r inside a K-major SWIZZLE_128B tile (rows of 128 B, 8-row groups of 1 KB)
__device__ __forceinline__ uint32_t pe_swz(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

// UMMA shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): K-major, SWIZZLE_128B,
// stride byte offset 1024 (8 rows x 128 B), leading byte offset unused (1), version 1 (Blackwell)
__device__ __forceinline__ uint64_t pe_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                               // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                               // version
  d |= (uint64_t)2 << 61;                               // layout type SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void pe_mbar_wait(uint32_t bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PE_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra PE_DONE_%=;\n"
      "bra PE_WAIT_%=;\n"
      "PE_DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

#ifdef B200_PE_TIMING
__device__ unsigned long long g_pe_timing[4];
#define PE_T(slot, since) do { if (tid == 0) atomicAdd(&g_pe_timing[slot], (unsigned long long)(clock64() - (since))); } while (0)
#else
#define PE_T(slot, since)
#endif

template <bool OUT_F16>
__global__ void __launch_bounds__(PE_THREADS, 1) patch_embed_kernel(const PatchEmbedParams p) {
  extern __shared__ uint8_t pe_smem_raw[];
  const uint32_t raw = (uint32_t)__cvta_generic_to_shared(pe_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                         // SWIZZLE_128B tiles need 1024-B alignment
  uint8_t* gen = pe_smem_raw + (base - raw);
  uint8_t* sB = gen;                                                    // [4][192 rows x 128 B]
  uint8_t* sA = gen + PE_NKB * PE_B_KB_BYTES;                           // [4][128 rows x 128 B]
  const uint32_t sB_a = base, sA_a = base + PE_NKB * PE_B_KB_BYTES;
  const uint32_t bar_a = sA_a + PE_NKB * PE_A_KB_BYTES;                 // mbarrier (8 B) + TMEM base address (4 B)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + PE_NKB * (PE_A_KB_BYTES + PE_B_KB_BYTES) + 16);
  float* sbias = reinterpret_cast<float*>(gen + PE_NKB * (PE_A_KB_BYTES + PE_B_KB_BYTES) + 64);   // [192] bias of the CTA's columns
  uint8_t* sStage = gen + PE_NKB * (PE_A_KB_BYTES + PE_B_KB_BYTES) + 64 + PE_N * 4;               // [8 warps][32 rows][208 B]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NT = p.D / PE_N;
  const int n_blk = (int)blockIdx.x % NT;
  const int n0 = n_blk * PE_N;
  const int64_t m_tiles = (p.M + PE_M - 1) / PE_M;
  const int cta_in_col = (int)blockIdx.x / NT, ctas_per_col = (int)gridDim.x / NT;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {                                                      // one warp allocates the accumulator columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bar_a + 16), "n"(PE_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // the CTA's weight tile: 192 rows x 256 fp16, resident for its whole life
  for (int idx = tid; idx < PE_N * 32; idx += PE_THREADS) {
    const int n = idx >> 5, ch = idx & 31;                              // 32 chunks of 8 halves per row
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.w + (size_t)(n0 + n) * PE_K) + ch);
    *reinterpret_cast<uint4*>(sB + (ch >> 3) * PE_B_KB_BYTES + pe_swz(n, ch & 7)) = v;
  }
  for (int i = tid; i < PE_N; i += PE_THREADS) sbias[i] = p.bias ? __ldg(p.bias + n0 + i) : 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D = F32, A = B = F16, both K-major, N = 192, M = 128
  constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(PE_N >> 3) << 17) | ((uint32_t)(PE_M >> 4) << 24);
  const int row = tid & 127;                                            // the A row this thread fills
  const int cbase = tid >> 7;                                           // it fills chunks cbase, cbase + 2, ... of the row's 32
  unsigned phase = 0;
  for (int64_t mt = cta_in_col; mt < m_tiles; mt += ctas_per_col) {
    // ---- A tile: 128 patches x 256 taps, fp32 -> fp16, swizzled ------------------------------------------------
#ifdef B200_PE_TIMING
    const long long t0_ = clock64();
#endif
    {
      const int64_t m = mt * PE_M + row;
      const bool live = m < p.M;
      const int per = p.Fp * p.Tp;
      const int b = live ? (int)(m / per) : 0;
      const int pp = live ? (int)(m - (int64_t)b * per) : 0;
      const int hp = pp / p.Tp, wp = pp - hp * p.Tp;
      const float* src = p.feat + ((size_t)b * p.F + (size_t)hp * p.stride) * p.T + (size_t)wp * p.stride;
      // all 16 x 4 loads of the thread are issued before the first conversion: the gather is latency-bound
      const bool al8 = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)(p.T * 4)) & 7) == 0;
      float2 f[16][4];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int ch = cbase + 2 * u;                                   // chunk of 8 taps: i = ch >> 1, j = 8 (ch & 1) ..
        const float* s = src + (size_t)(ch >> 1) * p.T + 8 * (ch & 1);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (!live) f[u][q] = make_float2(0.f, 0.f);
          else if (al8) f[u][q] = __ldg(reinterpret_cast<const float2*>(s) + q);
          else f[u][q] = make_float2(__ldg(s + 2 * q), __ldg(s + 2 * q + 1));
        }
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int ch = cbase + 2 * u;
        __half2 h0 = __floats2half2_rn(f[u][0].x, f[u][0].y), h1 = __floats2half2_rn(f[u][1].x, f[u][1].y);
        __half2 h2 = __floats2half2_rn(f[u][2].x, f[u][2].y), h3 = __floats2half2_rn(f[u][3].x, f[u][3].y);
        uint4 v;
        v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1);
        v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(sA + (ch >> 3) * PE_A_KB_BYTES + pe_swz(row, ch & 7)) = v;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy writes -> visible to the tensor core
    __syncthreads();
#ifdef B200_PE_TIMING
    PE_T(0, t0_);
    const long long t1_ = clock64();
#endif
    // ---- 16 MMAs (4 k-blocks x 4 x K16), one thread ------------------------------------------------------------
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int kb = 0; kb < PE_NKB; ++kb)
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const uint64_t da = pe_desc(sA_a + kb * PE_A_KB_BYTES + k4 * 32);
          const uint64_t db = pe_desc(sB_a + kb * PE_B_KB_BYTES + k4 * 32);
          const uint32_t acc = (kb | k4) ? 1u : 0u;
          asm volatile(
              "{\n"
              ".reg .pred p;\n"
              "setp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
              "}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a) : "memory");
    }
    pe_mbar_wait(bar_a, phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef B200_PE_TIMING
    PE_T(1, t1_);
    const long long t2_ = clock64();
#endif
    // ---- epilogue: warp w reads TMEM lanes 32 (w & 3) .., columns 96 (w >> 2) ..; row = lane ---------------------
    {
      const int r = 32 * (warp & 3) + lane;
      const int64_t m = mt * PE_M + r;
      const int c0 = 96 * (warp >> 2);
      uint32_t v[6][16];
#pragma unroll
      for (int cc = 0; cc < 6; ++cc)                                    // all six loads in flight, one wait
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[cc][0]), "=r"(v[cc][1]), "=r"(v[cc][2]), "=r"(v[cc][3]), "=r"(v[cc][4]), "=r"(v[cc][5]), "=r"(v[cc][6]),
                       "=r"(v[cc][7]), "=r"(v[cc][8]), "=r"(v[cc][9]), "=r"(v[cc][10]), "=r"(v[cc][11]), "=r"(v[cc][12]),
                       "=r"(v[cc][13]), "=r"(v[cc][14]), "=r"(v[cc][15])
                     : "r"(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(c0 + 16 * cc)));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      // bias, conversion, then a transpose through this warp's staging rows so that the global stores are whole
      // contiguous row segments (a thread owns a ROW of the accumulator; row pitch of the output is D elements).
      // Staging rows hold 192 B + 16 B pad (conflict-free 16-B writes): 96 fp16 columns at once, fp32 in two halves.
      constexpr int EB = OUT_F16 ? 2 : 4;                               // bytes per output element
      constexpr int HALVES = OUT_F16 ? 1 : 2, CCH = 6 / HALVES;         // 16-column groups per staging pass
      constexpr int SROW = 208;
      uint8_t* st = sStage + warp * (32 * SROW);
      const int64_t m0 = mt * PE_M + 32 * (warp & 3);
#pragma unroll
      for (int hf = 0; hf < HALVES; ++hf) {
#pragma unroll
        for (int c2 = 0; c2 < CCH; ++c2) {
          const int cc = hf * CCH + c2;
          const int col = c0 + 16 * cc;
          float f[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bq = *reinterpret_cast<const float4*>(sbias + col + 4 * q);
            f[4 * q] = __uint_as_float(v[cc][4 * q]) + bq.x; f[4 * q + 1] = __uint_as_float(v[cc][4 * q + 1]) + bq.y;
            f[4 * q + 2] = __uint_as_float(v[cc][4 * q + 2]) + bq.z; f[4 * q + 3] = __uint_as_float(v[cc][4 * q + 3]) + bq.w;
          }
          uint8_t* d = st + lane * SROW + 16 * c2 * EB;
          if (OUT_F16) {
            uint4 a, b2;
            __half2 h;
            h = __floats2half2_rn(f[0], f[1]); a.x = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2half2_rn(f[2], f[3]); a.y = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2half2_rn(f[4], f[5]); a.z = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2half2_rn(f[6], f[7]); a.w = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2half2_rn(f[8], f[9]); b2.x = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2half2_rn(f[10], f[11]); b2.y = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2half2_rn(f[12], f[13]); b2.z = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2half2_rn(f[14], f[15]); b2.w = *reinterpret_cast<uint32_t*>(&h);
            reinterpret_cast<uint4*>(d)[0] = a;
            reinterpret_cast<uint4*>(d)[1] = b2;
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(d)[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
          }
        }
        __syncwarp();
        // 32 rows x 12 chunks of 16 B (192 B per row); consecutive lanes take consecutive chunks of a row
#pragma unroll 4
        for (int idx = lane; idx < 32 * 12; idx += 32) {
          const int rr = idx / 12, ch = idx - rr * 12;
          if (m0 + rr < p.M) {
            const uint4 val = *reinterpret_cast<const uint4*>(st + rr * SROW + 16 * ch);
            uint8_t* o = reinterpret_cast<uint8_t*>(p.out) + ((size_t)(m0 + rr) * p.D + n0 + c0 + hf * (96 / HALVES)) * EB + 16 * ch;
            __stcs(reinterpret_cast<uint4*>(o), val);
          }
        }
        __syncwarp();
      }
      (void)m; (void)r;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                                    // accumulator drained, A tile free again
#ifdef B200_PE_TIMING
    PE_T(2, t2_);
    if (tid == 0) atomicAdd(&g_pe_timing[3], 1ull);
#endif
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(PE_TMEM_COLS) : "memory");
}

}  // namespace b200
