// AST patch embedding as a warp-specialised TMA + tcgen05 pipeline (same contract as
// patch_embed.cuh: Conv2d(1, D, 16, stride s) + flatten(2).transpose(1, 2) as the im2col GEMM C[M, D] = A[M, 256] W^T + bias,
// fp16 operands, fp32 accumulation; PatchEmbed.forward src/models/ast_mini.py:7-15, ASTModel.forward src/models/ast.py:30,50-56).
//
// The patches of one patch row overlap (16 wide, stride 10), and their stride of 40 bytes cannot be expressed by a tensor-map
// or a UMMA descriptor (both want multiples of 16 bytes), so the A operand is built in two steps:
//   1. TMA (cp.async.bulk.tensor.2d, SASS UTMALDG) pulls STRIPS of the feature map -- 4 feature rows x 512 frames of the
//      16-row band a patch row covers -- into a 4-stage shared-memory ring; every feature byte of a tile crosses L2 -> SM once
//      instead of once per overlapping patch, fully coalesced, asynchronously.
//   2. four "cutter" warps read the patches out of the strips (conflict-free 64-bit loads: lanes = consecutive patches,
//      40 B apart), round to fp16 and store the A chunk (128 rows x 64 taps) straight into TENSOR MEMORY (tcgen05.st, SASS STTM:
//      row = lane, two taps per 32-bit column); the MMA takes A from TMEM and only W from shared memory, which spares the
//      shared-memory data path a write + a read of every A byte and the cutters the generic -> async proxy fence.
// A tile = 2 segments of up to 64 patches of one patch row each (50 + 50 real rows of the 128 for AST's T = 512); its K = 256
// is cut in four chunks of 4 feature rows, and the chunks are the unit of the pipeline:
//
//   warp 0      TMA producer: the strips of chunks q + 1 .. q + 3 are in flight while chunk q is being cut
//   warps 8-11  cutters: strip stage -> A chunk slot in TMEM (2 slots of 32 columns), release the strip stage
//   warp 1      MMA issuer: 4 x tcgen05.mma (M 128, N 192, K 16) per chunk, tcgen05.commit frees the A chunk; the
//               accumulator (fp32, TMEM) is double buffered: tile k + 1 accumulates while tile k is drained
//   warps 4-7   epilogue: tcgen05.ld -> + bias -> fp16 / fp32 -> per-warp staging rows -> contiguous 16-byte row segments
//
// The CTA's 192 x 256 weight slice stays in shared memory for its whole life (persistent grid: #SMs CTAs, D / 192 column
// slices x #SMs / slices row walkers; the slices of one tile run on neighbouring CTAs at the same time, so three of the four
// strip reads hit L2).  Every hand-over is an mbarrier; nothing but the weights is loaded by plain global loads.
#pragma once
#include <cuda.h>
#include "patch_embed.cuh"

namespace b200 {

#ifndef B200_PP_CUT_WARPS
#define B200_PP_CUT_WARPS 4
#endif
constexpr int PP_CUT_WARPS = B200_PP_CUT_WARPS;         // 4: one cutter thread per A row and chunk; 8: two (feature rows 0-1 / 2-3 of the chunk)
constexpr int PP_THREADS = 256 + 32 * PP_CUT_WARPS;
constexpr int PP_SEG = 64;                               // row slots per segment (2 segments = the 128 rows of a tile)
constexpr int PP_BOXW = 256;                             // floats per TMA box row (the box limit)
constexpr int PP_BOX_BYTES = 4 * PP_BOXW * 4;            // 4 feature rows x 256 frames
constexpr int PP_SEG_BYTES = 2 * PP_BOX_BYTES;           // two boxes: 512 frames of 4 rows
constexpr int PP_STAGE_BYTES = 2 * PP_SEG_BYTES;         // two segments: 16 KB
#ifndef B200_PP_NSTAGE
#define B200_PP_NSTAGE 4
#endif
constexpr int PP_NSTAGE = B200_PP_NSTAGE;                // strip ring: the TMA latency (~1.5 k cycles) wants >= 48 KB in flight per SM
constexpr int PP_NA = 2;                                 // A chunk slots (TMEM, 32 columns each): one being cut, one being multiplied
constexpr int PP_ACOL0 = 192, PP_ACOL1 = 448;            // their columns: behind the two accumulators (0..191, 256..447)
constexpr int PP_STG_ROW = 208;                          // staging row: 192 B + 16 B pad
constexpr int PP_TMEM_COLS = 512;                        // two accumulators of 192 columns at 0 and 256
constexpr size_t PP_SMEM = 1024 + (size_t)PE_NKB * PE_B_KB_BYTES + PP_NSTAGE * PP_STAGE_BYTES +
                           4 * 32 * PP_STG_ROW + PE_N * 4 + 256;

// Optional pipeline timing (-DB200_PP_TIMING): cycles summed over all CTAs into g_pp_timing[12] =
// {producer wait s_empty, cutter wait s_full, cutter wait a_empty, cutter busy, mma wait a_full, mma wait acc_empty,
//  epilogue wait acc_full, epilogue busy, tiles, -, -, -}
#ifdef B200_PP_TIMING
__device__ unsigned long long g_pp_timing[16];
// (accumulated in registers and flushed once per role: a global atomic inside the loop would sit in front of the cutters'
// memory fence and distort what it measures)
#define PP_TDECL() unsigned long long tacc_[6] = {0ull, 0ull, 0ull, 0ull, 0ull, 0ull}
#define PP_TW(i, stmt) do { const long long t_ = clock64(); stmt; tacc_[i] += (unsigned long long)(clock64() - t_); } while (0)
#define PP_TADD(i, since) do { tacc_[i] += (unsigned long long)(clock64() - (since)); } while (0)
#define PP_TFLUSH(i, slot) do { if (lane == 0) atomicAdd(&g_pp_timing[slot], tacc_[i]); } while (0)
#define PP_NOW() clock64()
#else
#define PP_TDECL()
#define PP_TW(i, stmt) stmt
#define PP_TADD(i, since)
#define PP_TFLUSH(i, slot)
#define PP_NOW() 0ll
#endif

__device__ __forceinline__ void pp_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void pp_tma_box(uint32_t dst, const CUtensorMap* tm, int x, int y, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(bar) : "memory");
}

// nseg: segments per patch row, segp: patches per segment (<= 64, and (segp - 1) stride + 16 + 3 <= 512 frames)
template <bool OUT_F16>
__global__ void __launch_bounds__(PP_THREADS, 1) patch_embed_pipe_kernel(const PatchEmbedParams p, const int nseg, const int segp,
                                                                        const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ uint8_t pe_smem_raw[];
  const uint32_t raw = (uint32_t)__cvta_generic_to_shared(pe_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                         // SWIZZLE_128B tiles need 1024-B alignment
  uint8_t* gen = pe_smem_raw + (base - raw);
  constexpr uint32_t OFF_S = PE_NKB * PE_B_KB_BYTES, OFF_STG = OFF_S + PP_NSTAGE * PP_STAGE_BYTES,
                     OFF_BIAS = OFF_STG + 4 * 32 * PP_STG_ROW, OFF_BAR = OFF_BIAS + PE_N * 4;
  uint8_t* sB = gen;                                                    // [4][192 rows x 128 B] weights
  const uint8_t* sS = gen + OFF_S;                                      // [PP_NSTAGE][2 segments][2 boxes][4 rows][256] fp32 strips
  uint8_t* sStage = gen + OFF_STG;                                      // [4 warps][32 rows][208 B]
  float* sbias = reinterpret_cast<float*>(gen + OFF_BIAS);
  const uint32_t sB_a = base, sS_a = base + OFF_S, bar0 = base + OFF_BAR;
  // barriers (8 B each): s_full[8] s_empty[8] a_full[2] a_empty[2] acc_full[2] acc_empty[2]; then the TMEM base address
  static_assert(PP_NSTAGE <= 8 && PP_NA == 2, "barrier layout");
  const uint32_t s_full = bar0, s_empty = bar0 + 64, a_full = bar0 + 128, a_empty = bar0 + 144, acc_full = bar0 + 160, acc_empty = bar0 + 176;
  const uint32_t tmem_slot_a = bar0 + 192;
  const uint32_t* tmem_slot = reinterpret_cast<const uint32_t*>(gen + OFF_BAR + 192);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NT = p.D / PE_N;
  const int n_blk = (int)blockIdx.x % NT, n0 = n_blk * PE_N;
  const int cta_in_col = (int)blockIdx.x / NT, ctas_per_col = (int)gridDim.x / NT;
  const int64_t n_segs = (int64_t)p.B * p.Fp * nseg;
  const int64_t n_tiles = (n_segs + 1) >> 1;

  if (tid == 0) {
    auto init = [](uint32_t bar, unsigned cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(cnt) : "memory"); };
    for (int i = 0; i < PP_NSTAGE; ++i) { init(s_full + 8 * i, 1); init(s_empty + 8 * i, PP_CUT_WARPS); }
    for (int i = 0; i < 2; ++i) { init(acc_full + 8 * i, 1); init(acc_empty + 8 * i, 4); init(a_full + 8 * i, PP_CUT_WARPS); init(a_empty + 8 * i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot_a), "n"(PP_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int idx = tid; idx < PE_N * 32; idx += PP_THREADS) {             // the CTA's weight slice, resident
    const int n = idx >> 5, ch = idx & 31;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.w + (size_t)(n0 + n) * PE_K) + ch);
    *reinterpret_cast<uint4*>(sB + (ch >> 3) * PE_B_KB_BYTES + pe_swz(n, ch & 7)) = v;
  }
  for (int i = tid; i < PE_N; i += PP_THREADS) sbias[i] = p.bias ? __ldg(p.bias + n0 + i) : 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // weights: generic-proxy writes -> tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    unsigned q = 0;
    PP_TDECL();
    for (int64_t tile = cta_in_col; tile < n_tiles; tile += ctas_per_col) {
      int row[2], x0[2];
      bool valid[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int64_t G = 2 * tile + s;
        valid[s] = G < n_segs;
        const int64_t S = valid[s] ? G / nseg : 0;                      // patch row (b, fp)
        const int j = valid[s] ? (int)(G - S * nseg) : 0;
        const int b = (int)(S / p.Fp), fpr = (int)(S - (int64_t)b * p.Fp);
        row[s] = b * p.F + fpr * p.stride;
        x0[s] = (j * segp * p.stride) & ~3;                             // 16-byte aligned first frame of the segment
      }
      const unsigned bytes = (valid[0] ? PP_SEG_BYTES : 0) + (valid[1] ? PP_SEG_BYTES : 0);
      for (int c = 0; c < PE_NKB; ++c, ++q) {
        const unsigned stage = q % PP_NSTAGE, n = q / PP_NSTAGE;
        if (n >= 1) PP_TW(0, pe_mbar_wait(s_empty + 8 * stage, (n - 1) & 1u));
        if (lane == 0) {
          const uint32_t bar = s_full + 8 * stage;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
#pragma unroll
          for (int s = 0; s < 2; ++s)
            if (valid[s]) {
              const uint32_t dst = sS_a + stage * PP_STAGE_BYTES + s * PP_SEG_BYTES;
              pp_tma_box(dst, &tmap, x0[s], row[s] + 4 * c, bar);
              pp_tma_box(dst + PP_BOX_BYTES, &tmap, x0[s] + PP_BOXW, row[s] + 4 * c, bar);
            }
        }
        __syncwarp();
      }
    }
    PP_TFLUSH(0, 0);
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(PE_N >> 3) << 17) | ((uint32_t)(PE_M >> 4) << 24);
    unsigned k = 0, q = 0;
    PP_TDECL();
    for (int64_t tile = cta_in_col; tile < n_tiles; tile += ctas_per_col, ++k) {
      const unsigned buf = k & 1u, m = k >> 1;
      if (m >= 1) PP_TW(1, pe_mbar_wait(acc_empty + 8 * buf, (m - 1) & 1u));      // the epilogue has drained this accumulator
      const uint32_t acc_addr = tmem + buf * 256u;
      for (int c = 0; c < PE_NKB; ++c, ++q) {
        const unsigned slot = q & 1u;
        PP_TW(0, pe_mbar_wait(a_full + 8 * slot, (q >> 1) & 1u));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            // A straight from TMEM (row = lane, 8 columns = 16 fp16 taps per K step), W from shared memory
            const uint32_t ta = tmem + (slot ? PP_ACOL1 : PP_ACOL0) + 8u * k4;
            const uint64_t db = pe_desc(sB_a + c * PE_B_KB_BYTES + k4 * 32);
            const uint32_t acc = (c | k4) ? 1u : 0u;
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "setp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
                "}\n" ::"r"(acc_addr), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
          }
          // frees the A chunk once the MMAs that read it have completed; the last chunk also publishes the accumulator
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(a_empty + 8 * slot) : "memory");
          if (c == PE_NKB - 1)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(acc_full + 8 * buf) : "memory");
        }
        __syncwarp();
      }
    }
    PP_TFLUSH(0, 4); PP_TFLUSH(1, 5);
  } else if (warp >= 8) {
    // =============================== cutters: strips -> A chunks ===============================
    const int ct = (tid - 256) & 127;                                   // A row = TMEM lane
    constexpr int NFL = 16 / PP_CUT_WARPS;                              // feature rows of a chunk per cutter thread (4 or 2)
    const int fl0 = ((tid - 256) >> 7) * NFL;
    const int s = ct >> 6, slot = ct & (PP_SEG - 1);
    const bool even = (p.stride & 1) == 0;
    unsigned q = 0, k = 0;
    PP_TDECL();
    for (int64_t tile = cta_in_col; tile < n_tiles; tile += ctas_per_col, ++k) {
      const int64_t G = 2 * tile + s;
      const bool valid = G < n_segs;
      const int64_t S = valid ? G / nseg : 0;
      const int j = valid ? (int)(G - S * nseg) : 0;
      const bool live = valid && slot < segp && j * segp + slot < p.Tp;
      const int col0 = ((j * segp * p.stride) & 3) + slot * p.stride;   // first frame of the patch inside the staged strip
      int off[8];                                                       // byte offset of the pair (col0 + 2 u, + 1) in row 0 of the segment
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int col = col0 + 2 * u;                                   // even stride: a pair never straddles the two boxes
        off[u] = (col >> 8) * PP_BOX_BYTES + (col & 255) * 4;
      }
      for (int c = 0; c < PE_NKB; ++c, ++q) {
        const unsigned stage = q % PP_NSTAGE, n = q / PP_NSTAGE, slot = q & 1u, u_ = q >> 1;
        PP_TW(0, pe_mbar_wait(s_full + 8 * stage, n & 1u));
        if (u_ >= 1) PP_TW(1, pe_mbar_wait(a_empty + 8 * slot, (u_ - 1) & 1u));
        [[maybe_unused]] const long long tcut_ = PP_NOW();
        const uint8_t* src = sS + stage * PP_STAGE_BYTES + s * PP_SEG_BYTES;
        uint32_t r[8 * NFL];                                            // the row's taps of this chunk as fp16 pairs: column 8 fl + u
#pragma unroll
        for (int fi = 0; fi < NFL; ++fi) {                              // feature row 4 c + fl of the patch: taps 16 fl .. 16 fl + 15
          const int fl = fl0 + fi;
          float2 v[8];
          if (live) {
            if (even) {
#pragma unroll
              for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const float2*>(src + off[u] + fl * (PP_BOXW * 4));
            } else {
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int ca = col0 + 2 * u, cb = ca + 1;
                v[u].x = *reinterpret_cast<const float*>(src + (ca >> 8) * PP_BOX_BYTES + fl * (PP_BOXW * 4) + (ca & 255) * 4);
                v[u].y = *reinterpret_cast<const float*>(src + (cb >> 8) * PP_BOX_BYTES + fl * (PP_BOXW * 4) + (cb & 255) * 4);
              }
            }
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = make_float2(0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const __half2 h = __floats2half2_rn(v[u].x, v[u].y);
            r[8 * fi + u] = *reinterpret_cast<const uint32_t*>(&h);
          }
        }
        PP_TADD(3, tcut_);                                              // strip reads + conversion
        [[maybe_unused]] const long long tst_ = PP_NOW();
        // registers -> TMEM: lane = A row, 32 columns; no shared-memory round trip and no proxy fence for the A operand
        const uint32_t ta = tmem + ((uint32_t)(32 * (ct >> 5)) << 16) + (slot ? PP_ACOL1 : PP_ACOL0) + 8u * fl0;
        if constexpr (NFL == 4) {
          asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                       "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                       ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8 % (8 * NFL)]), "r"(r[9 % (8 * NFL)]),
                         "r"(r[10 % (8 * NFL)]), "r"(r[11 % (8 * NFL)]), "r"(r[12 % (8 * NFL)]), "r"(r[13 % (8 * NFL)]), "r"(r[14 % (8 * NFL)]), "r"(r[15 % (8 * NFL)]),
                         "r"(r[16 % (8 * NFL)]), "r"(r[17 % (8 * NFL)]), "r"(r[18 % (8 * NFL)]), "r"(r[19 % (8 * NFL)]), "r"(r[20 % (8 * NFL)]), "r"(r[21 % (8 * NFL)]),
                         "r"(r[22 % (8 * NFL)]), "r"(r[23 % (8 * NFL)]), "r"(r[24 % (8 * NFL)]), "r"(r[25 % (8 * NFL)]), "r"(r[26 % (8 * NFL)]), "r"(r[27 % (8 * NFL)]),
                         "r"(r[28 % (8 * NFL)]), "r"(r[29 % (8 * NFL)]), "r"(r[30 % (8 * NFL)]), "r"(r[31 % (8 * NFL)]) : "memory");
        } else {
          asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                       ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                         "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {                                                // one arrival per warp on either barrier
          pp_arrive(a_full + 8 * slot);
          pp_arrive(s_empty + 8 * stage);                               // the warp has read its part of the strips
        }
        PP_TADD(2, tcut_);
        PP_TADD(4, tst_);                                               // tcgen05.st + wait + arrivals
      }
    }
    PP_TFLUSH(0, 1); PP_TFLUSH(1, 2); PP_TFLUSH(2, 3); PP_TFLUSH(3, 9); PP_TFLUSH(4, 10);
  } else if (warp >= 4) {
    // =============================== epilogue ===============================
    const int ew = warp - 4;                                            // TMEM lanes 32 ew .. 32 ew + 31
    const int s = ew >> 1, slot0 = 32 * (ew & 1);
    constexpr int EB = OUT_F16 ? 2 : 4;
    constexpr int NPASS = OUT_F16 ? 2 : 4, CPP = PE_N / NPASS, LD = CPP / 16;   // columns per pass, x16 loads per pass
    uint8_t* st = sStage + ew * (32 * PP_STG_ROW);
    unsigned k = 0;
    PP_TDECL();
    for (int64_t tile = cta_in_col; tile < n_tiles; tile += ctas_per_col, ++k) {
      const unsigned buf = k & 1u, m = k >> 1;
      const int64_t G = 2 * tile + s;
      const bool valid = G < n_segs;
      const int64_t S = valid ? G / nseg : 0;
      const int j = valid ? (int)(G - S * nseg) : 0;
      const int tp0 = j * segp + slot0;                                 // first patch of this warp's 32 rows
      int n_rows = segp - slot0 < p.Tp - tp0 ? segp - slot0 : p.Tp - tp0;
      n_rows = !valid ? 0 : (n_rows < 0 ? 0 : (n_rows > 32 ? 32 : n_rows));
      const int64_t m0 = S * p.Tp + tp0;                                // output row of lane 0 (rows are consecutive)
      PP_TW(0, pe_mbar_wait(acc_full + 8 * buf, m & 1u));
      [[maybe_unused]] const long long tep_ = PP_NOW();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ps = 0; ps < NPASS; ++ps) {
        uint32_t v[LD][16];
        [[maybe_unused]] const long long tld_ = PP_NOW();
#pragma unroll
        for (int cc = 0; cc < LD; ++cc)
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                       : "=r"(v[cc][0]), "=r"(v[cc][1]), "=r"(v[cc][2]), "=r"(v[cc][3]), "=r"(v[cc][4]), "=r"(v[cc][5]), "=r"(v[cc][6]),
                         "=r"(v[cc][7]), "=r"(v[cc][8]), "=r"(v[cc][9]), "=r"(v[cc][10]), "=r"(v[cc][11]), "=r"(v[cc][12]),
                         "=r"(v[cc][13]), "=r"(v[cc][14]), "=r"(v[cc][15])
                       : "r"(tmem + ((uint32_t)(32 * ew) << 16) + buf * 256u + (uint32_t)(ps * CPP + 16 * cc)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        PP_TADD(3, tld_);
        if (ps == NPASS - 1) {                                          // the accumulator is in registers: hand it back
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) pp_arrive(acc_empty + 8 * buf);
        }
#pragma unroll
        for (int cc = 0; cc < LD; ++cc) {
          const int col = ps * CPP + 16 * cc;
          float f[16];
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            const float4 bq = *reinterpret_cast<const float4*>(sbias + col + 4 * qd);
            f[4 * qd] = __uint_as_float(v[cc][4 * qd]) + bq.x; f[4 * qd + 1] = __uint_as_float(v[cc][4 * qd + 1]) + bq.y;
            f[4 * qd + 2] = __uint_as_float(v[cc][4 * qd + 2]) + bq.z; f[4 * qd + 3] = __uint_as_float(v[cc][4 * qd + 3]) + bq.w;
          }
          uint8_t* d = st + lane * PP_STG_ROW + 16 * cc * EB;
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
            if (lane < n_rows)                                        // (slots past the segment's patches are not staged)
            { reinterpret_cast<uint4*>(d)[0] = a; reinterpret_cast<uint4*>(d)[1] = b2; }
          } else {
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) reinterpret_cast<float4*>(d)[qd] = make_float4(f[4 * qd], f[4 * qd + 1], f[4 * qd + 2], f[4 * qd + 3]);
          }
        }
        __syncwarp();
        // n_rows rows x 12 chunks of 16 B (192 B per row and pass); consecutive lanes take consecutive chunks of a row
        // (fully unrolled with a row predicate: the twelve staging reads are in flight together instead of one
        // read -> store round trip per iteration)
        uint8_t* const obase = reinterpret_cast<uint8_t*>(p.out) + ((size_t)m0 * p.D + n0 + ps * CPP) * EB;
        const size_t opitch = (size_t)p.D * EB;
        uint4 val[12];
        [[maybe_unused]] const long long tsg_ = PP_NOW();
#pragma unroll
        for (int it = 0; it < 12; ++it) {
          const int idx = lane + 32 * it, rr = idx / 12, ch = idx - rr * 12;
          val[it] = *reinterpret_cast<const uint4*>(st + rr * PP_STG_ROW + 16 * ch);
        }
#pragma unroll
        for (int it = 0; it < 12; ++it) {
          const int idx = lane + 32 * it, rr = idx / 12, ch = idx - rr * 12;
          if (rr < n_rows) __stcs(reinterpret_cast<uint4*>(obase + rr * opitch + 16 * ch), val[it]);
        }
        __syncwarp();
        PP_TADD(4, tsg_);
      }
      PP_TADD(1, tep_);
#ifdef B200_PP_TIMING
      tacc_[2] += 1ull;
#endif
    }
    PP_TFLUSH(0, 6); PP_TFLUSH(1, 7); PP_TFLUSH(3, 11);
#ifdef B200_PP_TIMING
    if (lane == 0) atomicAdd(&g_pp_timing[0] + 12, tacc_[4]);
#endif
#ifdef B200_PP_TIMING
    if (ew == 0) PP_TFLUSH(2, 8);
#endif
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(PP_TMEM_COLS) : "memory");
  }
}

}  // namespace b200
