// Tuned kernel for the reference-ACTUAL AST recipe: MelSpectrogram(n_fft 1024, win_length <= 416
// or any, hop, power 2, centred, reflect padding, periodic Hann, HTK mel) + AmplitudeToDB at the clips' own rate
// (ASTPreprocessor.preprocess, src/datasets/preprocessing.py:988-998, 1013-1027; melspectrogram(), src/utils/audio.py:60-84;
// torchaudio/functional/functional.py:123-137, 390-402).  The top_db clamp and the per-clip normalisation need the clip's
// maximum first and run as the second pass (clip_norm.cuh).
//
// No resampler here, so no producer warps: every warp is an independent worker that claims TILES of 8 consecutive
// frames of one clip from a global counter (neighbouring warps of a CTA hold neighbouring tiles, so the 2.5x re-read of
// the overlapping frames is served by L1) and runs four passes of two frames each:
//   stage 0   the frames' samples straight from global memory (coalesced, reflect index at the clip edges), window
//             multiply as packed pairs (frame a, frame b) -- the two frames are the real and imaginary part of ONE
//             complex 1024-point transform
//   stage 1   32-point DFT over n1 in registers (n = lane + 32 n1), pruned for the zero padding: with a 400-sample window
//             only n1 = 0..12 are non-zero, so the first radix-2 stage is a copy + twiddle and the second is half empty
//   exchange  W_1024^(lane k1) twiddles, 32 x 33 padded transpose through the warp's shared-memory buffer
//   stage 2   32-point DFT over n2 in registers (lane = k1), conjugate split of the two real spectra by shuffle, |.|^2
//             for the 513 bins as (frame a, frame b) pairs
//   mel       lane slots = mel bins (host-planned, LDS.64 per tap), 10 log10 through lg2.approx, running maximum
//   store     (B, 1, n_mels, T): the tile's 8 x n_mels values are transposed through shared memory so that every row
//             segment of 8 frames leaves as contiguous 8-byte stores; (B, T, n_mels): straight from the registers.
#pragma once
#include "fbank_fast.cuh"

namespace b200 {

#ifndef B200_MS_WARPS
#define B200_MS_WARPS 12
#endif
constexpr int MS_WARPS = B200_MS_WARPS, MS_THREADS = 32 * MS_WARPS;
constexpr int MS_N = 1024;
constexpr int MS_EROW = 33;                          // float2 per exchange row (32 + 1 pad)
constexpr int MS_EBUF = 32 * MS_EROW * 2;            // floats per warp: 32 x 33 complex; later 513 power pairs
constexpr int MS_TILE = 8;                           // frames per tile
constexpr int MS_TROW = 10;                          // floats per tile row: 8 frames + 2 pad (half-warp STS.64 conflict free)
constexpr int MS_TBUF = 128 * MS_TROW;               // floats per warp
constexpr float MS_DB_PER_LG2 = 3.0102999566398120f; // 10 log10(2)

struct MelFastParams {
  const float2* tw;        // [32][32] W_1024^(k1 * lane)
  const float* melw;       // [rows][32] zero-padded weights (x 1/4: the conjugate split leaves the factor out), lanes = slots
  const int* slot_bin;     // [32 groups'] mel bin of lane slot (group i, lane l), >= n_mel: none
  const int* slot_start;   // first FFT bin the slot reads
  int groups, maxcnt[4], woff[4], rows;
  int tiles_per_clip;
  int* counter;            // per launch: zeroed tile counter
};

// 32-point DIF over n1 for inputs with z[13..31] == 0 (in place; X[k1] ends up in z[bitrev5(k1)]).
__device__ __forceinline__ void ms_dft32_pruned13(float2 (&z)[32]) {
  // stage H = 16: the upper input of every butterfly is zero: v[j] = z[j], v[j + 16] = z[j] W32^j
  z[16] = z[0];
  z[17] = mul_w32<1>(z[1]);   z[18] = mul_w32<2>(z[2]);   z[19] = mul_w32<3>(z[3]);   z[20] = mul_w32<4>(z[4]);
  z[21] = mul_w32<5>(z[5]);   z[22] = mul_w32<6>(z[6]);   z[23] = mul_w32<7>(z[7]);   z[24] = mul_w32<8>(z[8]);
  z[25] = mul_w32<9>(z[9]);   z[26] = mul_w32<10>(z[10]); z[27] = mul_w32<11>(z[11]); z[28] = mul_w32<12>(z[12]);
  // stage H = 8 in both halves: inputs j + 8 exist for j <= 4 only
#pragma unroll
  for (int B = 0; B < 32; B += 16) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const float2 a = z[B + j], b = z[B + j + 8];
      z[B + j] = cadd(a, b);
      const float2 d = csub(a, b);
      z[B + j + 8] = j == 0 ? d : j == 1 ? mul_w32<2>(d) : j == 2 ? mul_w32<4>(d) : j == 3 ? mul_w32<6>(d) : mul_w32<8>(d);
    }
    z[B + 13] = mul_w32<10>(z[B + 5]); z[B + 14] = mul_w32<12>(z[B + 6]); z[B + 15] = mul_w32<14>(z[B + 7]);
  }
  DifStages<32, 4>::run(z);
}

// One frame pair -> P2[k] = (|A[k]|^2, |B[k]|^2) * 4 for k = 0..512 in the warp's buffer.  xa / xb: samples lane + 32 j of the
// two frames, w: window.  NJ = non-zero register inputs (13: pruned transform, 32: full).
template <int NJ>
__device__ __forceinline__ void ms_pair_power(const float (&xa)[NJ], const float (&xb)[NJ], const float (&w)[NJ],
                                              const float2* __restrict__ stw, float* __restrict__ Ebuf, int lane) {
  float2* E = reinterpret_cast<float2*>(Ebuf);
  {
    float2 z[32];
#pragma unroll
    for (int j = 0; j < NJ; ++j) z[j] = fk_upk(fk_mul2(fk_pk(xa[j], xb[j]), fk_pk(w[j], w[j])));
    if constexpr (NJ == 13) ms_dft32_pruned13(z);
    else fft_dif<32>(z);
    E[lane] = z[0];
#pragma unroll
    for (int k1 = 1; k1 < 32; ++k1) E[k1 * MS_EROW + lane] = fk_cmul(z[bitrev_n(k1, 5)], stw[k1 * 32 + lane]);
  }
  __syncwarp();
  float2 u[32];
  {
    const float2* row = E + lane * MS_EROW;
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) u[n2] = row[n2];
  }
  __syncwarp();
  fft_dif<32>(u);                                        // Z[lane + 32 k2] = u[bitrev5(k2)]
  const int src = (32 - lane) & 31;
  float2* P2 = E + lane;
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    const float2 zk = u[bitrev_n(k2, 5)];
    const float2 own = u[bitrev_n((32 - k2) & 31, 5)];     // lane 0: Z[1024 - 32 k2] sits in the same lane
    const float2 oth = u[bitrev_n(31 - k2, 5)];            // lane k1 > 0: Z[1024 - k] = lane 32 - k1, k2' = 31 - k2
    float px = __shfl_sync(0xffffffffu, oth.x, src);
    float py = __shfl_sync(0xffffffffu, oth.y, src);
    if (lane == 0) { px = own.x; py = own.y; }
    const float2 sm = cadd(zk, make_float2(px, py)), df = csub(zk, make_float2(px, py));
    P2[32 * k2] = fk_upk(fk_fma2(fk_pk(sm.x, sm.y), fk_pk(sm.x, sm.y), fk_mul2(fk_pk(df.y, df.x), fk_pk(df.y, df.x))));
  }
  if (lane == 0) {                                       // k = 512: its own partner
    const float2 zk = u[bitrev_n(16, 5)];
    P2[512] = make_float2(4.f * zk.x * zk.x, 4.f * zk.y * zk.y);
  }
  __syncwarp();
}

template <int NJ, int SJ>       // SJ > 0: hop == 32 SJ, frame b's sample j is frame a's sample j + SJ (shared loads)
__global__ void __launch_bounds__(MS_THREADS, 1) melspec_fast_kernel(const FbankParams p, const MelFastParams mp) {
  extern __shared__ __align__(16) float smem[];
  float2* stw = reinterpret_cast<float2*>(smem);                 // [1024]
  float* smelw = smem + 2 * MS_N;                                // [rows * 32]
  float* ebuf = smelw + ((mp.rows * 32 + 3) & ~3);               // [MS_WARPS][MS_EBUF]
  float* tbuf = ebuf + MS_WARPS * MS_EBUF;                       // [MS_WARPS][MS_TBUF]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < MS_N; i += MS_THREADS) stw[i] = __ldg(mp.tw + i);
  for (int i = tid; i < mp.rows * 32; i += MS_THREADS) smelw[i] = __ldg(mp.melw + i);
  __syncthreads();
  float* Ebuf = ebuf + warp * MS_EBUF;
  float* Tb = tbuf + warp * MS_TBUF;
  float w[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) w[j] = (lane + 32 * j < p.size) ? __ldg(p.window + lane + 32 * j) : 0.f;
  int mstart[4], mbin[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool have = i < mp.groups;
    mbin[i] = have ? __ldg(mp.slot_bin + lane + 32 * i) : p.n_mel;
    mstart[i] = have ? __ldg(mp.slot_start + lane + 32 * i) : 0;
  }
  const int total = p.B * mp.tiles_per_clip;
  const int half = p.size >> 1;
  int next = 0;
  if (lane == 0) next = atomicAdd(mp.counter, 1);
  for (;;) {
    const int id = __shfl_sync(0xffffffffu, next, 0);
    if (id >= total) break;
    if (lane == 0) next = atomicAdd(mp.counter, 1);              // claimed one tile ahead: the latency hides behind this tile
    const int b = id / mp.tiles_per_clip, tile = id - b * mp.tiles_per_clip;
    const ClipInfo c = clip_info(p, b);
    const int m_eff = (int)(c.m < p.out_frames ? c.m : p.out_frames);
    if (tile == 0 && lane == 0 && p.n_frames_out) p.n_frames_out[b] = m_eff;
    const int t0 = tile * MS_TILE;
    const int t_end = t0 + MS_TILE < p.out_frames ? t0 + MS_TILE : p.out_frames;
    const float* __restrict__ x = c.wav;
    const int64_t n = c.n_in;
    float vmax = -INFINITY;
#pragma unroll 1
    for (int pi = 0; pi < MS_TILE / 2; ++pi) {
      const int ta = t0 + 2 * pi;
      if (ta >= t_end) break;
      const bool live_a = ta < m_eff, live_b = ta + 1 < m_eff;
      float ya[4] = {0.f, 0.f, 0.f, 0.f}, yb[4] = {0.f, 0.f, 0.f, 0.f};
      if (live_a) {
        // ---- stage 0: samples lane + 32 j of the two frames (frame t starts at t * hop - size / 2) ----
        const int64_t base = (int64_t)ta * p.shift - half;
        constexpr int NL = SJ > 0 ? NJ + SJ : NJ;
        float xl[NL], xa[NJ], xb[NJ];            // (the reflect padding is a function of the absolute index: shared loads stay valid at the edges)
        const bool interior = base >= 0 && base + 32 * (SJ > 0 ? NL : NJ) + (SJ > 0 ? 0 : p.shift) <= n;
        if (interior) {
#pragma unroll
          for (int j = 0; j < NL; ++j) xl[j] = __ldg(x + base + lane + 32 * j);
          if constexpr (SJ == 0) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) xb[j] = __ldg(x + base + p.shift + lane + 32 * j);
          }
        } else {
#pragma unroll
          for (int j = 0; j < NL; ++j) {
            int64_t v = reflect_index(base + lane + 32 * j, n, 2);
            v = v < 0 ? 0 : (v >= n ? n - 1 : v);               // only cells under the zero part of the window can land here
            xl[j] = __ldg(x + v);
          }
          if constexpr (SJ == 0) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
              int64_t v = reflect_index(base + p.shift + lane + 32 * j, n, 2);
              v = v < 0 ? 0 : (v >= n ? n - 1 : v);
              xb[j] = __ldg(x + v);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          xa[j] = xl[j];
          if constexpr (SJ > 0) xb[j] = xl[j + SJ];
        }
        if (!live_b) {
#pragma unroll
          for (int j = 0; j < NJ; ++j) xb[j] = 0.f;
        }
        ms_pair_power<NJ>(xa, xb, w, stw, Ebuf, lane);
        // ---- mel (lane slots = bins), dB ----
        const float2* P2 = reinterpret_cast<const float2*>(Ebuf);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i >= mp.groups) continue;
          const float* wrow = smelw + mp.woff[i] * 32 + lane;
          const float2* pp = P2 + mstart[i];
          fk_u64 acc = 0ull;
          const int cnt = mp.maxcnt[i];
#pragma unroll 4
          for (int j = 0; j < cnt; ++j) {
            const float wj = wrow[j * 32];
            const float2 pv = pp[j];
            acc = fk_fma2(fk_pk(wj, wj), fk_pk(pv.x, pv.y), acc);
          }
          const float2 a = fk_upk(acc);
          float va = a.x, vb = a.y;
          if (p.db_mode) {                                      // 10 log10(max(x, 1e-10)), functional.py:390-396 (ref = 1, power)
            float la, lb;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(la) : "f"(fmaxf(va, 1e-10f)));
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lb) : "f"(fmaxf(vb, 1e-10f)));
            va = la * MS_DB_PER_LG2; vb = lb * MS_DB_PER_LG2;
          }
          ya[i] = va;
          yb[i] = live_b ? vb : 0.f;
          if (mbin[i] < p.n_mel) vmax = fmaxf(vmax, live_b ? fmaxf(va, vb) : va);
        }
      }
      // ---- hand the pair to the store stage ----
      if (p.layout == 0) {
        float* o = p.out + ((size_t)b * p.out_frames + ta) * p.n_cols;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i < mp.groups && mbin[i] < p.n_mel) {
            o[mbin[i]] = ya[i];
            if (ta + 1 < t_end) o[p.n_cols + mbin[i]] = yb[i];
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < mp.groups && mbin[i] < p.n_mel)
            *reinterpret_cast<float2*>(Tb + mbin[i] * MS_TROW + 2 * pi) = make_float2(ya[i], yb[i]);
      }
    }
    if (p.layout != 0) {
      // ---- (B, 1, n_mels, T): row segments of the tile, 8 rows x 4 column pairs per instruction ----
      __syncwarp();
      const int cpair = lane & 3;
      const int t = t0 + 2 * cpair;
      float* ob = p.out + (size_t)b * p.n_cols * p.out_frames + t;
      for (int r0 = 0; r0 < p.n_mel; r0 += 8) {
        const int r = r0 + (lane >> 2);
        if (r < p.n_mel && t < t_end) {
          const float2 v = *reinterpret_cast<const float2*>(Tb + r * MS_TROW + 2 * cpair);
          float* q = ob + (size_t)r * p.out_frames;
          if (t + 1 < t_end && (((uintptr_t)q) & 7) == 0) *reinterpret_cast<float2*>(q) = v;
          else { q[0] = v.x; if (t + 1 < t_end) q[1] = v.y; }
        }
      }
      __syncwarp();
    }
    if (p.db_mode && p.clip_max) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
      if (lane == 0 && vmax > -INFINITY) atomic_max_float(p.clip_max + b, vmax);
    }
  }
}

}  // namespace b200
