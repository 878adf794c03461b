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
//   mel       lane slots = the INTERVALS between neighbouring filter peaks: every FFT bin lies under exactly two triangles
//             (the down-slope of filter s and the up-slope of filter s + 1), so the slot of interval s + 1 reads each of its
//             bins ONCE (LDS.64 = both frames) and accumulates both slopes; filter s = its own down-slope sum + the
//             up-slope sum of the previous slot (one shuffle).  Half the shared-memory reads of a slot per filter.
//             (Banks whose first filter has an up-slope of its own keep one slot per filter.)  10 log10 through
//             lg2.approx, running maximum
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
#ifndef B200_MS_TROW
#define B200_MS_TROW 10
#endif
constexpr int MS_TROW = B200_MS_TROW;                // floats per tile row: 8 frames + 2 pad (half-warp STS.64 conflict free)
constexpr int MS_TBUF = 128 * MS_TROW;               // floats per warp
constexpr float MS_DB_PER_LG2 = 3.0102999566398120f; // 10 log10(2)

struct MelFastParams {
  const float2* tw;        // [32][32] W_1024^(k1 * lane)
  const float* melw;       // [rows][32] zero-padded weights (x 1/4: the conjugate split leaves the factor out), lanes = slots;
                           // interval form: [rows][32] PAIRS (down-slope weight of filter s, up-slope weight of filter s + 1)
  const int* slot_bin;     // [32 groups'] mel bin of lane slot (group i, lane l), >= n_mel: none
  const int* slot_start;   // first FFT bin the slot reads
  int interval;            // 1: slots are intervals (natural order, slot s = filter s), 0: one slot per filter
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

// Reflect padding of torch.stft (a function of the absolute index) + a clamp for the cells under the zero part of the window.
__device__ __forceinline__ float ms_ld_edge(const float* __restrict__ x, int64_t v, int64_t n) {
  v = reflect_index(v, n, 2);
  v = v < 0 ? 0 : (v >= n ? n - 1 : v);
  return __ldg(x + v);
}

// NJ / SJ: see above (SJ > 0: hop == 32 SJ, frame b's sample j is frame a's sample j + SJ: shared loads).
// IVAL: mel slots are intervals (melw2 = (down-slope of filter s, up-slope of filter s + 1) pairs), else one slot per filter.
template <int NJ, int SJ, bool IVAL>
__global__ void __launch_bounds__(MS_THREADS, 1) melspec_fast_kernel(const FbankParams p, const MelFastParams mp) {
  extern __shared__ __align__(16) float smem[];
  float2* stw = reinterpret_cast<float2*>(smem);                 // [1024]
  float* smelw = smem + 2 * MS_N;                                // IVAL: [rows][32] float2 (down, up); else [rows][32] float
  constexpr int WPS = IVAL ? 2 : 1;                              // floats per weight slot
  float* ebuf = smelw + ((WPS * mp.rows * 32 + 3) & ~3);         // [MS_WARPS][MS_EBUF]
  float* tbuf = ebuf + MS_WARPS * MS_EBUF;                       // [MS_WARPS][MS_TBUF]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < MS_N; i += MS_THREADS) stw[i] = __ldg(mp.tw + i);
  for (int i = tid; i < WPS * mp.rows * 32; i += MS_THREADS) smelw[i] = __ldg(mp.melw + i);
  __syncthreads();
  float* Ebuf = ebuf + warp * MS_EBUF;
  float* Tb = tbuf + warp * MS_TBUF;
  float w[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) w[j] = (lane + 32 * j < p.size) ? __ldg(p.window + lane + 32 * j) : 0.f;
  // per-lane constants of the mel / store stages: first power pair of the slot, its weight column, its tile row
  const float2* mp2[4];
  const float* mw[4];
  float* trow[4];
  int mbin[4], mcnt[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool have = i < mp.groups;
    mbin[i] = have ? __ldg(mp.slot_bin + lane + 32 * i) : p.n_mel;
    mp2[i] = reinterpret_cast<const float2*>(Ebuf) + (have ? __ldg(mp.slot_start + lane + 32 * i) : 0);
    mw[i] = smelw + WPS * (mp.woff[i] * 32 + lane);
    mcnt[i] = have ? mp.maxcnt[i] : 0;
    trow[i] = Tb + (mbin[i] < p.n_mel ? mbin[i] : 0) * MS_TROW;
  }
  const int total = p.B * mp.tiles_per_clip;
  const int half = p.size >> 1;
  const bool bft = p.layout != 0, db = p.db_mode != 0;
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
    // ---- stage 0: samples lane + 32 j of the frames (frame t starts at t * hop - size / 2).  With hop == 32 SJ the two frames
    // of a pass share their loads (frame b's sample j is frame a's sample j + SJ) and so do consecutive passes: the window
    // of NL = NJ + SJ registers slides by 2 SJ per pass, and the 2 SJ new samples are loaded one pass AHEAD.
    constexpr int NL = SJ > 0 ? NJ + SJ : NJ;
    constexpr int NS = 2 * SJ;                                   // registers the window slides by per pass
    const int64_t base0 = (int64_t)t0 * p.shift - half;
    // the whole tile (all four passes, look-ahead included) inside the clip: plain loads at constant offsets from one
    // pointer; otherwise every index goes through the reflect padding
    const bool interior = base0 >= 0 && base0 + 32 * (NL + 3 * (SJ > 0 ? NS : 0)) + (SJ > 0 ? 0 : 7 * p.shift) <= n;
    const float* __restrict__ xp = x + base0 + lane;             // sample `lane` of the pass's first frame
    float xl[NL];
    if (SJ > 0 && t0 < m_eff) {
      if (interior) {
#pragma unroll
        for (int j = 0; j < NL; ++j) xl[j] = __ldg(xp + 32 * j);
      } else {
#pragma unroll
        for (int j = 0; j < NL; ++j) xl[j] = ms_ld_edge(x, base0 + lane + 32 * j, n);
      }
    }
#pragma unroll 1
    for (int pi = 0; pi < MS_TILE / 2; ++pi) {
      const int ta = t0 + 2 * pi;
      if (ta >= t_end) break;
      const bool live_a = ta < m_eff, live_b = ta + 1 < m_eff;
      float2 y[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      if (live_a) {
        float xa[NJ], xb[NJ];
        if constexpr (SJ > 0) {
          float xn[NS];
          if (pi + 1 < MS_TILE / 2 && ta + 2 < m_eff) {          // the next pass's new samples, in flight during this pass
            if (interior) {
#pragma unroll
              for (int i = 0; i < NS; ++i) xn[i] = __ldg(xp + 32 * (NL + i));
            } else {
#pragma unroll
              for (int i = 0; i < NS; ++i) xn[i] = ms_ld_edge(x, base0 + (int64_t)(2 * pi) * p.shift + lane + 32 * (NL + i), n);
            }
          }
#pragma unroll
          for (int j = 0; j < NJ; ++j) { xa[j] = xl[j]; xb[j] = xl[j + SJ]; }
#pragma unroll
          for (int j = 0; j < NL - NS; ++j) xl[j] = xl[j + NS];
#pragma unroll
          for (int i = 0; i < NS; ++i) xl[NL - NS + i] = xn[i];
        } else {
          if (interior) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) { xa[j] = __ldg(xp + 32 * j); xb[j] = __ldg(xp + p.shift + 32 * j); }
          } else {
            const int64_t base = base0 + (int64_t)(2 * pi) * p.shift;
#pragma unroll
            for (int j = 0; j < NJ; ++j) { xa[j] = ms_ld_edge(x, base + lane + 32 * j, n); xb[j] = ms_ld_edge(x, base + p.shift + lane + 32 * j, n); }
          }
        }
        xp += 2 * p.shift;
        if (!live_b) {
#pragma unroll
          for (int j = 0; j < NJ; ++j) xb[j] = 0.f;
        }
        ms_pair_power<NJ>(xa, xb, w, stw, Ebuf, lane);
        // ---- mel (lane slots), dB ----
        fk_u64 prev_up = 0ull;                                  // interval form: up-slope sum of lane 31's slot in the previous group
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2* pp = mp2[i];
          fk_u64 acc = 0ull;
          if constexpr (IVAL) {
            const float2* wr = reinterpret_cast<const float2*>(mw[i]);
            fk_u64 up = 0ull;
#pragma unroll 2
            for (int j = 0; j < mcnt[i]; ++j) {                  // (host: counts padded to even)
              const float2 wd = wr[j * 32];
              const float2 pv = pp[j];
              const fk_u64 p2 = fk_pk(pv.x, pv.y);
              acc = fk_fma2(fk_pk(wd.x, wd.x), p2, acc);
              up = fk_fma2(fk_pk(wd.y, wd.y), p2, up);
            }
            // filter s = down-slope sum of slot s + up-slope sum of slot s - 1 (lane 0: lane 31 of the previous group)
            const fk_u64 recv = __shfl_sync(0xffffffffu, lane == 31 ? prev_up : up, (lane + 31) & 31);
            acc = fk_add2(acc, recv);
            prev_up = __shfl_sync(0xffffffffu, up, 31);
          } else {
            const float* wr = mw[i];
#pragma unroll 2
            for (int j = 0; j < mcnt[i]; ++j) {
              const float wj = wr[j * 32];
              const float2 pv = pp[j];
              acc = fk_fma2(fk_pk(wj, wj), fk_pk(pv.x, pv.y), acc);
            }
          }
          float2 v = fk_upk(acc);
          if (db) {                                             // 10 log10(max(x, 1e-10)), functional.py:390-396 (ref = 1, power)
            float la, lb;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(la) : "f"(fmaxf(v.x, 1e-10f)));
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lb) : "f"(fmaxf(v.y, 1e-10f)));
            v.x = la * MS_DB_PER_LG2; v.y = lb * MS_DB_PER_LG2;
          }
          if (!live_b) v.y = v.x;                               // (keeps the pad frame out of the maximum; stored as 0 below)
          if (mbin[i] < p.n_mel) vmax = fmaxf(vmax, fmaxf(v.x, v.y));
          if (!live_b) v.y = 0.f;
          y[i] = v;
        }
      } else {
        xp += 2 * p.shift;
      }
      // ---- hand the pair to the store stage ----
      if (!bft) {
        float* o = p.out + ((size_t)b * p.out_frames + ta) * p.n_cols;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (mbin[i] < p.n_mel) {
            o[mbin[i]] = y[i].x;
            if (ta + 1 < t_end) o[p.n_cols + mbin[i]] = y[i].y;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (mbin[i] < p.n_mel) *reinterpret_cast<float2*>(trow[i] + 2 * pi) = y[i];
      }
    }
    if (bft) {
      // ---- (B, 1, n_mels, T): row segments of the tile, 8 rows x 4 column pairs per instruction.  A lane's 8-byte alignment
      // is the same for all its rows (8 rows further = 32 T bytes), so the aligned / scalar choice is made once. ----
      __syncwarp();
      const int cpair = lane & 3;
      const int t = t0 + 2 * cpair;
      if (t < t_end) {
        const bool two = t + 1 < t_end;
        float* q = p.out + ((size_t)b * p.n_cols + (lane >> 2)) * p.out_frames + t;
        const float* tr = Tb + (lane >> 2) * MS_TROW + 2 * cpair;
        const size_t qstep = (size_t)8 * p.out_frames;
        const bool al = two && ((uintptr_t)q & 7) == 0;
        const int nr = (p.n_mel - (lane >> 2) + 7) >> 3;        // rows of this lane
        if (al) {
#pragma unroll 4
          for (int r = 0; r < nr; ++r) { *reinterpret_cast<float2*>(q) = *reinterpret_cast<const float2*>(tr); q += qstep; tr += 8 * MS_TROW; }
        } else {
#pragma unroll 4
          for (int r = 0; r < nr; ++r) {
            const float2 v = *reinterpret_cast<const float2*>(tr);
            q[0] = v.x;
            if (two) q[1] = v.y;
            q += qstep; tr += 8 * MS_TROW;
          }
        }
      }
      __syncwarp();
    }
    if (db && p.clip_max) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
      if (lane == 0 && vmax > -INFINITY) atomic_max_float(p.clip_max + b, vmax);
    }
  }
}

}  // namespace b200
