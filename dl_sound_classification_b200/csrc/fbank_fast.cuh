// Tuned fused kernel for the AST configuration: 25 ms / 10 ms frames at 16 kHz (400 / 160
// samples), 512-point FFT, up to 128 mel bins, snip_edges, no energy column.
//
// One CTA (256 threads, 2 CTAs per SM) streams a segment of one clip in chunks of 32 frames:
//
//   load      input chunk -> smem (float4 when 16-B aligned), zero-filled outside the clip
//   resample  44.1 kHz -> 16 kHz polyphase FIR: lanes = 32 hops (smem stride 441, odd ->
//             conflict-free), warps = groups of 5 phases whose 175 taps are broadcast from
//             smem as float4; 16 kHz samples go to a 34-hop ring (hop stride 161) and never
//             touch HBM.  Other rates use a per-sample loop into the same ring.
//   FFT       each warp takes 4 frames: DC removal + pre-emphasis + window on load, two
//             frames packed into one complex 512-point transform = radix-16 in registers ->
//             XOR-swizzled smem transpose -> radix-32 in registers; conjugate split by warp
//             shuffle; |.|^2
//   mel       lanes = mel bins, sparse triangular weights from smem, log, normalise,
//             SpecAugment zero-fill, coalesced store (each output written once).
#pragma once
#include "common.cuh"
#include "fbank_generic.cuh"
#include "fft_regs.cuh"

namespace b200 {

constexpr int FK_THREADS = 256;
constexpr int FK_CH = 32;                 // frames (= 16 kHz hops) per chunk
constexpr int FK_SIZE = 400, FK_SHIFT = 160, FK_N = 512;
constexpr int FK_RP = 5;                  // phases per group
constexpr int FK_NG = 32;                 // groups: 160 phases
constexpr int FK_LT = 35;                 // taps kept per phase (34 non-zero + alignment slack)
constexpr int FK_WIN = 45;                // input samples one lane reads per group
constexpr int FK_GROUP_FLOATS = 176;      // 175 taps in consumption order, padded to float4
constexpr int FK_ORIG = 441, FK_NEW = 160, FK_KLEN = 475, FK_WIDTH = 17;
constexpr int FK_RING_STRIDE = 161;       // hop stride in the ring: odd -> conflict-free column stores
constexpr int FK_RING_HOPS = 34;
constexpr int FK_RING_FLOATS = 5476;      // 34 * 161 = 5474, padded to a multiple of 4
constexpr int FK_EROW = 33;               // float2 per exchange row (32 + 1 pad: conflict-free, no index math)
constexpr int FK_EB_SKEW = 8;             // float2 between the two exchange buffers: rows of transform B sit 8 bank pairs from transform A's
constexpr int FK_EBUF = 2 * 16 * FK_EROW * 2 + 2 * FK_EB_SKEW;   // floats per warp: two 16 x 33 complex exchange buffers; later 256 x 4 powers
constexpr int FK_XFLOATS = 15040;         // 33 * 441 + 475 + slack, multiple of 4
constexpr int FK_TAPFLOATS = FK_NG * FK_GROUP_FLOATS;   // 5632
// Region AT = [x chunk | taps]; the FFT phase reuses its front as 8 x FK_EBUF exchange buffers, so the
// taps are re-staged (cp.async, L2-resident) together with every input chunk.
constexpr int FK_ATFLOATS = FK_XFLOATS + FK_TAPFLOATS;
static_assert(8 * FK_EBUF <= FK_ATFLOATS, "exchange buffers must fit the x|taps region");

__host__ __device__ constexpr int fk_off(int r) { return r == 0 ? 0 : r == 1 ? 2 : r == 2 ? 4 : r == 3 ? 7 : 10; }

struct FastParams {
  const float* taps;      // [32][176] consumption-ordered taps of the 441 -> 160 resampler
  const int* k0g;         // [32] first input index (dense tap coordinates) of each group
  int fast_rate_id;       // index in the rate table that is 44100 -> 16000, or -1
  const float2* tw;       // [16][32]  W_512^(k1 * lane)
  const float* melw;      // [sum maxcnt][32] zero-padded weights, lanes = bins
  int mel_groups;         // ceil(n_mel / 32)
  int mel_maxcnt[4];      // longest filter in each group of 32 bins
  int mel_woff[4];        // row offset of each group in melw
  int mel_rows;           // sum of mel_maxcnt
  // lane slot (group i, lane l) owns mel bin mel_slot_bin[32 i + l] (a permutation inside each group of 32 bins) and
  // starts its taps at FFT bin mel_slot_start[32 i + l] (<= the filter's first bin; leading weights are zero):
  // chosen on the host so the 8 lanes of every quarter-warp read 8 different 16-B bank groups (LDS.128 conflict free)
  const int* mel_slot_bin;
  const int* mel_slot_start;
  int gen_part[B200_MAX_RATES];   // outputs per staging pass for rates on the per-sample path
  int seg_frames, segs;
  int ast_bank;           // 1: filter lengths per group are (2,3,6,10) -> fully unrolled mel
  // warp-specialised kernel (fbank_ws.cuh): register-resident taps, phase r of group g starts at dense
  // index ws_k0g[g] + {0,2,4,6,10}[r] and keeps 36 taps (even offsets keep input pairs aligned for FFMA2)
  const float* ws_taps;   // [32][5][36]
  const int* ws_k0g;      // [32]
  int ws_ok;
  // resampler mode of each entry of the rate table (fbank_ws.cuh): 0 per-sample loop / identity, 1 = 441 -> 160,
  // 2 = 3 -> 1 (ws_t48: [2][21] tap pairs, even / odd window starts), 3 = 441 -> 320 (ws_t22: [2 parities][32][5][20],
  // ws_k22: [2][32] window starts)
  int ws_mode[B200_MAX_RATES];
  int ws_multi;           // some entry has mode >= 2: launch the MULTI instantiation
  int ws_persist;         // per launch: 0 = one item per CTA; 1 = grid of #SMs CTAs striding over the items (dense batches);
                          // 2 = the same with items claimed from ws_counter (ragged batches: uneven clips)
  int* ws_counter;        // per launch: zeroed device counter of the dynamic form
  int ws_counter_slot;    // host side: its slot in the plan's counter ring (-1: none)
  const float* ws_t48;
  const float* ws_t22;
  const int* ws_k22;
};

// Copy x[in_lo, in_lo + nx) of the clip into A[sh + i]; returns sh (0..3), chosen so that
// 16-B aligned global addresses land on 16-B aligned shared addresses.
__device__ __forceinline__ int fk_load_x(const ClipInfo& c, int64_t in_lo, int nx, float* A) {   // [phase: load_x]
  const int tid = threadIdx.x;
  const float* g = c.wav + in_lo;
  const int sh = (int)(((uintptr_t)g >> 2) & 3);
  float* d = A + sh;
  int head = (4 - sh) & 3;
  if (head > nx) head = nx;
  if (tid < head) {
    int64_t s = in_lo + tid;
    d[tid] = (s >= 0 && s < c.n_in) ? __ldg(g + tid) : 0.f;
  }
  const int nvec = (nx - head) >> 2;
  for (int v = tid; v < nvec; v += FK_THREADS) {
    const int i = head + 4 * v;
    const int64_t s = in_lo + i;
    float4 x;
    if (s >= 0 && s + 3 < c.n_in) {
      x = __ldg(reinterpret_cast<const float4*>(g + i));
    } else {
      x.x = (s >= 0 && s < c.n_in) ? __ldg(g + i) : 0.f;
      x.y = (s + 1 >= 0 && s + 1 < c.n_in) ? __ldg(g + i + 1) : 0.f;
      x.z = (s + 2 >= 0 && s + 2 < c.n_in) ? __ldg(g + i + 2) : 0.f;
      x.w = (s + 3 >= 0 && s + 3 < c.n_in) ? __ldg(g + i + 3) : 0.f;
    }
    *reinterpret_cast<float4*>(d + i) = x;
  }
  const int tail = head + 4 * nvec + tid;
  if (tail < nx) {
    int64_t s = in_lo + tail;
    d[tail] = (s >= 0 && s < c.n_in) ? __ldg(g + tail) : 0.f;
  }
  return sh;
}

// Per-sample polyphase loop (any rate): resampled samples [s0, s0 + count) -> ring, where
// ring sample 0 is absolute resampled index ring_base.  xs[i] = x[in_lo + i].
__device__ __forceinline__ void fk_resample_generic(const RateDev& R, const float* xs, int64_t in_lo,   // [phase: resample_generic]
                                                    int64_t s0, int count, float* ring, int64_t ring_base) {
  for (int t = threadIdx.x; t < count; t += FK_THREADS) {
    const int64_t s = s0 + t;
    const int64_t q = s / R.nw;
    const int ph = (int)(s - q * R.nw);
    const float* tp = R.taps + (size_t)ph * R.L;
    const float* x = xs + (q * R.orig + __ldg(R.k0 + ph) - R.width - in_lo);
    float acc = 0.f;
    for (int j = 0; j < R.L; ++j) acc = fmaf(__ldg(tp + j), x[j], acc);
    const int rel = (int)(s - ring_base);
    ring[rel + rel / FK_SHIFT] = acc;
  }
}

// One resampler task: 5 phases x 32 hops (lane = hop).  T4 = the group's 44 float4 of taps in
// consumption order, xs = this lane's first input sample, yo = ring slot of phase 5g of this hop.
__device__ __forceinline__ void fk_resample_group(const float4* __restrict__ T4, const float* __restrict__ xs,   // [phase: resample_group]
                                                  float* __restrict__ yo) {
  float acc[FK_RP];
#pragma unroll
  for (int r = 0; r < FK_RP; ++r) acc[r] = 0.f;
  float4 cur = make_float4(0.f, 0.f, 0.f, 0.f);
  int e = 0;
#pragma unroll
  for (int u = 0; u < FK_WIN; ++u) {
    const float xv = xs[u];
#pragma unroll
    for (int r = 0; r < FK_RP; ++r) {
      const int j = u - fk_off(r);
      if (j >= 0 && j < FK_LT) {
        if ((e & 3) == 0) cur = T4[e >> 2];
        const float tap = (e & 3) == 0 ? cur.x : (e & 3) == 1 ? cur.y : (e & 3) == 2 ? cur.z : cur.w;
        acc[r] = fmaf(tap, xv, acc[r]);
        ++e;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < FK_RP; ++r) yo[r] = acc[r];
}


__device__ __forceinline__ void fk_cp_async16(float* dst_smem, const float* src_gmem) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void fk_cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// Asynchronous variant of fk_load_x for chunks that lie strictly inside the clip (no zero
// fill): whole 16-B lines, global -> shared without passing through registers.   // [phase: load_x]
__device__ __forceinline__ int fk_load_x_async(const float* g /* = clip + in_lo */, int nx, float* A) {
  const int sh = (int)(((uintptr_t)g >> 2) & 3);
  const float* g0 = g - sh;                       // 16-B aligned
  const int nvec = (nx + sh + 3) >> 2;
  for (int v = threadIdx.x; v < nvec; v += FK_THREADS) fk_cp_async16(A + 4 * v, g0 + 4 * v);
  return sh;
}

// Per-lane constants of the frame pass (live for the whole kernel).   // [phase: -] (helpers: charged to their caller)
struct FkLane {
  float win[13];        // window at n = lane + 32 j
  int mstart[4];        // first FFT bin read by lane slot (i, lane)
  int mbin[4];          // output offset of the slot's mel bin m: m (kaldi layout, stats) or m * out_frames (AST layout); -1: no bin
  float nscale[4], nshift[4];   // epilogue: y = v * nscale + nshift, v = lg2(mel) (log output) or mel; see fk_fold_norm
  __device__ __forceinline__ float w(int j) const { return win[j]; }
  __device__ __forceinline__ int ms(int i) const { return mstart[i]; }
  __device__ __forceinline__ int bin(int i) const { return mbin[i]; }
  __device__ __forceinline__ float scale(int i) const { return nscale[i]; }
  __device__ __forceinline__ float shift(int i) const { return nshift[i]; }
};

// Epilogue constants of mel bin m: normalisation (x - mean) * (target_std / std) + target_mean folded with the ln 2 of
// the lg2-based log into ONE fma per output: y = lg2(mel) * (ln2 * s) + fma(-mean, s, target_mean).  Without
// normalisation this is lg2 * ln2 exactly as before; a frequency-masked bin has scale = shift = 0 (the column is 0.0).
__device__ __forceinline__ void fk_fold_norm(const FbankParams& p, bool stats, bool log_out, int m, bool fmask,
                                             float& scale, float& shift) {
  const bool norm = !stats && p.n_stats > 0 && m < p.n_mel;
  const int si = p.n_stats == 1 ? 0 : m;
  const float s = norm ? p.target_std / __ldg(p.std + si) : 1.f;
  const float k = log_out ? 0.69314718055994531f : 1.f;
  scale = fmask ? 0.f : (norm ? k * s : k);
  shift = (fmask || !norm) ? 0.f : fmaf(-__ldg(p.mean + si), s, p.target_mean);
}

constexpr int FK_LANE_ROWS = 29;

// Packed FP32 pairs: see fft_regs.cuh (fx_*).
typedef fx2 fk_u64;
__device__ __forceinline__ fk_u64 fk_pk(float lo, float hi) { return fx_pk(lo, hi); }
__device__ __forceinline__ float2 fk_upk(fk_u64 v) { return fx_upk(v); }
__device__ __forceinline__ fk_u64 fk_add2(fk_u64 a, fk_u64 b) { return fx_add(a, b); }
__device__ __forceinline__ fk_u64 fk_sub2(fk_u64 a, fk_u64 b) { return fx_sub(a, b); }
__device__ __forceinline__ fk_u64 fk_mul2(fk_u64 a, fk_u64 b) { return fx_mul(a, b); }
__device__ __forceinline__ fk_u64 fk_fma2(fk_u64 a, fk_u64 b, fk_u64 c) { return fx_fma(a, b, c); }

// DC removal + pre-emphasis + window for frames (row, row+1) -> packed complex z[n1], n = lane + 32 n1.
// yb = ring + row * RS + lane.  Loads are unconditional: every ring row holds finite data.   // [phase: stage0_frames]
// The two frames of the pair are the two halves of f32x2 registers all the way (z[j] = (frame a, frame b) is exactly
// such a pair): sum, DC subtraction and pre-emphasis are packed, in the reference's own order of roundings
// (kaldi.py:183-204; the per-frame sum order and the fused multiply-add of the pre-emphasis are those of round 1).
template <int RS, class LC>     // RS = ring hop stride in floats (160 = contiguous 16 kHz samples, 161 = padded rows)
__device__ __forceinline__ void fk_stage0_pair(const float* __restrict__ yb, const LC& L, float dc_scale,
                                               float preemph, int lane, float2 (&z)[16]) {
  constexpr int PADR = RS - FK_SHIFT;      // extra floats per hop row
  const int d0 = lane == 0 ? 0 : 1;        // j = 0: replicate pad at the frame start (kaldi.py:195-198)
  const int d5 = lane == 0 ? 1 + PADR : 1; // j = 5, 10: n - 1 sits in the previous hop row
  const float* ya = yb;                    // frame a
  const float* yc = yb + RS;               // frame b = the next hop row
  fk_u64 yv[13], s2 = 0ull;
#pragma unroll
  for (int j = 0; j < 13; ++j) {
    const int off = 32 * j + PADR * (j >= 10 ? 2 : (j >= 5 ? 1 : 0));
    yv[j] = fk_pk(ya[off], yc[off]);
    if (j < 12) s2 = fk_add2(s2, yv[j]);
  }
  s2 = fk_add2(s2, lane < 16 ? yv[12] : 0ull);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s2 = fk_add2(s2, __shfl_xor_sync(0xffffffffu, s2, o));
  const fk_u64 mean2 = fk_mul2(s2, fk_pk(dc_scale, dc_scale));     // dc_scale = 1 / 400, or 0 without DC removal
  const fk_u64 npre2 = fk_pk(-preemph, -preemph);
#pragma unroll
  for (int j = 0; j < 13; ++j) {
    const int off = 32 * j + PADR * (j >= 10 ? 2 : (j >= 5 ? 1 : 0));
    const int po = (j == 0) ? off - d0 : ((j == 5 || j == 10) ? off - d5 : off - 1);
    const fk_u64 prev = fk_pk(ya[po], yc[po]);
    // (y[n] - mean) - c (y[n-1] - mean), then the window
    const fk_u64 t = fk_fma2(npre2, fk_sub2(prev, mean2), fk_sub2(yv[j], mean2));
    z[j] = fk_upk(fk_mul2(t, fk_pk(L.w(j), L.w(j))));            // SASS: FMUL2 with a broadcast .F32 operand
  }
#pragma unroll
  for (int j = 13; j < 16; ++j) z[j] = make_float2(0.f, 0.f);
}

// 16-point DFT over n1 (in place, X[k1] ends up in z[bitrev(k1)]).   // [phase: fft_stage1]
__device__ __forceinline__ void fk_dft16_pruned(float2 (&z)[16]) {
  // z[13..15] are the zero padding 400 -> 512: the first radix-2 stage has nothing to add or subtract there
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const float2 a = z[j], b2 = z[j + 8];
    z[j] = cadd(a, b2);
    z[j + 8] = j == 0 ? csub(a, b2) : j == 1 ? mul_w32<2>(csub(a, b2)) : j == 2 ? mul_w32<4>(csub(a, b2))
             : j == 3 ? mul_w32<6>(csub(a, b2)) : mul_w32<8>(csub(a, b2));
  }
  z[13] = mul_w32<10>(z[5]); z[14] = mul_w32<12>(z[6]); z[15] = mul_w32<14>(z[7]);
  DifStages<16, 4>::run(z);
}

// a * w with the roundings of (fma(-a.y, w.y, a.x w.x), fma(a.y, w.x, a.x w.y)): FMUL2 + FFMA2
__device__ __forceinline__ float2 fk_cmul(float2 a, float2 w) {
  return fk_upk(fk_fma2(fk_pk(-w.y, w.x), fk_pk(a.y, a.y), fk_mul2(fk_pk(w.x, w.y), fk_pk(a.x, a.x))));
}

// Twiddle W_512^(lane k1) and store row k1 of the exchange buffers of BOTH packed transforms of the pass: the twiddle
// is the same for the two, so it is read once.
__device__ __forceinline__ void fk_stage1_store2(const float2 (&za)[16], const float2 (&zb)[16], const float2* __restrict__ stw,
                                                 int lane, float2* __restrict__ EA, float2* __restrict__ EB) {
  EA[lane] = za[0];
  EB[lane] = zb[0];
#pragma unroll
  for (int k1 = 1; k1 < 16; ++k1) {
    const float2 w = stw[k1 * 32 + lane];
    EA[k1 * FK_EROW + lane] = fk_cmul(za[bitrev_n(k1, 4)], w);
    EB[k1 * FK_EROW + lane] = fk_cmul(zb[bitrev_n(k1, 4)], w);
  }
}

// Mel filters lane + 32 i over the four frames of P4[k] = (P_f0, P_f1, P_f2, P_f3)[k].   // [phase: mel]
template <int MC>
__device__ __forceinline__ void fk_mel_group(const float4* __restrict__ P4, const float* __restrict__ wrow, int st,
                                             int mc_dyn, float (&acc)[4]) {
  // frames (0,1) and (2,3) are the halves of two f32x2 accumulators: FFMA2 with the weight broadcast
  fk_u64 a01 = 0ull, a23 = 0ull;
  if constexpr (MC >= 0) {
#pragma unroll
    for (int j = 0; j < MC; ++j) {
      const float w = wrow[j * 32];
      const float4 pv = P4[st + j];
      a01 = fk_fma2(fk_pk(w, w), fk_pk(pv.x, pv.y), a01);
      a23 = fk_fma2(fk_pk(w, w), fk_pk(pv.z, pv.w), a23);
    }
  } else {
    for (int j = 0; j < mc_dyn; ++j) {
      const float w = wrow[j * 32];
      const float4 pv = P4[st + j];
      a01 = fk_fma2(fk_pk(w, w), fk_pk(pv.x, pv.y), a01);
      a23 = fk_fma2(fk_pk(w, w), fk_pk(pv.z, pv.w), a23);
    }
  }
  const float2 r01 = fk_upk(a01), r23 = fk_upk(a23);
  acc[0] = r01.x; acc[1] = r01.y; acc[2] = r23.x; acc[3] = r23.y;
}

// One frame pass of a warp: four consecutive frames starting at output row t0, whose samples start at
// `rows` (ring rows of stride RS; 6 rows are touched) -> 4 x n_mel outputs.  n_live = how many of the four are
// real frames (the rest are pad rows).  Ebuf = this warp's FK_EBUF floats.  AST = compile-time filter lengths
// (2,3,6,10) of the AST bank.
template <bool STATS, bool AST, int RS, class LC, bool MIX = false>   // MIX: Mixup fused into the epilogue (own instantiation: the plain kernels carry none of its code)
__device__ __forceinline__ void fk_frame_pass(const FbankParams& p, const FastParams& fp, const LC& L,
                                              const float* __restrict__ rows, float* __restrict__ Ebuf,
                                              const float2* __restrict__ stw, const float* __restrict__ smelw,
                                              int b, int t0, int n_live, int row_end, int lane,
                                              int mk0, int mk1, int mk2, int mk3, double (&st_s)[4], double (&st_ss)[4]) {
  float2* EA = reinterpret_cast<float2*>(Ebuf);
  float2* EB = EA + 16 * FK_EROW + FK_EB_SKEW;
  constexpr int f0 = 0;
  const int nf = n_live;
  (void)mk2; (void)mk3;
  // Fused Mixup: the partner's cells of this pass (4 rows x n_cols) are pulled into L1 now, a whole pass ahead of the
  // epilogue that reads them
  if (MIX && !STATS && p.mix_bank != nullptr) {                                         // [phase: mixup_prefetch]
    const int j = __ldg(p.mix_partner + b);
    if (j >= 0 && t0 < row_end) {
      const float* mb = p.mix_bank + (size_t)j * p.out_frames * p.n_cols + (p.layout == 0 ? (size_t)t0 * p.n_cols : (size_t)t0);
      if (p.layout == 0) {
        const int r = lane >> 3, c = (lane & 7) * 32;                                  // 4 rows x up to 8 lines of 128 B
        if (c < p.n_cols && t0 + r < row_end) asm volatile("prefetch.global.L1 [%0];" ::"l"(mb + (size_t)r * p.n_cols + c));
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < fp.mel_groups && L.bin(i) >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(mb + L.bin(i)));
      }
    }
  }
  if (f0 < nf) {                                                                       // [phase: stage0_frames]
    const float dc_scale = p.remove_dc ? 1.f / (float)FK_SIZE : 0.f;
    // the two packed transforms (frames 0,1 and 2,3); unrolled so their dependency chains interleave (a rolled loop
    // halves the code but measured 3% slower: the pass is latency-bound, not instruction-cache bound)
    {
      float2 za[16], zb[16];
      fk_stage0_pair<RS, LC>(rows + lane, L, dc_scale, p.preemph, lane, za);
      fk_stage0_pair<RS, LC>(rows + 2 * RS + lane, L, dc_scale, p.preemph, lane, zb);
      fk_dft16_pruned(za);
      fk_dft16_pruned(zb);
      fk_stage1_store2(za, zb, stw, lane, EA, EB);
    }
    __syncwarp();
    // ---- stage 2: lane = 2 k1 + transform gathers its row, 32-point DFT over n2   // [phase: exchange]
    // (transform in the LOW lane bit: a half-warp then reads 8 rows of A and 8 of B, 16 distinct bank pairs thanks to
    // FK_EB_SKEW, and later writes the 4-frame power cells of 8 consecutive bins = 128 contiguous bytes)
    const int k1l = lane >> 1, trl = lane & 1;
    float2 u[32];
    {
      const float2* row = (trl ? EB : EA) + k1l * FK_EROW;
#pragma unroll
      for (int n2 = 0; n2 < 32; ++n2) u[n2] = row[n2];
    }
    __syncwarp();
    fft_dif<32>(u);                                                                    // [phase: fft_stage2]
    // ---- split the two real spectra (Z[k], conj Z[512-k]) and take |.|^2            // [phase: split_power]
    const int src = (((16 - k1l) & 15) << 1) | trl;
    float2* P2 = reinterpret_cast<float2*>(Ebuf) + lane;   // P4[k].{xy | zw}, k = k1 + 16 k2: float2 index 2 k + transform
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
      const float2 zk = u[bitrev_n(k2, 5)];
      const float2 own = u[bitrev_n((32 - k2) & 31, 5)];
      const float2 oth = u[bitrev_n(31 - k2, 5)];
      float px = __shfl_sync(0xffffffffu, oth.x, src);
      float py = __shfl_sync(0xffffffffu, oth.y, src);
      if (k1l == 0) { px = own.x; py = own.y; }
      // A = (Z[k] + conj Z[N-k]) / 2, B = (Z[k] - conj Z[N-k]) / 2i; the 1/4 (1/2 for magnitudes) is folded into
      // the mel weights on the host (exact: a power of two)
      const float2 sm = cadd(zk, make_float2(px, py)), df = csub(zk, make_float2(px, py));
      // (fma(sm.x, sm.x, df.y df.y), fma(sm.y, sm.y, df.x df.x)) as FMUL2 (swapped halves) + FFMA2
      float2 pw = fk_upk(fk_fma2(fk_pk(sm.x, sm.y), fk_pk(sm.x, sm.y), fk_mul2(fk_pk(df.y, df.x), fk_pk(df.y, df.x))));
      if (!AST && !p.use_power) { pw.x = sqrtf(pw.x); pw.y = sqrtf(pw.y); }      // AST variant: power spectrum only
      P2[32 * k2] = pw;
    }
    __syncwarp();
  }
  // ---- mel (lane slots = bins), log, normalise, mask, store                         // [phase: mel]
  const float4* P4 = reinterpret_cast<const float4*>(Ebuf);
  // warp-uniform output addressing, once per pass: row t0 of clip b; the lane adds its bin offset L.bin(i)
  const int ostep = p.layout == 0 ? p.n_cols : 1;
  float* const obase = STATS ? nullptr
                             : p.out + (p.layout == 0 ? ((size_t)b * p.out_frames + t0) * p.n_cols : (size_t)b * p.n_cols * p.out_frames + t0);
  // warp-uniform fast path: four live frames inside the segment and no time mask touching them
  const bool plain = (f0 + 4 <= nf) && (t0 + 4 <= row_end) && (mk1 <= 0 || t0 + 4 <= mk0 || t0 >= mk0 + mk1);
  // fused Mixup: partner cells sit at the same offsets of the partner's spectrogram (re-read per pass: L1 hits, no registers
  // held across the transform)
  const float* mbase = nullptr;
  float mlam = 1.f, moml = 0.f;
  if (MIX && !STATS && p.mix_bank != nullptr) {
    const int j = __ldg(p.mix_partner + b);
    if (j >= 0) {
      mlam = __ldg(p.mix_lam + b);
      moml = __fsub_rn(1.0f, mlam);
      mbase = p.mix_bank + (size_t)j * p.out_frames * p.n_cols + (p.layout == 0 ? (size_t)t0 * p.n_cols : (size_t)t0);
    }
  }
  // the partner's 4 x 4 cells of this lane are requested NOW, before the mel sums (the transform's registers are free
  // again), so their latency -- L1 hits after the prefetch, L2 otherwise -- hides behind the mel phase
  float mq[4][4];
  if (MIX) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int h = 0; h < 4; ++h)
        mq[i][h] = (mbase != nullptr && i < fp.mel_groups && L.bin(i) >= 0 && t0 + h < row_end) ? __ldg(mbase + L.bin(i) + h * ostep) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i >= fp.mel_groups) continue;
    const int moff = L.bin(i);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (f0 < nf) {
      const float* wrow = smelw + fp.mel_woff[i] * 32 + lane;
      if (AST) {
        if (i == 0) fk_mel_group<2>(P4, wrow, L.ms(i), 0, acc);
        else if (i == 1) fk_mel_group<3>(P4, wrow, L.ms(i), 0, acc);
        else if (i == 2) fk_mel_group<6>(P4, wrow, L.ms(i), 0, acc);
        else fk_mel_group<10>(P4, wrow, L.ms(i), 0, acc);
      } else {
        fk_mel_group<-1>(P4, wrow, L.ms(i), fp.mel_maxcnt[i], acc);
      }
    }
    if (moff >= 0) {                                                                   // [phase: epilogue_store]
      // v = lg2(max(mel, FLT_EPSILON)) (log output) or mel.  Floored cells take the constant -23 = lg2(FLT_EPSILON):
      // -23 * float(ln 2) is exactly the float32 log(FLT_EPSILON) of the reference (MUFU.LG2 itself is not exact there)
      float v[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        v[h] = acc[h];
        if (AST || p.use_log) {
          float l2;
          asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(acc[h]));
          v[h] = acc[h] > B200_FLT_EPSILON ? l2 : -23.0f;
        }
      }
      if (STATS) {
#pragma unroll
        for (int h = 0; h < 4; ++h)
          if ((f0 + h) < nf) {
            const float x = v[h] * L.scale(i);                    // STATS: scale = ln 2 (or 1), shift = 0
            st_s[i] += (double)x; st_ss[i] += (double)x * (double)x;
          }
      } else {
        float* o = obase + moff;
        if (plain) {
          const fk_u64 sc2 = fk_pk(L.scale(i), L.scale(i)), sh2 = fk_pk(L.shift(i), L.shift(i));
          float2 y01 = fk_upk(fk_fma2(fk_pk(v[0], v[1]), sc2, sh2)), y23 = fk_upk(fk_fma2(fk_pk(v[2], v[3]), sc2, sh2));
          if (MIX && mbase != nullptr) {                                                 // [phase: mixup_epilogue]
            y01.x = __fadd_rn(__fmul_rn(mlam, y01.x), __fmul_rn(moml, mq[i][0]));
            y01.y = __fadd_rn(__fmul_rn(mlam, y01.y), __fmul_rn(moml, mq[i][1]));
            y23.x = __fadd_rn(__fmul_rn(mlam, y23.x), __fmul_rn(moml, mq[i][2]));
            y23.y = __fadd_rn(__fmul_rn(mlam, y23.y), __fmul_rn(moml, mq[i][3]));
          }
          o[0] = y01.x; o[ostep] = y01.y; o[2 * ostep] = y23.x; o[3 * ostep] = y23.y;   // [phase: epilogue_store]
        } else {
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int t = t0 + h;
            if (t < row_end) {
              // H9: pad rows are 0.0 before normalisation = the folded shift
              float y = (f0 + h) < nf ? fmaf(v[h], L.scale(i), L.shift(i)) : L.shift(i);
              if (t >= mk0 && t < mk0 + mk1) y = 0.f;
              if (MIX && mbase != nullptr) y = __fadd_rn(__fmul_rn(mlam, y), __fmul_rn(moml, mq[i][h]));
              o[h * ostep] = y;
            }
          }
        }
      }
    }
  }
}

template <bool STATS, bool AST>
__global__ void __launch_bounds__(FK_THREADS, 2) fbank_fast_kernel(const FbankParams p, const FastParams fp) {   // [phase: setup]
  extern __shared__ __align__(16) float smem[];
  float* A = smem;                                   // [FK_XFLOATS] input chunk | [FK_TAPFLOATS] taps; FFT phase: exchange
  float* staps = A + FK_XFLOATS;
  float* ring = A + FK_ATFLOATS;                     // [FK_RING_FLOATS] 16 kHz samples, 34 hops x 161
  float2* stw = reinterpret_cast<float2*>(ring + FK_RING_FLOATS);   // [512]
  float* smelw = reinterpret_cast<float*>(stw + 512);               // [mel_rows * 32]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / fp.segs;
  const int seg = blockIdx.x - b * fp.segs;
  const ClipInfo c = clip_info(p, b);
  const int cap = STATS ? p.max_frames : p.out_frames;
  const int m_eff = (int)(c.m < cap ? c.m : cap);
  const int row_begin = seg * fp.seg_frames;
  if (row_begin >= cap) return;
  const int row_end = (row_begin + fp.seg_frames < cap) ? row_begin + fp.seg_frames : cap;
  if (!STATS && seg == 0 && tid == 0 && p.n_frames_out) p.n_frames_out[b] = m_eff;
  if (STATS && row_begin >= m_eff) return;

  const int rid = p.rate_id ? p.rate_id[b] : 0;
  const bool fast = (rid == fp.fast_rate_id);

  // ---- one-time staging of the small tables ---------------------------------------------------
  if (row_begin < m_eff) {
    for (int i = tid; i < 512; i += FK_THREADS) stw[i] = __ldg(fp.tw + i);
    for (int i = tid; i < fp.mel_rows * 32; i += FK_THREADS) smelw[i] = __ldg(fp.melw + i);
  }
  int mk0 = 0, mk1 = 0, mk2 = 0, mk3 = 0;
  if (!STATS && p.masks) {
    mk0 = __ldg(p.masks + (size_t)b * 4 + 0); mk1 = __ldg(p.masks + (size_t)b * 4 + 1);
    mk2 = __ldg(p.masks + (size_t)b * 4 + 2); mk3 = __ldg(p.masks + (size_t)b * 4 + 3);
  }
  FkLane L;
#pragma unroll
  for (int j = 0; j < 13; ++j) L.win[j] = (lane + 32 * j < FK_SIZE) ? __ldg(p.window + lane + 32 * j) : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool have = i < fp.mel_groups;
    const int m = have ? __ldg(fp.mel_slot_bin + lane + 32 * i) : p.n_mel;
    L.mbin[i] = m < p.n_mel ? m * ((STATS || p.layout == 0) ? 1 : p.out_frames) : -1;
    L.mstart[i] = have ? __ldg(fp.mel_slot_start + lane + 32 * i) : 0;
    fk_fold_norm(p, STATS, AST || p.use_log, m, m >= mk2 && m < mk2 + mk3, L.nscale[i], L.nshift[i]);
  }
  double st_s[4] = {0.0, 0.0, 0.0, 0.0}, st_ss[4] = {0.0, 0.0, 0.0, 0.0};
  __syncthreads();

  bool ring_valid = false;
  for (int r0 = row_begin; r0 < row_end; r0 += FK_CH) {   // [phase: chunk_control]
    int nf = m_eff - r0;
    nf = nf < 0 ? 0 : (nf > FK_CH ? FK_CH : nf);
    if (nf > 0) {
      // ================= resample: fill ring hops [hop_lo, 34) =============================
      const int hop_lo = ring_valid ? 2 : 0;
      const int64_t ring_base = (int64_t)r0 * FK_SHIFT;          // absolute index of ring sample 0
      if (fast) {
        const int64_t in_lo = (int64_t)(r0 + hop_lo) * FK_ORIG - FK_WIDTH;
        const int nx = (FK_RING_HOPS - hop_lo - 1) * FK_ORIG + FK_KLEN + 8;
        const float* g = c.wav + in_lo;
        // taps ride along with the chunk (they were overwritten by the exchange buffers)
        for (int v = tid; v < FK_TAPFLOATS / 4; v += FK_THREADS) fk_cp_async16(staps + 4 * v, fp.taps + 4 * v);
        int sh;
        const bool interior = in_lo >= 4 && in_lo + nx + 4 <= c.n_in && (g - 4 >= p.wav);
        if (interior) sh = fk_load_x_async(g, nx, A);
        else sh = fk_load_x(c, in_lo, nx, A);
        if (ring_valid) {          // carry the last two hops to the front while the copies fly
          float keep0 = 0.f, keep1 = 0.f;
          if (tid < 2 * FK_RING_STRIDE) keep0 = ring[32 * FK_RING_STRIDE + tid];
          if (tid + FK_THREADS < 2 * FK_RING_STRIDE) keep1 = ring[32 * FK_RING_STRIDE + tid + FK_THREADS];
          __syncthreads();
          if (tid < 2 * FK_RING_STRIDE) ring[tid] = keep0;
          if (tid + FK_THREADS < 2 * FK_RING_STRIDE) ring[tid + FK_THREADS] = keep1;
        }
        fk_cp_async_wait_all();
        __syncthreads();
        const float* xs = A + sh;
        if (!ring_valid)       // prologue: hops 0 and 1 on the per-sample path
          fk_resample_generic(c.R, xs, in_lo, ring_base, 2 * FK_SHIFT, ring, ring_base);
        const float* xl = xs + (ring_valid ? 0 : 2 * FK_ORIG) + lane * FK_ORIG;
        float* yl = ring + (2 + lane) * FK_RING_STRIDE;
#pragma unroll 1
        for (int gi = 0; gi < FK_NG / 8; ++gi) {
          const int g2 = warp * (FK_NG / 8) + gi;
          fk_resample_group(reinterpret_cast<const float4*>(staps + g2 * FK_GROUP_FLOATS),
                            xl + __ldg(fp.k0g + g2), yl + FK_RP * g2);
        }
      } else {
        if (ring_valid) {
          float keep0 = 0.f, keep1 = 0.f;
          if (tid < 2 * FK_RING_STRIDE) keep0 = ring[32 * FK_RING_STRIDE + tid];
          if (tid + FK_THREADS < 2 * FK_RING_STRIDE) keep1 = ring[32 * FK_RING_STRIDE + tid + FK_THREADS];
          __syncthreads();
          if (tid < 2 * FK_RING_STRIDE) ring[tid] = keep0;
          if (tid + FK_THREADS < 2 * FK_RING_STRIDE) ring[tid + FK_THREADS] = keep1;
        }
        if (c.R.identity) {
          const int cnt = (FK_RING_HOPS - hop_lo) * FK_SHIFT;
          const int64_t s0 = ring_base + hop_lo * FK_SHIFT;
          for (int t = tid; t < cnt; t += FK_THREADS) {
            const int64_t s = s0 + t;
            const int rel = (int)(s - ring_base);
            ring[rel + rel / FK_SHIFT] = (s < c.n_in) ? __ldg(c.wav + s) : 0.f;
          }
        } else {
          const int total = (FK_RING_HOPS - hop_lo) * FK_SHIFT;
          const int part = fp.gen_part[rid];
          for (int done = 0; done < total; done += part) {
            const int cnt = (total - done < part) ? total - done : part;
            const int64_t s0 = ring_base + hop_lo * FK_SHIFT + done;
            const int64_t q_lo = s0 / c.R.nw, q_hi = (s0 + cnt - 1) / c.R.nw;
            const int64_t in_lo = q_lo * c.R.orig - c.R.width;
            const int nx = (int)((q_hi - q_lo) * c.R.orig + c.R.klen);
            if (done > 0) __syncthreads();
            const int sh = fk_load_x(c, in_lo, nx, A);
            __syncthreads();
            fk_resample_generic(c.R, A + sh, in_lo, s0, cnt, ring, ring_base);
          }
        }
      }
      __syncthreads();
      ring_valid = true;
    }

    // ================= FFT + mel: warp w owns frames 4w .. 4w+3 of the chunk ===============
    {
      int n_live = nf - 4 * warp;
      n_live = n_live < 0 ? 0 : (n_live > 4 ? 4 : n_live);
      fk_frame_pass<STATS, AST, FK_RING_STRIDE, FkLane>(p, fp, L, ring + 4 * warp * FK_RING_STRIDE, A + warp * FK_EBUF, stw, smelw,
                                                b, r0 + 4 * warp, n_live, row_end, lane, mk0, mk1, mk2, mk3, st_s, st_ss);
    }
    __syncthreads();      // the x|taps region (exchange / power) and the ring are reused by the next chunk   // [phase: chunk_control]
  }

  if (STATS) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = L.mbin[i];
      if (i < fp.mel_groups && m >= 0) {
        atomicAdd(p.sums + m, st_s[i]);
        atomicAdd(p.sums + p.n_cols + m, st_ss[i]);
      }
    }
    if (tid == 0) {
      int real = m_eff - row_begin;
      real = real > (row_end - row_begin) ? (row_end - row_begin) : real;
      if (real > 0) atomicAdd(p.sums + 2 * p.n_cols, (double)real);
    }
  }
}

}  // namespace b200
