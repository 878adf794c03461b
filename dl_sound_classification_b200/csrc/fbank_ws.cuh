// Warp-specialised, persistent fused kernel for the AST configuration (same envelope as fbank_fast.cuh).
//
// One CTA per SM, 384 threads = three warpgroups working as a producer/consumer pipeline:
//
//   warpgroup 0 (4 "R" warps, 232 registers/thread after setmaxnreg.inc): polyphase resampler with the taps in
//     REGISTERS, one mode per input rate (one chunk loop per mode, so only that mode's state is live):
//       441 -> 160 (44.1 kHz): lane = phase group (5 phases, 180 taps); 23 LDS.64 feed 90 packed FFMA2 (fma.rn.f32x2) per
//         5 outputs.  A 64-bit load needs an even sample address and 441 is odd, so the R warps come in two classes
//         whose tap windows start at dense indices of opposite parity (R warps 0-1: ws_k0g[g] + {0,2,4,6,10}; R warps
//         2-3: ws_k0g[32 + g] + {0,2,4,8,10}): for every hop exactly one class reads lane g's window from an even
//         address.  Inside a class lane g visits its 16 hops in the order m = (slot + skew[g]) mod 16, q = 2m + parity;
//         skew[g] = 9*((g mod 16) - base[g]) mod 16 puts the 16 lanes of each half-warp on 16 distinct 8-byte bank
//         pairs in every load (441 = 9 mod 16, 9*9 = 1 mod 16), and the ring stride of 160 keeps the 5-wide column
//         stores conflict free whatever the hop.
//       3 -> 1 (48 kHz): one phase of 41 taps shared by all lanes, held as even- and odd-aligned pairs.
//       441 -> 320 (22.05 kHz): 320 phases = two hops; even / odd R warps hold the taps of even / odd hops.
//       anything else: per-sample loop (16 kHz input: plain copy).
//     Input chunks (32 hops) arrive by ONE TMA bulk copy (cp.async.bulk + mbarrier complete_tx), the next chunk is pulled
//     into L2 by a bulk prefetch meanwhile; at a clip edge the lines outside / across the edge are assembled by hand.
//   warpgroups 1-2 (8 "F" warps, 136 registers/thread after setmaxnreg.dec): frame passes of fbank_fast.cuh
//     (DC/pre-emphasis/window, packed 512-point FFT, |.|^2, mel, log, epilogue, store), pass slot p -> warp p mod 8.
//
// The 16 kHz signal lives in a shared-memory ring of three 32-hop slots (+6 mirrored rows so every pass reads 6
// contiguous rows); slots are handed over with mbarriers (full: 128 R arrivals, empty: 8 pass slots x 32 lanes).
// Nothing intermediate touches HBM.
//
// Work distribution (fp.ws_persist): 0 = one item (clip segment) per CTA; 1 = grid of #SMs CTAs striding over the items
// (dense batches); 2 = the same with items claimed from a global counter through a shared-memory queue (ragged
// batches).  In the persistent forms ring slots, barrier phases and pass slots are numbered per CTA, so the pipeline
// runs straight through clip boundaries.
#pragma once
#include "fbank_fast.cuh"

namespace b200 {

// Optional pipeline timing (compile with -DB200_WS_TIMING): per-role cycle counters summed over all CTAs into
// g_ws_timing[8] = {R load-wait, R empty-wait, R compute, R chunks, F full-wait, F pass, F passes, -}.
#ifdef B200_WS_TIMING
__device__ unsigned long long g_ws_timing[8];
#define WS_T0() const long long t0_ = clock64()
#define WS_TACC(slot, since) do { if (lane == 0) atomicAdd(&g_ws_timing[slot], (unsigned long long)(clock64() - (since))); } while (0)
#else
#define WS_T0()
#define WS_TACC(slot, since)
#endif

constexpr int WS_THREADS = 384;
#ifndef B200_WS_HINT_NS
#define B200_WS_HINT_NS 2000
#endif
constexpr unsigned WS_WAIT_HINT_NS = B200_WS_HINT_NS;                    // try_wait suspend-time hint: a waiting warp sleeps instead of polling every ~40 cycles
constexpr int WS_R_WARPS = 4, WS_F_WARPS = 8;                // warpgroup 0 = R warps, warpgroups 1-2 = F warps (2 R + 10 F measured slower: 0.82 vs 0.75 ms)
constexpr int WS_R_ITERS = 32 / WS_R_WARPS;                   // hops per R warp per chunk
constexpr int WS_R_THREADS = 32 * WS_R_WARPS;
#ifndef B200_WS_SLOTS
#define B200_WS_SLOTS 3
#endif
static_assert(B200_WS_SLOTS >= 2 && B200_WS_SLOTS <= 3, "bars[8]: full[], empty[], the TMA barrier; a fourth slot needs a ninth");
constexpr int WS_SLOTS = B200_WS_SLOTS;                      // ring slots of 32 hops (2 measured 16 % slower: the R warps must be able to run two chunks ahead)
constexpr int WS_RING_ROWS = 32 * WS_SLOTS;                   // 96 hops
constexpr int WS_MIRROR = 6;                                  // rows 96..101 repeat rows 0..5
constexpr int WS_RING_FLOATS = (WS_RING_ROWS + WS_MIRROR) * FK_SHIFT;   // 16320
#ifndef B200_WS_XFLOATS
#define B200_WS_XFLOATS 15424
#endif
constexpr int WS_XFLOATS = B200_WS_XFLOATS;                             // 48 kHz chunk: 3*(32*160 - 1) + 41 + 8 slack + 3 shift (44.1 kHz: 31*441 + 475 + 11), multiple of 32
constexpr int WS_W48 = 19, WS_P48 = 21;                       // 3 -> 1: width 19, 41 taps = 21 pairs (one zero)
constexpr int WS_W22 = 9, WS_LT22 = 20;                       // 441 -> 320: width 9, 16-17 taps per phase kept as 20
__host__ __device__ constexpr int ws_off22(int r) { return r < 2 ? 0 : r == 2 ? 2 : 4; }
constexpr int WS_LT = 36;                                     // taps per phase (34 non-zero + even alignment)
constexpr int WS_GROUP_FLOATS = FK_RP * WS_LT;                // 180

// window start of phase r of a group relative to the group's k0, per R-warp class (all even: input pairs stay aligned)
__host__ __device__ constexpr int ws_off(int r) { return r == 0 ? 0 : r == 1 ? 2 : r == 2 ? 4 : r == 3 ? 6 : 10; }
__host__ __device__ constexpr int ws_off1(int r) { return r == 0 ? 0 : r == 1 ? 2 : r == 2 ? 4 : r == 3 ? 8 : 10; }
__host__ __device__ constexpr int ws_offc(int cls, int r) { return cls ? ws_off1(r) : ws_off(r); }

// ---- mbarrier / named barrier / register reallocation --------------------------------------------   // [phase: -]
__device__ __forceinline__ void ws_mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ws_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void ws_mbar_wait(uint64_t* bar, unsigned parity) {   // [phase: mbar_wait]
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WS_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra WS_DONE_%=;\n"
      "bra WS_WAIT_%=;\n"
      "WS_DONE_%=:\n"
      "}\n" ::"r"(a), "r"(parity), "r"(WS_WAIT_HINT_NS) : "memory");
}
// [phase: -]
__device__ __forceinline__ void ws_bar_r() { asm volatile("bar.sync 1, %0;" ::"n"(WS_R_THREADS) : "memory"); }

// 1-D bulk copy global -> shared through the TMA unit (UBLKCP): one instruction moves the whole input chunk and
// signals `bar` with the byte count; src/dst 16-B aligned, bytes a multiple of 16.
__device__ __forceinline__ void ws_tma_load(float* dst_smem, const float* src_gmem, unsigned bytes, uint64_t* bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem), m = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(d), "l"(src_gmem), "r"(bytes), "r"(m) : "memory");
}

__device__ __forceinline__ unsigned long long ws_pack(float a, float b) {
  return (unsigned long long)__float_as_uint(a) | ((unsigned long long)__float_as_uint(b) << 32);
}
__device__ __forceinline__ unsigned long long ws_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// One hop of one lane: 5 phases from 46 input samples = 23 aligned 64-bit loads; T[r][jj] = taps (2jj, 2jj+1) of
// phase r, whose window starts ws_offc(CLS, r) samples after xs.   // [phase: ws_resample]
template <int CLS>
__device__ __forceinline__ void ws_resample_hop(const float* __restrict__ xs, const unsigned long long (&T)[FK_RP * WS_LT / 2],
                                                float (&y)[FK_RP]) {
  unsigned long long acc[FK_RP] = {0ull, 0ull, 0ull, 0ull, 0ull};
  const unsigned long long* __restrict__ xp2 = reinterpret_cast<const unsigned long long*>(xs);   // 8-byte aligned by construction
#pragma unroll
  for (int u = 0; u < ws_offc(CLS, FK_RP - 1) + WS_LT; u += 2) {
    const unsigned long long xp = xp2[u >> 1];
#pragma unroll
    for (int r = 0; r < FK_RP; ++r) {
      const int j = u - ws_offc(CLS, r);
      if (j >= 0 && j < WS_LT) acc[r] = ws_fma2(xp, T[r * (WS_LT / 2) + (j >> 1)], acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < FK_RP; ++r) y[r] = __uint_as_float((unsigned)acc[r]) + __uint_as_float((unsigned)(acc[r] >> 32));
}

// Alignment shift of a chunk: x[in_lo + i] lands at xbuf[sh + i], so 16-B aligned global lines are 16-B aligned
// in shared memory.
__device__ __forceinline__ int ws_shift(const float* g) { return (int)(((uintptr_t)g >> 2) & 3); }

// R-side load of an EDGE chunk x[in_lo, in_lo+nx) (128 threads): whole lines inside the clip go through cp.async,
// lines that straddle a clip edge are assembled by hand, samples outside the clip are zeros (the zero padding
// of torchaudio's conv, functional.py:1424).  Interior chunks use one TMA bulk copy instead.   // [phase: ws_load]
__device__ __forceinline__ void ws_load_edge_issue(const ClipInfo& c, int64_t in_lo, int nx, int sh, float* xbuf, int rt) {
  const float* g = c.wav + in_lo;
  const float* g0 = g - sh;                                  // 16-B aligned; slot v <-> elements i = 4v - sh + {0..3}
  const int nvec = (nx + sh + 3) >> 2;
  const int lo = in_lo < 0 ? (int)(-in_lo) : 0;              // first element index that exists in the clip
  const int64_t hi64 = c.n_in - in_lo;                       // one past the last
  const int hi = hi64 > nx + 8 ? nx + 8 : (hi64 < 0 ? 0 : (int)hi64);
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(xbuf);
  for (int v = rt; v < nvec; v += WS_R_THREADS) {
    const int i0 = 4 * v - sh;
    if (i0 >= lo && i0 + 4 <= hi) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + 16u * v), "l"(g0 + 4 * v) : "memory");
    } else {
      float4 x;
      x.x = (i0 >= lo && i0 < hi) ? __ldg(g + i0) : 0.f;
      x.y = (i0 + 1 >= lo && i0 + 1 < hi) ? __ldg(g + i0 + 1) : 0.f;
      x.z = (i0 + 2 >= lo && i0 + 2 < hi) ? __ldg(g + i0 + 2) : 0.f;
      x.w = (i0 + 3 >= lo && i0 + 3 < hi) ? __ldg(g + i0 + 3) : 0.f;
      *reinterpret_cast<float4*>(xbuf + 4 * v) = x;
    }
  }
}
__device__ __forceinline__ void ws_load_edge(const ClipInfo& c, int64_t in_lo, int nx, int sh, float* xbuf, int rt) {
  ws_load_edge_issue(c, in_lo, nx, sh, xbuf, rt);
  fk_cp_async_wait_all();
}

constexpr int WS_PAD_COLS = 128;                              // columns of the per-warp pad-row table (the envelope's n_mel <= 128)
constexpr int WS_QN = 8;                                      // depth of the item queue of the dynamic persistent form
__device__ __forceinline__ int ws_claim(const FastParams& fp, int total) {
  const int id = (int)gridDim.x + atomicAdd(fp.ws_counter, 1);
  return id < total ? id : -1;
}

// Geometry of one work item = one segment of one clip (recomputed by each role from the item number).
struct WsItem {
  int b, seg, row_begin, row_end, m_eff;
  int n_pass;        // passes that own output rows
  int n_real;        // real frames of the segment
  int last_hop;      // last ring row (relative hop) any pass reads
  int n_chunks;      // 32-hop chunks the R warps produce
  int pp_total;      // pass slots of the item: max(n_pass, 8 n_chunks); slots >= n_pass only release ring rows
  bool valid;
};
template <bool STATS>
__device__ __forceinline__ WsItem ws_item(const FbankParams& p, const FastParams& fp, const ClipInfo& c, int item) {
  WsItem it;
  it.b = item / fp.segs;
  it.seg = item - it.b * fp.segs;
  const int cap = STATS ? p.max_frames : p.out_frames;
  it.m_eff = (int)(c.m < cap ? c.m : cap);
  it.row_begin = it.seg * fp.seg_frames;
  it.row_end = (it.row_begin + fp.seg_frames < cap) ? it.row_begin + fp.seg_frames : cap;
  it.valid = it.row_begin < cap && !(STATS && it.row_begin >= it.m_eff);
  const int rows = it.row_end - it.row_begin;
  it.n_pass = (rows + 3) >> 2;
  int n_real = it.m_eff - it.row_begin;
  it.n_real = n_real < 0 ? 0 : (n_real > rows ? rows : n_real);
  const int n_real_pass = (it.n_real + 3) >> 2;                  // passes with at least one real frame
  it.last_hop = 4 * (n_real_pass - 1) + 5;
  it.n_chunks = n_real_pass > 0 ? it.last_hop / 32 + 1 : 0;
  it.pp_total = it.n_pass > 8 * it.n_chunks ? it.n_pass : 8 * it.n_chunks;
  return it;
}

// MIX: Mixup fused into the epilogue (b200fbank_execute_mixup).
// MULTI: the 48 kHz / 22.05 kHz resampler modes are compiled in (plans whose rate table needs them); single-rate
// 44.1 kHz plans run the lean instantiation.
//
// Persistent form (fp.ws_persist, dense batches): grid = #SMs, CTA x works on items x, x + grid, ...; the R/F pipeline
// runs straight through the item boundaries (ring slots, barrier phases and the pass -> warp assignment are numbered
// per CTA, not per clip), so the R warps resample the head of the next clip while the F warps still finish the tail of
// the current one.  Otherwise grid = #items and every CTA takes exactly one.
template <bool STATS, bool AST, bool MULTI, bool MIX = false>
__global__ void __launch_bounds__(WS_THREADS, 1) fbank_ws_kernel(const FbankParams p, const FastParams fp) {   // [phase: ws_setup]
  extern __shared__ __align__(16) float smem[];
  float* xbuf = smem;                                            // [WS_XFLOATS] input chunk (R warps only)
  float* ring = xbuf + WS_XFLOATS;                               // [WS_RING_FLOATS] 16 kHz samples, contiguous
  float* ebuf = ring + WS_RING_FLOATS;                           // [8][FK_EBUF] exchange / power buffers of the F warps
  float2* stw = reinterpret_cast<float2*>(ebuf + WS_F_WARPS * FK_EBUF);   // [512]
  float* smelw = reinterpret_cast<float*>(stw + 512);            // [mel_rows * 32]
  float* slane = smelw + ((fp.mel_rows * 32 + 3) & ~3);          // [FK_LANE_ROWS][32] static per-lane constants of the F warps
  float* spad = slane + FK_LANE_ROWS * 32;                       // [8 F warps][WS_PAD_COLS] pad-row values of the current clip by column
  uint64_t* bars = reinterpret_cast<uint64_t*>(spad + WS_F_WARPS * WS_PAD_COLS);   // full[3], empty[3], xfull

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int total = p.B * fp.segs;
  const int stride = fp.ws_persist ? (int)gridDim.x : total;     // items of this CTA: blockIdx.x, + stride, ...
  // ws_persist == 2 (ragged batches): items come from a global counter instead; R thread 0 claims them one ahead and
  // hands them to everybody through a small shared-memory queue: sq[0..7] item ids (-1 = no more), sq[8] = number
  // published, sq[16..23] = per F warp, the item step it has reached (so a queue slot is never overwritten early)
  const bool dyn = fp.ws_persist == 2;
  volatile int* sq = reinterpret_cast<volatile int*>(bars + 8);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < WS_SLOTS; ++s) {
      ws_mbar_init(bars + s, WS_R_THREADS);                      // full[s]
      ws_mbar_init(bars + WS_SLOTS + s, 32 * 8);                 // empty[s]: the 8 pass slots of a chunk x 32 lanes
    }
    ws_mbar_init(bars + 2 * WS_SLOTS, 1);                        // xfull: TMA transaction barrier of the input chunk
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 24; ++i) sq[i] = 0;
  }
  for (int i = tid; i < 512; i += WS_THREADS) stw[i] = __ldg(fp.tw + i);
  for (int i = tid; i < fp.mel_rows * 32; i += WS_THREADS) smelw[i] = __ldg(fp.melw + i);
  // static lane constants, one copy per CTA: rows 0..12 window, 13..16 first FFT bin of the slot, 17..20 output offset of
  // its mel bin, 21..24 the mel bin itself (the per-clip epilogue constants are folded from it by each F warp)
  for (int e = tid; e < 25 * 32; e += WS_THREADS) {
    const int row = e >> 5, ln = e & 31;
    float v;
    if (row < 13) {
      v = (ln + 32 * row < FK_SIZE) ? __ldg(p.window + ln + 32 * row) : 0.f;
    } else {
      const int i = (row - 13) & 3, kind = (row - 13) >> 2;      // kind 0 mstart, 1 output offset, 2 mel bin
      const bool have = i < fp.mel_groups;
      const int m = have ? __ldg(fp.mel_slot_bin + ln + 32 * i) : p.n_mel;
      if (kind == 0) v = __int_as_float(have ? __ldg(fp.mel_slot_start + ln + 32 * i) : 0);
      else if (kind == 1) v = __int_as_float(m < p.n_mel ? m * ((STATS || p.layout == 0) ? 1 : p.out_frames) : -1);
      else v = __int_as_float(m);
    }
    slane[e] = v;
  }
  __syncthreads();

  // register reallocation is per warpgroup: warpgroup 0 (the R warps) grows to 232, the F warpgroups shrink to 136
  // (R warps last, where the issue arbiter would favour them, measured 3% slower)
  const bool is_r = warp < WS_R_WARPS;
  if (is_r) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  else asm volatile("setmaxnreg.dec.sync.aligned.u32 136;");

  if (is_r) {
    // =============================== R warps: resampler ==========================================
    // mode 1: 441 -> 160 (44.1 kHz), 2: 3 -> 1 (48 kHz), 3: 441 -> 320 (22.05 kHz): register-resident taps, one TMA
    // bulk copy per chunk; mode 0: any other ratio (per-sample loop) or no resampling at all
    const int rt = tid;                                          // 0..127
    const int rw = warp;                                         // R warp index
    const int g = lane;
    unsigned x_parity = 0;
    // the taps of the current mode (pairs): mode 1 T[r][jj] = TT[18 r + jj]; mode 2 He[i] = TT[i], Ho[i] = TT[21 + i];
    // mode 3 T[r][i] = TT[10 r + i].  Reloaded only when the rate of the next clip differs.
    unsigned long long TT[FK_RP * WS_LT / 2];
#pragma unroll
    for (int i = 0; i < FK_RP * WS_LT / 2; ++i) TT[i] = 0ull;
    int cur_mode = -1, k0 = 0;
    const int cls = rw >> 1;                                     // mode 1: parity class of this warp's tap windows
    if (!MULTI) {                                                // single tuned rate: its taps are loaded once per CTA
      const float2* tp = reinterpret_cast<const float2*>(fp.ws_taps + (cls * 32 + g) * WS_GROUP_FLOATS);
#pragma unroll
      for (int i = 0; i < FK_RP * WS_LT / 2; ++i) { const float2 t2 = __ldg(tp + i); TT[i] = ws_pack(t2.x, t2.y); }
      k0 = __ldg(fp.ws_k0g + cls * 32 + g);
    }
    int gch = 0;                                                 // chunks this CTA has produced so far
#ifdef B200_WS_TIMING
    long long tc_ = 0;
#endif
    int pend = -1;                                               // thread 0: the item after the next one, claimed early
    if (dyn && rt == 0) {
      sq[0] = (int)blockIdx.x;
      __threadfence_block();
      sq[8] = 1;
      pend = ws_claim(fp, total);
    }
    for (int k = 0;; ++k) {                                      // [phase: ws_item_ctl]
      int item;
      if (dyn) {
        if (rt == 0) {                                           // publish item k + 1, start claiming item k + 2
          if (k + 1 >= WS_QN) {
            int lo;
            do {
              lo = sq[16];
#pragma unroll
              for (int w = 1; w < WS_F_WARPS; ++w) lo = min(lo, sq[16 + w]);
              if (lo < k + 2 - WS_QN) __nanosleep(64);
            } while (lo < k + 2 - WS_QN);
          }
          sq[(k + 1) & (WS_QN - 1)] = pend;
          __threadfence_block();
          sq[8] = k + 2;
          if (pend >= 0) pend = ws_claim(fp, total);
        }
        ws_bar_r();
        item = sq[k & (WS_QN - 1)];
        if (item < 0) break;
      } else {
        item = (int)blockIdx.x + k * stride;
        if (item >= total) break;
      }
      const int b = item / fp.segs;
      const ClipInfo c = clip_info(p, b);
      const WsItem it = ws_item<STATS>(p, fp, c, item);
      if (!it.valid) continue;
      if (!STATS && it.seg == 0 && tid == 0 && p.n_frames_out) p.n_frames_out[b] = it.m_eff;
      if (it.n_chunks == 0) continue;
      const int rid = p.rate_id ? p.rate_id[b] : 0;
      const int mode = fp.ws_mode[rid];
      const int row_begin = it.row_begin, last_hop = it.last_hop, n_chunks = it.n_chunks;
      if (MULTI && mode != cur_mode) {
        cur_mode = mode;
        if (mode == 1) {
          const float2* tp = reinterpret_cast<const float2*>(fp.ws_taps + (cls * 32 + g) * WS_GROUP_FLOATS);
#pragma unroll
          for (int i = 0; i < FK_RP * WS_LT / 2; ++i) { const float2 t2 = __ldg(tp + i); TT[i] = ws_pack(t2.x, t2.y); }
          k0 = __ldg(fp.ws_k0g + cls * 32 + g);
        } else if (MULTI && mode == 2) {
          const float2* tp = reinterpret_cast<const float2*>(fp.ws_t48);
#pragma unroll
          for (int i = 0; i < 2 * WS_P48; ++i) { const float2 t2 = __ldg(tp + i); TT[i] = ws_pack(t2.x, t2.y); }
        } else if (MULTI && mode == 3) {
          const int par = rw & 1;
          const float2* tp = reinterpret_cast<const float2*>(fp.ws_t22 + (size_t)(par * 32 + g) * (FK_RP * WS_LT22));
#pragma unroll
          for (int i = 0; i < FK_RP * WS_LT22 / 2; ++i) { const float2 t2 = __ldg(tp + i); TT[i] = ws_pack(t2.x, t2.y); }
          k0 = __ldg(fp.ws_k22 + par * 32 + g);
        }
      }
      // stage the input of a chunk: x[in_lo, in_lo + nx) -> xbuf[sh + i]; interior chunks take ONE TMA bulk copy   // [phase: ws_stage_x]
      auto stage_x = [&](int64_t in_lo, int nx) -> int {
        const float* gsrc = c.wav + in_lo;
        const int sh = ws_shift(gsrc);
        if (in_lo >= 4 && in_lo + nx + 4 <= c.n_in) {
          if (rt == 0) ws_tma_load(xbuf, gsrc - sh, (unsigned)(((nx + sh + 3) >> 2) << 4), bars + 2 * WS_SLOTS);
          ws_mbar_wait(bars + 2 * WS_SLOTS, x_parity);
          x_parity ^= 1u;
        } else {
          // Chunk at a clip edge (the first and the last of every clip): the 16-byte lines that lie inside the clip still
          // come by ONE bulk copy; only the few lines that straddle the edge or lie outside it (zeros: the padding of
          // torchaudio's conv, functional.py:1424) are assembled by hand.  (All by hand, this chunk cost the R warps as many
          // instructions as resampling it.)
          const int nvec = (nx + sh + 3) >> 2;
          const int lo = in_lo < 0 ? (int)(-in_lo) : 0;            // first element index that exists in the clip
          const int64_t hi64 = c.n_in - in_lo;                     // one past the last
          const int hi = hi64 > nx + 8 ? nx + 8 : (hi64 < 0 ? 0 : (int)hi64);
          int v_lo = (lo + sh + 3) >> 2, v_hi = (hi + sh) >> 2;    // vector slots [v_lo, v_hi) lie entirely inside the clip
          v_hi = v_hi > nvec ? nvec : v_hi;
          const bool bulk = v_hi >= v_lo + 64;
          if (!bulk) v_lo = v_hi = 0;
          if (bulk && rt == 0) ws_tma_load(xbuf + 4 * v_lo, gsrc - sh + 4 * v_lo, (unsigned)((v_hi - v_lo) << 4), bars + 2 * WS_SLOTS);
          const int n_hand = v_lo + (nvec - v_hi);
          for (int w = rt; w < n_hand; w += WS_R_THREADS) {
            const int v = w < v_lo ? w : v_hi + (w - v_lo);
            const int i0 = 4 * v - sh;
            float4 x;
            x.x = (i0 >= lo && i0 < hi) ? __ldg(gsrc + i0) : 0.f;
            x.y = (i0 + 1 >= lo && i0 + 1 < hi) ? __ldg(gsrc + i0 + 1) : 0.f;
            x.z = (i0 + 2 >= lo && i0 + 2 < hi) ? __ldg(gsrc + i0 + 2) : 0.f;
            x.w = (i0 + 3 >= lo && i0 + 3 < hi) ? __ldg(gsrc + i0 + 3) : 0.f;
            *reinterpret_cast<float4*>(xbuf + 4 * v) = x;
          }
          if (bulk) {
            ws_mbar_wait(bars + 2 * WS_SLOTS, x_parity);
            x_parity ^= 1u;
          }
          ws_bar_r();
        }
        return sh;
      };
      // warm the L2 with the NEXT chunk's input while this one is being resampled: the xbuf itself cannot be loaded ahead
      // (every lane visits every hop of the chunk), but a bulk L2 prefetch takes the DRAM latency out of the next TMA copy
      auto prefetch_x = [&](int64_t in_lo, int nx) {
        if (rt == 32 && in_lo >= 4 && in_lo + nx + 4 <= c.n_in) {
          const float* gsrc = c.wav + in_lo;
          const int sh = ws_shift(gsrc);
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc - sh), "r"((unsigned)(((nx + sh + 3) >> 2) << 4)) : "memory");
        }
      };
      auto put_hop = [&](float* rb, int slot, int q, const float (&y)[FK_RP]) {   // [phase: ws_put_hop]
        float* o = rb + q * FK_SHIFT + FK_RP * g;
#pragma unroll
        for (int r = 0; r < FK_RP; ++r) o[r] = y[r];
        if (slot == 0 && q < WS_MIRROR) {
          float* om = ring + (WS_RING_ROWS + q) * FK_SHIFT + FK_RP * g;
#pragma unroll
          for (int r = 0; r < FK_RP; ++r) om[r] = y[r];
        }
      };
      // [phase: ws_chunk_ctl]
      // one loop per mode (not one loop with the mode switch inside): only the state of the running mode stays live in
      // registers, which is what lets ptxas issue the input loads of a hop well ahead of the FFMA2 chain
      auto run_chunks = [&](auto&& chunk_body) {
        for (int ch = 0; ch < n_chunks; ++ch) {                  // [phase: ws_resample_loop]
          const int gc = gch + ch;                               // chunk number of this CTA: ring slot and barrier phase
          const int slot = gc % WS_SLOTS;
          float* rb = ring + slot * 32 * FK_SHIFT;
          int nh = last_hop - 32 * ch + 1;                       // hops of this chunk anyone reads
          nh = nh > 32 ? 32 : nh;
          int nh2 = last_hop - 32 * (ch + 1) + 1;                // ... and of the next one (0: none)
          nh2 = ch + 1 < n_chunks ? (nh2 > 32 ? 32 : nh2) : 0;
          const int64_t hop0 = (int64_t)row_begin + 32 * ch;     // absolute hop (= 16 kHz sample / 160) of ring row 0 of the slot
          chunk_body(gc, slot, rb, nh, nh2, hop0);
          ws_mbar_arrive(bars + slot);                           // full[slot]: release the 32 hops to the F warps
          ws_bar_r();                                            // everyone is done with xbuf before the next load
        }
      };
      if (mode == 1) {
        run_chunks([&](int gc, int slot, float* rb, int nh, int nh2, int64_t hop0) {
          // ------------ 441 -> 160: lane = 5 phases, 180 taps in registers, hop rotation (see the header) ------------
#ifdef B200_WS_TIMING
          const long long ta_ = clock64();
#endif
          // (the prefetch of the NEXT chunk goes out after this chunk's copy has landed: issued before it, the copy queued
          // behind the prefetch's DRAM fetch in the SM's bulk-copy engine and took 2.6 k cycles instead of an L2 hit's ~1 k)
          const int sh = stage_x(hop0 * FK_ORIG - FK_WIDTH, (nh - 1) * FK_ORIG + FK_KLEN + 8);
          if (nh2 > 0) prefetch_x((hop0 + 32) * FK_ORIG - FK_WIDTH, (nh2 - 1) * FK_ORIG + FK_KLEN + 8);
#ifdef B200_WS_TIMING
          const long long tb_ = clock64();
          WS_TACC(0, ta_);
#endif
          if (gc >= WS_SLOTS) ws_mbar_wait(bars + WS_SLOTS + slot, (unsigned)((gc / WS_SLOTS - 1) & 1));   // [phase: ws_resample_441]
#ifdef B200_WS_TIMING
          tc_ = clock64();
          WS_TACC(1, tb_);
#endif
          // lane g of this class owns the hops q = 2m + pg (m = 0..15) whose window starts on an even address; slot
          // s = 8 (warp & 1) + i takes m = (s + skew) mod 16 (see the header).  Hops past the chunk's last needed one are
          // computed from whatever the buffer holds and land in ring rows nobody reads.
          (void)nh;
          const int pg = (sh + k0) & 1;
          const int skew = (9 * ((g & 15) - (((sh + k0 + FK_ORIG * pg) >> 1) & 15))) & 15;
          const float* xs = xbuf + sh + k0 + pg * FK_ORIG;
          const int s0 = WS_R_ITERS * (rw & 1) + skew;
#ifdef B200_WS_SKIP_R                                                                     // timing experiment only: no resampling
          if (false)
#endif
          if (cls == 0) {
#pragma unroll 1
            for (int i = 0; i < WS_R_ITERS; ++i) {
              const int q2 = 2 * ((s0 + i) & 15);
              float y[FK_RP];
              ws_resample_hop<0>(xs + q2 * FK_ORIG, TT, y);
              put_hop(rb, slot, q2 + pg, y);
            }
          } else {
#pragma unroll 1
            for (int i = 0; i < WS_R_ITERS; ++i) {
              const int q2 = 2 * ((s0 + i) & 15);
              float y[FK_RP];
              ws_resample_hop<1>(xs + q2 * FK_ORIG, TT, y);
              put_hop(rb, slot, q2 + pg, y);
            }
          }
#ifdef B200_WS_TIMING
          WS_TACC(2, tc_); if (lane == 0) atomicAdd(&g_ws_timing[3], 1ull);
#endif
        });
      } else if (MULTI && mode == 2) {
        run_chunks([&](int gc, int slot, float* rb, int nh, int nh2, int64_t hop0) {
          // ------------ 3 -> 1 (48 kHz): ONE phase of 41 taps shared by every lane; lane = outputs 5g..5g+4 of a hop, whose
          // windows start 3 samples apart: even starts pair the taps as (h[2i], h[2i+1]) = He, odd starts as (h[2i-1], h[2i])
          // = Ho.  Lanes read x at stride 15 (odd): conflict free without any rotation.       // [phase: ws_resample_48k]
          const int sh = stage_x(hop0 * (3 * FK_SHIFT) - WS_W48, (nh * FK_SHIFT - 1) * 3 + 2 * WS_W48 + 3 + 8);
          if (nh2 > 0) prefetch_x((hop0 + 32) * (3 * FK_SHIFT) - WS_W48, (nh2 * FK_SHIFT - 1) * 3 + 2 * WS_W48 + 3 + 8);
          if (gc >= WS_SLOTS) ws_mbar_wait(bars + WS_SLOTS + slot, (unsigned)((gc / WS_SLOTS - 1) & 1));
          const float* xs = xbuf + sh + 3 * FK_RP * g;
#pragma unroll 1
          for (int q = rw; q < nh; q += WS_R_WARPS) {
            const float* xq = xs + q * (3 * FK_SHIFT);
            unsigned long long acc[FK_RP] = {0ull, 0ull, 0ull, 0ull, 0ull};
#pragma unroll
            for (int j = 0; j < WS_P48 + 6; ++j) {               // x pairs (2j, 2j+1); phase r starts at pair {0,1,3,4,6}[r]
              const unsigned long long xp = ws_pack(xq[2 * j], xq[2 * j + 1]);
              if (j < WS_P48) acc[0] = ws_fma2(xp, TT[j], acc[0]);
              if (j >= 1 && j < WS_P48 + 1) acc[1] = ws_fma2(xp, TT[WS_P48 + j - 1], acc[1]);
              if (j >= 3 && j < WS_P48 + 3) acc[2] = ws_fma2(xp, TT[j - 3], acc[2]);
              if (j >= 4 && j < WS_P48 + 4) acc[3] = ws_fma2(xp, TT[WS_P48 + j - 4], acc[3]);
              if (j >= 6) acc[4] = ws_fma2(xp, TT[j - 6], acc[4]);
            }
            float y[FK_RP];
#pragma unroll
            for (int r = 0; r < FK_RP; ++r) y[r] = __uint_as_float((unsigned)acc[r]) + __uint_as_float((unsigned)(acc[r] >> 32));
            put_hop(rb, slot, q, y);
          }
        });
      } else if (MULTI && mode == 3) {
        run_chunks([&](int gc, int slot, float* rb, int nh, int nh2, int64_t hop0) {
          // ------------ 441 -> 320 (22.05 kHz): 320 phases = two hops; even R warps hold the taps of phases 0..159 (even
          // hops), odd R warps those of 160..319; lane = 5 phases x 20 taps.  The host picks each lane's window start inside
          // its slack so that the 32 starts are distinct mod 32: conflict free without rotation.  // [phase: ws_resample_22k]
          const int par = rw & 1;
          const int np = (nh + 1) >> 1;                          // periods (pairs of hops) of this chunk; hop0 is even
          const int sh = stage_x((hop0 >> 1) * FK_ORIG - WS_W22, (np - 1) * FK_ORIG + FK_ORIG + 2 * WS_W22 + 8);
          if (nh2 > 0) prefetch_x(((hop0 >> 1) + 16) * FK_ORIG - WS_W22, (((nh2 + 1) >> 1) - 1) * FK_ORIG + FK_ORIG + 2 * WS_W22 + 8);
          if (gc >= WS_SLOTS) ws_mbar_wait(bars + WS_SLOTS + slot, (unsigned)((gc / WS_SLOTS - 1) & 1));
          const float* xs = xbuf + sh + k0;
#pragma unroll 1
          for (int Q = rw >> 1; 2 * Q + par < nh; Q += WS_R_WARPS / 2) {
            const float* xq = xs + Q * FK_ORIG;
            unsigned long long acc[FK_RP] = {0ull, 0ull, 0ull, 0ull, 0ull};
#pragma unroll
            for (int j = 0; j < WS_LT22 / 2 + 2; ++j) {          // phase r starts at pair ws_off22(r) / 2
              const unsigned long long xp = ws_pack(xq[2 * j], xq[2 * j + 1]);
#pragma unroll
              for (int r = 0; r < FK_RP; ++r) {
                const int i = j - ws_off22(r) / 2;
                if (i >= 0 && i < WS_LT22 / 2) acc[r] = ws_fma2(xp, TT[r * (WS_LT22 / 2) + i], acc[r]);
              }
            }
            float y[FK_RP];
#pragma unroll
            for (int r = 0; r < FK_RP; ++r) y[r] = __uint_as_float((unsigned)acc[r]) + __uint_as_float((unsigned)(acc[r] >> 32));
            put_hop(rb, slot, 2 * Q + par, y);
          }
        });
      } else {
        run_chunks([&](int gc, int slot, float* rb, int nh, int nh2, int64_t hop0) {
          (void)nh2;
          // ------------ other rates: per-sample polyphase loop (or plain copy) into the same ring slot ------------
          if (gc >= WS_SLOTS) ws_mbar_wait(bars + WS_SLOTS + slot, (unsigned)((gc / WS_SLOTS - 1) & 1));
          const int total_s = nh * FK_SHIFT;
          const int64_t s_base = hop0 * FK_SHIFT;                // absolute resampled index of slot row 0
          if (c.R.identity) {
            for (int t = rt; t < total_s; t += WS_R_THREADS) {
              const int64_t s2 = s_base + t;
              const float v = (s2 < c.n_in) ? __ldg(c.wav + s2) : 0.f;
              rb[t] = v;
              if (slot == 0 && t < WS_MIRROR * FK_SHIFT) ring[WS_RING_ROWS * FK_SHIFT + t] = v;
            }
          } else {
            const int part = fp.gen_part[rid];
            for (int done = 0; done < total_s; done += part) {
              const int cnt = (total_s - done < part) ? total_s - done : part;
              const int64_t s0 = s_base + done;
              const int64_t q_lo = s0 / c.R.nw, q_hi = (s0 + cnt - 1) / c.R.nw;
              const int64_t in_lo = q_lo * c.R.orig - c.R.width;
              const int nx = (int)((q_hi - q_lo) * c.R.orig + c.R.klen);
              if (done > 0) ws_bar_r();
              const int sh = ws_shift(c.wav + in_lo);
              ws_load_edge(c, in_lo, nx, sh, xbuf, rt);
              ws_bar_r();
              const float* xsg = xbuf + sh;
              for (int t = rt; t < cnt; t += WS_R_THREADS) {
                const int64_t s2 = s0 + t;
                const int64_t q = s2 / c.R.nw;
                const int ph = (int)(s2 - q * c.R.nw);
                const float* tpp = c.R.taps + (size_t)ph * c.R.L;
                const float* x = xsg + (q * c.R.orig + __ldg(c.R.k0 + ph) - c.R.width - in_lo);
                float acc = 0.f;
                for (int j = 0; j < c.R.L; ++j) acc = fmaf(__ldg(tpp + j), x[j], acc);
                const int rel = done + t;
                rb[rel] = acc;
                if (slot == 0 && rel < WS_MIRROR * FK_SHIFT) ring[WS_RING_ROWS * FK_SHIFT + rel] = acc;
              }
            }
          }
        });
      }
      gch += n_chunks;
    }
  } else {
    // =============================== F warps: frames =============================================
    const int wf = warp - WS_R_WARPS;
    FkLane L;                                                    // registers (a shared-memory copy measured 3% slower)
#pragma unroll
    for (int j = 0; j < 13; ++j) L.win[j] = slane[j * 32 + lane];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      L.mstart[i] = __float_as_int(slane[(13 + i) * 32 + lane]); L.mbin[i] = __float_as_int(slane[(17 + i) * 32 + lane]);
      L.nscale[i] = 1.f; L.nshift[i] = 0.f;
    }
    double st_s[4] = {0.0, 0.0, 0.0, 0.0}, st_ss[4] = {0.0, 0.0, 0.0, 0.0};
    long long st_n = 0;
    float* Ebuf = ebuf + wf * FK_EBUF;
    int gch = 0;                                                 // chunks / pass slots of the items this CTA has finished
    unsigned gpp = 0;
    for (int k = 0;; ++k) {
      int item;
      if (dyn) {
        if (lane == 0) {
          sq[16 + wf] = k;                                       // items before step k are finished
          while (sq[8] <= k) __nanosleep(64);                    // published one item ahead: normally no wait at all
        }
        __syncwarp();
        item = sq[k & (WS_QN - 1)];
        if (item < 0) break;
      } else {
        item = (int)blockIdx.x + k * stride;
        if (item >= total) break;
      }
      const int b = item / fp.segs;
      const ClipInfo c = clip_info(p, b);
      const WsItem it = ws_item<STATS>(p, fp, c, item);
      if (!it.valid) continue;
      int mk0 = 0, mk1 = 0, mk2 = 0, mk3 = 0;
      if (!STATS && p.masks) {
        mk0 = __ldg(p.masks + (size_t)b * 4 + 0); mk1 = __ldg(p.masks + (size_t)b * 4 + 1);
        mk2 = __ldg(p.masks + (size_t)b * 4 + 2); mk3 = __ldg(p.masks + (size_t)b * 4 + 3);
      }
      // per-clip epilogue constants: normalisation folded with ln 2, frequency mask = zero scale and shift
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = __float_as_int(slane[(21 + i) * 32 + lane]);
        fk_fold_norm(p, STATS, AST || p.use_log, m, m >= mk2 && m < mk2 + mk3, L.nscale[i], L.nshift[i]);
      }
      // Pure pad passes (four rows past the clip's last frame: 76 % of the rows of a US8K batch padded to 1024 frames) skip the
      // frame pass: the pad value of a column is the folded shift (0.0 normalised), identical for every pad row of the clip,
      // so such a pass is four 16-byte stores per lane.  (B, T, n_cols): values by column from a per-warp table;
      // (B, 1, n_cols, T): the lane's own slot values, four frames of a bin are contiguous.
      const bool pad_fast = !STATS && !MIX && (p.n_cols & 3) == 0 && p.n_cols <= WS_PAD_COLS && (p.out_frames & 3) == 0 &&
                            (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
      float* padv = spad + wf * WS_PAD_COLS;
      if (pad_fast && p.layout == 0) {
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < fp.mel_groups && L.mbin[i] >= 0) padv[L.mbin[i]] = L.nshift[i];
        __syncwarp();
      }
      const int row0 = (gch % WS_SLOTS) * 32;                    // ring row of the item's first hop
      for (int pp = (int)((wf - gpp) & (WS_F_WARPS - 1)); pp < it.pp_total; pp += WS_F_WARPS) {   // [phase: ws_frame_loop]
        bool waited = false;                                     // this slot has waited for (at least) its own chunk
        if (pp < it.n_pass) {
          const int t0 = it.row_begin + 4 * pp;
          int n_live = it.m_eff - t0;
          n_live = n_live < 0 ? 0 : (n_live > 4 ? 4 : n_live);
          waited = n_live > 0;
#ifdef B200_WS_TIMING
          const long long tf0_ = clock64();
#endif
          if (n_live > 0) {
            const int gc = gch + ((4 * pp + 5) >> 5);            // newest chunk this pass reads
            ws_mbar_wait(bars + gc % WS_SLOTS, (unsigned)((gc / WS_SLOTS) & 1));
          }
#ifdef B200_WS_TIMING
          const long long tf1_ = clock64();
          WS_TACC(4, tf0_);
#endif
          const int row = (row0 + 4 * pp) % WS_RING_ROWS;
          if (pad_fast && n_live == 0 && t0 + 4 <= it.row_end) {                       // [phase: ws_pad_rows]
            if (p.layout == 0) {
              const int c4 = 4 * lane;
              if (c4 < p.n_cols) {
                const float4 pv = *reinterpret_cast<const float4*>(padv + c4);
                float* o = p.out + ((size_t)b * p.out_frames + t0) * p.n_cols + c4;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                  const int t = t0 + r;
                  *reinterpret_cast<float4*>(o + (size_t)r * p.n_cols) = (t >= mk0 && t < mk0 + mk1) ? make_float4(0.f, 0.f, 0.f, 0.f) : pv;
                }
              }
            } else {
              float* o = p.out + (size_t)b * p.n_cols * p.out_frames + t0;
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (i < fp.mel_groups && L.mbin[i] >= 0) {
                  const float sv = L.nshift[i];
                  float4 v;
                  v.x = (t0 >= mk0 && t0 < mk0 + mk1) ? 0.f : sv;         v.y = (t0 + 1 >= mk0 && t0 + 1 < mk0 + mk1) ? 0.f : sv;
                  v.z = (t0 + 2 >= mk0 && t0 + 2 < mk0 + mk1) ? 0.f : sv; v.w = (t0 + 3 >= mk0 && t0 + 3 < mk0 + mk1) ? 0.f : sv;
                  *reinterpret_cast<float4*>(o + L.mbin[i]) = v;
                }
            }
          } else                                                                          // [phase: ws_frame_loop]
#ifdef B200_WS_SKIP_F                                                                     // timing experiment only: the F warps do nothing
          if (false)
#endif
          fk_frame_pass<STATS, AST, FK_SHIFT, FkLane, MIX>(p, fp, L, ring + row * FK_SHIFT, Ebuf, stw, smelw, b, t0, n_live, it.row_end, lane,
                                              mk0, mk1, mk2, mk3, st_s, st_ss);
#ifdef B200_WS_TIMING
          WS_TACC(5, tf1_); if (lane == 0) atomicAdd(&g_ws_timing[6], 1ull);
#endif
        }
        if (pp < 8 * it.n_chunks) {
          // A slot without live frames (pad rows, or beyond the item's passes) only releases ring rows.  It must not do so
          // before the R warps have produced its chunk: an early arrival would be counted in the PREVIOUS phase of the
          // slot's `empty` barrier (the chunk three earlier, whose passes a slower warp may still be running), that phase
          // would complete one arrival short of its own passes and the R warps would overwrite rows still being read.
          const int gc = gch + (pp >> 3);
          if (!waited) ws_mbar_wait(bars + gc % WS_SLOTS, (unsigned)((gc / WS_SLOTS) & 1));
          ws_mbar_arrive(bars + WS_SLOTS + gc % WS_SLOTS);       // empty[slot of the pass's own rows]
        }
      }
      if (STATS && wf == 0) st_n += it.n_real;
      gch += it.n_chunks;
      gpp += (unsigned)it.pp_total;
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
      if (wf == 0 && lane == 0 && st_n > 0) atomicAdd(p.sums + 2 * p.n_cols, (double)st_n);
    }
  }
}

}  // namespace b200
