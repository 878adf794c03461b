// libb200fbank.so -- C ABI (include/b200fbank.h) over the sm_100a fbank kernels.
//
// Host side: builds the immutable tables of a plan in float64 (polyphase taps, window,
// twiddles, sparse mel weights), uploads them once, validates arguments the way
// torchaudio asserts them, and launches kernels on the caller's stream.  No CPU compute
// path exists: device calls on a host-only plan fail with B200FBANK_ERR_NO_DEVICE.
#include "../../include/b200fbank.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include "common.cuh"
#include "fbank_generic.cuh"
#include "fbank_fast.cuh"
#include "fbank_ws.cuh"
#include "mixup.cuh"
#include "clip_norm.cuh"
#include "melspec_fast.cuh"
#include "patch_embed.cuh"
#include "patch_embed_pipe.cuh"
#include <cstdlib>

namespace {

// The opt-in limit is a per-function attribute shared by every plan in the process, so it is
// always raised to the sm_100 maximum (227 KB); occupancy depends on the launch size only.
constexpr int kMaxSmemOptin = 227 * 1024;

thread_local std::string g_err;
thread_local int64_t g_launches = 0;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(expr)                                                              \
  do {                                                                              \
    cudaError_t e_ = (expr);                                                        \
    if (e_ != cudaSuccess)                                                          \
      return fail(B200FBANK_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_));     \
  } while (0)

struct RateHost {
  int hz = 0, orig = 0, nw = 0, width = 0, klen = 0, L = 0, identity = 0;
  std::vector<float> dense;   // [nw][klen]
  std::vector<float> sparse;  // [nw][L]
  std::vector<int> k0;        // [nw]
};

// Hann-windowed sinc taps: torchaudio/functional/functional.py:1305-1402 as called by
// transforms.Resample (dtype=None): float64 grid, float32 phase term, float32 result.
void build_rate(RateHost& r, int hz, int target_hz, int lpw, double rolloff) {
  r.hz = hz;
  if (hz == target_hz) {
    r.identity = 1; r.orig = r.nw = 1; r.width = 0; r.klen = 1; r.L = 1;
    r.dense = {1.f}; r.sparse = {1.f}; r.k0 = {0};
    return;
  }
  int g = std::gcd(hz, target_hz);
  r.orig = hz / g; r.nw = target_hz / g;
  double base_freq = std::min(r.orig, r.nw) * rolloff;
  r.width = (int)std::ceil(lpw * r.orig / base_freq);
  r.klen = 2 * r.width + r.orig;
  r.dense.assign((size_t)r.nw * r.klen, 0.f);
  const double scale = base_freq / r.orig;
  for (int p = 0; p < r.nw; ++p) {
    const double phase = (double)((float)(-p) / (float)r.nw);   // int64 tensor / int -> float32
    for (int k = 0; k < r.klen; ++k) {
      double t = (phase + (double)(k - r.width) / r.orig) * base_freq;
      t = std::min(std::max(t, (double)-lpw), (double)lpw);
      double c = std::cos(t * M_PI / lpw / 2);
      double w = c * c;
      t *= M_PI;
      double s = (t == 0.0) ? 1.0 : std::sin(t) / t;
      r.dense[(size_t)p * r.klen + k] = (float)(s * (w * scale));
    }
  }
  // keep the run of taps that are not (numerically) zero; the clamp makes everything outside
  // +-lpw zero crossings cos^2(pi/2) ~ 1e-33
  std::vector<int> first(r.nw), last(r.nw);
  r.L = 1;
  for (int p = 0; p < r.nw; ++p) {
    int a = r.klen, b = -1;
    for (int k = 0; k < r.klen; ++k)
      if (std::fabs(r.dense[(size_t)p * r.klen + k]) > 1e-25f) { a = std::min(a, k); b = std::max(b, k); }
    if (b < 0) { a = 0; b = 0; }
    first[p] = a; last[p] = b;
    r.L = std::max(r.L, b - a + 1);
  }
  r.k0.resize(r.nw);
  r.sparse.assign((size_t)r.nw * r.L, 0.f);
  for (int p = 0; p < r.nw; ++p) {
    int a = std::min(first[p], r.klen - r.L);
    r.k0[p] = a;
    for (int j = 0; j < r.L; ++j) {
      float v = r.dense[(size_t)p * r.klen + a + j];
      r.sparse[(size_t)p * r.L + j] = (a + j >= first[p] && a + j <= last[p]) ? v : 0.f;
    }
  }
}

double mel_scale(double f) { return 1127.0 * std::log(1.0 + f / 700.0); }
double inv_mel_scale(double m) { return 700.0 * (std::exp(m / 1127.0) - 1.0); }

// vtln_warp_freq, torchaudio/compliance/kaldi.py:334-407 (assignment order preserved)
double vtln_warp_freq(double vl, double vh, double lo, double hi, double warp, double f) {
  const double l = vl * std::max(1.0, warp), h = vh * std::min(1.0, warp);
  const double scale = 1.0 / warp, Fl = scale * l, Fh = scale * h;
  const double sl = (Fl - lo) / (l - lo), sr = (hi - Fh) / (hi - h);
  double res = 0.0;
  if (f >= h) res = hi + sr * (f - hi);
  if (f < h) res = scale * f;
  if (f < l) res = lo + sl * (f - lo);
  if (f < lo || f > hi) res = f;
  return res;
}

}  // namespace

// Per-launch work counters of the persistent kernels (items / tiles are claimed with atomicAdd).  A ring of device ints,
// each guarded by an event recorded behind the kernel that used it: a launch takes the first slot whose last user has
// finished (whatever stream that was on), so launches in flight on any number of streams never share a counter, and
// the launch path makes no allocation.  (cudaMallocAsync per launch was tried: correct, but the stream-ordered pool
// cost up to 0.5 ms per launch once other work had synchronised the device.)
struct CounterRing {
  static constexpr int kSlots = 64;
  int* base = nullptr;
  cudaEvent_t ev[kSlots] = {};
  bool used[kSlots] = {};
  int next = 0;
  std::mutex mu;
  int init() {
    if (cudaMalloc((void**)&base, kSlots * sizeof(int)) != cudaSuccess) return -1;
    for (int i = 0; i < kSlots; ++i)
      if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) return -1;
    return 0;
  }
  // returns the slot index; the counter is zeroed on `st`
  int acquire(cudaStream_t st, int** out) {
    std::lock_guard<std::mutex> lock(mu);
    int idx = -1;
    for (int i = 0; i < kSlots && idx < 0; ++i) {
      const int c = (next + i) % kSlots;
      if (!used[c] || cudaEventQuery(ev[c]) == cudaSuccess) idx = c;
    }
    if (idx < 0) { idx = next; cudaEventSynchronize(ev[idx]); }
    next = (idx + 1) % kSlots;
    used[idx] = true;
    *out = base + idx;
    return cudaMemsetAsync(base + idx, 0, sizeof(int), st) == cudaSuccess ? idx : -1;
  }
  void release(int idx, cudaStream_t st) { if (idx >= 0) cudaEventRecord(ev[idx], st); }
  ~CounterRing() {
    for (int i = 0; i < kSlots; ++i) if (ev[i]) cudaEventDestroy(ev[i]);
    if (base) cudaFree(base);
  }
};

struct b200fbank_plan {
  b200fbank_opts o;
  int device = -1;
  int shift = 0, size = 0, padded = 0, log2n = 0, n_mel = 0, n_cols = 0, frame_mode = 0;
  std::vector<RateHost> rates;
  std::vector<float> window;          // [size]
  std::vector<float> mel_dense;       // [n_mel][padded/2]
  std::vector<int> mel_start, mel_cnt, mel_off;
  std::vector<float> mel_w;
  std::vector<float> twiddle;         // [padded/2][2]
  // device copies
  void* d_blob = nullptr;
  std::vector<void*> owned;           // further device allocations
  b200::FbankParams base{};           // table pointers + scalars filled once
  int generic_threads = 256;
  size_t generic_smem = 0;
  // fast (AST-configuration) kernel
  bool fast_ok = false;
  b200::FastParams fast;
  size_t fast_smem = 0;
  // warp-specialised kernel (fbank_ws.cuh); shares FastParams
  bool ws_ok = false;
  size_t ws_smem = 0;
  // debugging override, read ONCE when the plan is created (never on the launch path): B200FBANK_PERSIST = the
  // work-distribution mode of the warp-specialised kernel (-1 = unset)
  int env_persist = -1;
  mutable CounterRing counters;
  // tuned MelSpectrogram-dB kernel (melspec_fast.cuh)
  int melfast = 0;              // 0: not available, 1: <13, 5> (window <= 416, hop 160), 2: <32, 0> (any window / hop)
  b200::MelFastParams mfast;
  size_t mfast_smem = 0;
};

namespace {

int build_tables(b200fbank_plan* p) {
  const b200fbank_opts& o = p->o;
  // _get_waveform_and_window_properties, kaldi.py:125-151 (same float64 expression order)
  const bool melspec = o.frontend == B200FBANK_FRONTEND_MELSPEC_DB;
  p->shift = (int)(o.sample_frequency * o.frame_shift * 0.001);
  p->size = (int)(o.sample_frequency * o.frame_length * 0.001);
  if (melspec) {     // MelSpectrogram(n_fft, win_length, hop_length), torchaudio/transforms/_transforms.py:515-631
    p->shift = o.hop_length;
    p->size = o.win_length > 0 ? o.win_length : o.n_fft;
    if (o.n_fft < 2 || (o.n_fft & (o.n_fft - 1))) return fail(B200FBANK_ERR_UNSUPPORTED, "n_fft %d: only power-of-two FFT sizes are implemented", o.n_fft);
    if (p->size > o.n_fft) return fail(B200FBANK_ERR_INVALID, "win_length %d must be <= n_fft %d", p->size, o.n_fft);
  }
  if (o.sample_frequency <= 0) return fail(B200FBANK_ERR_INVALID, "`sample_frequency` must be greater than zero");
  if (p->size < 2) return fail(B200FBANK_ERR_INVALID, "choose a window size %d that is [2, len(waveform)]", p->size);
  if (p->shift <= 0) return fail(B200FBANK_ERR_INVALID, "`window_shift` must be greater than 0");
  int pow2 = 1;
  while (pow2 < p->size) pow2 <<= 1;
  p->padded = o.round_to_power_of_two ? pow2 : p->size;
  if (melspec) { p->padded = o.n_fft; pow2 = o.n_fft; }
  if (p->padded % 2 != 0)
    return fail(B200FBANK_ERR_INVALID, "the padded `window_size` must be divisible by two. use `round_to_power_of_two` or change `frame_length`");
  if (p->padded != pow2)
    return fail(B200FBANK_ERR_UNSUPPORTED, "round_to_power_of_two=False with window size %d: only power-of-two FFT sizes are implemented", p->size);
  if (p->padded > 4096) return fail(B200FBANK_ERR_UNSUPPORTED, "FFT size %d > 4096", p->padded);
  p->log2n = 0;
  while ((1 << p->log2n) < p->padded) ++p->log2n;
  if (!(o.preemphasis_coefficient >= 0.0 && o.preemphasis_coefficient <= 1.0))
    return fail(B200FBANK_ERR_INVALID, "`preemphasis_coefficient` must be between [0,1]");
  if (o.energy_floor < 0.0) return fail(B200FBANK_ERR_INVALID, "energy_floor must be >= 0");
  if (o.frontend != B200FBANK_FRONTEND_KALDI_FBANK && !melspec)
    return fail(B200FBANK_ERR_INVALID, "unknown frontend %d", o.frontend);
  p->frame_mode = melspec ? 2 : (o.snip_edges ? 0 : 1);

  // window, kaldi.py:86-113
  p->window.resize(p->size);
  const double a = 2.0 * M_PI / (p->size - 1);
  for (int i = 0; i < p->size; ++i) {
    double w;
    if (melspec) { p->window[i] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * i / p->size)); continue; }   // torch.hann_window (periodic)
    switch (o.window_type) {
      case B200FBANK_WINDOW_HANNING: w = 0.5 - 0.5 * std::cos(a * i); break;
      case B200FBANK_WINDOW_HAMMING: w = 0.54 - 0.46 * std::cos(a * i); break;
      case B200FBANK_WINDOW_POVEY:   w = std::pow((double)(float)(0.5 - 0.5 * std::cos(a * i)), 0.85); break;
      case B200FBANK_WINDOW_RECTANGULAR: w = 1.0; break;
      case B200FBANK_WINDOW_BLACKMAN:
        w = o.blackman_coeff - 0.5 * std::cos(a * i) + (0.5 - o.blackman_coeff) * std::cos(2 * a * i); break;
      default: return fail(B200FBANK_ERR_INVALID, "Invalid window type %d", o.window_type);
    }
    p->window[i] = (float)w;
  }

  // twiddles
  const int N = p->padded, NB = N / 2;
  p->twiddle.resize((size_t)NB * 2);
  for (int k = 0; k < NB; ++k) {
    p->twiddle[2 * k] = (float)std::cos(2.0 * M_PI * k / N);
    p->twiddle[2 * k + 1] = (float)(-std::sin(2.0 * M_PI * k / N));
  }

  // get_mel_banks, kaldi.py:436-511, evaluated in float64
  p->n_mel = o.num_mel_bins;
  if (p->n_mel <= 3) return fail(B200FBANK_ERR_INVALID, "Must have at least 3 mel bins");
  p->n_cols = p->n_mel + ((o.use_energy && !melspec) ? 1 : 0);
  if (melspec) {
    // melscale_fbanks(n_fft/2+1, f_min=0, f_max=sr//2, n_mels, sr, norm=None, "htk"), functional.py:518-587, in float64.
    // The Nyquist bin gets weight 0 from every filter (the last triangle ends at f_max), so n_fft/2 bins suffice.
    const int NBm = N / 2, n_freqs = NBm + 1;
    const double f_max = (double)((int)o.sample_frequency / 2);
    const double m_min = 2595.0 * std::log10(1.0 + 0.0 / 700.0), m_max = 2595.0 * std::log10(1.0 + f_max / 700.0);
    std::vector<double> f_pts(p->n_mel + 2);
    for (int i = 0; i < p->n_mel + 2; ++i) {
      const double m = m_min + (m_max - m_min) * i / (p->n_mel + 1);
      f_pts[i] = 700.0 * (std::pow(10.0, m / 2595.0) - 1.0);
    }
    p->mel_dense.assign((size_t)p->n_mel * NBm, 0.f);
    p->mel_start.resize(p->n_mel); p->mel_cnt.resize(p->n_mel); p->mel_off.resize(p->n_mel);
    p->mel_w.clear();
    for (int m = 0; m < p->n_mel; ++m) {
      int first = NBm, last = -1;
      for (int k = 0; k < NBm; ++k) {
        const double f = f_max * k / (n_freqs - 1);
        const double down = (f - f_pts[m]) / (f_pts[m + 1] - f_pts[m]), up = (f_pts[m + 2] - f) / (f_pts[m + 2] - f_pts[m + 1]);
        const float wf = (float)std::max(0.0, std::min(down, up));
        p->mel_dense[(size_t)m * NBm + k] = wf;
        if (wf != 0.f) { first = std::min(first, k); last = std::max(last, k); }
      }
      p->mel_off[m] = (int)p->mel_w.size();
      if (last < 0) { p->mel_start[m] = 0; p->mel_cnt[m] = 0; continue; }
      p->mel_start[m] = first; p->mel_cnt[m] = last - first + 1;
      for (int k = first; k <= last; ++k) p->mel_w.push_back(p->mel_dense[(size_t)m * NBm + k]);
    }
  }
  const double nyquist = 0.5 * o.sample_frequency;
  double high = o.high_freq, low = o.low_freq;
  if (melspec) { low = 0.0; high = nyquist; }
  if (high <= 0.0) high += nyquist;
  if (!((0.0 <= low && low < nyquist) && (0.0 < high && high <= nyquist) && (low < high)))
    return fail(B200FBANK_ERR_INVALID, "Bad values in options: low-freq %g and high-freq %g vs. nyquist %g", low, high, nyquist);
  const double bin_width = o.sample_frequency / N;
  const double mel_low = mel_scale(low), mel_high = mel_scale(high);
  const double delta = (mel_high - mel_low) / (p->n_mel + 1);
  double vh = o.vtln_high;
  if (vh < 0.0) vh += nyquist;
  const bool warp = o.vtln_warp != 1.0;
  if (warp && !((low < o.vtln_low && o.vtln_low < high) && (0.0 < vh && vh < high) && (o.vtln_low < vh)))
    return fail(B200FBANK_ERR_INVALID, "Bad values in options: vtln-low %g and vtln-high %g, versus low-freq %g and high-freq %g", o.vtln_low, vh, low, high);
  if (!melspec) {
    p->mel_dense.assign((size_t)p->n_mel * NB, 0.f);
    p->mel_start.resize(p->n_mel); p->mel_cnt.resize(p->n_mel); p->mel_off.resize(p->n_mel);
    p->mel_w.clear();
  }
  for (int m = 0; m < p->n_mel && !melspec; ++m) {
    double left = mel_low + m * delta, center = mel_low + (m + 1.0) * delta, right = mel_low + (m + 2.0) * delta;
    if (warp) {
      left = mel_scale(vtln_warp_freq(o.vtln_low, vh, low, high, o.vtln_warp, inv_mel_scale(left)));
      center = mel_scale(vtln_warp_freq(o.vtln_low, vh, low, high, o.vtln_warp, inv_mel_scale(center)));
      right = mel_scale(vtln_warp_freq(o.vtln_low, vh, low, high, o.vtln_warp, inv_mel_scale(right)));
    }
    int first = NB, last = -1;
    for (int k = 0; k < NB; ++k) {
      const double mel = mel_scale(bin_width * k);
      const double up = (mel - left) / (center - left), down = (right - mel) / (right - center);
      double w;
      if (!warp) w = std::max(0.0, std::min(up, down));
      else w = (mel > left && mel <= center) ? up : ((mel > center && mel < right) ? down : 0.0);
      const float wf = (float)w;
      p->mel_dense[(size_t)m * NB + k] = wf;
      if (wf != 0.f) { first = std::min(first, k); last = std::max(last, k); }
    }
    p->mel_off[m] = (int)p->mel_w.size();
    if (last < 0) { p->mel_start[m] = 0; p->mel_cnt[m] = 0; continue; }
    p->mel_start[m] = first; p->mel_cnt[m] = last - first + 1;
    for (int k = first; k <= last; ++k) p->mel_w.push_back(p->mel_dense[(size_t)m * NB + k]);
  }

  // rate table
  if (o.n_rates < 1 || o.n_rates > B200FBANK_MAX_RATES)
    return fail(B200FBANK_ERR_INVALID, "n_rates must be in [1, %d]", B200FBANK_MAX_RATES);
  if (o.lowpass_filter_width <= 0) return fail(B200FBANK_ERR_INVALID, "Low pass filter width should be positive.");
  const int target = (int)o.sample_frequency;
  p->rates.resize(o.n_rates);
  for (int i = 0; i < o.n_rates; ++i) {
    if (o.orig_rates[i] <= 0) return fail(B200FBANK_ERR_INVALID, "orig_rates[%d] must be positive", i);
    if (o.orig_rates[i] != target && (double)target != o.sample_frequency)
      return fail(B200FBANK_ERR_INVALID, "Frequencies must be of integer type to ensure quality resampling computation.");
    build_rate(p->rates[i], o.orig_rates[i], target, o.lowpass_filter_width, o.rolloff);
  }
  return 0;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Shared-memory carve-up of the generic kernel for a tile of F frames.
void generic_smem_layout(const b200fbank_plan* p, int F, int* sy, int* sx, int* sz) {
  const int ny = (F - 1) * p->shift + p->size;
  int nx = 0;
  for (const RateHost& r : p->rates) {
    if (r.identity) continue;
    int q = ny / r.nw + 2;
    nx = std::max(nx, q * r.orig + r.klen);
  }
  nx = std::max(nx, F * (p->padded / 2));
  *sy = (int)align_up(ny, 4);
  *sx = (int)align_up(nx, 4);
  *sz = F * p->padded;
}

// Tables of the fast kernel (fbank_fast.cuh); leaves p->fast_ok = false when the
// configuration is outside its envelope (the generic kernel then serves every call).
// Wavefronts of one LDS.128 by a quarter-warp whose lanes read power cells a[0..7] (16 B each, -1 = idle lane):
// cells in the same 16-B bank group (index mod 8) serialise unless they are the same cell.
static int quarter_wavefronts(const int* a, int nl = 8) {      // nl = 16: one LDS.64 by a half-warp (8-B cells, 16 bank pairs)
  int worst = 1;
  for (int r = 0; r < nl; ++r) {
    int seen[16], n = 0;
    for (int l = 0; l < nl; ++l) {
      if (a[l] < 0 || (a[l] & (nl - 1)) != r) continue;
      bool dup = false;
      for (int q = 0; q < n; ++q) dup |= seen[q] == a[l];
      if (!dup) seen[n++] = a[l];
    }
    worst = std::max(worst, n);
  }
  return worst;
}

// Lane slots of mel group `gi` (bins 32 gi .. 32 gi + 31): a deterministic annealing over (bin permutation, early
// start of short filters) that minimises the LDS.128 wavefronts of the group's tap loop; identity if nothing better.
static void plan_mel_slots(const b200fbank_plan* p, int gi, int maxcnt, int* bin, int* start, int nl = 8) {
  int slack[32], sh[32], cell[32];
  for (int l = 0; l < 32; ++l) {
    const int m = 32 * gi + l;
    bin[l] = m;
    sh[l] = 0;
    slack[l] = m < p->n_mel ? std::min(maxcnt - p->mel_cnt[m], p->mel_start[m]) : 0;
  }
  auto cost = [&]() {
    int c = 0;
    for (int q = 0; q < 32 / nl; ++q) {
      for (int l = 0; l < nl; ++l) {
        const int m = bin[nl * q + l];
        cell[l] = (m < p->n_mel && p->mel_cnt[m] > 0) ? p->mel_start[m] - sh[m - 32 * gi] : -1;
      }
      c += quarter_wavefronts(cell, nl);
    }
    return c;
  };
  uint64_t rng = 0x9E3779B97F4A7C15ull + (uint64_t)gi;
  auto next = [&]() { rng = rng * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(rng >> 33); };
  int cur = cost(), best = cur, best_bin[32], best_sh[32];
  std::copy(bin, bin + 32, best_bin); std::copy(sh, sh + 32, best_sh);
  double T = 1.0;
  for (int it = 0; it < 60000 && best > 32 / nl; ++it, T = std::max(0.05, T * 0.9999)) {
    const int a = (int)(next() & 31), b = (int)(next() & 31);
    const bool swap_move = (next() & 1) != 0;
    int old_sh = 0;
    if (swap_move) {
      if (a / nl == b / nl) continue;
      std::swap(bin[a], bin[b]);
    } else {
      if (slack[a] == 0) continue;
      old_sh = sh[a];
      sh[a] = (int)(next() % (uint32_t)(slack[a] + 1));
    }
    const int c2 = cost();
    const double u = (double)(next() & 0xFFFFFF) / (double)0x1000000;
    if (c2 <= cur || u < std::exp((double)(cur - c2) / T)) {
      cur = c2;
      if (cur < best) { best = cur; std::copy(bin, bin + 32, best_bin); std::copy(sh, sh + 32, best_sh); }
    } else if (swap_move) {
      std::swap(bin[a], bin[b]);
    } else {
      sh[a] = old_sh;
    }
  }
  for (int l = 0; l < 32; ++l) {
    bin[l] = best_bin[l];
    const int m = bin[l];
    start[l] = m < p->n_mel ? p->mel_start[m] - best_sh[m - 32 * gi] : 0;
  }
}


// Tables of the tuned MelSpectrogram-dB kernel (melspec_fast.cuh): 1024-point transforms at the clips' own rate.
int setup_melfast(b200fbank_plan* p, std::vector<void*>& owned) {
  using namespace b200;
  const b200fbank_opts& o = p->o;
  p->melfast = 0;
  const char* env = getenv("B200FBANK_KERNEL");
  if (env && strcmp(env, "generic") == 0) return 0;
  if (o.frontend != B200FBANK_FRONTEND_MELSPEC_DB || p->padded != MS_N || p->n_mel > 128 || p->n_mel < 1) return 0;
  for (const RateHost& r : p->rates)
    if (!r.identity) return 0;                                   // a clip that needs resampling takes the generic kernel
  MelFastParams& f = p->mfast;
  f.groups = (p->n_mel + 31) / 32;
  int rows = 0;
  for (int i = 0; i < 4; ++i) { f.maxcnt[i] = 0; f.woff[i] = 0; }
  for (int i = 0; i < f.groups; ++i) {
    int mc = 1;
    for (int m = 32 * i; m < std::min(32 * i + 32, p->n_mel); ++m) mc = std::max(mc, p->mel_cnt[m]);
    f.maxcnt[i] = mc; f.woff[i] = rows; rows += mc;
  }
  f.rows = rows;
  std::vector<float> tw((size_t)2 * MS_N);
  for (int k1 = 0; k1 < 32; ++k1)
    for (int l = 0; l < 32; ++l) {
      const double a = 2.0 * M_PI * (double)(k1 * l) / MS_N;
      tw[2 * (k1 * 32 + l)] = (float)std::cos(a);
      tw[2 * (k1 * 32 + l) + 1] = (float)(-std::sin(a));
    }
  std::vector<int> slot_bin(32 * f.groups), slot_start(32 * f.groups, 0);
  std::vector<float> melw;
  // ---- interval form: slot s = the bins between the peaks of filters s and s + 1 (down-slope of s, up-slope of s + 1) ----
  const int NBk = p->padded / 2;
  auto W = [&](int m, int k) -> float {
    return (m >= 0 && m < p->n_mel && k >= p->mel_start[m] && k < p->mel_start[m] + p->mel_cnt[m]) ? p->mel_w[p->mel_off[m] + k - p->mel_start[m]] : 0.f;
  };
  bool ival = !(getenv("B200FBANK_MEL_SLOTS") && strcmp(getenv("B200FBANK_MEL_SLOTS"), "filters") == 0);
  std::vector<int> s_first(p->n_mel, -1), s_last(p->n_mel, -1), s_cnt(p->n_mel, 0);
  for (int k = 0; k < NBk && ival; ++k) {
    int lo = -1, hi = -1;
    for (int m = 0; m < p->n_mel; ++m)
      if (W(m, k) != 0.f) { if (lo < 0) lo = m; hi = m; }
    if (lo < 0) continue;
    int slot;
    if (hi == lo + 1) slot = lo;
    else if (hi == lo) {
      int pk = p->mel_start[lo];
      for (int j = 0; j < p->mel_cnt[lo]; ++j)
        if (p->mel_w[p->mel_off[lo] + j] > W(lo, pk)) pk = p->mel_start[lo] + j;
      slot = k <= pk ? lo - 1 : lo;                              // up-slope of filter lo (interval lo) or its down-slope
    } else { ival = false; break; }
    if (slot < 0) { ival = false; break; }                      // the first filter has an up-slope of its own: one slot per filter
    if (s_first[slot] < 0) s_first[slot] = k;
    if (s_last[slot] >= 0 && k != s_last[slot] + 1) { ival = false; break; }
    s_last[slot] = k; ++s_cnt[slot];
  }
  if (ival) {
    rows = 0;
    for (int i = 0; i < f.groups; ++i) {
      int mc = 1;
      for (int sI = 32 * i; sI < std::min(32 * i + 32, p->n_mel); ++sI) mc = std::max(mc, s_cnt[sI]);
      mc = (mc + 1) & ~1;                                        // even: the kernel's tap loop is unrolled by two
      f.maxcnt[i] = mc; f.woff[i] = rows; rows += mc;
    }
    f.rows = rows;
    melw.assign((size_t)rows * 32 * 2, 0.f);                     // pairs (down-slope of filter s, up-slope of filter s + 1)
    for (int i = 0; i < f.groups; ++i) {
      // early starts inside each slot's slack so that the 16 lanes of a half-warp read 16 different 8-byte bank pairs
      int sh[32] = {}, slack[32], cell[16];
      for (int l = 0; l < 32; ++l) {
        const int sI = 32 * i + l;
        slack[l] = (sI < p->n_mel && s_cnt[sI] > 0) ? std::min(f.maxcnt[i] - s_cnt[sI], s_first[sI]) : 0;
      }
      auto cost = [&]() {
        int c = 0;
        for (int q = 0; q < 2; ++q) {
          for (int l = 0; l < 16; ++l) {
            const int sI = 32 * i + 16 * q + l;
            cell[l] = (sI < p->n_mel && s_cnt[sI] > 0) ? s_first[sI] - sh[16 * q + l] : -1;
          }
          c += quarter_wavefronts(cell, 16);
        }
        return c;
      };
      uint64_t rng = 0x9E3779B97F4A7C15ull + (uint64_t)i;
      auto next = [&]() { rng = rng * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(rng >> 33); };
      int cur = cost(), best = cur, best_sh[32];
      std::copy(sh, sh + 32, best_sh);
      double T = 1.0;
      for (int it = 0; it < 40000 && best > 2; ++it, T = std::max(0.05, T * 0.9998)) {
        const int a = (int)(next() & 31);
        if (slack[a] == 0) continue;
        const int old = sh[a];
        sh[a] = (int)(next() % (uint32_t)(slack[a] + 1));
        const int c2 = cost();
        const double u = (double)(next() & 0xFFFFFF) / (double)0x1000000;
        if (c2 <= cur || u < std::exp((double)(cur - c2) / T)) {
          cur = c2;
          if (cur < best) { best = cur; std::copy(sh, sh + 32, best_sh); }
        } else sh[a] = old;
      }
      for (int l = 0; l < 32; ++l) {
        const int sI = 32 * i + l;
        slot_bin[32 * i + l] = sI < p->n_mel ? sI : p->n_mel;
        if (sI >= p->n_mel || s_cnt[sI] == 0) { slot_start[32 * i + l] = 0; continue; }
        const int st = s_first[sI] - best_sh[l];
        slot_start[32 * i + l] = st;
        for (int j = 0; j < s_cnt[sI]; ++j) {                  // x 1/4: the conjugate split leaves the factor of the power out
          const int k = s_first[sI] + j;
          melw[((size_t)(f.woff[i] + best_sh[l] + j) * 32 + l) * 2] = W(sI, k) * 0.25f;
          melw[((size_t)(f.woff[i] + best_sh[l] + j) * 32 + l) * 2 + 1] = W(sI + 1, k) * 0.25f;
        }
      }
    }
  } else {
    for (int i = 0; i < f.groups; ++i) plan_mel_slots(p, i, f.maxcnt[i], slot_bin.data() + 32 * i, slot_start.data() + 32 * i, 16);
    melw.assign((size_t)rows * 32, 0.f);
    for (int i = 0; i < f.groups; ++i)
      for (int l = 0; l < 32; ++l) {
        const int m = slot_bin[32 * i + l];
        if (m >= p->n_mel) continue;
        const int lead = p->mel_start[m] - slot_start[32 * i + l];
        for (int j = 0; j < p->mel_cnt[m]; ++j)        // x 1/4: the conjugate split leaves the factor of the power out
          melw[(size_t)(f.woff[i] + lead + j) * 32 + l] = p->mel_w[p->mel_off[m] + j] * 0.25f;
      }
  }
  f.interval = ival ? 1 : 0;
  if (getenv("B200FBANK_VERBOSE"))
    fprintf(stderr, "b200fbank: melspec fast kernel, %s slots, taps per group %d %d %d %d (%d rows)\n", ival ? "interval" : "filter",
            f.maxcnt[0], f.maxcnt[1], f.maxcnt[2], f.maxcnt[3], f.rows);
  auto dev_copy = [&](const void* src, size_t bytes, const void** dst) -> int {
    void* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, std::max<size_t>(bytes, 16)));
    CUDA_TRY(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
    owned.push_back(d);
    *dst = d;
    return 0;
  };
  if (p->device >= 0) {
    if (int rc = dev_copy(tw.data(), tw.size() * 4, (const void**)&f.tw)) return rc;
    if (int rc = dev_copy(melw.data(), melw.size() * 4, (const void**)&f.melw)) return rc;
    if (int rc = dev_copy(slot_bin.data(), slot_bin.size() * 4, (const void**)&f.slot_bin)) return rc;
    if (int rc = dev_copy(slot_start.data(), slot_start.size() * 4, (const void**)&f.slot_start)) return rc;
    CUDA_TRY(cudaFuncSetAttribute(b200::melspec_fast_kernel<13, 5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
    CUDA_TRY(cudaFuncSetAttribute(b200::melspec_fast_kernel<13, 5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
    CUDA_TRY(cudaFuncSetAttribute(b200::melspec_fast_kernel<32, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
    CUDA_TRY(cudaFuncSetAttribute(b200::melspec_fast_kernel<32, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
  }
  p->mfast_smem = (size_t)(2 * MS_N + (((ival ? 2 : 1) * rows * 32 + 3) & ~3) + MS_WARPS * (MS_EBUF + MS_TBUF)) * 4;
  if (p->mfast_smem > (size_t)kMaxSmemOptin) return 0;
  if (p->device >= 0 && !p->counters.base && p->counters.init() != 0) return fail(B200FBANK_ERR_CUDA, "work-counter ring: allocation failed");
  p->melfast = (p->size <= 13 * 32 && p->shift == 5 * 32) ? 1 : 2;
  return 0;
}

int setup_fast(b200fbank_plan* p, std::vector<void*>& owned) {
  using namespace b200;
  const b200fbank_opts& o = p->o;
  p->fast_ok = false;
  const char* env = getenv("B200FBANK_KERNEL");
  if (env && strcmp(env, "generic") == 0) return 0;
  if (o.frontend != B200FBANK_FRONTEND_KALDI_FBANK) return 0;
  if (!(p->size == FK_SIZE && p->shift == FK_SHIFT && p->padded == FK_N && o.snip_edges && !o.use_energy &&
        p->n_mel <= 128))
    return 0;
  FastParams& f = p->fast;
  memset(&f, 0, sizeof f);
  f.fast_rate_id = -1;
  std::vector<float> taps((size_t)FK_NG * FK_GROUP_FLOATS, 0.f);
  std::vector<int> k0g(FK_NG, 0);
  for (size_t ri = 0; ri < p->rates.size(); ++ri) {
    const RateHost& r = p->rates[ri];
    if (r.identity || r.orig != FK_ORIG || r.nw != FK_NEW || r.klen != FK_KLEN || r.width != FK_WIDTH) continue;
    // group g = phases 5g..5g+4; phase r starts at dense index k0g + off(r) and keeps FK_LT taps
    bool ok = true;
    std::vector<int> first(r.nw), last(r.nw);
    for (int ph = 0; ph < r.nw; ++ph) {
      int a = r.klen, b = -1;
      for (int k = 0; k < r.klen; ++k)
        if (std::fabs(r.dense[(size_t)ph * r.klen + k]) > 1e-25f) { a = std::min(a, k); b = std::max(b, k); }
      first[ph] = a; last[ph] = b;
    }
    for (int g = 0; g < FK_NG && ok; ++g) {
      int k0 = r.klen;
      for (int q = 0; q < FK_RP; ++q) k0 = std::min(k0, first[FK_RP * g + q] - fk_off(q));
      k0 = std::max(k0, 0);
      k0g[g] = k0;
      for (int q = 0; q < FK_RP; ++q) {
        const int s0 = k0 + fk_off(q);
        if (s0 > first[FK_RP * g + q] || last[FK_RP * g + q] >= s0 + FK_LT) ok = false;
      }
      int e = 0;
      for (int u = 0; u < FK_WIN; ++u)
        for (int q = 0; q < FK_RP; ++q) {
          const int j = u - fk_off(q);
          if (j < 0 || j >= FK_LT) continue;
          const int k = k0 + u;
          taps[(size_t)g * FK_GROUP_FLOATS + e++] = (k < r.klen) ? r.dense[(size_t)(FK_RP * g + q) * r.klen + k] : 0.f;
        }
      if (e != FK_RP * FK_LT) ok = false;
    }
    if (ok) f.fast_rate_id = (int)ri;
  }
  // stage-1 twiddles W_512^(k1 * lane)
  std::vector<float> tw(512 * 2);
  for (int k1 = 0; k1 < 16; ++k1)
    for (int l = 0; l < 32; ++l) {
      const double a = -2.0 * M_PI * (double)(k1 * l) / 512.0;
      tw[2 * (k1 * 32 + l)] = (float)std::cos(a);
      tw[2 * (k1 * 32 + l) + 1] = (float)std::sin(a);
    }
  // mel weights, lanes = bins
  f.mel_groups = (p->n_mel + 31) / 32;
  int rows = 0;
  for (int i = 0; i < f.mel_groups; ++i) {
    int mc = 0;
    for (int m = 32 * i; m < std::min(p->n_mel, 32 * i + 32); ++m) mc = std::max(mc, p->mel_cnt[m]);
    f.mel_maxcnt[i] = mc; f.mel_woff[i] = rows; rows += mc;
  }
  f.mel_rows = rows;
  // Lane slots: inside each group of 32 bins, permute the bins over the lanes and let short filters start up to
  // (group length - own length) FFT bins early (zero leading weights) so that the 8 lanes of every quarter-warp
  // read 8 different 16-B bank groups of the 4-frame power cells (or the very same cell): LDS.128 conflict free.
  std::vector<int> slot_bin(32 * f.mel_groups), slot_start(32 * f.mel_groups, 0);
  for (int i = 0; i < f.mel_groups; ++i) plan_mel_slots(p, i, f.mel_maxcnt[i], slot_bin.data() + 32 * i, slot_start.data() + 32 * i);
  std::vector<float> melw((size_t)std::max(rows, 1) * 32, 0.f);
  for (int i = 0; i < f.mel_groups; ++i)
    for (int l = 0; l < 32; ++l) {
      const int m = slot_bin[32 * i + l];
      if (m >= p->n_mel) continue;
      const int lead = p->mel_start[m] - slot_start[32 * i + l];
      // the kernel leaves out the 1/4 (power) or 1/2 (magnitude) of the conjugate split: exact power-of-two fold
      for (int j = 0; j < p->mel_cnt[m]; ++j)
        melw[(size_t)(f.mel_woff[i] + lead + j) * 32 + l] = p->mel_w[p->mel_off[m] + j] * (o.use_power ? 0.25f : 0.5f);
    }
  for (size_t ri = 0; ri < p->rates.size(); ++ri) {
    const RateHost& r = p->rates[ri];
    int part = FK_RING_HOPS * FK_SHIFT;
    if (!r.identity) {
      // largest staging pass whose input tile fits the x region
      const int xcap = (FK_XFLOATS < WS_XFLOATS ? FK_XFLOATS : WS_XFLOATS) - 4;
      while (part > 32 && ((int64_t)(part / r.nw + 2) * r.orig + r.klen) > xcap) part -= 32;
      if (((int64_t)(part / r.nw + 2) * r.orig + r.klen) > xcap) return 0;   // ratio too extreme
    }
    f.gen_part[ri] = part;
  }
  p->fast_smem = (size_t)(FK_ATFLOATS + FK_RING_FLOATS + 1024 + rows * 32) * 4;
  // the AST-specialised variants: (2,3,6,10)-tap mel groups, power spectrum, log output
  f.ast_bank = (f.mel_groups == 4 && f.mel_maxcnt[0] == 2 && f.mel_maxcnt[1] == 3 && f.mel_maxcnt[2] == 6 && f.mel_maxcnt[3] == 10 &&
                o.use_power && o.use_log_fbank);
  if (p->fast_smem > 113 * 1024) return 0;
  auto dev_copy = [&](const void* src, size_t bytes, const void** dst) -> int {
    void* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, std::max<size_t>(bytes, 16)));
    CUDA_TRY(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
    owned.push_back(d);
    *dst = d;
    return 0;
  };
  if (int rc = dev_copy(taps.data(), taps.size() * 4, (const void**)&f.taps)) return rc;
  if (int rc = dev_copy(k0g.data(), k0g.size() * 4, (const void**)&f.k0g)) return rc;
  if (int rc = dev_copy(tw.data(), tw.size() * 4, (const void**)&f.tw)) return rc;
  if (int rc = dev_copy(melw.data(), melw.size() * 4, (const void**)&f.melw)) return rc;
  if (int rc = dev_copy(slot_bin.data(), slot_bin.size() * 4, (const void**)&f.mel_slot_bin)) return rc;
  if (int rc = dev_copy(slot_start.data(), slot_start.size() * 4, (const void**)&f.mel_slot_start)) return rc;
  // ---- warp-specialised kernel: register-resident taps [32][5][36], even phase offsets ----------
  {
    // two parity classes of tap windows (fbank_ws.cuh): class c of group g starts phase r at dense index
    // wk[32 c + g] + ws_offc(c, r) and keeps WS_LT taps; the two starts of a group have opposite parity, so for every
    // hop one of them sits on an even (8-byte aligned) sample address
    std::vector<float> wt((size_t)2 * FK_NG * WS_GROUP_FLOATS, 0.f);
    std::vector<int> wk(2 * FK_NG, 0);
    bool ok = f.fast_rate_id >= 0;
    if (ok) {
      const RateHost& r = p->rates[f.fast_rate_id];
      for (int g = 0; g < FK_NG && ok; ++g) {
        int first[FK_RP], last[FK_RP];
        for (int q = 0; q < FK_RP; ++q) {
          int a = r.klen, bb = -1;
          for (int k = 0; k < r.klen; ++k)
            if (std::fabs(r.dense[(size_t)(FK_RP * g + q) * r.klen + k]) > 1e-25f) { a = std::min(a, k); bb = std::max(bb, k); }
          first[q] = a; last[q] = bb;
        }
        for (int c = 0; c < 2 && ok; ++c) {
          // every phase's non-zero taps must fall inside its window: k0 + off(r) in [last - (WS_LT - 1), first]
          int lo = 0, hi = r.klen;
          for (int q = 0; q < FK_RP; ++q) {
            lo = std::max(lo, last[q] - (WS_LT - 1) - ws_offc(c, q));
            hi = std::min(hi, first[q] - ws_offc(c, q));
          }
          hi = std::min(hi, r.klen + 8 - (ws_offc(c, FK_RP - 1) + WS_LT));      // stays inside the staged input tile
          int k0 = hi;
          if (c == 1 && ((k0 ^ wk[g]) & 1) == 0) --k0;                          // opposite parity to class 0
          if (k0 < lo) { ok = false; break; }
          wk[32 * c + g] = k0;
          for (int q = 0; q < FK_RP; ++q) {
            const int s0 = k0 + ws_offc(c, q);
            for (int j = 0; j < WS_LT; ++j)
              wt[((size_t)(32 * c + g) * FK_RP + q) * WS_LT + j] =
                  (s0 + j >= 0 && s0 + j < r.klen) ? r.dense[(size_t)(FK_RP * g + q) * r.klen + s0 + j] : 0.f;
          }
        }
      }
    }
    const char* kenv = getenv("B200FBANK_KERNEL");
    if (kenv && strcmp(kenv, "fast") == 0) ok = false;
    if (int rc = dev_copy(wt.data(), wt.size() * 4, (const void**)&f.ws_taps)) return rc;
    if (int rc = dev_copy(wk.data(), wk.size() * 4, (const void**)&f.ws_k0g)) return rc;
    // ---- resampler mode per rate: 1 = 441 -> 160 (above), 2 = 3 -> 1, 3 = 441 -> 320, 0 = per-sample loop / identity
    for (int i = 0; i < B200_MAX_RATES; ++i) f.ws_mode[i] = 0;
    if (ok) f.ws_mode[f.fast_rate_id] = 1;
    std::vector<float> t48(4 * WS_P48, 0.f), t22((size_t)2 * 32 * FK_RP * WS_LT22, 0.f);
    std::vector<int> k22(64, 0);
    const bool tuned_other = !(getenv("B200FBANK_WS_GENERIC_RATES") && atoi(getenv("B200FBANK_WS_GENERIC_RATES")));
    for (size_t ri = 0; ri < p->rates.size() && tuned_other; ++ri) {
      const RateHost& r = p->rates[ri];
      if (r.identity) continue;
      if (r.orig == 3 && r.nw == 1 && r.width == WS_W48 && r.klen == 2 * WS_W48 + 3) {
        // 3 -> 1: tap pairs for even window starts (h[2i], h[2i+1]) and odd ones (h[2i-1], h[2i])
        auto h = [&](int k) { return (k >= 0 && k < r.klen) ? r.dense[k] : 0.f; };
        for (int i = 0; i < WS_P48; ++i) {
          t48[2 * i] = h(2 * i); t48[2 * i + 1] = h(2 * i + 1);
          t48[2 * (WS_P48 + i)] = h(2 * i - 1); t48[2 * (WS_P48 + i) + 1] = h(2 * i);
        }
        f.ws_mode[ri] = 2;
      } else if (r.orig == FK_ORIG && r.nw == 2 * FK_NEW && r.width == WS_W22) {
        // 441 -> 320: per hop parity and lane, 5 phases x WS_LT22 taps from a window start chosen inside the lane's slack
        // so that the 32 starts of a warp are distinct mod 32 (bipartite matching lanes -> banks)
        bool fit = true;
        for (int par = 0; par < 2 && fit; ++par) {
          int lo[32], hi[32], owner[32], pick[32];
          for (int g = 0; g < 32 && fit; ++g) {
            lo[g] = 0; hi[g] = r.klen;
            for (int q = 0; q < FK_RP; ++q) {
              const float* d = &r.dense[(size_t)(160 * par + FK_RP * g + q) * r.klen];
              int a = r.klen, bb = -1;
              for (int k = 0; k < r.klen; ++k)
                if (std::fabs(d[k]) > 1e-25f) { a = std::min(a, k); bb = std::max(bb, k); }
              hi[g] = std::min(hi[g], a - ws_off22(q));
              lo[g] = std::max(lo[g], bb - ws_off22(q) - WS_LT22 + 1);
            }
            hi[g] = std::min(hi[g], r.klen + 8 - (ws_off22(FK_RP - 1) + WS_LT22));   // stays inside the staged input tile
            if (lo[g] > hi[g]) fit = false;
          }
          if (!fit) break;
          for (int b = 0; b < 32; ++b) owner[b] = -1;
          // Kuhn's augmenting paths; visited[] per search
          std::vector<char> visited(32);
          std::function<bool(int)> place = [&](int g) -> bool {
            for (int k = hi[g]; k >= lo[g]; --k) {
              const int b = k & 31;
              if (visited[b]) continue;
              visited[b] = 1;
              if (owner[b] < 0 || place(owner[b])) { owner[b] = g; pick[g] = k; return true; }
            }
            return false;
          };
          for (int g = 0; g < 32; ++g) {
            std::fill(visited.begin(), visited.end(), 0);
            if (!place(g)) pick[g] = hi[g];          // no perfect matching: keep a legal start, accept the conflict
          }
          for (int g = 0; g < 32; ++g) {
            k22[par * 32 + g] = pick[g];
            for (int q = 0; q < FK_RP; ++q) {
              const float* d = &r.dense[(size_t)(160 * par + FK_RP * g + q) * r.klen];
              for (int j = 0; j < WS_LT22; ++j) {
                const int k = pick[g] + ws_off22(q) + j;
                t22[((size_t)(par * 32 + g) * FK_RP + q) * WS_LT22 + j] = (k >= 0 && k < r.klen) ? d[k] : 0.f;
              }
            }
          }
        }
        if (fit) f.ws_mode[ri] = 3;
      }
    }
    if (int rc = dev_copy(t48.data(), t48.size() * 4, (const void**)&f.ws_t48)) return rc;
    if (int rc = dev_copy(t22.data(), t22.size() * 4, (const void**)&f.ws_t22)) return rc;
    if (int rc = dev_copy(k22.data(), k22.size() * 4, (const void**)&f.ws_k22)) return rc;
    p->ws_smem = (size_t)(WS_XFLOATS + WS_RING_FLOATS + WS_F_WARPS * FK_EBUF + 1024 + ((rows * 32 + 3) & ~3) + FK_LANE_ROWS * 32 + WS_F_WARPS * WS_PAD_COLS) * 4 + 256;
    // rates without the 44.1 kHz structure still run through this kernel's per-sample path
    p->ws_ok = (ok || f.fast_rate_id < 0) && !(kenv && strcmp(kenv, "fast") == 0) && p->ws_smem <= 227 * 1024;
    f.ws_ok = p->ws_ok;
    if (const char* e = getenv("B200FBANK_PERSIST")) p->env_persist = atoi(e);
    if (!p->counters.base && p->counters.init() != 0) return fail(B200FBANK_ERR_CUDA, "work-counter ring: allocation failed");
    f.ws_multi = 0;
    for (int i = 0; i < B200_MAX_RATES; ++i) f.ws_multi |= (f.ws_mode[i] >= 2);
#define B200_WS_ATTR(S, A, M) CUDA_TRY(cudaFuncSetAttribute(b200::fbank_ws_kernel<S, A, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin))
    B200_WS_ATTR(false, false, false); B200_WS_ATTR(false, true, false); B200_WS_ATTR(true, false, false); B200_WS_ATTR(true, true, false);
    B200_WS_ATTR(false, false, true); B200_WS_ATTR(false, true, true); B200_WS_ATTR(true, false, true); B200_WS_ATTR(true, true, true);
#undef B200_WS_ATTR
#define B200_WS_ATTR_MIX(A, M) CUDA_TRY(cudaFuncSetAttribute(b200::fbank_ws_kernel<false, A, M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin))
    B200_WS_ATTR_MIX(false, false); B200_WS_ATTR_MIX(true, false); B200_WS_ATTR_MIX(false, true); B200_WS_ATTR_MIX(true, true);
#undef B200_WS_ATTR_MIX
  }
  const char* seg = getenv("B200FBANK_SEG");
  f.seg_frames = seg ? std::max(32, atoi(seg) / 32 * 32) : 0;      // 0 = pick per launch (pick_seg_frames)
  CUDA_TRY(cudaFuncSetAttribute(b200::fbank_fast_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
  CUDA_TRY(cudaFuncSetAttribute(b200::fbank_fast_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
  CUDA_TRY(cudaFuncSetAttribute(b200::fbank_fast_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
  CUDA_TRY(cudaFuncSetAttribute(b200::fbank_fast_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
  p->fast_ok = true;
  return 0;
}

int upload(b200fbank_plan* p) {
  CUDA_TRY(cudaSetDevice(p->device));
  std::vector<void*>& blob_host_extra = p->owned;
  // one blob: [window][twiddle][mel_w][mel_start][mel_cnt][mel_off]{[taps][k0]}*
  std::vector<char> blob;
  auto put = [&](const void* src, size_t bytes) {
    size_t off = align_up(blob.size(), 16);
    blob.resize(off + bytes);
    if (bytes) memcpy(blob.data() + off, src, bytes);
    return off;
  };
  size_t o_win = put(p->window.data(), p->window.size() * 4);
  size_t o_tw = put(p->twiddle.data(), p->twiddle.size() * 4);
  size_t o_mw = put(p->mel_w.data(), p->mel_w.size() * 4);
  size_t o_ms = put(p->mel_start.data(), p->mel_start.size() * 4);
  size_t o_mc = put(p->mel_cnt.data(), p->mel_cnt.size() * 4);
  size_t o_mo = put(p->mel_off.data(), p->mel_off.size() * 4);
  std::vector<size_t> o_taps, o_k0;
  for (const RateHost& r : p->rates) {
    o_taps.push_back(put(r.sparse.data(), r.sparse.size() * 4));
    o_k0.push_back(put(r.k0.data(), r.k0.size() * 4));
  }
  CUDA_TRY(cudaMalloc(&p->d_blob, blob.size()));
  CUDA_TRY(cudaMemcpy(p->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  char* d = (char*)p->d_blob;
  b200::FbankParams& k = p->base;
  memset(&k, 0, sizeof k);
  const b200fbank_opts& o = p->o;
  for (size_t i = 0; i < p->rates.size(); ++i) {
    const RateHost& r = p->rates[i];
    k.rates[i] = b200::RateDev{r.orig, r.nw, r.width, r.klen, r.L, r.identity,
                               (const float*)(d + o_taps[i]), (const int*)(d + o_k0[i])};
  }
  k.shift = p->shift; k.size = p->size; k.padded = p->padded; k.log2n = p->log2n;
  const bool melspec = o.frontend == B200FBANK_FRONTEND_MELSPEC_DB;
  k.frame_mode = p->frame_mode; k.db_mode = melspec ? 1 : 0;
  k.snip_edges = p->frame_mode == 0; k.remove_dc = melspec ? 0 : o.remove_dc_offset; k.raw_energy = o.raw_energy;
  k.use_energy = melspec ? 0 : o.use_energy; k.htk_compat = o.htk_compat; k.use_power = melspec ? 1 : o.use_power; k.use_log = o.use_log_fbank;
  k.preemph = melspec ? 0.f : (float)o.preemphasis_coefficient;
  k.has_energy_floor = o.energy_floor != 0.0;
  k.log_energy_floor = k.has_energy_floor ? (float)std::log(o.energy_floor) : 0.f;
  k.window = (const float*)(d + o_win);
  k.twiddle = (const float2*)(d + o_tw);
  k.n_mel = p->n_mel; k.n_cols = p->n_cols;
  k.mel_start = (const int*)(d + o_ms); k.mel_cnt = (const int*)(d + o_mc); k.mel_off = (const int*)(d + o_mo);
  k.mel_w = (const float*)(d + o_mw);

  if (int rc = setup_fast(p, blob_host_extra)) return rc;
  if (int rc = setup_melfast(p, blob_host_extra)) return rc;

  // pick the largest tile that fits comfortably (two CTAs per SM when possible)
  int F = 16;
  for (; F >= 2; F >>= 1) {
    int sy, sx, sz;
    generic_smem_layout(p, F, &sy, &sx, &sz);
    size_t bytes = (size_t)(sy + sx + sz + F) * 4;
    if (bytes <= 110 * 1024 || (F == 2 && bytes <= 227 * 1024)) {
      k.tile_frames = F; k.smem_y = sy; k.smem_x = sx; k.smem_z = sz;
      p->generic_smem = bytes;
      break;
    }
  }
  if (F < 2) return fail(B200FBANK_ERR_UNSUPPORTED, "configuration does not fit in shared memory");
  CUDA_TRY(cudaFuncSetAttribute(b200::fbank_generic_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
  CUDA_TRY(cudaFuncSetAttribute(b200::fbank_generic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
  return 0;
}

// Frames per CTA of the fast kernel: long segments amortise the per-segment set-up (table staging,
// two-hop prologue), short ones keep every SM busy for small batches.  2 CTAs/SM x 148 SMs = 296 slots.
int pick_seg_frames(const b200::FastParams& f, int B, int frames, int slots = 296) {
  if (f.seg_frames > 0) return f.seg_frames;
  for (int seg = 512; seg > 32; seg >>= 1)
    if ((int64_t)B * ((frames + seg - 1) / seg) >= 2 * slots) return seg;
  return 32;
}

// Persistent launch of the warp-specialised kernel: dense batches (equal clips) larger than one wave run as one CTA
// per SM that walks its items with the R/F pipeline carried across clip boundaries; ragged batches keep one CTA per
// item so the hardware scheduler balances the uneven clips.  B200FBANK_PERSIST=0/1 overrides.
int ws_pick_grid(const b200fbank_plan* p, const int64_t* d_offsets, b200::FastParams& f, int64_t& grid, cudaStream_t st) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms = n;
  }
  int mode = d_offsets == nullptr ? 1 : 2;           // dense: static stride; ragged: dynamic claims
  if (p->env_persist >= 0) mode = p->env_persist;
  f.ws_persist = (mode != 0 && grid > sms) ? mode : 0;
  f.ws_counter = nullptr;
  if (f.ws_persist) grid = sms;
  f.ws_counter_slot = -1;
  if (f.ws_persist == 2) {
    // one work counter PER LAUNCH in flight (CounterRing): released behind the kernel by ws_release_counter
    f.ws_counter_slot = p->counters.acquire(st, &f.ws_counter);
    if (f.ws_counter_slot < 0) return fail(B200FBANK_ERR_CUDA, "work counter: cudaMemsetAsync failed");
  }
  return 0;
}

int ws_release_counter(const b200fbank_plan* p, b200::FastParams& f, cudaStream_t st) {
  p->counters.release(f.ws_counter_slot, st);
  f.ws_counter_slot = -1;
  return 0;
}

int check_device_call(const b200fbank_plan* p, const void* wav, const int64_t* offsets, int64_t clip_samples, int B) {
  if (!p) return fail(B200FBANK_ERR_INVALID, "plan is NULL");
  if (p->device < 0) return fail(B200FBANK_ERR_NO_DEVICE, "host-only plan (device=-1): no CPU compute path exists");
  if (B < 0) return fail(B200FBANK_ERR_INVALID, "B must be >= 0");
  if (B > 0 && !wav) return fail(B200FBANK_ERR_INVALID, "d_wav is NULL");
  if (!offsets && clip_samples <= 0 && B > 0) return fail(B200FBANK_ERR_INVALID, "clip_samples must be > 0 when d_offsets is NULL");
  return 0;
}

}  // namespace

extern "C" {

int b200fbank_abi_version(void) { return B200FBANK_ABI_VERSION; }
int b200fbank_sizeof_opts(void) { return (int)sizeof(b200fbank_opts); }

void b200fbank_default_opts(b200fbank_opts* o) {
  memset(o, 0, sizeof *o);
  o->blackman_coeff = 0.42; o->energy_floor = 1.0; o->frame_length = 25.0; o->frame_shift = 10.0;
  o->high_freq = 0.0; o->low_freq = 20.0; o->preemphasis_coefficient = 0.97; o->sample_frequency = 16000.0;
  o->vtln_high = -500.0; o->vtln_low = 100.0; o->vtln_warp = 1.0;
  o->num_mel_bins = 23; o->window_type = B200FBANK_WINDOW_POVEY;
  o->htk_compat = 0; o->raw_energy = 1; o->remove_dc_offset = 1; o->round_to_power_of_two = 1;
  o->snip_edges = 1; o->subtract_mean = 0; o->use_energy = 0; o->use_log_fbank = 1; o->use_power = 1;
  o->n_rates = 1; o->orig_rates[0] = 16000; o->lowpass_filter_width = 6; o->rolloff = 0.99;
  o->frontend = B200FBANK_FRONTEND_KALDI_FBANK; o->n_fft = 1024; o->hop_length = 160; o->win_length = 400;
  o->top_db = 80.0;
}

int b200fbank_plan_create(const b200fbank_opts* o, int device, b200fbank_plan** out) {
  if (!o || !out) return fail(B200FBANK_ERR_INVALID, "NULL argument");
  *out = nullptr;
  b200fbank_plan* p = new b200fbank_plan();
  p->o = *o;
  p->device = device;
  int rc = build_tables(p);
  if (rc == 0 && device >= 0) rc = upload(p);
  if (rc != 0) { b200fbank_plan_destroy(p); return rc; }
  *out = p;
  return 0;
}

void b200fbank_plan_destroy(b200fbank_plan* p) {
  if (!p) return;
  if (p->d_blob) cudaFree(p->d_blob);
  for (void* d : p->owned) cudaFree(d);
  delete p;
}

const char* b200fbank_last_error(void) { return g_err.c_str(); }

int64_t b200fbank_resampled_length(const b200fbank_plan* p, int64_t n, int rate_id) {
  if (!p || rate_id < 0 || rate_id >= (int)p->rates.size()) return fail(B200FBANK_ERR_INVALID, "bad rate_id %d", rate_id);
  const RateHost& r = p->rates[rate_id];
  return r.identity ? n : b200::resampled_length(n, r.orig, r.nw);
}

int64_t b200fbank_num_frames(const b200fbank_plan* p, int64_t n, int rate_id) {
  int64_t n_rs = b200fbank_resampled_length(p, n, rate_id);
  if (n_rs < 0) return n_rs;
  return b200::num_frames(n_rs, p->size, p->shift, p->frame_mode, p->padded);
}

int b200fbank_num_cols(const b200fbank_plan* p) { return p ? p->n_cols : fail(B200FBANK_ERR_INVALID, "plan is NULL"); }

int64_t b200fbank_plan_table(const b200fbank_plan* p, int table, int arg, float* dst, int64_t cap) {
  if (!p) return fail(B200FBANK_ERR_INVALID, "plan is NULL");
  const std::vector<float>* v = nullptr;
  switch (table) {
    case B200FBANK_TABLE_WINDOW: v = &p->window; break;
    case B200FBANK_TABLE_MEL_DENSE: v = &p->mel_dense; break;
    case B200FBANK_TABLE_TAPS_DENSE:
      if (arg < 0 || arg >= (int)p->rates.size()) return fail(B200FBANK_ERR_INVALID, "bad rate_id %d", arg);
      v = &p->rates[arg].dense; break;
    default: return fail(B200FBANK_ERR_INVALID, "bad table %d", table);
  }
  int64_t n = (int64_t)v->size();
  if (dst) memcpy(dst, v->data(), (size_t)std::min(n, cap) * 4);
  return n;
}

int b200fbank_plan_info(const b200fbank_plan* p, int arg, int64_t info[8]) {
  if (!p || !info) return fail(B200FBANK_ERR_INVALID, "NULL argument");
  if (arg < 0 || arg >= (int)p->rates.size()) return fail(B200FBANK_ERR_INVALID, "bad rate_id %d", arg);
  const RateHost& r = p->rates[arg];
  info[0] = p->shift; info[1] = p->size; info[2] = p->padded; info[3] = (int64_t)p->rates.size();
  info[4] = r.orig; info[5] = r.nw; info[6] = r.width; info[7] = r.L;
  return 0;
}

static int execute_impl(const b200fbank_plan* p, const float* d_wav, const int64_t* d_offsets,
                        int64_t clip_samples, const int32_t* d_rate_id, int B,
                        const int32_t* d_masks, const float* d_mean, const float* d_std, int n_stats,
                        float target_mean, float target_std, int out_frames, int layout,
                        const float* d_bank, const int32_t* d_partner, const float* d_lam,
                        float* d_out, int32_t* d_n_frames, void* stream) {
  int rc = check_device_call(p, d_wav, d_offsets, clip_samples, B);
  if (rc) return rc;
  if (out_frames <= 0) return fail(B200FBANK_ERR_INVALID, "out_frames must be > 0");
  if (layout != B200FBANK_LAYOUT_BTF && layout != B200FBANK_LAYOUT_BFT) return fail(B200FBANK_ERR_INVALID, "bad layout %d", layout);
  if (p->o.frontend != B200FBANK_FRONTEND_KALDI_FBANK) return fail(B200FBANK_ERR_INVALID, "plan was created for the MELSPEC_DB frontend: call b200fbank_melspec_db");
  if (n_stats != 0 && n_stats != 1 && n_stats != p->n_cols)
    return fail(B200FBANK_ERR_INVALID, "n_stats must be 0, 1 or n_cols=%d (got %d)", p->n_cols, n_stats);
  if (n_stats != 0 && (!d_mean || !d_std)) return fail(B200FBANK_ERR_INVALID, "d_mean/d_std are NULL");
  if (!d_out) return fail(B200FBANK_ERR_INVALID, "d_out is NULL");
  if (B == 0) return 0;
  CUDA_TRY(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  b200::FbankParams k = p->base;
  k.wav = d_wav; k.offsets = d_offsets; k.clip_samples = clip_samples; k.rate_id = d_rate_id; k.B = B;
  k.out_frames = out_frames; k.layout = layout; k.out = d_out; k.n_frames_out = d_n_frames;
  k.masks = d_masks; k.mean = d_mean; k.std = d_std; k.n_stats = n_stats;
  k.target_mean = target_mean; k.target_std = target_std;
  const bool cms = p->o.subtract_mean != 0;
  if (cms) { k.masks = nullptr; k.n_stats = 0; }     // raw features first, cms_kernel finishes
  // Mixup rides in the epilogue of the tuned kernels; the generic kernel (and CMS, which finishes in a second kernel)
  // mix with one more launch of the stand-alone mixup kernel, in place
  const bool mix = d_bank != nullptr;
  const bool mix_fused = mix && p->fast_ok && p->ws_ok && !cms;
  if (mix_fused) { k.mix_bank = d_bank; k.mix_partner = d_partner; k.mix_lam = d_lam; }
  if (p->fast_ok && p->ws_ok) {
    b200::FastParams f = p->fast;
    f.seg_frames = pick_seg_frames(f, B, out_frames, 148);
    f.segs = (out_frames + f.seg_frames - 1) / f.seg_frames;
    int64_t grid = (int64_t)B * f.segs;
    if (grid > 0x7fffffffLL) return fail(B200FBANK_ERR_INVALID, "B * segments = %lld exceeds the grid limit", (long long)grid);
    if (int rc = ws_pick_grid(p, d_offsets, f, grid, st)) return rc;
    if (mix_fused) {
      if (f.ws_multi) {
        if (f.ast_bank) b200::fbank_ws_kernel<false, true, true, true><<<(unsigned)grid, b200::WS_THREADS, p->ws_smem, st>>>(k, f);
        else b200::fbank_ws_kernel<false, false, true, true><<<(unsigned)grid, b200::WS_THREADS, p->ws_smem, st>>>(k, f);
      } else {
        if (f.ast_bank) b200::fbank_ws_kernel<false, true, false, true><<<(unsigned)grid, b200::WS_THREADS, p->ws_smem, st>>>(k, f);
        else b200::fbank_ws_kernel<false, false, false, true><<<(unsigned)grid, b200::WS_THREADS, p->ws_smem, st>>>(k, f);
      }
    } else if (f.ws_multi) {
      if (f.ast_bank) b200::fbank_ws_kernel<false, true, true><<<(unsigned)grid, b200::WS_THREADS, p->ws_smem, st>>>(k, f);
      else b200::fbank_ws_kernel<false, false, true><<<(unsigned)grid, b200::WS_THREADS, p->ws_smem, st>>>(k, f);
    } else {
      if (f.ast_bank) b200::fbank_ws_kernel<false, true, false><<<(unsigned)grid, b200::WS_THREADS, p->ws_smem, st>>>(k, f);
      else b200::fbank_ws_kernel<false, false, false><<<(unsigned)grid, b200::WS_THREADS, p->ws_smem, st>>>(k, f);
    }
    if (int rc = ws_release_counter(p, f, st)) return rc;
  } else if (p->fast_ok) {
    b200::FastParams f = p->fast;
    f.seg_frames = pick_seg_frames(f, B, out_frames);
    f.segs = (out_frames + f.seg_frames - 1) / f.seg_frames;
    const int64_t grid = (int64_t)B * f.segs;
    if (grid > 0x7fffffffLL) return fail(B200FBANK_ERR_INVALID, "B * segments = %lld exceeds the grid limit", (long long)grid);
    if (f.ast_bank) b200::fbank_fast_kernel<false, true><<<(unsigned)grid, b200::FK_THREADS, p->fast_smem, st>>>(k, f);
    else b200::fbank_fast_kernel<false, false><<<(unsigned)grid, b200::FK_THREADS, p->fast_smem, st>>>(k, f);
  } else {
    k.tiles = (out_frames + k.tile_frames - 1) / k.tile_frames;
    const int64_t grid = (int64_t)B * k.tiles;
    if (grid > 0x7fffffffLL) return fail(B200FBANK_ERR_INVALID, "B * tiles = %lld exceeds the grid limit", (long long)grid);
    b200::fbank_generic_kernel<false><<<(unsigned)grid, p->generic_threads, p->generic_smem, st>>>(k);
  }
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  if (cms) {
    k.masks = d_masks; k.n_stats = n_stats;
    b200::cms_kernel<<<B, 128, 0, st>>>(k);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
  }
  if (mix && !mix_fused)
    return b200fbank_mixup(d_out, d_bank, d_partner, d_lam, B, (int64_t)out_frames * p->n_cols, d_out, stream);
  return 0;
}

int b200fbank_execute(const b200fbank_plan* p, const float* d_wav, const int64_t* d_offsets,
                      int64_t clip_samples, const int32_t* d_rate_id, int B,
                      const int32_t* d_masks, const float* d_mean, const float* d_std, int n_stats,
                      float target_mean, float target_std, int out_frames, int layout,
                      float* d_out, int32_t* d_n_frames, void* stream) {
  return execute_impl(p, d_wav, d_offsets, clip_samples, d_rate_id, B, d_masks, d_mean, d_std, n_stats, target_mean,
                      target_std, out_frames, layout, nullptr, nullptr, nullptr, d_out, d_n_frames, stream);
}

int b200fbank_execute_mixup(const b200fbank_plan* p, const float* d_wav, const int64_t* d_offsets,
                            int64_t clip_samples, const int32_t* d_rate_id, int B,
                            const int32_t* d_masks, const float* d_mean, const float* d_std, int n_stats,
                            float target_mean, float target_std, int out_frames, int layout,
                            const float* d_bank, const int32_t* d_partner, const float* d_lam,
                            float* d_out, int32_t* d_n_frames, void* stream) {
  if (!d_bank || !d_partner || !d_lam) return fail(B200FBANK_ERR_INVALID, "d_bank / d_partner / d_lam are NULL");
  if (d_bank == d_out) return fail(B200FBANK_ERR_INVALID, "the bank must not alias the output");
  return execute_impl(p, d_wav, d_offsets, clip_samples, d_rate_id, B, d_masks, d_mean, d_std, n_stats, target_mean,
                      target_std, out_frames, layout, d_bank, d_partner, d_lam, d_out, d_n_frames, stream);
}

int b200fbank_melspec_db(const b200fbank_plan* p, const float* d_wav, const int64_t* d_offsets, int64_t clip_samples,
                         const int32_t* d_rate_id, int B, const int32_t* d_masks, int to_db, int normalize,
                         float target_mean, float target_std, int out_frames, int layout, float* d_out,
                         int32_t* d_n_frames, float* d_clip_max, void* stream) {
  int rc = check_device_call(p, d_wav, d_offsets, clip_samples, B);
  if (rc) return rc;
  if (p->o.frontend != B200FBANK_FRONTEND_MELSPEC_DB) return fail(B200FBANK_ERR_INVALID, "plan was not created for the MELSPEC_DB frontend");
  if (out_frames <= 0) return fail(B200FBANK_ERR_INVALID, "out_frames must be > 0");
  if (layout != B200FBANK_LAYOUT_BTF && layout != B200FBANK_LAYOUT_BFT) return fail(B200FBANK_ERR_INVALID, "bad layout %d", layout);
  if (!d_out) return fail(B200FBANK_ERR_INVALID, "d_out is NULL");
  if (to_db && !d_clip_max) return fail(B200FBANK_ERR_INVALID, "d_clip_max ([B] floats of workspace) is NULL");
  if ((to_db || d_masks) && !d_n_frames) return fail(B200FBANK_ERR_INVALID, "d_n_frames is required with to_db or masks (the per-clip pass reads it)");
  if (!to_db && normalize) return fail(B200FBANK_ERR_INVALID, "normalize requires to_db (the reference normalises dB values)");
  if (B == 0) return 0;
  CUDA_TRY(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  b200::FbankParams k = p->base;
  k.wav = d_wav; k.offsets = d_offsets; k.clip_samples = clip_samples; k.rate_id = d_rate_id; k.B = B;
  k.out_frames = out_frames; k.layout = layout; k.out = d_out; k.n_frames_out = d_n_frames;
  k.masks = nullptr; k.n_stats = 0; k.target_mean = target_mean; k.target_std = target_std;
  k.db_mode = to_db ? 1 : 0; k.use_log = 0;
  k.clip_max = to_db ? d_clip_max : nullptr;
  if (to_db) {      // -inf: the ordered-int atomicMax identity
    static const unsigned kNegInf = 0xff800000u;
    CUDA_TRY(cudaMemsetAsync(d_clip_max, 0, sizeof(float) * B, st));
    b200::fill_u32_kernel<<<(B + 255) / 256, 256, 0, st>>>(reinterpret_cast<unsigned*>(d_clip_max), kNegInf, B);
    ++g_launches;
  }
  if (p->melfast) {
    // tuned kernel: independent warps claim 8-frame tiles from a per-launch counter (stream-ordered allocation)
    b200::MelFastParams f = p->mfast;
    f.tiles_per_clip = (out_frames + b200::MS_TILE - 1) / b200::MS_TILE;
    if ((int64_t)B * f.tiles_per_clip > 0x7fffffffLL) return fail(B200FBANK_ERR_INVALID, "B * tiles exceeds 2^31 - 1");
    const int cslot = p->counters.acquire(st, &f.counter);
    if (cslot < 0) return fail(B200FBANK_ERR_CUDA, "work counter: cudaMemsetAsync failed");
    static int sms = 0;
    if (sms == 0) {
      int n = 0;
      if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, p->device) != cudaSuccess || n <= 0) n = 148;
      sms = n;
    }
    const int64_t warps_needed = (int64_t)B * f.tiles_per_clip;
    const int grid = (int)std::min<int64_t>(sms, (warps_needed + b200::MS_WARPS - 1) / b200::MS_WARPS);
    if (p->melfast == 1) {
      if (f.interval) b200::melspec_fast_kernel<13, 5, true><<<grid, b200::MS_THREADS, p->mfast_smem, st>>>(k, f);
      else b200::melspec_fast_kernel<13, 5, false><<<grid, b200::MS_THREADS, p->mfast_smem, st>>>(k, f);
    } else {
      if (f.interval) b200::melspec_fast_kernel<32, 0, true><<<grid, b200::MS_THREADS, p->mfast_smem, st>>>(k, f);
      else b200::melspec_fast_kernel<32, 0, false><<<grid, b200::MS_THREADS, p->mfast_smem, st>>>(k, f);
    }
    ++g_launches;
    p->counters.release(cslot, st);
    CUDA_TRY(cudaGetLastError());
  } else {
    k.tiles = (out_frames + k.tile_frames - 1) / k.tile_frames;
    const int64_t grid = (int64_t)B * k.tiles;
    if (grid > 0x7fffffffLL) return fail(B200FBANK_ERR_INVALID, "B * tiles exceeds the grid limit");
    b200::fbank_generic_kernel<false><<<(unsigned)grid, p->generic_threads, p->generic_smem, st>>>(k);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
  }
  if (to_db || d_masks) {
    b200::ClipNormParams q;
    q.x = d_out; q.n_frames = d_n_frames; q.B = B; q.out_frames = out_frames; q.n_cols = k.n_cols; q.layout = layout;
    q.clip_max = to_db ? d_clip_max : nullptr; q.top_db = to_db ? (float)p->o.top_db : -1.f; q.normalize = normalize;
    q.target_mean = target_mean; q.target_std = target_std; q.masks = d_masks;
    b200::clip_normalize_kernel<<<B, b200::CN_THREADS, 0, st>>>(q);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
  }
  return 0;
}

int b200fbank_clip_normalize(float* d_x, const int32_t* d_n_frames, int B, int out_frames, int n_cols, int layout,
                             const float* d_clip_max, float top_db, int normalize, float target_mean, float target_std,
                             const int32_t* d_masks, void* stream) {
  if (B < 0 || out_frames <= 0 || n_cols <= 0) return fail(B200FBANK_ERR_INVALID, "bad shape");
  if (layout != B200FBANK_LAYOUT_BTF && layout != B200FBANK_LAYOUT_BFT) return fail(B200FBANK_ERR_INVALID, "bad layout %d", layout);
  if (B == 0) return 0;
  if (!d_x || !d_n_frames) return fail(B200FBANK_ERR_INVALID, "d_x and d_n_frames must not be NULL");
  b200::ClipNormParams q;
  q.x = d_x; q.n_frames = d_n_frames; q.B = B; q.out_frames = out_frames; q.n_cols = n_cols; q.layout = layout;
  q.clip_max = d_clip_max; q.top_db = d_clip_max ? top_db : -1.f; q.normalize = normalize;
  q.target_mean = target_mean; q.target_std = target_std; q.masks = d_masks;
  b200::clip_normalize_kernel<<<B, b200::CN_THREADS, 0, (cudaStream_t)stream>>>(q);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int b200fbank_pcm16_to_float(const int16_t* d_pcm, const int64_t* d_offsets, int64_t clip_samples, int B,
                             const float* d_divisor, int64_t max_clip_samples, float* d_out, void* stream) {
  if (B < 0) return fail(B200FBANK_ERR_INVALID, "B must be >= 0");
  if (B == 0) return 0;
  if (!d_pcm || !d_out) return fail(B200FBANK_ERR_INVALID, "NULL device pointer");
  const int64_t longest = d_offsets ? max_clip_samples : clip_samples;
  if (longest <= 0) return fail(B200FBANK_ERR_INVALID, "clip_samples (dense) or max_clip_samples (ragged) must be > 0");
  const int64_t per_cta = (int64_t)b200::PCM_THREADS * b200::PCM_PER_THREAD;
  const int64_t tiles = (longest + per_cta - 1) / per_cta;
  if (tiles * B > 0x7fffffffLL) return fail(B200FBANK_ERR_INVALID, "B * tiles exceeds the grid limit");
  b200::pcm16_to_float_kernel<<<(unsigned)(tiles * B), b200::PCM_THREADS, 0, (cudaStream_t)stream>>>(
      d_pcm, d_offsets, clip_samples, d_divisor, (int)tiles, d_out);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int b200fbank_remove_clip_mean(const float* d_wav, const int64_t* d_offsets, int64_t clip_samples, int B, float* d_out,
                               float* d_mean, void* stream) {
  if (B < 0) return fail(B200FBANK_ERR_INVALID, "B must be >= 0");
  if (B == 0) return 0;
  if (!d_wav || !d_out) return fail(B200FBANK_ERR_INVALID, "NULL device pointer");
  if (!d_offsets && clip_samples <= 0) return fail(B200FBANK_ERR_INVALID, "clip_samples must be > 0 when d_offsets is NULL");
  b200::clip_mean_kernel<<<B, b200::CN_THREADS, 0, (cudaStream_t)stream>>>(d_wav, d_offsets, clip_samples, d_out, d_mean);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int b200fbank_stats_accumulate(const b200fbank_plan* p, const float* d_wav, const int64_t* d_offsets,
                               int64_t clip_samples, const int32_t* d_rate_id, int B, int max_frames,
                               double* d_sums, void* stream) {
  int rc = check_device_call(p, d_wav, d_offsets, clip_samples, B);
  if (rc) return rc;
  if (max_frames <= 0) return fail(B200FBANK_ERR_INVALID, "max_frames must be > 0");
  if (!d_sums) return fail(B200FBANK_ERR_INVALID, "d_sums is NULL");
  if (p->o.subtract_mean) return fail(B200FBANK_ERR_UNSUPPORTED, "stats with subtract_mean=True are identically zero-mean; not implemented");
  if (B == 0) return 0;
  CUDA_TRY(cudaSetDevice(p->device));
  b200::FbankParams k = p->base;
  k.wav = d_wav; k.offsets = d_offsets; k.clip_samples = clip_samples; k.rate_id = d_rate_id; k.B = B;
  k.max_frames = max_frames; k.sums = d_sums;
  if (p->fast_ok && p->ws_ok) {
    b200::FastParams f = p->fast;
    f.seg_frames = pick_seg_frames(f, B, max_frames, 148);
    f.segs = (max_frames + f.seg_frames - 1) / f.seg_frames;
    int64_t grid = (int64_t)B * f.segs;
    if (grid > 0x7fffffffLL) return fail(B200FBANK_ERR_INVALID, "B * segments exceeds the grid limit");
    if (int rc = ws_pick_grid(p, d_offsets, f, grid, (cudaStream_t)stream)) return rc;
    if (f.ws_multi) {
      if (f.ast_bank) b200::fbank_ws_kernel<true, true, true><<<(unsigned)grid, b200::WS_THREADS, p->ws_smem, (cudaStream_t)stream>>>(k, f);
      else b200::fbank_ws_kernel<true, false, true><<<(unsigned)grid, b200::WS_THREADS, p->ws_smem, (cudaStream_t)stream>>>(k, f);
    } else {
      if (f.ast_bank) b200::fbank_ws_kernel<true, true, false><<<(unsigned)grid, b200::WS_THREADS, p->ws_smem, (cudaStream_t)stream>>>(k, f);
      else b200::fbank_ws_kernel<true, false, false><<<(unsigned)grid, b200::WS_THREADS, p->ws_smem, (cudaStream_t)stream>>>(k, f);
    }
    if (int rc = ws_release_counter(p, f, (cudaStream_t)stream)) return rc;
  } else if (p->fast_ok) {
    b200::FastParams f = p->fast;
    f.seg_frames = pick_seg_frames(f, B, max_frames);
    f.segs = (max_frames + f.seg_frames - 1) / f.seg_frames;
    const int64_t grid = (int64_t)B * f.segs;
    if (grid > 0x7fffffffLL) return fail(B200FBANK_ERR_INVALID, "B * segments exceeds the grid limit");
    if (f.ast_bank) b200::fbank_fast_kernel<true, true><<<(unsigned)grid, b200::FK_THREADS, p->fast_smem, (cudaStream_t)stream>>>(k, f);
    else b200::fbank_fast_kernel<true, false><<<(unsigned)grid, b200::FK_THREADS, p->fast_smem, (cudaStream_t)stream>>>(k, f);
  } else {
    k.tiles = (max_frames + k.tile_frames - 1) / k.tile_frames;
    const int64_t grid = (int64_t)B * k.tiles;
    if (grid > 0x7fffffffLL) return fail(B200FBANK_ERR_INVALID, "B * tiles exceeds the grid limit");
    b200::fbank_generic_kernel<true><<<(unsigned)grid, p->generic_threads, p->generic_smem, (cudaStream_t)stream>>>(k);
  }
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int b200fbank_resample(const b200fbank_plan* p, const float* d_wav, const int64_t* d_offsets,
                       int64_t clip_samples, const int32_t* d_rate_id, int B, float* d_out,
                       const int64_t* d_out_offsets, int64_t out_clip_samples, void* stream) {
  int rc = check_device_call(p, d_wav, d_offsets, clip_samples, B);
  if (rc) return rc;
  if (!d_out) return fail(B200FBANK_ERR_INVALID, "d_out is NULL");
  if (out_clip_samples <= 0) return fail(B200FBANK_ERR_INVALID, "out_clip_samples (longest resampled clip) must be > 0");
  if (B == 0) return 0;
  CUDA_TRY(cudaSetDevice(p->device));
  const int chunk = 4096;
  int nx = 0;
  for (const RateHost& r : p->rates)
    if (!r.identity) nx = std::max(nx, (chunk / r.nw + 2) * r.orig + r.klen);
  const size_t smem = (size_t)(chunk + align_up(std::max(nx, 4), 4)) * 4;
  if (smem > 227 * 1024) return fail(B200FBANK_ERR_UNSUPPORTED, "resample ratio needs %zu B of shared memory", smem);
  CUDA_TRY(cudaFuncSetAttribute(b200::resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
  b200::FbankParams k = p->base;
  k.wav = d_wav; k.offsets = d_offsets; k.clip_samples = clip_samples; k.rate_id = d_rate_id; k.B = B;
  k.tiles = (int)((out_clip_samples + chunk - 1) / chunk);
  const int64_t grid = (int64_t)B * k.tiles;
  if (grid > 0x7fffffffLL) return fail(B200FBANK_ERR_INVALID, "B * tiles exceeds the grid limit");
  b200::resample_kernel<<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(k, d_out, d_out_offsets, out_clip_samples, chunk);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

#ifdef B200_WS_TIMING
// debug builds only: read and clear the pipeline timing counters of fbank_ws.cuh
extern "C" int b200fbank_debug_ws_timing(unsigned long long out[8]) {
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(out, b200::g_ws_timing, sizeof z) != cudaSuccess) return -1;
  cudaMemcpyToSymbol(b200::g_ws_timing, z, sizeof z);
  return 0;
}
#endif

int b200fbank_mixup(const float* d_x, const float* d_bank, const int32_t* d_partner, const float* d_lam, int B,
                    int64_t clip_elems, float* d_out, void* stream) {
  if (B < 0 || clip_elems <= 0) return fail(B200FBANK_ERR_INVALID, "B must be >= 0 and clip_elems > 0");
  if (B == 0) return 0;
  if (!d_x || !d_bank || !d_partner || !d_lam || !d_out) return fail(B200FBANK_ERR_INVALID, "NULL device pointer");
  const int64_t per_cta = (int64_t)b200::MIX_THREADS * b200::MIX_VEC_PER_THREAD * 4;
  const int64_t tiles = (clip_elems + per_cta - 1) / per_cta;
  if (tiles * B > 0x7fffffffLL) return fail(B200FBANK_ERR_INVALID, "B * tiles exceeds the grid limit (2^31 - 1 CTAs)");
  b200::mixup_kernel<<<(unsigned)(tiles * B), b200::MIX_THREADS, 0, (cudaStream_t)stream>>>(
      d_x, d_bank, d_partner, d_lam, clip_elems, (int)tiles, d_out);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int b200fbank_mixup_labels(const int64_t* d_label, const int64_t* d_partner_label, const int32_t* d_partner,
                           const float* d_lam, int B, int num_classes, float* d_soft, void* stream) {
  if (B < 0 || num_classes <= 0) return fail(B200FBANK_ERR_INVALID, "B must be >= 0 and num_classes > 0");
  if (B == 0) return 0;
  if (!d_label || !d_partner_label || !d_partner || !d_lam || !d_soft) return fail(B200FBANK_ERR_INVALID, "NULL device pointer");
  b200::mixup_labels_kernel<<<(unsigned)B, 64, 0, (cudaStream_t)stream>>>(d_label, d_partner_label, d_partner, d_lam, B,
                                                                          num_classes, d_soft);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int b200fbank_patch_embed(const float* d_feat, int B, int F, int T, const void* d_weight_f16, const float* d_bias, int D,
                          int patch, int stride, void* d_out, int out_f16, void* stream) {
  if (B < 0 || F <= 0 || T <= 0 || D <= 0 || stride <= 0) return fail(B200FBANK_ERR_INVALID, "bad shape");
  if (patch != 16) return fail(B200FBANK_ERR_UNSUPPORTED, "patch size %d: only the 16 x 16 patches of AST are built", patch);
  if (D % b200::PE_N) return fail(B200FBANK_ERR_UNSUPPORTED, "embedding dim %d is not a multiple of %d", D, b200::PE_N);
  if (F < patch || T < patch) return fail(B200FBANK_ERR_INVALID, "spectrogram %d x %d is smaller than one patch", F, T);
  if (B == 0) return 0;
  if (!d_feat || !d_weight_f16 || !d_out) return fail(B200FBANK_ERR_INVALID, "NULL device pointer");
  b200::PatchEmbedParams k;
  k.feat = d_feat; k.w = (const __half*)d_weight_f16; k.bias = d_bias; k.out = d_out;
  k.B = B; k.F = F; k.T = T; k.D = D; k.stride = stride; k.out_f16 = out_f16;
  k.Fp = (F - patch) / stride + 1; k.Tp = (T - patch) / stride + 1;
  k.M = (int64_t)B * k.Fp * k.Tp;
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int NT = D / b200::PE_N;
  // The TMA + tcgen05 pipeline (patch_embed_pipe.cuh) needs a tensor map over the features: rows of T floats with a
  // 16-byte aligned pitch and base.  Anything else (and B200FBANK_PE=gather) runs the gather kernel below.
  static const bool force_gather = [] { const char* e = getenv("B200FBANK_PE"); return e && !strcmp(e, "gather"); }();
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static const EncodeFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess) fn = nullptr;
    return (EncodeFn)fn;
  }();
  const int seg_cap = std::min(b200::PP_SEG, (2 * b200::PP_BOXW - 3 - patch) / stride + 1);
  if (!force_gather && encode && seg_cap >= 1 && T % 4 == 0 && ((uintptr_t)d_feat & 15) == 0 && ((uintptr_t)d_out & 15) == 0 &&
      (int64_t)B * F < (1ll << 31)) {
    const int nseg = (k.Tp + seg_cap - 1) / seg_cap, segp = (k.Tp + nseg - 1) / nseg;
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)T, (cuuint64_t)B * (cuuint64_t)F};
    const cuuint64_t strides[1] = {(cuuint64_t)T * 4};
    const cuuint32_t box[2] = {(cuuint32_t)b200::PP_BOXW, 4}, estr[2] = {1, 1};
    const CUresult er = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(d_feat), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (er != CUDA_SUCCESS) return fail(B200FBANK_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)er);
    const int64_t n_tiles = ((int64_t)B * k.Fp * nseg + 1) / 2;
    const int per_col = (int)std::min<int64_t>(n_tiles, std::max(1, sms / NT));
    static std::atomic<unsigned long long> pipe_attr_done{0};
    if (dev >= 64 || !((pipe_attr_done.load() >> dev) & 1ull)) {
      CUDA_TRY(cudaFuncSetAttribute(b200::patch_embed_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b200::PP_SMEM));
      CUDA_TRY(cudaFuncSetAttribute(b200::patch_embed_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b200::PP_SMEM));
      if (dev < 64) pipe_attr_done.fetch_or(1ull << dev);
    }
    const unsigned pgrid = (unsigned)(per_col * NT);
    if (out_f16) b200::patch_embed_pipe_kernel<true><<<pgrid, b200::PP_THREADS, b200::PP_SMEM, (cudaStream_t)stream>>>(k, nseg, segp, tm);
    else b200::patch_embed_pipe_kernel<false><<<pgrid, b200::PP_THREADS, b200::PP_SMEM, (cudaStream_t)stream>>>(k, nseg, segp, tm);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return 0;
  }
  const int64_t m_tiles = (k.M + b200::PE_M - 1) / b200::PE_M;
  int per_col = (int)std::min<int64_t>(m_tiles, std::max(1, sms / NT));
  const unsigned grid = (unsigned)(per_col * NT);
  static std::atomic<unsigned long long> attr_done{0};          // bit d: the opt-in shared-memory size is set on device d
  if (dev >= 64 || !((attr_done.load() >> dev) & 1ull)) {
    CUDA_TRY(cudaFuncSetAttribute(b200::patch_embed_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b200::PE_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(b200::patch_embed_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b200::PE_SMEM));
    if (dev < 64) attr_done.fetch_or(1ull << dev);
  }
  if (out_f16) b200::patch_embed_kernel<true><<<grid, b200::PE_THREADS, b200::PE_SMEM, (cudaStream_t)stream>>>(k);
  else b200::patch_embed_kernel<false><<<grid, b200::PE_THREADS, b200::PE_SMEM, (cudaStream_t)stream>>>(k);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

#ifdef B200_PE_TIMING
extern "C" int b200fbank_debug_pe_timing(unsigned long long out[4]) {
  unsigned long long z[4] = {0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(out, b200::g_pe_timing, sizeof z) != cudaSuccess) return -1;
  cudaMemcpyToSymbol(b200::g_pe_timing, z, sizeof z);
  return 0;
}
#endif

#ifdef B200_PP_TIMING
extern "C" int b200fbank_debug_pp_timing(unsigned long long out[16]) {
  unsigned long long z[16] = {0};
  if (cudaMemcpyFromSymbol(out, b200::g_pp_timing, sizeof z) != cudaSuccess) return -1;
  cudaMemcpyToSymbol(b200::g_pp_timing, z, sizeof z);
  return 0;
}
#endif

int64_t b200fbank_launch_count(int reset) {
  int64_t n = g_launches;
  if (reset) g_launches = 0;
  return n;
}

}  // extern "C"
