// Per-clip passes around the fused frontend: one CTA per clip, the clip read twice (the second read comes from L2).
//
//   clip_normalize_kernel  AmplitudeToDB's top_db clamp (torchaudio/functional/functional.py:398-402) and the reference's
//                          per-clip normalisation (ASTPreprocessor.preprocess, src/datasets/preprocessing.py:1027-1037):
//                          global mean and UNBIASED std over the clip's own frames, (x - mean) / std * target_std +
//                          target_mean, skipped when std == 0; SpecAugment zero-fill afterwards
//                          (src/datasets/esc50.py:267-273).  In place; rows past the clip's frame count are not
//                          touched (they hold the 0.0 the main kernel wrote).
//   clip_mean_kernel       whole-clip DC removal of the waveform, `waveform - waveform.mean()` (SURVEY.md section 8a H2,
//                          the AST recipe's convention), before the resampler.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace b200 {

constexpr int CN_THREADS = 1024;

struct ClipNormParams {
  float* x;                  // (B, out_frames, n_cols) or (B, 1, n_cols, out_frames)
  const int32_t* n_frames;   // [B] real frames of each clip (<= out_frames)
  int B, out_frames, n_cols, layout;
  const float* clip_max;     // [B] per-clip maximum (dB) or nullptr
  float top_db;              // < 0: no clamp
  int normalize;
  float target_mean, target_std;
  const int32_t* masks;      // [B][4] or nullptr
};

__device__ __forceinline__ void cn_block_sum(double& s, double& ss, double (*red)[32]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, k); ss += __shfl_xor_sync(0xffffffffu, ss, k); }
  if (lane == 0) { red[0][warp] = s; red[1][warp] = ss; }
  __syncthreads();
  if (warp == 0) {
    s = lane < (int)(blockDim.x >> 5) ? red[0][lane] : 0.0;
    ss = lane < (int)(blockDim.x >> 5) ? red[1][lane] : 0.0;
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, k); ss += __shfl_xor_sync(0xffffffffu, ss, k); }
  }
}

// The clip's real cells as `rows` runs of `len` floats, `stride` floats apart: BTF = one run of m * n_cols, BFT = n_cols
// runs of m.  VEC: every run starts 16-B aligned and len % 4 == 0 is not required (the tail is scalar).
__global__ void __launch_bounds__(CN_THREADS) clip_normalize_kernel(const ClipNormParams p) {
  __shared__ double red[2][32];
  __shared__ float s_mu, s_sd;
  const int b = blockIdx.x, tid = threadIdx.x;
  int m = p.n_frames[b];
  m = m < 0 ? 0 : (m > p.out_frames ? p.out_frames : m);
  float* o = p.x + (size_t)b * p.out_frames * p.n_cols;
  const int rows = p.layout == 0 ? 1 : p.n_cols;
  const int len = p.layout == 0 ? m * p.n_cols : m;
  const int stride = p.layout == 0 ? 0 : p.out_frames;
  const bool vec = ((uintptr_t)o & 15) == 0 && (stride & 3) == 0;
  const int nv = vec ? (len >> 2) : 0;                       // float4 per run
  const int tail0 = nv << 2;                                 // scalar cells [tail0, len) of each run
  const bool clamp = p.top_db >= 0.f && p.clip_max != nullptr;
  const float floor_db = clamp ? p.clip_max[b] - p.top_db : -INFINITY;
  double s = 0.0, ss = 0.0;
  if (p.normalize || clamp) {
    for (int w = tid; w < rows * nv; w += CN_THREADS) {
      const int r = w / nv, v = w - r * nv;
      float4* q = reinterpret_cast<float4*>(o + (size_t)r * stride) + v;
      float4 x = *q;
      if (clamp) {
        x.x = fmaxf(x.x, floor_db); x.y = fmaxf(x.y, floor_db); x.z = fmaxf(x.z, floor_db); x.w = fmaxf(x.w, floor_db);
        *q = x;
      }
      s += (double)((x.x + x.y) + (x.z + x.w));
      ss += (double)(fmaf(x.x, x.x, x.y * x.y) + fmaf(x.z, x.z, x.w * x.w));
    }
    const int nt = len - tail0;
    for (int w = tid; w < rows * nt; w += CN_THREADS) {
      const int r = w / nt, e = tail0 + (w - r * nt);
      float* q = o + (size_t)r * stride + e;
      float x = *q;
      if (clamp) { x = fmaxf(x, floor_db); *q = x; }
      s += (double)x; ss += (double)x * (double)x;
    }
  }
  cn_block_sum(s, ss, red);
  if (tid == 0) {
    const double n = (double)rows * (double)len;
    const double mu = n > 0 ? s / n : 0.0;
    const double var = n > 1 ? (ss - n * mu * mu) / (n - 1.0) : 0.0;      // torch .std(): unbiased
    s_mu = (float)mu;
    s_sd = (p.normalize && var > 0.0) ? (float)sqrt(var) : 0.f;
  }
  __syncthreads();
  int mk0 = 0, mk1 = 0, mk2 = 0, mk3 = 0;
  if (p.masks) {
    mk0 = __ldg(p.masks + (size_t)b * 4 + 0); mk1 = __ldg(p.masks + (size_t)b * 4 + 1);
    mk2 = __ldg(p.masks + (size_t)b * 4 + 2); mk3 = __ldg(p.masks + (size_t)b * 4 + 3);
  }
  const bool do_norm = s_sd > 0.f, do_mask = mk1 > 0 || mk3 > 0;
  if (!do_norm && !do_mask) return;
  const float mu = s_mu, sd = s_sd, ts = p.target_std, tm = p.target_mean;
  // the reference's own three roundings: (x - mean) / std, * target_std, + target_mean
  auto cell = [&](float x, int e, int r) -> float {
    if (do_norm) x = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(x, mu), sd), ts), tm);
    if (do_mask) {
      const int t = p.layout == 0 ? e / p.n_cols : e;
      const int col = p.layout == 0 ? e - t * p.n_cols : r;
      if ((t >= mk0 && t < mk0 + mk1) || (col >= mk2 && col < mk2 + mk3)) x = 0.f;
    }
    return x;
  };
  for (int w = tid; w < rows * nv; w += CN_THREADS) {
    const int r = w / nv, v = w - r * nv;
    float4* q = reinterpret_cast<float4*>(o + (size_t)r * stride) + v;
    float4 x = *q;
    x.x = cell(x.x, 4 * v, r); x.y = cell(x.y, 4 * v + 1, r); x.z = cell(x.z, 4 * v + 2, r); x.w = cell(x.w, 4 * v + 3, r);
    __stcs(q, x);
  }
  const int nt = len - tail0;
  for (int w = tid; w < rows * nt; w += CN_THREADS) {
    const int r = w / nt, e = tail0 + (w - r * nt);
    float* q = o + (size_t)r * stride + e;
    *q = cell(*q, e, r);
  }
  // a mask may reach into the pad rows (t >= m): they are 0.0 already, nothing to do
}

// out[b] = wav[b] - mean(wav[b]); out may alias wav.  Clip b = [offsets[b], offsets[b+1]) or row b of a dense batch.
__global__ void __launch_bounds__(CN_THREADS) clip_mean_kernel(const float* wav, const int64_t* __restrict__ offsets,
                                                               int64_t clip_samples, float* out, float* __restrict__ mean_out) {
  __shared__ double red[2][32];
  __shared__ float s_mu;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int64_t o0 = offsets ? offsets[b] : (int64_t)b * clip_samples;
  const int64_t n = (offsets ? offsets[b + 1] : o0 + clip_samples) - o0;
  const float* x = wav + o0;
  float* y = out + o0;
  const int head = (int)((4 - (((uintptr_t)x >> 2) & 3)) & 3);          // scalars up to the first 16-B aligned sample
  const bool vec = (((uintptr_t)x ^ (uintptr_t)y) & 15) == 0;
  const int64_t h = vec ? (head < n ? head : n) : n;
  const int64_t nv = vec ? (n - h) >> 2 : 0;
  const int64_t t0 = h + (nv << 2);
  double s = 0.0, ss = 0.0;
  for (int64_t i = tid; i < h; i += CN_THREADS) s += (double)x[i];
  for (int64_t v = tid; v < nv; v += CN_THREADS) {
    const float4 a = reinterpret_cast<const float4*>(x + h)[v];
    s += (double)((a.x + a.y) + (a.z + a.w));
  }
  for (int64_t i = t0 + tid; i < n; i += CN_THREADS) s += (double)x[i];
  cn_block_sum(s, ss, red);
  if (tid == 0) {
    s_mu = n > 0 ? (float)(s / (double)n) : 0.f;
    if (mean_out) mean_out[b] = s_mu;
  }
  __syncthreads();
  const float mu = s_mu;
  for (int64_t i = tid; i < h; i += CN_THREADS) y[i] = x[i] - mu;
  for (int64_t v = tid; v < nv; v += CN_THREADS) {
    float4 a = reinterpret_cast<const float4*>(x + h)[v];
    a.x -= mu; a.y -= mu; a.z -= mu; a.w -= mu;
    reinterpret_cast<float4*>(y + h)[v] = a;
  }
  for (int64_t i = t0 + tid; i < n; i += CN_THREADS) y[i] = x[i] - mu;
}

// 16-bit PCM -> float32 waveforms on the device: out = float(pcm) / divisor[clip] (correctly rounded division).
// divisor = 32768 is what torchaudio.load returns for a 16-bit WAV (src/utils/audio.py:42); divisor = max |pcm| of the
// clip reproduces, bit for bit, the peak normalisation of scripts/prepare_esc50.py:94-101
// (wave / wave.abs().max() on the loaded float32 samples: both scalings by 2^-15 are exact, so the quotient is the same).
// Halves the host -> device bytes of the end-to-end path.  grid = (tiles, B).
constexpr int PCM_THREADS = 256, PCM_PER_THREAD = 8;
__global__ void __launch_bounds__(PCM_THREADS) pcm16_to_float_kernel(const int16_t* __restrict__ pcm, const int64_t* __restrict__ offsets,
                                                                    int64_t clip_samples, const float* __restrict__ divisor,
                                                                    int tiles, float* __restrict__ out) {
  const int b = (int)(blockIdx.x / (unsigned)tiles), tile = (int)(blockIdx.x - (unsigned)b * (unsigned)tiles);
  const int64_t o0 = offsets ? offsets[b] : (int64_t)b * clip_samples;
  const int64_t n = (offsets ? offsets[b + 1] : o0 + clip_samples) - o0;
  const float d = divisor ? __ldg(divisor + b) : 32768.f;
  const int16_t* x = pcm + o0;
  float* y = out + o0;
  const int64_t e0 = ((int64_t)tile * PCM_THREADS + threadIdx.x) * PCM_PER_THREAD;
  if (e0 >= n) return;
  if (e0 + PCM_PER_THREAD <= n && ((uintptr_t)(x + e0) & 15) == 0 && ((uintptr_t)(y + e0) & 15) == 0) {
    const int4 v = __ldcs(reinterpret_cast<const int4*>(x + e0));
    const int w[4] = {v.x, v.y, v.z, v.w};
    float4 a, c;
    a.x = __fdiv_rn((float)(short)(w[0] & 0xffff), d); a.y = __fdiv_rn((float)(short)(w[0] >> 16), d);
    a.z = __fdiv_rn((float)(short)(w[1] & 0xffff), d); a.w = __fdiv_rn((float)(short)(w[1] >> 16), d);
    c.x = __fdiv_rn((float)(short)(w[2] & 0xffff), d); c.y = __fdiv_rn((float)(short)(w[2] >> 16), d);
    c.z = __fdiv_rn((float)(short)(w[3] & 0xffff), d); c.w = __fdiv_rn((float)(short)(w[3] >> 16), d);
    reinterpret_cast<float4*>(y + e0)[0] = a;
    reinterpret_cast<float4*>(y + e0)[1] = c;
  } else {
    for (int i = 0; i < PCM_PER_THREAD && e0 + i < n; ++i) y[e0 + i] = __fdiv_rn((float)x[e0 + i], d);
  }
}

}  // namespace b200
