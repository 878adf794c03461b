// Per-clip passes around the fused frontend: one CTA per clip, the clip read twice (the second read comes from L2) and
// written once (the top_db clamp is applied on the fly by both passes).
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

// The clip's real cells as R rows of Lr floats, S floats apart: BTF = m rows of n_cols, BFT = n_cols rows of m.  One
// warp per row (rows strided over the CTA's warps), lanes along the row: scalar cells up to the first 16-B aligned
// address of the row, float4 body, scalar tail -- no index division anywhere, coalesced whatever the row stride.
template <class F>
__device__ __forceinline__ void cn_for_rows(float* o, int R, int Lr, int S, F&& f) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = warp; r < R; r += nw) {
    float* row = o + (size_t)r * S;
    int head = (int)((4 - (((uintptr_t)row >> 2) & 3)) & 3);
    head = head < Lr ? head : Lr;
    if (lane < head) f.scalar(row + lane, r, lane);
    const int nv = (Lr - head) >> 2;
    float4* row4 = reinterpret_cast<float4*>(row + head);
    for (int v = lane; v < nv; v += 32) f.vec(row4 + v, r, head + 4 * v);
    const int t0 = head + 4 * nv;
    if (lane < Lr - t0) f.scalar(row + t0 + lane, r, t0 + lane);
  }
}

struct CnSum {
  float floor_db; bool clamp; double s, ss;
  __device__ __forceinline__ void scalar(float* q, int, int) {
    float x = *q;
    if (clamp) x = fmaxf(x, floor_db);                    // (written by the second pass only: one store per cell)
    s += (double)x; ss += (double)x * (double)x;
  }
  __device__ __forceinline__ void vec(float4* q, int, int) {
    float4 x = *q;
    if (clamp) { x.x = fmaxf(x.x, floor_db); x.y = fmaxf(x.y, floor_db); x.z = fmaxf(x.z, floor_db); x.w = fmaxf(x.w, floor_db); }
    s += (double)((x.x + x.y) + (x.z + x.w));
    ss += (double)(fmaf(x.x, x.x, x.y * x.y) + fmaf(x.z, x.z, x.w * x.w));
  }
};

struct CnApply {
  float mu, sd, ts, tm, floor_db; bool do_norm, do_mask, clamp; int layout, mk0, mk1, mk2, mk3;
  // the reference's own three roundings: (x - mean) / std, * target_std, + target_mean; (r, e) = (row, position in the row)
  __device__ __forceinline__ float cell(float x, int r, int e) const {
    if (clamp) x = fmaxf(x, floor_db);
    if (do_norm) x = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(x, mu), sd), ts), tm);
    if (do_mask) {
      const int t = layout == 0 ? r : e, col = layout == 0 ? e : r;
      if ((t >= mk0 && t < mk0 + mk1) || (col >= mk2 && col < mk2 + mk3)) x = 0.f;
    }
    return x;
  }
  __device__ __forceinline__ void scalar(float* q, int r, int e) { *q = cell(*q, r, e); }
  __device__ __forceinline__ void vec(float4* q, int r, int e) {
    float4 x = *q;
    x.x = cell(x.x, r, e); x.y = cell(x.y, r, e + 1); x.z = cell(x.z, r, e + 2); x.w = cell(x.w, r, e + 3);
    __stcs(q, x);
  }
};

__global__ void __launch_bounds__(CN_THREADS) clip_normalize_kernel(const ClipNormParams p) {
  __shared__ double red[2][32];
  __shared__ float s_mu, s_sd;
  const int b = blockIdx.x, tid = threadIdx.x;
  int m = p.n_frames[b];
  m = m < 0 ? 0 : (m > p.out_frames ? p.out_frames : m);
  float* o = p.x + (size_t)b * p.out_frames * p.n_cols;
  const int R = p.layout == 0 ? m : p.n_cols, Lr = p.layout == 0 ? p.n_cols : m, S = p.layout == 0 ? p.n_cols : p.out_frames;
  const bool clamp = p.top_db >= 0.f && p.clip_max != nullptr;
  CnSum acc{clamp ? p.clip_max[b] - p.top_db : -INFINITY, clamp, 0.0, 0.0};
  if (p.normalize) cn_for_rows(o, R, Lr, S, acc);        // statistics of the clamped values; clamp-only runs need no first pass
  double s = acc.s, ss = acc.ss;
  cn_block_sum(s, ss, red);
  if (tid == 0) {
    const double n = (double)R * (double)Lr;
    const double mu = n > 0 ? s / n : 0.0;
    const double var = n > 1 ? (ss - n * mu * mu) / (n - 1.0) : 0.0;      // torch .std(): unbiased
    s_mu = (float)mu;
    s_sd = (p.normalize && var > 0.0) ? (float)sqrt(var) : 0.f;
  }
  __syncthreads();
  CnApply ap;
  ap.mk0 = ap.mk1 = ap.mk2 = ap.mk3 = 0;
  if (p.masks) {
    ap.mk0 = __ldg(p.masks + (size_t)b * 4 + 0); ap.mk1 = __ldg(p.masks + (size_t)b * 4 + 1);
    ap.mk2 = __ldg(p.masks + (size_t)b * 4 + 2); ap.mk3 = __ldg(p.masks + (size_t)b * 4 + 3);
  }
  ap.do_norm = s_sd > 0.f; ap.do_mask = ap.mk1 > 0 || ap.mk3 > 0;
  ap.clamp = clamp; ap.floor_db = acc.floor_db;
  if (!ap.do_norm && !ap.do_mask && !clamp) return;
  ap.mu = s_mu; ap.sd = s_sd; ap.ts = p.target_std; ap.tm = p.target_mean; ap.layout = p.layout;
  cn_for_rows(o, R, Lr, S, ap);
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
