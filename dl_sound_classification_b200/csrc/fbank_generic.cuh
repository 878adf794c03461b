// Generic fused fbank kernel: any power-of-two FFT size, any rate table, every
// kaldi.fbank option.  One CTA = one tile of `tile_frames` frames of one clip; all
// intermediates (input tile, resampled tile, FFT workspace, power spectrum) live in shared
// memory, so HBM sees the waveform once and each output once.
//
// This is the correctness-first path; fbank_fast.cuh holds the tuned kernel for the AST
// configuration (512-point FFT).  Both share the staging code below.
#pragma once
#include "common.cuh"

namespace b200 {

struct ClipInfo {
  const float* wav;   // clip base
  int64_t n_in;       // input samples
  int64_t n_rs;       // samples after resampling
  int64_t m;          // frames (uncropped)
  RateDev R;
};

__device__ inline ClipInfo clip_info(const FbankParams& p, int b) {
  ClipInfo c;
  int64_t o0 = p.offsets ? p.offsets[b] : (int64_t)b * p.clip_samples;
  int64_t o1 = p.offsets ? p.offsets[b + 1] : o0 + p.clip_samples;
  c.wav = p.wav + o0;
  c.n_in = o1 - o0;
  int r = p.rate_id ? p.rate_id[b] : 0;
  c.R = p.rates[r];
  c.n_rs = c.R.identity ? c.n_in : resampled_length(c.n_in, c.R.orig, c.R.nw);
  c.m = num_frames(c.n_rs, p.size, p.shift, p.frame_mode, p.padded);
  return c;
}

// Stage resampled samples [s_lo, s_hi) of the clip into ybuf[0 .. s_hi-s_lo).
// H1: torchaudio/functional/functional.py:1405-1432 -- y[q*new + p] = sum_k taps[p][k] *
// xpad[q*orig + k], xpad = zero-pad(width, width+orig); only the L non-zero taps of each
// phase are stored.  xin is scratch for the input tile.
__device__ inline void stage_resampled(const ClipInfo& c, int64_t s_lo, int64_t s_hi, float* xin,
                                       float* ybuf) {
  const RateDev& R = c.R;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int ny = (int)(s_hi - s_lo);
  if (ny <= 0) return;
  if (R.identity) {
    for (int i = tid; i < ny; i += nt) {
      int64_t s = s_lo + i;
      ybuf[i] = (s >= 0 && s < c.n_in) ? __ldg(c.wav + s) : 0.f;
    }
    __syncthreads();
    return;
  }
  const int64_t q_lo = s_lo / R.nw, q_hi = (s_hi - 1) / R.nw;
  const int64_t in_lo = q_lo * R.orig - R.width;             // real index of xin[0]
  const int nx = (int)((q_hi - q_lo) * R.orig + R.klen);
  // coalesced load of the input tile; float4 body when the global address is 16-B aligned
  {
    const float* g = c.wav + in_lo;                           // may point before the clip
    int head = (int)((4 - (((uintptr_t)g >> 2) & 3)) & 3);    // scalars until g is 16-B aligned
    if (head > nx) head = nx;
    for (int i = tid; i < head; i += nt) {
      int64_t s = in_lo + i;
      xin[i] = (s >= 0 && s < c.n_in) ? __ldg(g + i) : 0.f;
    }
    const int nvec = (nx - head) >> 2;
    for (int v = tid; v < nvec; v += nt) {
      int i = head + 4 * v;
      int64_t s = in_lo + i;
      float4 x;
      if (s >= 0 && s + 3 < c.n_in) {
        x = __ldg(reinterpret_cast<const float4*>(g + i));
      } else {
        x.x = (s >= 0 && s < c.n_in) ? __ldg(g + i) : 0.f;
        x.y = (s + 1 >= 0 && s + 1 < c.n_in) ? __ldg(g + i + 1) : 0.f;
        x.z = (s + 2 >= 0 && s + 2 < c.n_in) ? __ldg(g + i + 2) : 0.f;
        x.w = (s + 3 >= 0 && s + 3 < c.n_in) ? __ldg(g + i + 3) : 0.f;
      }
      xin[i] = x.x; xin[i + 1] = x.y; xin[i + 2] = x.z; xin[i + 3] = x.w;
    }
    for (int i = head + 4 * nvec + tid; i < nx; i += nt) {
      int64_t s = in_lo + i;
      xin[i] = (s >= 0 && s < c.n_in) ? __ldg(g + i) : 0.f;
    }
  }
  __syncthreads();
  for (int i = tid; i < ny; i += nt) {
    int64_t s = s_lo + i;
    int q = (int)(s / R.nw - q_lo);
    int ph = (int)(s % R.nw);
    const float* t = R.taps + (size_t)ph * R.L;
    const float* x = xin + q * R.orig + __ldg(R.k0 + ph);
    float acc = 0.f;
    for (int j = 0; j < R.L; ++j) acc = fmaf(__ldg(t + j), x[j], acc);
    ybuf[i] = acc;
  }
  __syncthreads();
}

// atomicMax on a float through the ordered-integer view (valid for any finite mix of signs)
__device__ inline void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__device__ inline float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Epilogue for one output cell: pad value, normalisation, SpecAugment zero-fill.
__device__ inline float finish_cell(const FbankParams& p, float x, int t, int c, const int* mk) {
  if (p.n_stats > 0) {
    int si = p.n_stats == 1 ? 0 : c;
    x = (x - __ldg(p.mean + si)) * (p.target_std / __ldg(p.std + si)) + p.target_mean;
  }
  if (mk) {
    if ((t >= mk[0] && t < mk[0] + mk[1]) || (c >= mk[2] && c < mk[2] + mk[3])) x = 0.f;
  }
  return x;
}

template <bool STATS>
__global__ void fbank_generic_kernel(const FbankParams p) {
  extern __shared__ float smem[];
  float* ybuf = smem;                         // [smem_y]
  float* xin = ybuf + p.smem_y;               // [smem_x]   (input tile, later power spectrum)
  float2* zbuf = reinterpret_cast<float2*>(xin + p.smem_x);   // [F/2][N] complex
  float* ebuf = reinterpret_cast<float*>(zbuf) + p.smem_z;    // [F] log energies
  float* pbuf = xin;

  const int b = blockIdx.x / p.tiles;
  const int F = p.tile_frames;
  const int t0 = (blockIdx.x - b * p.tiles) * F;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  const ClipInfo c = clip_info(p, b);
  const int cap = STATS ? p.max_frames : p.out_frames;
  const int m_eff = (int)(c.m < cap ? c.m : cap);
  if (!STATS && t0 == 0 && tid == 0 && p.n_frames_out) p.n_frames_out[b] = m_eff;
  const int t_end = (t0 + F < cap) ? t0 + F : cap;           // rows this CTA owns
  const int f_end = (t_end < m_eff) ? t_end : m_eff;         // real frames among them
  const int nf = f_end - t0;                                 // may be <= 0
  const int N = p.padded, NB = N >> 1;

  if (nf > 0) {
    // ---- H1 + H3: stage the resampled samples this tile's frames touch ------------------
    int64_t v_lo = (int64_t)t0 * p.shift, v_hi = (int64_t)(f_end - 1) * p.shift + p.size;
    int64_t s_lo = v_lo, s_hi = v_hi;
    if (p.frame_mode != 0) {
      const int64_t fo = frame_offset(p.size, p.shift, p.frame_mode);
      v_lo += fo; v_hi += fo;
      // hull of the mirrored index set {reflect(v) : v_lo <= v < v_hi}; at most one sample longer than the span
      s_lo = v_lo < 0 ? 0 : v_lo;
      s_hi = v_hi > c.n_rs ? c.n_rs : v_hi;
      if (v_hi > c.n_rs) { const int64_t r = reflect_index(v_hi - 1, c.n_rs, p.frame_mode); if (r < s_lo) s_lo = r; }
      if (v_lo < 0) { const int64_t r = reflect_index(v_lo, c.n_rs, p.frame_mode) + 1; if (r > s_hi) s_hi = r; }
      if (s_lo < 0) s_lo = 0;
      if (s_hi > c.n_rs) s_hi = c.n_rs;
    }
    stage_resampled(c, s_lo, s_hi, xin, ybuf);

    // ---- H4-H7: DC removal, pre-emphasis, window, zero-pad; two frames per complex row --
    const int npair = (nf + 1) >> 1;
    for (int g = warp; g < npair; g += nwarp) {
      float mean[2] = {0.f, 0.f}, e_raw[2] = {0.f, 0.f}, e_win[2] = {0.f, 0.f};
      int64_t base[2];
      bool live[2];
      for (int h = 0; h < 2; ++h) {
        int f = t0 + 2 * g + h;
        live[h] = f < f_end;
        base[h] = (int64_t)f * p.shift + frame_offset(p.size, p.shift, p.frame_mode);
      }
      auto sample = [&](int h, int j) -> float {
        int64_t v = base[h] + j;
        if (p.frame_mode != 0) v = reflect_index(v, c.n_rs, p.frame_mode);
        return ybuf[v - s_lo];
      };
      for (int h = 0; h < 2; ++h) {
        if (!live[h]) continue;
        if (p.remove_dc) {
          float s = 0.f;
          for (int j = lane; j < p.size; j += 32) s += sample(h, j);
          mean[h] = warp_sum(s) / (float)p.size;
        }
        if (p.use_energy && p.raw_energy) {
          float s = 0.f;
          for (int j = lane; j < p.size; j += 32) { float x = sample(h, j) - mean[h]; s = fmaf(x, x, s); }
          e_raw[h] = warp_sum(s);
        }
      }
      float2* z = zbuf + (size_t)g * N;
      for (int j = lane; j < N; j += 32) {
        float v[2] = {0.f, 0.f};
        if (j < p.size) {
          const float w = __ldg(p.window + j);
          for (int h = 0; h < 2; ++h) {
            if (!live[h]) continue;
            float x = sample(h, j) - mean[h];
            float xp = sample(h, j > 0 ? j - 1 : 0) - mean[h];      // replicate pad, kaldi.py:195-198
            v[h] = (x - p.preemph * xp) * w;
          }
        }
        e_win[0] = fmaf(v[0], v[0], e_win[0]);
        e_win[1] = fmaf(v[1], v[1], e_win[1]);
        z[__brev((unsigned)j) >> (32 - p.log2n)] = make_float2(v[0], v[1]);
      }
      if (p.use_energy) {
        for (int h = 0; h < 2; ++h) {
          float e = p.raw_energy ? e_raw[h] : warp_sum(e_win[h]);
          e = logf(fmaxf(e, B200_FLT_EPSILON));                       // kaldi.py:116-122
          if (p.has_energy_floor) e = fmaxf(e, p.log_energy_floor);
          if (lane == 0 && live[h]) ebuf[2 * g + h] = e;
        }
      }
    }
    __syncthreads();

    // ---- H8: radix-2 DIT FFT of the packed rows (input already bit-reversed) ------------
    const int total = npair * NB;
    for (int s = 1; s <= p.log2n; ++s) {
      const int half = 1 << (s - 1);
      for (int idx = tid; idx < total; idx += nt) {
        int g = idx / NB, t = idx - g * NB;
        int k = t & (half - 1);
        int i0 = ((t >> (s - 1)) << s) + k;
        float2 w = __ldg(p.twiddle + ((size_t)k << (p.log2n - s)));
        float2* z = zbuf + (size_t)g * N;
        float2 a = z[i0], bq = z[i0 + half];
        float2 tw = make_float2(bq.x * w.x - bq.y * w.y, bq.x * w.y + bq.y * w.x);
        z[i0] = make_float2(a.x + tw.x, a.y + tw.y);
        z[i0 + half] = make_float2(a.x - tw.x, a.y - tw.y);
      }
      __syncthreads();
    }

    // ---- split the two real spectra, |.|^2 (kaldi.py:616-618) ---------------------------
    for (int idx = tid; idx < total; idx += nt) {
      int g = idx / NB, k = idx - g * NB;
      const float2* z = zbuf + (size_t)g * N;
      float2 zk = z[k], zn = z[(N - k) & (N - 1)];
      float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
      float br = 0.5f * (zk.y + zn.y), bi = 0.5f * (zn.x - zk.x);
      float pa = ar * ar + ai * ai, pb = br * br + bi * bi;
      if (!p.use_power) { pa = sqrtf(pa); pb = sqrtf(pb); }
      pbuf[(size_t)(2 * g) * NB + k] = pa;
      pbuf[(size_t)(2 * g + 1) * NB + k] = pb;
    }
    __syncthreads();
  }

  // ---- H8c mel + log, H9 pad, H11 normalise, H10 mask, store ---------------------------
  const int mel_c0 = (p.use_energy && !p.htk_compat) ? 1 : 0;
  const int e_col = p.htk_compat ? p.n_mel : 0;
  auto cell = [&](int f, int col) -> float {   // f = frame index within the tile
    if (p.use_energy && col == e_col) return ebuf[f];
    int m = col - mel_c0;
    int st = __ldg(p.mel_start + m), cn = __ldg(p.mel_cnt + m);
    const float* w = p.mel_w + __ldg(p.mel_off + m);
    const float* P = pbuf + (size_t)f * NB + st;
    float acc = 0.f;
    for (int j = 0; j < cn; ++j) acc = fmaf(__ldg(w + j), P[j], acc);
    if (p.db_mode) return 10.f * log10f(fmaxf(acc, 1e-10f));           // amplitude_to_DB, functional.py:390-396 (ref = 1, power)
    if (p.use_log) acc = logf(fmaxf(acc, B200_FLT_EPSILON));           // kaldi.py:633
    return acc;
  };

  if (STATS) {
    if (nf <= 0) return;
    for (int col = tid; col < p.n_cols; col += nt) {
      double s = 0.0, ss = 0.0;
      for (int f = 0; f < nf; ++f) { double x = (double)cell(f, col); s += x; ss += x * x; }
      atomicAdd(p.sums + col, s);
      atomicAdd(p.sums + p.n_cols + col, ss);
    }
    if (tid == 0) atomicAdd(p.sums + 2 * p.n_cols, (double)nf);
    return;
  }

  const int rows = t_end - t0;
  if (rows <= 0) return;
  int mk_local[4];
  const int* mk = nullptr;
  if (p.masks) {
    for (int i = 0; i < 4; ++i) mk_local[i] = __ldg(p.masks + (size_t)b * 4 + i);
    mk = mk_local;
  }
  const int ncell = rows * p.n_cols;
  float vmax = -INFINITY;                      // db_mode: running maximum over this tile's real cells
  if (p.layout == 0) {
    float* o = p.out + ((size_t)b * p.out_frames + t0) * p.n_cols;
    for (int idx = tid; idx < ncell; idx += nt) {
      int f = idx / p.n_cols, col = idx - f * p.n_cols;
      float x = (f < nf) ? cell(f, col) : 0.f;
      if (f < nf) vmax = fmaxf(vmax, x);
      o[idx] = finish_cell(p, x, t0 + f, col, mk);
    }
  } else {
    float* o = p.out + (size_t)b * p.n_cols * p.out_frames + t0;
    for (int idx = tid; idx < ncell; idx += nt) {
      int col = idx / rows, f = idx - col * rows;
      float x = (f < nf) ? cell(f, col) : 0.f;
      if (f < nf) vmax = fmaxf(vmax, x);
      o[(size_t)col * p.out_frames + f] = finish_cell(p, x, t0 + f, col, mk);
    }
  }
  if (p.db_mode && p.clip_max) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0 && vmax > -INFINITY) atomic_max_float(p.clip_max + b, vmax);
  }
}

// subtract_mean (CMS, kaldi.py:220-226): column mean over the clip's real frames.  Runs after
// the main kernel wrote RAW features; re-applies pad / normalise / mask.
__global__ void cms_kernel(const FbankParams p) {
  const int b = blockIdx.x;
  const ClipInfo c = clip_info(p, b);
  const int m_eff = (int)(c.m < p.out_frames ? c.m : p.out_frames);
  int mk_local[4];
  const int* mk = nullptr;
  if (p.masks) {
    for (int i = 0; i < 4; ++i) mk_local[i] = __ldg(p.masks + (size_t)b * 4 + i);
    mk = mk_local;
  }
  for (int col = threadIdx.x; col < p.n_cols; col += blockDim.x) {
    float* o = p.out + (size_t)b * p.out_frames * p.n_cols;
    const size_t st_t = p.layout == 0 ? p.n_cols : 1, st_c = p.layout == 0 ? 1 : p.out_frames;
    float s = 0.f;
    for (int t = 0; t < m_eff; ++t) s += o[t * st_t + col * st_c];
    const float mu = m_eff > 0 ? s / (float)m_eff : 0.f;
    for (int t = 0; t < p.out_frames; ++t) {
      float x = t < m_eff ? o[t * st_t + col * st_c] - mu : 0.f;
      o[t * st_t + col * st_c] = finish_cell(p, x, t, col, mk);
    }
  }
}

__global__ void fill_u32_kernel(unsigned* dst, unsigned v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}

// Resample-only kernel (resample_waveform, src/datasets/preprocessing.py:61-76).
__global__ void resample_kernel(const FbankParams p, float* out, const int64_t* out_offsets,
                                int64_t out_clip_samples, int chunk) {
  extern __shared__ float smem[];
  float* ybuf = smem;
  float* xin = smem + chunk;
  const int b = blockIdx.x / p.tiles;
  const ClipInfo c = clip_info(p, b);
  int64_t s_lo = (int64_t)(blockIdx.x - b * p.tiles) * chunk;
  if (s_lo >= c.n_rs) return;
  int64_t s_hi = s_lo + chunk < c.n_rs ? s_lo + chunk : c.n_rs;
  stage_resampled(c, s_lo, s_hi, xin, ybuf);
  float* o = out + (out_offsets ? out_offsets[b] : (int64_t)b * out_clip_samples) + s_lo;
  for (int i = threadIdx.x; i < (int)(s_hi - s_lo); i += blockDim.x) o[i] = ybuf[i];
}

}  // namespace b200
