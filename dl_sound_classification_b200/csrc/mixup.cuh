// On-device Mixup of a batch of spectrograms with partners from a feature bank.
//
// Reference: MixupAugmentation.__call__  src/datasets/preprocessing.py:933-968
//   mixed = lam * spec1 + (1 - lam) * spec2        (float32 tensor arithmetic: three roundings per element)
//   soft[label1] = lam; soft[label2] = 1 - lam     (in this order: equal labels leave 1 - lam)
// and MixupDataset.apply_mixup  src/datasets/esc50.py:43-76 (which samples are mixed, with whom).
// The random draws are replayed on the host (dl_sound_classification_b200/mixup.py); the kernel is the arithmetic,
// with the reference's own order of roundings (explicit __fmul_rn / __fadd_rn: no fma contraction), so the mixed
// batch is bit-identical to the per-sample reference loop.  Pure streaming: 2 reads + 1 write per element.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr int MIX_THREADS = 256;
constexpr int MIX_VEC_PER_THREAD = 4;          // float4 per thread: 4 independent 16-B loads per stream in flight

// grid = B * tiles CTAs, clip-major (a linear grid: B is not bound by the 65 535 limit of gridDim.y).
// partner[i] < 0: sample i is not mixed (spec returned unchanged, one-hot label).
__global__ void __launch_bounds__(MIX_THREADS) mixup_kernel(const float* __restrict__ x, const float* __restrict__ bank,
                                                            const int32_t* __restrict__ partner, const float* __restrict__ lam,
                                                            int64_t clip_elems, int tiles, float* __restrict__ out) {
  const int i = (int)(blockIdx.x / (unsigned)tiles);
  const int tile = (int)(blockIdx.x - (unsigned)i * (unsigned)tiles);
  const int j = __ldg(partner + i);
  const float l = __ldg(lam + i);
  const float oml = __fsub_rn(1.0f, l);
  const float* a = x + (size_t)i * clip_elems;
  const float* b = bank + (size_t)(j < 0 ? 0 : j) * clip_elems;
  float* o = out + (size_t)i * clip_elems;
  const int64_t nvec = clip_elems >> 2;
  const int64_t v0 = ((int64_t)tile * MIX_VEC_PER_THREAD) * MIX_THREADS + threadIdx.x;
  if ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)o) & 15) == 0) {
    float4 va[MIX_VEC_PER_THREAD], vb[MIX_VEC_PER_THREAD];
#pragma unroll
    for (int u = 0; u < MIX_VEC_PER_THREAD; ++u) {
      const int64_t v = v0 + (int64_t)u * MIX_THREADS;
      if (v < nvec) {
        va[u] = __ldcs(reinterpret_cast<const float4*>(a) + v);
        if (j >= 0) vb[u] = __ldg(reinterpret_cast<const float4*>(b) + v);
      }
    }
#pragma unroll
    for (int u = 0; u < MIX_VEC_PER_THREAD; ++u) {
      const int64_t v = v0 + (int64_t)u * MIX_THREADS;
      if (v < nvec) {
        float4 r = va[u];
        if (j >= 0) {
          r.x = __fadd_rn(__fmul_rn(l, va[u].x), __fmul_rn(oml, vb[u].x));
          r.y = __fadd_rn(__fmul_rn(l, va[u].y), __fmul_rn(oml, vb[u].y));
          r.z = __fadd_rn(__fmul_rn(l, va[u].z), __fmul_rn(oml, vb[u].z));
          r.w = __fadd_rn(__fmul_rn(l, va[u].w), __fmul_rn(oml, vb[u].w));
        }
        __stcs(reinterpret_cast<float4*>(o) + v, r);
      }
    }
    // tail (clip_elems not a multiple of 4): first tile only
    if (tile == 0 && threadIdx.x < (clip_elems & 3)) {
      const int64_t e = (nvec << 2) + threadIdx.x;
      o[e] = j >= 0 ? __fadd_rn(__fmul_rn(l, a[e]), __fmul_rn(oml, b[e])) : a[e];
    }
  } else {                                       // unaligned rows: scalar path
    const int64_t e0 = v0 * 4, stride = (int64_t)MIX_THREADS * 4;
#pragma unroll
    for (int u = 0; u < MIX_VEC_PER_THREAD; ++u)
      for (int q = 0; q < 4; ++q) {
        const int64_t e = e0 + (int64_t)u * stride + q;
        if (e < clip_elems) o[e] = j >= 0 ? __fadd_rn(__fmul_rn(l, a[e]), __fmul_rn(oml, b[e])) : a[e];
      }
  }
}

// soft labels (B, C): zeros; [label1] = lam, then [label2] = 1 - lam (mixed) or [label1] = 1 (not mixed).
__global__ void mixup_labels_kernel(const int64_t* __restrict__ label, const int64_t* __restrict__ partner_label,
                                    const int32_t* __restrict__ partner, const float* __restrict__ lam, int B, int C,
                                    float* __restrict__ soft) {
  const int i = blockIdx.x;
  if (i >= B) return;
  const int j = __ldg(partner + i);
  const int l1 = (int)__ldg(label + i);
  const int l2 = j >= 0 ? (int)__ldg(partner_label + i) : -1;
  const float l = __ldg(lam + i);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = 0.f;
    if (j < 0) v = c == l1 ? 1.f : 0.f;
    else {
      if (c == l1) v = l;
      if (c == l2) v = __fsub_rn(1.0f, l);       // assigned second: wins when label1 == label2 (preprocessing.py:964-966)
    }
    soft[(size_t)i * C + c] = v;
  }
}

}  // namespace b200
