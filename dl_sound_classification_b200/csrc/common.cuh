// Shared device-side parameter blocks and index arithmetic for the fbank kernels.
//
// Index arithmetic that the reference leaves to torch views is restated here once, as
// __host__ __device__ functions, so the CPU test-suite can pin it through the C ABI
// (b200fbank_num_frames / b200fbank_resampled_length) without a GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define B200_MAX_RATES 8
#define B200_FLT_EPSILON 1.1920928955078125e-07f   // torchaudio/compliance/kaldi.py:26
#define B200_LOG_FLT_EPSILON -15.942384719848633f  // float32(log(FLT_EPSILON))

namespace b200 {

// One entry of the polyphase table (torchaudio/functional/functional.py:1305-1402).
struct RateDev {
  int orig;          // gcd-reduced input rate  (441 for 44.1 kHz -> 16 kHz)
  int nw;            // gcd-reduced output rate (160)
  int width;         // left zero padding of the torchaudio conv (17)
  int klen;          // 2*width + orig: dense taps per phase (475)
  int L;             // taps kept per phase after dropping exact zeros (34)
  int identity;      // 1: this rate needs no resampling
  const float* taps; // [nw][L]
  const int* k0;     // [nw] dense index of the first kept tap of each phase
};

struct FbankParams {
  // input
  const float* wav;
  const int64_t* offsets;    // [B+1] or nullptr
  int64_t clip_samples;      // dense row length when offsets == nullptr
  const int32_t* rate_id;    // [B] or nullptr
  int B;
  RateDev rates[B200_MAX_RATES];
  // framing (torchaudio/compliance/kaldi.py:125-151)
  int shift, size, padded, log2n;
  int snip_edges, remove_dc, raw_energy, use_energy, htk_compat, use_power, use_log;
  int frame_mode;            // 0 snip_edges, 1 kaldi mirrored, 2 stft centred reflect (snip_edges == (frame_mode == 0))
  int db_mode;               // 1: epilogue = 10*log10(max(x, 1e-10)) and a per-clip running maximum (AmplitudeToDB)
  float* clip_max;           // db_mode: [B] per-clip maximum of the dB values (ordered-int atomicMax)
  float preemph, log_energy_floor; int has_energy_floor;
  const float* window;       // [size]
  const float2* twiddle;     // [padded/2]  W_N^k = (cos 2 pi k/N, -sin 2 pi k/N)
  // sparse mel (get_mel_banks, kaldi.py:436-511): bin m = sum_j w[off[m]+j] * P[start[m]+j]
  int n_mel, n_cols;
  const int* mel_start; const int* mel_cnt; const int* mel_off; const float* mel_w;
  // epilogue
  int out_frames, layout;
  const int32_t* masks;      // [B][4] or nullptr
  const float* mean; const float* std; int n_stats;
  float target_mean, target_std;
  float* out;
  int32_t* n_frames_out;
  // Mixup fused into the epilogue (tuned kernels only): out = lam * y + (1 - lam) * bank[partner] cell by cell, with the
  // reference's three roundings (csrc/mixup.cuh); the bank holds spectrograms in the SAME layout and out_frames as `out`
  const float* mix_bank;     // (N, out_frames, n_cols) / (N, 1, n_cols, out_frames), or nullptr
  const int32_t* mix_partner;   // [B] index into the bank, < 0: clip left alone
  const float* mix_lam;      // [B]
  double* sums;              // stats mode: [2*n_cols+1]
  int max_frames;            // stats mode frame cap
  int tile_frames;           // F: frames per CTA
  int tiles;                 // CTAs per clip (grid = B * tiles, clip-major)
  // dynamic shared memory carve-up (in floats)
  int smem_y, smem_x, smem_z;
};

__host__ __device__ inline int64_t resampled_length(int64_t n, int orig, int nw) {
  // ceil(new * length / orig), functional.py:1427
  return (n * (int64_t)nw + orig - 1) / orig;
}

// Framing modes.  0: kaldi snip_edges=True; 1: kaldi snip_edges=False (mirrored edges, kaldi.py:70-80);
// 2: torch.stft(center=True, pad_mode="reflect") as used by MelSpectrogram (torchaudio/functional/functional.py:123-137).
// `padded` (the FFT length) bounds the clips the mirrored framings accept: torch's reflect padding needs
// n > n_fft / 2 (torch.stft raises otherwise) and kaldi's mirrored left edge needs n >= size/2 - shift/2 samples to
// mirror; shorter clips yield NO frames here (a batch cannot raise per clip) instead of reading outside the clip.
__host__ __device__ inline int64_t num_frames(int64_t n, int size, int shift, int frame_mode, int padded = 0) {
  if (frame_mode == 0) return n < size ? 0 : 1 + (n - size) / shift;     // _get_strided, kaldi.py:63-69
  if (frame_mode == 1) return n < size / 2 - shift / 2 ? 0 : (n + shift / 2) / shift;
  return n <= padded / 2 ? 0 : 1 + n / shift;                            // stft: 1 + floor((n + 2*(n_fft/2) - n_fft) / hop)
}

// First (virtual) sample of frame 0.  Mode 2: the win_length window sits in the middle of the n_fft buffer, so its
// non-zero part starts win/2 before the frame centre; |DFT|^2 does not depend on where the zeros are.
__host__ __device__ inline int frame_offset(int size, int shift, int frame_mode) {
  return frame_mode == 0 ? 0 : (frame_mode == 1 ? -(size / 2 - shift / 2) : -(size / 2));
}

// Virtual sample index -> real index for snip_edges=False (kaldi.py:70-80): the signal is
// mirrored about -1/2 on the left and about n-1/2 on the right.
__host__ __device__ inline int64_t reflect_index(int64_t v, int64_t n, int frame_mode = 1) {
  if (frame_mode == 2) {            // torch 'reflect': [2,1,0,1,2] -- the edge sample is not repeated
    if (v < 0) return -v;
    if (v >= n) return 2 * n - 2 - v;
    return v;
  }
  if (v < 0) return -1 - v;         // kaldi: [1,0,0,1] -- mirrored about -1/2 and n-1/2
  if (v >= n) return 2 * n - 1 - v;
  return v;
}

}  // namespace b200
