/*
 * b200fbank.h -- C ABI of the B200-native waveform -> log-mel fbank frontend.
 *
 * Drop-in boundary.  The reference has NO FFI for this path: its
 * boundary is a pair of Python call signatures, so every entry point below names the
 * Python interface it replaces.  Paths are relative to the reference repo unless they
 * start with "torchaudio/" (site-packages/torchaudio, the third-party library that
 * holds the arithmetic).
 *
 * Conventions: plain pointers and sizes only (no torch types).  Every `d_*` pointer is
 * DEVICE memory on the plan's device; everything else is host memory.  `stream` is a
 * cudaStream_t passed as void*.  Functions return 0 on success or a negative
 * b200fbank_status; the message is available from b200fbank_last_error() (thread local).
 * A plan is immutable after creation: execute() is thread-safe per (plan, stream).
 * The library never allocates user-visible memory and never falls back to the CPU.
 */
#ifndef B200FBANK_H_
#define B200FBANK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200FBANK_ABI_VERSION 1
#define B200FBANK_MAX_RATES 8

typedef enum {
  B200FBANK_OK = 0,
  B200FBANK_ERR_INVALID = -1,      /* bad option / argument (torchaudio would assert or raise)   */
  B200FBANK_ERR_UNSUPPORTED = -2,  /* valid in torchaudio but not implemented here (e.g. dither) */
  B200FBANK_ERR_CUDA = -3,         /* CUDA runtime error                                          */
  B200FBANK_ERR_NO_DEVICE = -4     /* host-only plan used for a device call                       */
} b200fbank_status;

typedef enum {                      /* torchaudio/compliance/kaldi.py:86-113 */
  B200FBANK_WINDOW_POVEY = 0,
  B200FBANK_WINDOW_HANNING = 1,
  B200FBANK_WINDOW_HAMMING = 2,
  B200FBANK_WINDOW_RECTANGULAR = 3,
  B200FBANK_WINDOW_BLACKMAN = 4
} b200fbank_window;

typedef enum {
  B200FBANK_LAYOUT_BTF = 0,         /* (B, T, n_cols): kaldi-native rows = frames                */
  B200FBANK_LAYOUT_BFT = 1          /* (B, 1, n_cols, T): what src/models/ast.py:50-56 consumes  */
} b200fbank_layout;

typedef enum {
  B200FBANK_FRONTEND_KALDI_FBANK = 0,  /* torchaudio/compliance/kaldi.py:514-645 (north_star)    */
  B200FBANK_FRONTEND_MELSPEC_DB = 1    /* MelSpectrogram + AmplitudeToDB(top_db) + per-clip norm,
                                          src/datasets/preprocessing.py:988-998,1013-1039        */
} b200fbank_frontend;

/*
 * Options.  The first block is every keyword of torchaudio.compliance.kaldi.fbank
 * (torchaudio/compliance/kaldi.py:514-541) except `waveform`, `channel` (the caller
 * passes the selected channel), `dither` (must be 0: RNG parity is impossible) and
 * `min_duration` (a host-side early-out, handled by the Python wrapper).  The second
 * block is torchaudio.transforms.Resample (torchaudio/transforms/_transforms.py:945-979)
 * as called by resample_waveform (src/datasets/preprocessing.py:61-76).
 */
typedef struct {
  /* kaldi.fbank */
  double blackman_coeff;            /* 0.42  */
  double energy_floor;              /* 1.0   */
  double frame_length;              /* 25.0 ms */
  double frame_shift;               /* 10.0 ms */
  double high_freq;                 /* 0.0   */
  double low_freq;                  /* 20.0  */
  double preemphasis_coefficient;   /* 0.97  */
  double sample_frequency;          /* 16000.0 -- the rate the fbank runs at (resample target)  */
  double vtln_high;                 /* -500.0 */
  double vtln_low;                  /* 100.0 */
  double vtln_warp;                 /* 1.0   */
  int32_t num_mel_bins;             /* 23    */
  int32_t window_type;              /* b200fbank_window, default POVEY                          */
  int32_t htk_compat;               /* 0 */
  int32_t raw_energy;               /* 1 */
  int32_t remove_dc_offset;         /* 1 */
  int32_t round_to_power_of_two;    /* 1 (0 is B200FBANK_ERR_UNSUPPORTED unless already 2^k)    */
  int32_t snip_edges;               /* 1 */
  int32_t subtract_mean;            /* 0 */
  int32_t use_energy;               /* 0 */
  int32_t use_log_fbank;            /* 1 */
  int32_t use_power;                /* 1 */
  /* Resample: input rates a clip may arrive at; rate_id indexes this table.  A rate equal
     to sample_frequency means "no resampling" for that id. */
  int32_t n_rates;                  /* 1..B200FBANK_MAX_RATES */
  int32_t orig_rates[B200FBANK_MAX_RATES];
  int32_t lowpass_filter_width;     /* 6    */
  double rolloff;                   /* 0.99 */
  /* frontend selector + MELSPEC_DB parameters (ignored for KALDI_FBANK) */
  int32_t frontend;                 /* b200fbank_frontend */
  int32_t n_fft;                    /* 1024 (AST_N_FFT, src/datasets/preprocessing.py:56)       */
  int32_t hop_length;               /* 160  */
  int32_t win_length;               /* 400  */
  double top_db;                    /* 80.0; < 0 = None */
} b200fbank_opts;

typedef struct b200fbank_plan b200fbank_plan;

int b200fbank_abi_version(void);
/* sizeof(b200fbank_opts) as compiled into the library (binding sanity check). */
int b200fbank_sizeof_opts(void);

/* Fill `o` with the defaults of kaldi.fbank (torchaudio/compliance/kaldi.py:514-541) and
   Resample (torchaudio/transforms/_transforms.py:945-953); n_rates = 1, orig_rates[0] = 16000. */
void b200fbank_default_opts(b200fbank_opts* o);

/* Build the immutable tables (polyphase taps, window, twiddles, sparse mel weights) and,
   when device >= 0, upload them.  device = -1 makes a host-only plan: table queries and
   length arithmetic work, device calls return B200FBANK_ERR_NO_DEVICE.  Replaces the
   per-call table construction of kaldi.fbank (get_mel_banks, kaldi.py:621-624;
   _feature_window_function, :201) and Resample.__init__ (_transforms.py:969-979). */
int b200fbank_plan_create(const b200fbank_opts* o, int device, b200fbank_plan** out);
void b200fbank_plan_destroy(b200fbank_plan* p);
const char* b200fbank_last_error(void);

/* ---- host-side length arithmetic ------------------------------------------------- */
/* ceil(new*n/orig), torchaudio/functional/functional.py:1427 */
int64_t b200fbank_resampled_length(const b200fbank_plan* p, int64_t n_samples, int rate_id);
/* frames produced for a clip of n_samples input samples at rate_id
   (_get_strided, torchaudio/compliance/kaldi.py:63-69, after resampling) */
int64_t b200fbank_num_frames(const b200fbank_plan* p, int64_t n_samples, int rate_id);
/* num_mel_bins + use_energy */
int b200fbank_num_cols(const b200fbank_plan* p);

/* ---- table queries (host; used by the CPU test-suite to pin the host logic) -------- */
typedef enum {
  B200FBANK_TABLE_WINDOW = 0,       /* [window_size] float                                    */
  B200FBANK_TABLE_MEL_DENSE = 1,    /* [num_mel_bins][padded/2] float (get_mel_banks layout)   */
  B200FBANK_TABLE_TAPS_DENSE = 2    /* [new][2*width+orig] float for rate `arg`                */
} b200fbank_table;
/* Copies up to `cap` floats into dst; returns the table length in floats (or <0). */
int64_t b200fbank_plan_table(const b200fbank_plan* p, int table, int arg, float* dst, int64_t cap);
/* info[0..7] = window_shift, window_size, padded_window_size, n_rates, orig_reduced(arg),
   new_reduced(arg), width(arg), taps_per_phase(arg) */
int b200fbank_plan_info(const b200fbank_plan* p, int arg, int64_t info[8]);

/* ---- device calls ---------------------------------------------------------------- */
/*
 * The fused path: resample -> frame -> DC removal -> pre-emphasis -> window -> real FFT
 * -> power -> mel -> log-floor -> pad/crop to out_frames -> normalise -> SpecAugment
 * zero-fill -> store (each output written once).  Replaces, per clip,
 * ASTPreprocessor.preprocess (src/datasets/preprocessing.py:1013-1039),
 * resample_waveform (:61-76), kaldi.fbank (torchaudio/compliance/kaldi.py:514-645) and
 * ASTPreprocessor.apply_specaugment (:1075-1104) -- batched.
 *
 *  d_wav        concatenated mono float32 clips
 *  d_offsets    [B+1] sample offsets into d_wav, or NULL for a dense (B, clip_samples) batch
 *  clip_samples row length when d_offsets == NULL
 *  d_rate_id    [B] index into orig_rates, or NULL (all clips at orig_rates[0])
 *  d_masks      [B][4] = t_start, t_len, f_start, f_len (len 0 = none), or NULL; cells are
 *               set to 0.0 AFTER normalisation (src/datasets/esc50.py:267-273)
 *  d_mean/d_std [n_stats] with n_stats in {0, 1, n_cols}; 0 = no normalisation;
 *               out = (x - mean) / std * target_std + target_mean
 *               (src/datasets/preprocessing.py:1035-1037; north_star (x-mean)/(2*std))
 *  d_out        (B, out_frames, n_cols) or (B, 1, n_cols, out_frames) float32; rows past a
 *               clip's frame count hold 0.0 before normalisation
 *  d_n_frames   [B] min(frames, out_frames) per clip, or NULL
 */
int b200fbank_execute(const b200fbank_plan* p, const float* d_wav, const int64_t* d_offsets,
                      int64_t clip_samples, const int32_t* d_rate_id, int B,
                      const int32_t* d_masks, const float* d_mean, const float* d_std, int n_stats,
                      float target_mean, float target_std, int out_frames, int layout,
                      float* d_out, int32_t* d_n_frames, void* stream);

/*
 * b200fbank_execute with Mixup fused into the epilogue (SURVEY.md section 8f N3 moved into the producing kernel): after
 * normalisation and the SpecAugment zero-fill -- the reference's order, src/datasets/esc50.py:267-285 -- clip b becomes
 * lam[b] * y + (1 - lam[b]) * bank[partner[b]] with the three float32 roundings of MixupAugmentation.__call__
 * (src/datasets/preprocessing.py:960, csrc/mixup.cuh), pad rows included (the reference mixes whole tensors), so the
 * result is bit-identical to b200fbank_execute followed by b200fbank_mixup while the batch is written once instead of
 * written, re-read and written again.
 *  d_bank      (N, out_frames, n_cols) or (N, 1, n_cols, out_frames) float32 spectrograms in the SAME layout and
 *              out_frames as d_out (the reference's `_cached_data`, esc50.py:280-282); must not alias d_out
 *  d_partner   [B] index into the bank, < 0: the clip is left alone
 *  d_lam       [B] mixing coefficients
 * Plans that run on the generic kernel (or with subtract_mean) mix with a second launch, same result.
 */
int b200fbank_execute_mixup(const b200fbank_plan* p, const float* d_wav, const int64_t* d_offsets,
                            int64_t clip_samples, const int32_t* d_rate_id, int B,
                            const int32_t* d_masks, const float* d_mean, const float* d_std, int n_stats,
                            float target_mean, float target_std, int out_frames, int layout,
                            const float* d_bank, const int32_t* d_partner, const float* d_lam,
                            float* d_out, int32_t* d_n_frames, void* stream);

/*
 * The reference-ACTUAL recipe on a MELSPEC_DB plan: resample-if-needed ->
 * MelSpectrogram(n_fft, win_length, hop_length, power=2, center, reflect, periodic Hann, HTK mel, norm=None)
 * -> AmplitudeToDB(top_db) -> per-clip (x - mean) / unbiased_std * target_std + target_mean.  Replaces
 * ASTPreprocessor.preprocess (src/datasets/preprocessing.py:1013-1039) and melspectrogram()
 * (src/utils/audio.py:60-84).  to_db = 0 returns the raw mel power (log_scale=False); normalize = 0 skips
 * the per-clip normalisation.  d_clip_max: [B] floats of workspace (receives each clip's dB maximum); d_n_frames is
 * required with to_db or masks (the per-clip pass reads it).  Output rows past a clip's frame count (1 + n/hop) hold
 * 0.0; clips of at most n_fft/2 samples (torch.stft's reflect padding raises for them) produce no frames.
 */
int b200fbank_melspec_db(const b200fbank_plan* p, const float* d_wav, const int64_t* d_offsets,
                         int64_t clip_samples, const int32_t* d_rate_id, int B, const int32_t* d_masks,
                         int to_db, int normalize, float target_mean, float target_std, int out_frames,
                         int layout, float* d_out, int32_t* d_n_frames, float* d_clip_max, void* stream);

/* Per-clip normalisation of features already on the device, in place (the reference's ASTPreprocessor.preprocess,
   src/datasets/preprocessing.py:1030-1037, applied to a batch): for every clip the global mean and UNBIASED standard
   deviation over its own d_n_frames[b] x n_cols cells, x = (x - mean) / std * target_std + target_mean (skipped when
   std == 0 or normalize == 0), then the SpecAugment cells of d_masks (may be NULL) are zeroed as
   src/datasets/esc50.py:267-273 does after the transform.  d_clip_max ([B], may be NULL) with top_db >= 0 applies
   AmplitudeToDB's clamp max(x, clip_max - top_db) first (torchaudio/functional/functional.py:398-402).  Rows past a
   clip's frame count are left as they are (the 0.0 padding).  Used by b200fbank_melspec_db itself and, after
   b200fbank_execute without statistics, for the kaldi recipe with per-clip statistics.  No plan is needed. */
int b200fbank_clip_normalize(float* d_x, const int32_t* d_n_frames, int B, int out_frames, int n_cols, int layout,
                             const float* d_clip_max, float top_db, int normalize, float target_mean, float target_std,
                             const int32_t* d_masks, void* stream);

/* Whole-clip DC removal of the waveforms, `waveform - waveform.mean()` (the AST recipe's
   convention before kaldi.fbank): d_out[clip] = d_wav[clip] - mean(d_wav[clip]) per clip, same indexing as
   b200fbank_execute (d_offsets / clip_samples); d_out may alias d_wav; d_mean ([B], may be NULL) receives the means. */
int b200fbank_remove_clip_mean(const float* d_wav, const int64_t* d_offsets, int64_t clip_samples, int B, float* d_out,
                               float* d_mean, void* stream);

/* 16-bit PCM -> float32 waveforms on the device: d_out = float(d_pcm) / d_divisor[clip] (correctly rounded), same
   indexing as b200fbank_execute; d_divisor NULL = 32768, i.e. what torchaudio.load returns for a 16-bit WAV
   (src/utils/audio.py:42); d_divisor[b] = max |pcm| of clip b reproduces scripts/prepare_esc50.py:94-101 (peak
   normalisation of the loaded samples) bit for bit.  max_clip_samples: the longest clip of a ragged batch (sizes the
   launch; ignored for dense batches).  Halves the host -> device bytes of the end-to-end path. */
int b200fbank_pcm16_to_float(const int16_t* d_pcm, const int64_t* d_offsets, int64_t clip_samples, int B,
                             const float* d_divisor, int64_t max_clip_samples, float* d_out, void* stream);

/* Dataset-statistics pass (north_star config 4; no reference code): the same fused path
   with the epilogue replaced by float64 accumulation of per-column sum / sum of squares
   over the REAL frames [0, min(frames, max_frames)) of every clip.  d_sums is
   [2*n_cols + 1] doubles (sum, sumsq, frame count) and is ADDED to (zero it first).
   The cross-GPU reduction is one all-reduce of d_sums (torch.distributed / NCCL). */
int b200fbank_stats_accumulate(const b200fbank_plan* p, const float* d_wav, const int64_t* d_offsets,
                               int64_t clip_samples, const int32_t* d_rate_id, int B, int max_frames,
                               double* d_sums, void* stream);

/* Resample only: torchaudio.transforms.Resample(orig, sample_frequency)(wave) as called by
   resample_waveform (src/datasets/preprocessing.py:61-76).  Output clip b is written at
   d_out + (d_out_offsets ? d_out_offsets[b] : b * out_clip_samples). */
int b200fbank_resample(const b200fbank_plan* p, const float* d_wav, const int64_t* d_offsets,
                       int64_t clip_samples, const int32_t* d_rate_id, int B, float* d_out,
                       const int64_t* d_out_offsets, int64_t out_clip_samples, void* stream);

/* Mixup of a batch with partners from a feature bank, the arithmetic of
   MixupAugmentation.__call__ (src/datasets/preprocessing.py:933-968) applied to sample i of d_x with
   d_bank[d_partner[i]] as the reference's MixupDataset.apply_mixup pairs them (src/datasets/esc50.py:43-76):
     out[i] = lam[i] * x[i] + (1 - lam[i]) * bank[partner[i]]     (float32, the reference's order of roundings)
   d_partner[i] < 0: sample i is not mixed, out[i] = x[i].  Rows are clip_elems floats; d_out may alias d_x.
   The random draws (who is mixed, with whom, lambda ~ Beta) are host logic: dl_sound_classification_b200/mixup.py
   replays the reference's generators in the reference's order.  No plan is needed. */
int b200fbank_mixup(const float* d_x, const float* d_bank, const int32_t* d_partner, const float* d_lam, int B,
                    int64_t clip_elems, float* d_out, void* stream);

/* Soft labels of the same mix: d_soft[B][num_classes] = 0, then [label] = lam and [partner_label] = 1 - lam, in
   this order (equal labels leave 1 - lam, preprocessing.py:962-966); one-hot for unmixed samples
   (create_one_hot_labels, preprocessing.py:39-52). */
int b200fbank_mixup_labels(const int64_t* d_label, const int64_t* d_partner_label, const int32_t* d_partner,
                           const float* d_lam, int B, int num_classes, float* d_soft, void* stream);

/* AST patch embedding fed from the frontend's features: Conv2d(1, D, patch, stride) on
   d_feat (B, 1, F, T) float32 -> d_out (B, Fp * Tp, D), patch index Fp-major as flatten(2).transpose(1, 2) leaves it
   (PatchEmbed.forward, src/models/ast_mini.py:7-15; ASTModel.forward, src/models/ast.py:30,50-56).  An im2col GEMM on
   the 5th-generation tensor cores (tcgen05.mma, fp16 operands as under the reference's "16-mixed" AST setting, fp32
   accumulation in TMEM).  d_weight_f16: (D, patch * patch) fp16 = weight[d][0][i][j] at patch * i + j; d_bias (D) float32 or
   NULL; d_out fp16 (out_f16 != 0) or float32.  Built for patch == 16 and D a multiple of 192 (192 / 384 / 768). */
int b200fbank_patch_embed(const float* d_feat, int B, int F, int T, const void* d_weight_f16, const float* d_bias, int D,
                          int patch, int stride, void* d_out, int out_f16, void* stream);

/* Number of kernel launches the calls above issued on this thread since the last reset
   (bench.py's gpu_launches claim). */
int64_t b200fbank_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* B200FBANK_H_ */
