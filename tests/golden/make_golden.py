"""Generate tests/golden/*.npz from the INSTALLED torchaudio and the reference repo.

Run in the authoring container only (needs /root/reference and torchaudio):

    python tests/golden/make_golden.py

The fixtures pin ``oracle/fbank_oracle.py`` (CPU tests) and the CUDA path (GPU
tests).  Inputs are never stored: every test regenerates them with the same
``torch.manual_seed`` + ``torch.rand`` calls (``tests/inputs.py``), which are
deterministic for torch's CPU generator.  torch/torchaudio versions are
recorded in every file.
"""
import os
import random
import sys

import numpy as np
import torch
import torchaudio
import torchaudio.compliance.kaldi as kaldi
import torchaudio.transforms as T

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from inputs import config1_clips, us8k_small_clips, short_clip  # noqa: E402

META = dict(torch=torch.__version__, torchaudio=torchaudio.__version__)
torch.set_num_threads(1)


def save(name, **arrays):
    np.savez_compressed(os.path.join(HERE, name), **arrays, **{"meta_" + k: np.array(v) for k, v in META.items()})
    print(name, {k: getattr(v, "shape", None) for k, v in arrays.items()})


def ast_kwargs():
    return dict(htk_compat=True, sample_frequency=16000, use_energy=False, window_type="hanning",
                num_mel_bins=128, dither=0.0, frame_shift=10)


def gen_resample_kernels():
    out = {}
    for orig in (44100, 22050, 48000, 8000, 32000):
        r = T.Resample(orig, 16000)
        k = r.kernel[:, 0, :].numpy()
        out[f"k{orig}_shape"] = np.array(k.shape)
        out[f"k{orig}_width"] = np.array(r.width)
        out[f"k{orig}_sum"] = np.array(k.astype(np.float64).sum())
        out[f"k{orig}_sumsq"] = np.array((k.astype(np.float64) ** 2).sum())
        nz = (np.abs(k) > 1e-30).sum(axis=1)
        out[f"k{orig}_nnz_minmax"] = np.array([nz.min(), nz.max()])
        rows = sorted(set(min(r, k.shape[0] - 1) for r in (0, 1, k.shape[0] // 2, k.shape[0] - 1)))
        out[f"k{orig}_rows"] = np.array(rows)
        out[f"k{orig}_rowvals"] = k[rows]
    save("resample_kernels.npz", **out)


def gen_mel_banks():
    b, _ = kaldi.get_mel_banks(128, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
    b23, _ = kaldi.get_mel_banks(23, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
    bv, _ = kaldi.get_mel_banks(40, 512, 16000.0, 20.0, -400.0, 100.0, -500.0, 1.1)
    save("mel_banks.npz", b128=b.numpy(), b23=b23.numpy(), b40_vtln=bv.numpy())


def gen_config1():
    clips = config1_clips(40)
    rs = T.Resample(44100, 16000)
    feats = []
    res0 = None
    for i, w in enumerate(clips):
        y = rs(w)
        if i == 0:
            res0 = y[0].numpy().copy()
        feats.append(kaldi.fbank(y, **ast_kwargs()).numpy())
    feats = np.stack(feats)                                   # (40, 498, 128)
    f64 = feats.astype(np.float64)
    sums = np.concatenate([f64.sum((0, 1)), (f64 ** 2).sum((0, 1)), [feats.shape[0] * feats.shape[1]]])
    save("config1.npz",
         full_idx=np.array([0, 1, 39]), full=feats[[0, 1, 39]],
         frame_idx=np.array([0, 249, 497]), frames=feats[:, [0, 249, 497], :],
         colsum=f64.sum(1), colsumsq=(f64 ** 2).sum(1),
         stats_sums=sums,
         res0_head=res0[:2000], res0_tail=res0[-2000:], res0_len=np.array(res0.shape[0]),
         res0_sum=np.array(res0.astype(np.float64).sum()), res0_sumsq=np.array((res0.astype(np.float64) ** 2).sum()))


from make_golden_variants import KALDI_VARIANTS  # noqa: E402


def gen_kaldi_variants():
    w = short_clip(8000 * 3)            # 1.5 s @ 16 k, (1, n)
    out = {}
    for name, kw in KALDI_VARIANTS.items():
        out[name] = kaldi.fbank(w, **kw).numpy()
    two = torch.cat([w, -0.5 * w.flip(1)], 0)
    out["channel1"] = kaldi.fbank(two, channel=1, num_mel_bins=40).numpy()
    save("kaldi_variants.npz", **out)


def gen_specaugment():
    sys.path.insert(0, "/root/reference")
    from src.datasets.preprocessing import ASTPreprocessor, PreprocessingConfig
    from src.utils.audio import SpecAugment
    pre = ASTPreprocessor(PreprocessingConfig(sample_rate=44100, n_mels=128))

    def interval(mask_1d):
        idx = np.nonzero(mask_1d)[0]
        return (0, 0) if idx.size == 0 else (int(idx[0]), int(idx[-1] - idx[0] + 1))

    shapes = [(128, 1379), (128, 512), (128, 1024), (128, 100), (40, 300)]
    seeds = [0, 1, 42, 1234, 2**31 + 5, 2**40 + 17]
    ref_rows, ta_rows = [], []
    for (F, Tn) in shapes:
        spec = torch.ones(1, F, Tn)
        for s in seeds:
            random.seed(s)
            for draw in range(3):     # consecutive calls share one RNG stream
                o = pre.apply_specaugment(spec, 192, 48)[0].numpy()
                z = (o == 0)
                t0, tl = interval(z.all(axis=0))
                f0, fl = interval(z.all(axis=1))
                ref_rows.append([F, Tn, s, draw, t0, tl, f0, fl])
            torch.manual_seed(s)
            aug = SpecAugment(192, 48)
            for draw in range(3):
                o = aug(spec)[0].numpy()
                z = (o == 0)
                t0, tl = interval(z.all(axis=0))
                f0, fl = interval(z.all(axis=1))
                ta_rows.append([F, Tn, s, draw, t0, tl, f0, fl])
    save("specaugment.npz", reference=np.array(ref_rows, dtype=np.int64), torchaudio=np.array(ta_rows, dtype=np.int64))


def gen_us8k_small():
    clips, rates = us8k_small_clips(9)
    feats, ms = [], []
    for w, r in zip(clips, rates):
        y = T.Resample(int(r), 16000)(w)
        f = kaldi.fbank(y, **ast_kwargs()).numpy()
        ms.append(f.shape[0])
        pad = np.zeros((1024, 128), np.float32)
        pad[: f.shape[0]] = f
        feats.append(pad[:400])      # store the first 400 rows (>= max m = 398), rest is zero
    save("us8k_small.npz", feats=np.stack(feats), n_frames=np.array(ms), rates=np.array(rates),
         lengths=np.array([c.shape[1] for c in clips]))


def gen_reference_actual():
    sys.path.insert(0, "/root/reference")
    from src.datasets.preprocessing import ASTPreprocessor, PreprocessingConfig
    from src.utils.audio import melspectrogram
    cfg = PreprocessingConfig(sample_rate=44100, n_mels=128, bc_mixing=False, normalize=True,
                              target_mean=0.0, target_std=0.5)
    pre = ASTPreprocessor(cfg)
    w = short_clip(44100, seed=77)                       # 1 s @ 44.1 k
    a = pre.preprocess(w, 44100).numpy()
    w2 = short_clip(22050, seed=78)                      # 1 s @ 22.05 k -> resampled to 44.1 k
    b = pre.preprocess(w2, 22050).numpy()
    pre_nn = ASTPreprocessor(PreprocessingConfig(sample_rate=44100, n_mels=128, normalize=False))
    c = pre_nn.preprocess(w, 44100).numpy()
    d = melspectrogram(w, 44100, 128, 1024, 160, log_scale=True).numpy()
    save("reference_actual.npz", ast_44k=a, ast_22k=b, ast_44k_nonorm=c, fallback_melspec=d)


if __name__ == "__main__":
    gen_resample_kernels()
    gen_mel_banks()
    gen_config1()
    gen_kaldi_variants()
    gen_specaugment()
    gen_us8k_small()
    gen_reference_actual()
