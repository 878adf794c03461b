"""Pins oracle/fbank_oracle.py against the golden fixtures made from the installed
torchaudio and the reference repo (tests/golden/make_golden.py) and against the
stage-level known answers of SURVEY.md section 8c.  CPU only."""
import math
import random

import numpy as np
import pytest

import oracle
from oracle import fbank_oracle as O
from inputs import config1_clips, short_clip, us8k_small_clips
from parity import assert_logmel_close

# float32 op-order noise between numpy and torch CPU kernels (BLAS / pocketfft vs MKL / SLEEF log)
ORACLE_TOL = 5e-4


def test_resample_kernels_match_torchaudio(golden):
    g = golden("resample_kernels.npz")
    for orig, shape, width in ((44100, (160, 475), 17), (22050, (320, 459), 9), (48000, (1, 41), 19),
                               (8000, (2, 15), None), (32000, (1, 28), None)):
        k, w, o, n = O.sinc_resample_kernel(orig, 16000)
        assert tuple(g[f"k{orig}_shape"]) == k.shape == shape
        assert int(g[f"k{orig}_width"]) == w
        if width is not None:
            assert w == width
        rows = g[f"k{orig}_rows"]
        np.testing.assert_allclose(k[rows], g[f"k{orig}_rowvals"], rtol=0, atol=1e-9)
        assert abs(k.astype(np.float64).sum() - float(g[f"k{orig}_sum"])) < 1e-6
        assert abs((k.astype(np.float64) ** 2).sum() - float(g[f"k{orig}_sumsq"])) < 1e-7
        nz = (np.abs(k) > 1e-30).sum(axis=1)
        assert [nz.min(), nz.max()] == list(g[f"k{orig}_nnz_minmax"])
    k, *_ = O.sinc_resample_kernel(44100, 16000)
    nz = (np.abs(k) > 1e-30).sum(axis=1)
    assert 33 <= nz.min() and nz.max() <= 34


def test_resampled_lengths():
    assert O.resampled_length(220500, 44100, 16000) == 80000
    assert O.resampled_length(88200, 22050, 16000) == 64000
    assert O.resampled_length(192000, 48000, 16000) == 64000
    assert O.resampled_length(100001, 44100, 16000) == math.ceil(160 * 100001 / 441)


def test_mel_banks_match_torchaudio(golden):
    g = golden("mel_banks.npz")
    b = O.kaldi_mel_banks(128, 512, 16000.0, 20.0, 0.0)
    assert b.shape == (128, 256)
    np.testing.assert_allclose(b, g["b128"], rtol=0, atol=5e-5)   # fp32 log ulp noise / bin width
    assert (g["b128"] != 0).sum() == 504 and (b != 0).sum() == 504
    assert ((b != 0).sum(axis=0) <= 2).all()
    assert (b[3] == 0).all()
    np.testing.assert_allclose(O.kaldi_mel_banks(23, 512, 16000.0, 20.0, 0.0), g["b23"], rtol=0, atol=5e-5)
    bv = O.kaldi_mel_banks(40, 512, 16000.0, 20.0, -400.0, 100.0, -500.0, 1.1)
    np.testing.assert_allclose(bv, g["b40_vtln"], rtol=0, atol=1e-4)
    # exact (float64) evaluation is what the C library ships; must stay within 2e-5 of torch fp32
    b64 = O.kaldi_mel_banks(128, 512, 16000.0, 20.0, 0.0, dtype=np.float64)
    assert np.abs(b64 - g["b128"]).max() < 5e-5


def test_config1_pipeline(golden):
    g = golden("config1.npz")
    clips = config1_clips(40)
    y0 = O.resample(clips[0][0].numpy(), 44100, 16000)
    assert y0.shape[0] == int(g["res0_len"]) == 80000
    np.testing.assert_allclose(y0[:2000], g["res0_head"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(y0[-2000:], g["res0_tail"], rtol=0, atol=2e-6)
    assert abs(y0.astype(np.float64).sum() - float(g["res0_sum"])) < 1e-3
    opts = O.ast_fbank_options()
    for j, i in enumerate(g["full_idx"]):
        f = O.kaldi_fbank(O.resample(clips[int(i)][0].numpy(), 44100, 16000), opts)
        assert f.shape == (498, 128)
        assert np.abs(f - g["full"][j]).max() < ORACLE_TOL
        assert (f[:, 3] == np.float32(O.LOG_FLT_EPSILON)).all()
    for i in (5, 17):
        f = O.kaldi_fbank(O.resample(clips[i][0].numpy(), 44100, 16000), opts)
        assert np.abs(f[g["frame_idx"]] - g["frames"][i]).max() < ORACLE_TOL
        np.testing.assert_allclose(f.astype(np.float64).sum(0), g["colsum"][i], rtol=1e-5, atol=1e-3)


def test_fp64_truth_noise_floor():
    """the fp32 path sits ~5e-4 from the fp64 evaluation on broadband noise."""
    w = config1_clips(1)[0][0].numpy()
    opts = O.ast_fbank_options()
    f32 = O.kaldi_fbank(O.resample(w, 44100, 16000), opts)
    f64 = O.kaldi_fbank(O.resample(w, 44100, 16000, dtype=np.float64), opts, dtype=np.float64)
    d = np.abs(f32 - f64).max()
    assert d < 1e-3, d


def test_kaldi_variants(golden):
    from make_golden_variants import KALDI_VARIANTS
    g = golden("kaldi_variants.npz")
    w = short_clip(8000 * 3).numpy()
    for name, kw in KALDI_VARIANTS.items():
        f = O.kaldi_fbank(w, **kw)
        assert f.shape == g[name].shape, name
        tol = ORACLE_TOL if kw.get("use_log_fbank", True) else None
        if tol is None:      # linear energies: relative check
            np.testing.assert_allclose(f, g[name], rtol=2e-4, atol=1e-5, err_msg=name)
        else:
            assert_logmel_close(f, g[name], tol, name)
    two = np.concatenate([w, -0.5 * w[:, ::-1]], 0)
    f = O.kaldi_fbank(two, channel=1, num_mel_bins=40)
    assert np.abs(f - g["channel1"]).max() < ORACLE_TOL


def test_kaldi_preconditions():
    w = short_clip(300).numpy()
    with pytest.raises(AssertionError):
        O.kaldi_fbank(w)                                       # window 400 > 300 samples
    with pytest.raises(AssertionError):
        O.kaldi_fbank(short_clip(1000).numpy(), channel=2)
    with pytest.raises(ValueError):
        O.kaldi_fbank(short_clip(1000).numpy(), dither=1.0)
    assert O.kaldi_fbank(short_clip(1000).numpy(), min_duration=1.0).shape == (0,)
    assert O.kaldi_fbank(np.zeros(1600, np.float32), num_mel_bins=128).min() == np.float32(O.LOG_FLT_EPSILON)


def test_silence_is_floor():
    f = O.kaldi_fbank(np.zeros(16000, np.float32), O.ast_fbank_options())
    assert (f == np.float32(O.LOG_FLT_EPSILON)).all()
    assert abs(O.LOG_FLT_EPSILON - (-15.942385)) < 1e-5


def test_specaugment_replay(golden):
    g = golden("specaugment.npz")
    rows = g["reference"]
    state = {}
    for F, T, s, draw, t0, tl, f0, fl in rows:
        key = (int(F), int(T), int(s))
        if draw == 0:
            state[key] = O.PyRandom(int(s))
        got = O.specaugment_intervals_reference(state[key], int(T), int(F), 192, 48)
        assert got == (t0, tl, f0, fl), (key, draw, got)
    state = {}
    for F, T, s, draw, t0, tl, f0, fl in g["torchaudio"]:
        key = (int(F), int(T), int(s))
        if draw == 0:
            state[key] = O.TorchCPUGenerator(int(s))
        got = O.specaugment_intervals_torchaudio(state[key], int(T), int(F), 192, 48)
        if tl == T or fl == F:         # one axis fully masked: the all-zero probe hides the other interval
            assert got[1] == T or got[3] == F, (key, draw, got)
            continue
        assert got == (t0, tl, f0, fl), (key, draw, got)
    # SURVEY 8c (6), (7)
    assert O.specaugment_intervals_reference(O.PyRandom(42), 1379, 128) == (228, 164, 94, 2)
    assert O.specaugment_intervals_torchaudio(O.TorchCPUGenerator(42), 1379, 128, 192, 48) == (1106, 169, 105, 18)


def test_pyrandom_matches_stdlib():
    for seed in (0, 7, 42, 2**33 + 1):
        r = random.Random(seed)
        o = O.PyRandom(seed)
        for _ in range(50):
            a, b = 1, 1 + r.getrandbits(6)
            r2 = random.Random(seed)   # keep streams aligned: compare fresh draws
        r = random.Random(seed)
        o = O.PyRandom(seed)
        for hi in (1, 2, 3, 47, 48, 127, 128, 191, 1378, 2**20):
            assert r.randint(0, hi) == o.randint(0, hi)


def test_us8k_small(golden):
    g = golden("us8k_small.npz")
    clips, rates = us8k_small_clips(9)
    assert list(g["rates"]) == rates
    assert set(rates) == {22050, 44100, 48000}
    for i in (0, 3, 8):
        f, m = O.ast_frontend(clips[i][0].numpy(), rates[i], target_frames=1024)
        assert m == int(g["n_frames"][i])
        assert f.shape == (1024, 128)
        assert np.abs(f[:400] - g["feats"][i]).max() < ORACLE_TOL
        assert (f[m:] == 0).all()


def test_normalize_and_pad():
    x = np.arange(12, dtype=np.float32).reshape(3, 4)
    p = O.pad_or_crop_frames(x, 5)
    assert p.shape == (5, 4) and (p[3:] == 0).all() and (p[:3] == x).all()
    assert (O.pad_or_crop_frames(x, 2) == x[:2]).all()
    n = O.normalize(x, -6.6268, 5.0613)
    np.testing.assert_allclose(n, (x + 6.6268) / (2 * 5.0613), rtol=1e-6)
    nb = O.normalize(x, np.arange(4, dtype=np.float32), np.full(4, 2.0, np.float32), 1.0, 1.0)
    np.testing.assert_allclose(nb, (x - np.arange(4)) / 2.0 + 1.0, rtol=1e-6)


def test_dataset_stats(golden):
    g = golden("config1.npz")
    clips = config1_clips(3)
    opts = O.ast_fbank_options()
    feats = [O.kaldi_fbank(O.resample(c[0].numpy(), 44100, 16000), opts) for c in clips]
    s = O.dataset_stats(feats)
    assert s.shape == (257,) and s[256] == 3 * 498
    ref = g["colsum"][:3].sum(0)
    np.testing.assert_allclose(s[:128], ref, rtol=1e-4)
    mean_b, std_b, gm, gs = O.stats_finalize(g["stats_sums"])
    assert mean_b.shape == (128,) and (std_b[np.arange(128) != 3] > 0).all() and std_b[3] < 1e-3


def test_reference_actual(golden):
    g = golden("reference_actual.npz")
    w = short_clip(44100, seed=77).numpy()
    a = O.reference_ast_preprocess(w, 44100)
    assert a.shape == g["ast_44k"].shape == (1, 128, 276)
    assert np.abs(a - g["ast_44k"]).max() < 2e-4
    c = O.reference_ast_preprocess(w, 44100, normalize_flag=False)
    assert np.abs(c - g["ast_44k_nonorm"]).max() < 2e-3       # dB units
    b = O.reference_ast_preprocess(short_clip(22050, seed=78).numpy(), 22050)
    assert np.abs(b - g["ast_22k"]).max() < 2e-4
    d = O.reference_melspectrogram_db(w[0], 44100, 1024, 160, None, 128, 80.0)
    assert np.abs(d[None] - g["fallback_melspec"]).max() < 2e-3


def test_oracle_header_says_test_only():
    assert "TEST INFRASTRUCTURE ONLY" in oracle.__doc__
    assert "TEST INFRASTRUCTURE ONLY" in O.__doc__
