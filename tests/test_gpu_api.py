"""GPU tests of the reference-facing API (ASTPreprocessor mirror, resample_waveform, stats) and
of size-independent properties at BASELINE.json's full sizes.  Runs on the B200 box (-m gpu)."""
import math
import random

import numpy as np
import pytest
import torch

from inputs import config1_clips, short_clip, us8k_small_clips
from parity import LOGMEL_TOL, assert_logmel_close

pytestmark = pytest.mark.gpu
AST_MEAN, AST_STD = -6.6268, 5.0613


@pytest.fixture(scope="module")
def b2():
    import dl_sound_classification_b200 as m
    assert torch.cuda.is_available()
    return m


@pytest.fixture(scope="module")
def O():
    from oracle import fbank_oracle
    return fbank_oracle


def test_ast_preprocessor_per_clip_contract(b2, O):
    """preprocess(waveform[1,N] CPU, sr) -> [1, n_mels, T] on the input's device (preprocessing.py:1013)."""
    pre = b2.create_preprocessor("ast", dict(sample_rate=44100, n_mels=128, normalize=True, target_mean=0.0,
                                             target_std=0.5, norm_mean=AST_MEAN, norm_std=AST_STD,
                                             target_frames=512, frontend="kaldi_fbank"), "/tmp/unused")
    w = config1_clips(1)[0]
    out = pre.preprocess(w, 44100)
    assert out.device.type == "cpu" and tuple(out.shape) == (1, 128, 512) and out.dtype == torch.float32
    ref, m = O.ast_frontend(w[0].numpy(), 44100, target_frames=512, mean=AST_MEAN, std=AST_STD)
    assert np.abs(out[0].numpy().T - ref).max() < LOGMEL_TOL / (2 * AST_STD) * 1.01
    assert torch.equal(pre.preprocess_with_cache(w, 44100, None), out)
    g = pre.preprocess(w.cuda(), 44100)
    assert g.is_cuda and torch.equal(g.cpu(), out)
    # natural frame count when target_frames is not configured
    pre2 = b2.ASTPreprocessor(b2.PreprocessingConfig(sample_rate=44100, n_mels=128, normalize=False, target_sample_rate=16000))
    assert tuple(pre2.preprocess(w, 44100).shape) == (1, 128, 498)
    # per-clip normalisation (the reference's own convention: global mean / unbiased std, * 0.5)
    pre3 = b2.ASTPreprocessor(b2.PreprocessingConfig(sample_rate=44100, n_mels=128, normalize=True, frontend="kaldi_fbank"))
    x = pre3.preprocess(w, 44100)
    assert abs(float(x.mean())) < 1e-4 and abs(float(x.std()) - 0.5) < 1e-4
    with pytest.raises(AssertionError):
        pre2.preprocess(torch.zeros(1, 500), 44100)            # shorter than one 25 ms window


def test_batch_with_fused_masks_matches_sequential_reference_calls(b2, O):
    pre = b2.ASTPreprocessor(b2.PreprocessingConfig(sample_rate=44100, n_mels=128, norm_mean=AST_MEAN, norm_std=AST_STD,
                                                    target_frames=512, frontend="kaldi_fbank"))
    clips = config1_clips(4, length=110250)
    random.seed(5)
    masks = pre.draw_specaugment_masks(4, 512, 192, 48)
    out, nfr = pre.preprocess_batch(torch.cat(clips, 0), 44100, masks=masks)
    assert tuple(out.shape) == (4, 1, 128, 512)
    random.seed(5)
    for i, c in enumerate(clips):
        plain = pre.preprocess(c, 44100)
        want = pre.apply_specaugment(plain, 192, 48)             # the reference's call sequence (esc50.py:254-273)
        assert torch.equal(out[i].cpu(), want), i
        assert int(want.eq(0).sum()) > 0


def test_multi_crop_test(b2):
    pre = b2.ASTPreprocessor(b2.PreprocessingConfig(sample_rate=44100, n_mels=128, normalize=False, frontend="kaldi_fbank"))
    w = short_clip(44100 * 7, seed=5)
    crops = pre.multi_crop_test(w)
    assert len(crops) == 10 and all(tuple(c.shape) == (1, 128, 498) for c in crops)
    starts = torch.linspace(0, 44100 * 2, 10).long()
    one = pre.preprocess(w[..., int(starts[3]):int(starts[3]) + 220500], 44100)
    assert torch.equal(crops[3], one)
    assert len(pre.multi_crop_test(short_clip(44100 * 2))) == 1


def test_resample_waveform(b2, O):
    for rate in (44100, 22050, 48000, 8000):
        w = short_clip(rate // 2 + 37, seed=rate)
        y = b2.resample_waveform(w, rate, 16000)
        ref = O.resample(w[0].numpy(), rate, 16000, dtype=np.float64)
        assert y.device.type == "cpu" and tuple(y.shape) == (1, ref.shape[0])
        assert np.abs(y[0].numpy() - ref).max() < 3e-6, rate
    w = short_clip(1000)
    assert b2.resample_waveform(w, 44100, 44100) is w
    up = b2.resample_waveform(w.cuda(), 16000, 44100)                         # the reference also upsamples (prepare_esc50.py:53-57)
    ref = O.resample(w[0].numpy(), 16000, 44100, dtype=np.float64)
    assert up.is_cuda and np.abs(up[0].cpu().numpy() - ref).max() < 3e-6
    live = pytest.importorskip("torchaudio")
    import torchaudio.transforms as T
    w = short_clip(50000, seed=3)
    assert np.abs(b2.resample_waveform(w, 44100, 16000).numpy() - T.Resample(44100, 16000)(w).numpy()).max() < 3e-6


def test_dataset_stats_api(b2, O, golden):
    g = golden("config1.npz")
    clips = config1_clips(40)
    fe = b2.FbankFrontend(orig_rates=(44100,), **b2.AST_FBANK_KWARGS)
    ds = b2.DatasetStats(fe, max_frames=512)
    wav = torch.cat(clips, 0).cuda()
    ds.update(wav[:13]).update(wav[13:]).all_reduce()
    st = ds.finalize()
    rm, rs, rgm, rgs = O.stats_finalize(g["stats_sums"])
    assert st.frames == 40 * 498
    assert (np.abs(st.mean_per_bin.numpy() - rm) <= 1e-4 * (np.abs(rm) + rs)).all()
    keep = np.arange(128) != 3
    np.testing.assert_allclose(st.std_per_bin.numpy()[keep], rs[keep], rtol=1e-4)
    assert abs(st.mean - rgm) <= 1e-4 * abs(rgm) and abs(st.std - rgs) <= 1e-4 * rgs
    # the stats pass and the feature pass agree with each other (same kernel, different epilogue)
    feats, _ = fe(wav, out_frames=498)
    f64 = feats.double()
    s = torch.cat([f64.sum((0, 1)), (f64 * f64).sum((0, 1))]).cpu().numpy()
    np.testing.assert_allclose(ds.sums.cpu().numpy()[:256], s, rtol=1e-9, atol=1e-6)


# ---------------------------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties
# ---------------------------------------------------------------------------------------------
def test_config2_batch1024_properties(b2, O):
    """1024 ESC-50 clips -> (1024, 512, 128): batch invariance, pad rows, one-hop shift, gain."""
    B, N = 1024, 220500
    fe = b2.FbankFrontend(orig_rates=(44100,), **b2.AST_FBANK_KWARGS)
    gen = torch.Generator(device="cuda").manual_seed(99)
    wav = torch.rand((B, N), generator=gen, device="cuda") * 2 - 1
    out, nfr = fe(wav, out_frames=512, mean=AST_MEAN, std=AST_STD)
    assert tuple(out.shape) == (B, 512, 128) and bool((nfr == 498).all())
    assert bool(torch.isfinite(out).all())
    pad = (0.0 - AST_MEAN) * (0.5 / AST_STD)
    assert bool((out[:, 498:, :] - pad).abs().max() < 1e-6)
    idx = [0, 1, 511, 777, 1023]
    alone, _ = fe(wav[idx], out_frames=512, mean=AST_MEAN, std=AST_STD)
    assert torch.equal(alone, out[idx])                                   # batch invariance, bit-exact
    # a 441-sample (= one 16 kHz hop) delay shifts the features by one frame; a 4-hop delay keeps the
    # frame pairing of the packed FFT and the lane assignment, so it is bit-exact
    a, _ = fe(wav[:4], out_frames=512)
    sh = torch.zeros((4, N), device="cuda")
    sh[:, 441:] = wav[:4, :-441]
    b_, _ = fe(sh, out_frames=512)
    live1 = a[:, 3:497] > -12
    assert float((b_[:, 4:498] - a[:, 3:497])[live1].abs().max()) < LOGMEL_TOL   # two float32 evaluations of the same frames
    sh = torch.zeros((4, N), device="cuda")
    sh[:, 4 * 441:] = wav[:4, :-4 * 441]
    b_, _ = fe(sh, out_frames=512)
    assert torch.equal(b_[:, 8:498], a[:, 4:494])
    # gain 2^-3 lowers every log-mel value by 6 ln 2 (power-of-two scaling is exact in float32)
    c_, _ = fe(wav[:4] * 0.125, out_frames=512)
    live = a[:, :498] > -10          # stay clear of the FLT_EPSILON floor after the -4.16 shift
    assert float(((a[:, :498] - c_[:, :498]) - 6 * math.log(2.0))[live].abs().max()) < 5e-6
    # oracle on a sample of clips (the oracle needs ~15 ms per clip)
    for i in (3, 500, 1023):
        ref, _ = O.ast_frontend(wav[i].cpu().numpy(), 44100, target_frames=512, mean=AST_MEAN, std=AST_STD)
        assert np.abs(out[i].cpu().numpy() - ref).max() < LOGMEL_TOL / (2 * AST_STD) * 1.01


def test_config3_us8k_4096_ragged_masks(b2, O):
    """4096 ragged clips at 22.05/44.1/48 kHz -> (4096, 1024, 128) with SpecAugment masks."""
    B = 4096
    table = (22050, 44100, 48000)
    g = torch.Generator().manual_seed(31)
    rid = torch.randint(0, 3, (B,), generator=g)
    dur = 1.0 + 3.0 * torch.rand(B, generator=g)
    lens = (dur * torch.tensor(table)[rid]).long()
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)])
    gen = torch.Generator(device="cuda").manual_seed(32)
    flat = torch.rand(int(offsets[-1]), generator=gen, device="cuda") * 2 - 1
    fe = b2.FbankFrontend(orig_rates=table, **b2.AST_FBANK_KWARGS)
    random.seed(77)
    masks = b2.specaugment.draw_masks(B, 1024, 128, 192, 48, variant="reference")
    out, nfr = fe(flat, out_frames=1024, offsets=offsets, rate_ids=rid.int(), masks=masks, mean=AST_MEAN, std=AST_STD)
    assert tuple(out.shape) == (B, 1024, 128) and bool(torch.isfinite(out).all())
    want_frames = [fe.num_frames(int(n), int(r)) for n, r in zip(lens[:64], rid[:64])]
    assert nfr[:64].cpu().tolist() == want_frames and int(nfr.max()) <= 398
    # masked cells: exactly the drawn intervals, exactly 0.0 (bit-exact per seed)
    rng = O.PyRandom(77)
    zeros = out.eq(0)
    for i in range(B):
        t0, tl, f0, fl = O.specaugment_intervals_reference(rng, 1024, 128, 192, 48)
        assert masks[i].tolist() == [t0, tl, f0, fl]
        if i % 257 == 0:
            z = zeros[i].cpu().numpy()
            want = np.zeros((1024, 128), bool)
            want[t0:t0 + tl] = True
            want[:, f0:f0 + fl] = True
            assert (z == want).all(), i
    for i in (0, 1234, 4095):
        w = flat[int(offsets[i]):int(offsets[i + 1])].cpu().numpy()
        ref, m = O.ast_frontend(w, table[int(rid[i])], target_frames=1024, mean=AST_MEAN, std=AST_STD,
                                mask=masks[i].tolist())
        assert m == int(nfr[i])
        assert np.abs(out[i].cpu().numpy() - ref).max() < LOGMEL_TOL / (2 * AST_STD) * 1.01, i


def test_config4_stats_sharded_partials_add_up(b2, O):
    """Stats pass over 2048 synthetic clips in 8 contiguous shards == one pass over all of them, and a
    64-clip subset matches the float64 oracle to rel 1e-4 (the cross-rank all-reduce is a plain SUM)."""
    from dl_sound_classification_b200 import stats as ST
    B, N = 2048, 220500
    fe = b2.FbankFrontend(orig_rates=(44100,), **b2.AST_FBANK_KWARGS)
    gen = torch.Generator(device="cuda").manual_seed(4)
    wav = torch.rand((B, N), generator=gen, device="cuda") * 2 - 1
    whole = b2.DatasetStats(fe, 512).update(wav).sums
    parts = torch.zeros_like(whole)
    for r in range(8):
        lo, hi = ST.shard_bounds(B, r, 8)
        parts += b2.DatasetStats(fe, 512).update(wav[lo:hi]).sums
    assert float(whole[256]) == B * 498 == float(parts[256])
    np.testing.assert_allclose(parts.cpu().numpy(), whole.cpu().numpy(), rtol=1e-12)
    sub = b2.DatasetStats(fe, 512).update(wav[:64]).finalize()
    feats = [O.kaldi_fbank(O.resample(wav[i].cpu().numpy(), 44100, 16000), O.ast_fbank_options()) for i in range(64)]
    rm, rs, rgm, rgs = O.stats_finalize(O.dataset_stats(feats))
    assert (np.abs(sub.mean_per_bin.numpy() - rm) <= 1e-4 * (np.abs(rm) + rs)).all()
    keep = np.arange(128) != 3
    np.testing.assert_allclose(sub.std_per_bin.numpy()[keep], rs[keep], rtol=1e-4)


# ---------------------------------------------------------------------------------------------
# N1: the reference-actual recipe (MelSpectrogram + AmplitudeToDB + per-clip normalisation)
# ---------------------------------------------------------------------------------------------
def test_reference_actual_frontend_vs_golden(b2, O, golden):
    g = golden("reference_actual.npz")
    w = short_clip(44100, seed=77)
    cfg = dict(sample_rate=44100, n_mels=128, bc_mixing=False, normalize=True, target_mean=0.0, target_std=0.5,
               frontend="melspectrogram")
    pre = b2.create_preprocessor("ast", cfg, "/tmp/unused")
    a = pre.preprocess(w, 44100)
    assert a.device.type == "cpu" and tuple(a.shape) == (1, 128, 276)
    assert np.abs(a.numpy() - g["ast_44k"]).max() < 5e-4              # output of the reference's own ASTPreprocessor
    assert abs(float(a.mean())) < 1e-4 and abs(float(a.std()) - 0.5) < 1e-4
    w2 = short_clip(22050, seed=78)                                    # 22.05 kHz clip: resampled to 44.1 kHz first
    pre2 = b2.create_preprocessor("ast", dict(cfg, extra_rates=[22050]), "/tmp/unused")
    b = pre2.preprocess(w2, 22050)
    assert np.abs(b.numpy() - g["ast_22k"]).max() < 5e-4
    nn = b2.create_preprocessor("ast", dict(cfg, normalize=False), "/tmp/unused").preprocess(w, 44100)
    assert np.abs(nn.numpy() - g["ast_44k_nonorm"]).max() < 5e-3      # dB units
    assert float(nn.max() - nn.min()) <= 80.0 + 1e-3                   # top_db clamp
    fb = b2.melspectrogram(w, 44100, 128, 1024, 160, log_scale=True)   # fallback frontend, win_length = n_fft
    assert tuple(fb.shape) == (1, 128, 276) and np.abs(fb.numpy() - g["fallback_melspec"]).max() < 5e-3
    ref = O.reference_ast_preprocess(w.numpy(), 44100)
    assert np.abs(a.numpy() - ref).max() < 5e-4


def test_reference_actual_batch_and_masks(b2, O):
    pre = b2.ASTPreprocessor(b2.PreprocessingConfig(sample_rate=44100, n_mels=128, frontend="melspectrogram"))
    clips = config1_clips(3, length=66150)
    random.seed(9)
    masks = pre.draw_specaugment_masks(3, 414, 192, 48)
    out, nfr = pre.preprocess_batch(torch.cat(clips, 0), 44100, masks=masks)
    assert tuple(out.shape) == (3, 1, 128, 414) and nfr.cpu().tolist() == [414] * 3
    random.seed(9)
    for i, c in enumerate(clips):
        want = pre.apply_specaugment(pre.preprocess(c, 44100), 192, 48)
        assert torch.equal(out[i].cpu(), want), i
        ref = O.reference_ast_preprocess(c.numpy(), 44100)
        ref = O.apply_mask_intervals(ref, masks[i].tolist(), layout="ft")
        assert np.abs(out[i].cpu().numpy() - ref).max() < 5e-4
    live = pytest.importorskip("torchaudio")
    import torchaudio.transforms as T
    w = clips[0]
    db = T.AmplitudeToDB(top_db=80)(T.MelSpectrogram(44100, n_fft=1024, hop_length=160, win_length=400, n_mels=128, power=2.0)(w))
    mine = b2.ASTPreprocessor(b2.PreprocessingConfig(sample_rate=44100, n_mels=128, frontend="melspectrogram",
                                                     normalize=False)).preprocess(w, 44100)
    assert np.abs(mine.numpy() - db.numpy()).max() < 5e-3


def _frontends_per_launch_form(b2, monkeypatch, flags, **kw):
    """B200FBANK_PERSIST is read once, when a plan is created: one frontend per launch form."""
    fes = {}
    for flag in flags:
        monkeypatch.setenv("B200FBANK_PERSIST", flag)
        fes[flag] = b2.FbankFrontend(**kw)
    monkeypatch.delenv("B200FBANK_PERSIST")
    return fes


@pytest.mark.parametrize("rate,seconds,frames", [(44100, 2.3, 1024), (48000, 1.1, 256), (22050, 0.9, 128), (16000, 1.5, 512)])
def test_persistent_launch_is_bit_identical_to_one_cta_per_item(b2, O, monkeypatch, rate, seconds, frames):
    """Dense batches larger than one wave run as ONE CTA per SM that walks its items with the resampler / frame pipeline
    carried across clip (and segment) boundaries; B200FBANK_PERSIST=0 forces one CTA per item.  Same arithmetic per
    item either way: features, frame counts and masks must be bit-identical, and match the oracle."""
    B = 333                                            # > 2 waves of 148; 1024-frame outputs are split into segments
    n = int(rate * seconds) + 5
    g = torch.Generator().manual_seed(rate + frames)
    wav = (torch.rand((B, n), generator=g) * 2 - 1).cuda()
    table = (22050, 44100, 48000, 16000)
    fes = _frontends_per_launch_form(b2, monkeypatch, ("1", "0"), orig_rates=table, **b2.AST_FBANK_KWARGS)
    rid = torch.full((B,), table.index(rate), dtype=torch.int32)
    random.seed(5)
    masks = b2.specaugment.draw_masks(B, frames, 128, 48, 24, variant="reference")
    kw = dict(out_frames=frames, rate_ids=rid, masks=masks, mean=AST_MEAN, std=AST_STD)
    out_p, nfr_p = fes["1"](wav, **kw)
    out_1, nfr_1 = fes["0"](wav, **kw)
    torch.cuda.synchronize()
    assert torch.equal(nfr_p, nfr_1) and torch.equal(out_p, out_1)
    raw, _ = fes["1"](wav, out_frames=frames, rate_ids=rid)         # un-normalised, un-masked: the oracle's units
    for i in (0, 147, 148, 332):
        ref = O.kaldi_fbank(O.resample(wav[i].cpu().numpy(), rate, 16000), O.ast_fbank_options())
        m = min(ref.shape[0], frames)
        assert m == int(nfr_p[i])
        assert_logmel_close(raw[i, :m].cpu().numpy(), ref[:m], LOGMEL_TOL, f"persistent clip {i} @ {rate}")
    # the stats pass takes the same two launch forms (float64 atomics: equal up to summation order)
    sums = []
    for flag in ("1", "0"):
        ds = b2.DatasetStats(fes[flag], max_frames=frames)
        ds.update(wav, rate_ids=rid)
        sums.append(ds.sums.cpu().numpy())
    assert sums[0][-1] == sums[1][-1] == float(nfr_p.sum().item())
    np.testing.assert_allclose(sums[0], sums[1], rtol=1e-11, atol=1e-6)


def test_dynamic_persistent_launch_on_ragged_batches(b2, O, monkeypatch):
    """Ragged batches larger than one wave: one CTA per SM, items claimed from a global counter and handed from the
    resampler warps to the frame warps through a shared-memory queue (B200FBANK_PERSIST=2, the default for ragged
    input) == one CTA per item (=0), bit for bit, including empty and too-short clips in the middle of the batch."""
    B = 700
    table = (22050, 44100, 48000, 16000)
    g = torch.Generator().manual_seed(404)
    rid = torch.randint(0, 4, (B,), generator=g)
    dur = 0.2 + 0.9 * torch.rand(B, generator=g)
    lens = (dur * torch.tensor(table)[rid]).long()
    lens[5] = 0                                           # an empty clip
    lens[300:310] = 100                                   # shorter than one 25 ms window: zero frames, pad rows only
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)])
    flat = (torch.rand(int(offsets[-1]), generator=g) * 2 - 1).cuda()
    fes = _frontends_per_launch_form(b2, monkeypatch, ("2", "0"), orig_rates=table, **b2.AST_FBANK_KWARGS)
    fe = b2.FbankFrontend(orig_rates=table, **b2.AST_FBANK_KWARGS)
    random.seed(9)
    masks = b2.specaugment.draw_masks(B, 128, 128, 32, 16, variant="reference")
    kw = dict(out_frames=128, offsets=offsets, rate_ids=rid.int(), masks=masks, mean=AST_MEAN, std=AST_STD)
    outs = {}
    for flag in ("2", "0", "2"):
        out, nfr = fes[flag](flat, **kw)
        torch.cuda.synchronize()
        if flag in outs:
            assert torch.equal(out, outs[flag][0])            # the claim order varies from run to run, the result does not
        outs[flag] = (out, nfr)
    assert torch.equal(outs["2"][1], outs["0"][1]) and torch.equal(outs["2"][0], outs["0"][0])
    nfr = outs["2"][1].cpu()
    assert int(nfr[5]) == 0 and bool((nfr[300:310] == 0).all())
    want = [min(128, fe.num_frames(int(n), int(r))) for n, r in zip(lens, rid)]
    assert nfr.tolist() == want
    raw, _ = fe(flat, out_frames=128, offsets=offsets, rate_ids=rid.int())
    for i in (0, 1, 299, 311, 699):
        w = flat[int(offsets[i]):int(offsets[i + 1])].cpu().numpy()
        ref = O.kaldi_fbank(O.resample(w, table[int(rid[i])], 16000), O.ast_fbank_options())
        m = min(ref.shape[0], 128)
        assert m == int(nfr[i])
        assert_logmel_close(raw[i, :m].cpu().numpy(), ref[:m], LOGMEL_TOL, f"ragged clip {i}")


def test_persistent_launch_with_a_non_ast_filterbank(b2, monkeypatch):
    """The general (non-unrolled mel) instantiation of the warp-specialised kernel in its persistent form: 40 povey-windowed
    bins at 16 kHz (kaldi defaults), no resampling: == one CTA per item, and == the single-clip kaldi.fbank drop-in."""
    B, n = 400, 16000 + 123
    g = torch.Generator().manual_seed(77)
    wav = (torch.rand((B, n), generator=g) * 2 - 1).cuda()
    fes = _frontends_per_launch_form(b2, monkeypatch, ("1", "0"), orig_rates=(16000,), num_mel_bins=40, window_type="povey",
                                     sample_frequency=16000.0)
    outs = []
    for flag in ("1", "0"):
        outs.append(fes[flag](wav, out_frames=128))
    torch.cuda.synchronize()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    m = int(outs[0][1][0])
    assert m == 99
    for i in (0, 399):
        one = b2.fbank(wav[i:i + 1], num_mel_bins=40)
        assert torch.equal(outs[0][0][i, :m], one)


def test_process_host_matches_device_call(b2):
    """Host buffers in / host buffers out (the e2e form bench.py times): the returned CPU tensor is complete when the
    call returns (the side streams are synchronised) and equals the device-resident call bit for bit."""
    fe = b2.FbankFrontend(orig_rates=(44100,), device="cuda:0", **b2.AST_FBANK_KWARGS)
    clips = torch.cat(config1_clips(37, length=44100), 0)
    h_wav = clips.pin_memory()
    for layout in ("btf", "bft"):
        want, _ = fe(clips.cuda(), out_frames=128, mean=AST_MEAN, std=AST_STD, layout=layout)
        for chunk in (8, 64):
            got = fe.process_host(h_wav, 128, chunk_clips=chunk, mean=AST_MEAN, std=AST_STD, layout=layout)
            assert got.device.type == "cpu"
            assert torch.equal(got, want.cpu()), (layout, chunk)


@pytest.mark.parametrize("seed", [31, 32])
def test_us8k_shaped_dynamic_launch_is_race_free(b2, monkeypatch, seed):
    """Regression (round 2): pass slots without live frames used to release their ring rows without waiting for the R
    warps, so their arrivals could be counted in the previous phase of the slot's `empty` barrier; with US8K-shaped
    batches (every second segment is padding only) and time masks the resampler then overwrote rows still being read:
    results differed from one CTA per item and the seed-32 batch deadlocked.  Both launch forms must agree bit for bit."""
    B, table = 4096, (22050, 44100, 48000)
    g = torch.Generator().manual_seed(seed)
    rid = torch.randint(0, 3, (B,), generator=g)
    lens = ((1.0 + 3.0 * torch.rand(B, generator=g)) * torch.tensor(table)[rid]).long()
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)]).cuda()
    flat = torch.rand(int(offsets[-1]), device="cuda", generator=torch.Generator(device="cuda").manual_seed(seed + 46)) * 2 - 1
    random.seed(77)
    masks = b2.specaugment.draw_masks(B, 1024, 128, 192, 48).cuda()
    fes = _frontends_per_launch_form(b2, monkeypatch, ("2", "0"), orig_rates=table, **b2.AST_FBANK_KWARGS)
    kw = dict(out_frames=1024, offsets=offsets, rate_ids=rid.int().cuda(), masks=masks, return_n_frames=False)
    ref = fes["0"](flat, **kw)[0]
    for _ in range(3):
        got = fes["2"](flat, **kw)[0]
        torch.cuda.synchronize()
        assert torch.equal(got, ref)
