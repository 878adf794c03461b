"""CUDA path vs oracle / golden fixtures / live torchaudio.  Runs on the B200 box (-m gpu)."""
import numpy as np
import pytest
import torch

from inputs import config1_clips, short_clip, us8k_small_clips
from make_golden_variants import KALDI_VARIANTS
from parity import LOGMEL_TOL, assert_logmel_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b2():
    import dl_sound_classification_b200 as m
    assert torch.cuda.is_available()
    return m


@pytest.fixture(scope="module")
def O():
    from oracle import fbank_oracle
    return fbank_oracle


def test_resample_only_vs_oracle_and_golden(b2, O, golden):
    g = golden("config1.npz")
    clips = config1_clips(2)
    fe = b2.FbankFrontend(orig_rates=(44100,), **b2.AST_FBANK_KWARGS)
    y = fe.resample(torch.cat(clips, 0).cuda()).cpu().numpy()
    assert y.shape == (2, 80000)
    np.testing.assert_allclose(y[0][:2000], g["res0_head"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(y[0][-2000:], g["res0_tail"], rtol=0, atol=2e-6)
    ref = O.resample(clips[1][0].numpy(), 44100, 16000, dtype=np.float64)
    assert np.abs(y[1] - ref).max() < 2e-6


def test_config1_vs_golden_and_oracle(b2, O, golden):
    g = golden("config1.npz")
    clips = config1_clips(40)
    fe = b2.FbankFrontend(orig_rates=(44100,), **b2.AST_FBANK_KWARGS)
    out, nfr = fe(torch.cat(clips, 0).cuda(), out_frames=512)
    got = out.cpu().numpy()
    assert got.shape == (40, 512, 128)
    assert (nfr.cpu().numpy() == 498).all()
    assert (got[:, 498:, :] == 0).all()                          # H9 pad rows, un-normalised
    for j, i in enumerate(g["full_idx"]):
        assert_logmel_close(got[int(i), :498], g["full"][j], LOGMEL_TOL, f"clip {i} vs torchaudio golden")
    assert_logmel_close(got[:, g["frame_idx"], :], g["frames"], LOGMEL_TOL, "frames vs torchaudio golden")
    np.testing.assert_allclose(got[:, :498].astype(np.float64).sum(1), g["colsum"], rtol=1e-4, atol=2e-2)
    # fp64 truth: our float32 path must sit no further from it than the bar
    w = clips[7][0].numpy()
    f64 = O.kaldi_fbank(O.resample(w, 44100, 16000, dtype=np.float64), O.ast_fbank_options(), dtype=np.float64)
    assert_logmel_close(got[7, :498], f64, LOGMEL_TOL, "clip 7 vs fp64 truth")
    assert (got[:, :498, 3] == np.float32(O.LOG_FLT_EPSILON)).all()       # empty mel filter #3


def test_normalisation_masks_and_layout(b2, O):
    clips = config1_clips(3, length=50000)
    fe = b2.FbankFrontend(orig_rates=(44100,), **b2.AST_FBANK_KWARGS)
    wav = torch.cat(clips, 0).cuda()
    masks = torch.tensor([[10, 30, 5, 20], [0, 0, 100, 28], [100, 28, 0, 0]], dtype=torch.int32)
    mean = torch.linspace(-7, -6, 128)
    std = torch.linspace(4, 5, 128)
    a, _ = fe(wav, out_frames=128, mean=mean, std=std, masks=masks, layout="btf")
    b, _ = fe(wav, out_frames=128, mean=mean, std=std, masks=masks, layout="bft")
    assert b.shape == (3, 1, 128, 128)
    assert torch.equal(a, b[:, 0].transpose(1, 2))
    for i, c in enumerate(clips):
        ref, m = O.ast_frontend(c[0].numpy(), 44100, target_frames=128, mean=mean.numpy(), std=std.numpy(),
                                mask=masks[i].tolist())
        got = a[i].cpu().numpy()
        assert (got == 0).sum() == (ref == 0).sum()                       # masked cells bit-exact
        assert ((got == 0) == (ref == 0)).all()
        assert np.abs(got - ref).max() < LOGMEL_TOL / 4.0
    s, _ = fe(wav, out_frames=100, mean=-6.6268, std=5.0613)              # crop + scalar stats
    ref, m = O.ast_frontend(clips[0][0].numpy(), 44100, target_frames=100, mean=-6.6268, std=5.0613)
    assert np.abs(s[0].cpu().numpy() - ref).max() < LOGMEL_TOL / 5.0


def test_us8k_ragged_mixed_rate(b2, O, golden):
    g = golden("us8k_small.npz")
    clips, rates = us8k_small_clips(9)
    table = (22050, 44100, 48000)
    fe = b2.FbankFrontend(orig_rates=table, **b2.AST_FBANK_KWARGS)
    flat = torch.cat([c[0] for c in clips]).cuda()
    lens = torch.tensor([c.shape[1] for c in clips])
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)])
    rid = torch.tensor([table.index(r) for r in rates], dtype=torch.int32)
    out, nfr = fe(flat, out_frames=1024, offsets=offsets, rate_ids=rid)
    got = out.cpu().numpy()
    assert (nfr.cpu().numpy() == g["n_frames"]).all()
    for i in range(9):
        m = int(g["n_frames"][i])
        assert_logmel_close(got[i, :m], g["feats"][i, :m], LOGMEL_TOL, f"us8k clip {i} rate {rates[i]}")
        assert (got[i, m:] == 0).all()


@pytest.mark.parametrize("name", sorted(KALDI_VARIANTS))
def test_kaldi_fbank_variants(b2, golden, name):
    g = golden("kaldi_variants.npz")
    kw = KALDI_VARIANTS[name]
    w = short_clip(8000 * 3)
    if name == "no_pow2":
        with pytest.raises(NotImplementedError):
            b2.fbank(w.cuda(), **kw)
        return
    got = b2.fbank(w.cuda(), **kw)
    assert got.is_cuda and tuple(got.shape) == g[name].shape
    got = got.cpu().numpy()
    if kw.get("use_log_fbank", True):
        assert_logmel_close(got, g[name], LOGMEL_TOL, name)
    else:
        np.testing.assert_allclose(got, g[name], rtol=1e-3, atol=1e-5)
    cpu = b2.fbank(w, **kw)                                                # CPU tensor in -> CPU tensor out
    assert cpu.device.type == "cpu" and np.array_equal(cpu.numpy(), got)


def test_kaldi_fbank_channel_and_errors(b2, golden):
    g = golden("kaldi_variants.npz")
    w = short_clip(8000 * 3)
    two = torch.cat([w, -0.5 * w.flip(1)], 0).cuda()
    assert_logmel_close(b2.fbank(two, channel=1, num_mel_bins=40).cpu().numpy(), g["channel1"], LOGMEL_TOL, "channel1")
    with pytest.raises(AssertionError):
        b2.fbank(short_clip(300).cuda())
    with pytest.raises(AssertionError):
        b2.fbank(w.cuda(), channel=3)
    with pytest.raises(NotImplementedError):
        b2.fbank(w.cuda(), dither=1.0)
    assert b2.fbank(w.cuda(), min_duration=10.0).numel() == 0
    sil = b2.fbank(torch.zeros(1, 16000).cuda(), **{k: v for k, v in b2.AST_FBANK_KWARGS.items()})
    assert (sil.cpu().numpy() == np.float32(-15.942385)).all()


def test_live_torchaudio_if_available(b2):
    ta = pytest.importorskip("torchaudio")
    import torchaudio.compliance.kaldi as tk
    w = short_clip(16000 * 2, seed=99)
    for kw in (dict(), dict(num_mel_bins=80, window_type="hamming"), dict(b2.AST_FBANK_KWARGS)):
        ref = tk.fbank(w, **kw).numpy()
        assert_logmel_close(b2.fbank(w.cuda(), **kw).cpu().numpy(), ref, LOGMEL_TOL, str(kw))


def test_stats_pass(b2, O, golden):
    g = golden("config1.npz")
    clips = config1_clips(40)
    fe = b2.FbankFrontend(orig_rates=(44100,), **b2.AST_FBANK_KWARGS)
    sums = torch.zeros(257, dtype=torch.float64, device="cuda")
    wav = torch.cat(clips, 0).cuda()
    fe.accumulate_stats(wav[:25], sums, max_frames=512)
    fe.accumulate_stats(wav[25:], sums, max_frames=512)
    s = sums.cpu().numpy()
    assert s[256] == 40 * 498
    ref = g["stats_sums"]
    mean_b, std_b, gm, gs = O.stats_finalize(s)
    rm, rs, rgm, rgs = O.stats_finalize(ref)
    # rel 1e-4 on the normalisation statistics, relative to the scale they normalise by
    assert (np.abs(mean_b - rm) <= 1e-4 * (np.abs(rm) + rs)).all()
    keep = np.arange(128) != 3                                             # filter #3 is constant: std == 0
    np.testing.assert_allclose(std_b[keep], rs[keep], rtol=1e-4)
    assert abs(gm - rgm) <= 1e-4 * abs(rgm) and abs(gs - rgs) <= 1e-4 * rgs


@pytest.mark.parametrize("seed,table", [(1, (44100,)), (2, (44100,)), (3, (22050, 44100, 48000)), (4, (16000, 44100))])
def test_tuned_kernel_agrees_with_the_generic_kernel_on_random_ragged_batches(b2, seed, table, monkeypatch):
    """Fuzz of the persistent kernel's edge handling (clip-edge chunks by bulk copy + hand-assembled lines, pad-row fast
    path, crops, masks reaching into pad rows, clips shorter than a window) against the correctness-first generic kernel
    of the same plan options: same frame counts, same zero pattern, log-mel within the bar."""
    g = torch.Generator().manual_seed(100 + seed)
    B = 96
    rid = torch.randint(0, len(table), (B,), generator=g)
    rates = torch.tensor(table)[rid]
    dur = torch.exp(torch.rand(B, generator=g) * 5.0 - 4.6)                      # 0.01 .. 1.5 s, log-uniform
    lens = (dur * rates).long().clamp(min=1)
    lens[0] = 399 * int(rates[0]) // 16000                                       # shorter than one analysis window
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)])
    flat = (torch.rand(int(offsets[-1]), generator=g) * 2 - 1).cuda()
    fe = b2.FbankFrontend(orig_rates=table, **b2.AST_FBANK_KWARGS)
    monkeypatch.setenv("B200FBANK_KERNEL", "generic")
    ref_fe = b2.FbankFrontend(orig_rates=table, **b2.AST_FBANK_KWARGS)
    monkeypatch.delenv("B200FBANK_KERNEL")
    for T, layout in ((64, "btf"), (100, "bft"), (130, "btf"), (152, "bft")):
        masks = torch.stack([torch.randint(0, T, (B,), generator=g), torch.randint(0, T // 3, (B,), generator=g),
                             torch.randint(0, 100, (B,), generator=g), torch.randint(0, 28, (B,), generator=g)], 1).int()
        masks[:, 1] = torch.minimum(masks[:, 1], T - masks[:, 0])
        kw = dict(offsets=offsets, rate_ids=rid.int(), masks=masks, mean=-4.27, std=4.57, layout=layout)
        got, n1 = fe(flat, T, **kw)
        want, n2 = ref_fe(flat, T, **kw)
        assert torch.equal(n1, n2) and int(n1[0]) == 0
        a, b_ = got.cpu().numpy(), want.cpu().numpy()
        assert ((a == 0) == (b_ == 0)).all(), (T, layout)
        if layout == "bft":
            a, b_ = a[:, 0].transpose(0, 2, 1), b_[:, 0].transpose(0, 2, 1)
        # back to log-mel units for the parity bar: y = (x - mean) / (2 std); pad rows and masked cells (0.0 in both) drop out
        live = a != 0
        assert_logmel_close(np.where(live, a * (2 * 4.57) - 4.27, 0.0), np.where(live, b_ * (2 * 4.57) - 4.27, 0.0), LOGMEL_TOL,
                            f"fuzz seed {seed} T {T} {layout}", 1e-3)
