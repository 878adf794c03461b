"""GPU tests of the per-clip passes: per-clip normalisation (reference src/datasets/preprocessing.py:1030-1037), whole-clip
DC removal, Mixup beyond 65 535 clips."""
import numpy as np
import pytest
import torch

from inputs import config1_clips, us8k_small_clips

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b2():
    import dl_sound_classification_b200 as m
    return m


def _reference_per_clip(x, nfr, masks, target_mean=0.0, target_std=0.5):
    """The reference's arithmetic on one clip at a time (torch CPU): log_mel.mean(), log_mel.std() (unbiased),
    (x - mean) / std * target_std + target_mean, then the zero-fill of apply_specaugment."""
    out = x.clone()
    for i in range(x.shape[0]):
        m = int(nfr[i])
        v = x[i, :, :, :m]
        if v.numel() >= 2 and float(v.std()) > 0:
            out[i, :, :, :m] = (v - v.mean()) / v.std() * target_std + target_mean
        if masks is not None:
            t0, tl, f0, fl = [int(a) for a in masks[i]]
            out[i, :, :, t0:t0 + tl] = 0
            out[i, :, f0:f0 + fl, :] = 0
    return out


@pytest.mark.parametrize("layout", ["bft", "btf"])
def test_per_clip_normalisation_matches_the_reference_loop(b2, layout):
    clips, rates = us8k_small_clips(9)
    table = (22050, 44100, 48000)
    fe = b2.FbankFrontend(orig_rates=table, **b2.AST_FBANK_KWARGS)
    lens = torch.tensor([c.shape[1] for c in clips])
    offs = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)])
    rid = torch.tensor([table.index(r) for r in rates], dtype=torch.int32)
    flat = torch.cat([c[0] for c in clips])
    masks = torch.tensor([[10 + i, 7 * (i % 3), 5 * i, 4 * (i % 4)] for i in range(9)], dtype=torch.int32)
    raw, nfr = fe(flat, 512, offsets=offs, rate_ids=rid, layout=layout)
    got, nfr2 = fe(flat, 512, offsets=offs, rate_ids=rid, layout=layout, masks=masks, per_clip_norm=True)
    assert torch.equal(nfr, nfr2)
    raw, got = raw.cpu(), got.cpu()
    if layout == "btf":
        raw, got = raw.transpose(1, 2).unsqueeze(1), got.transpose(1, 2).unsqueeze(1)
    want = _reference_per_clip(raw, nfr.cpu(), masks)
    # float64 statistics on the device vs torch's float32 reductions: a few ulp of the normalised values
    assert float((got - want).abs().max()) < 2e-6
    for i in range(9):
        m = int(nfr[i])
        assert torch.count_nonzero(got[i, :, :, m:]) == 0                   # pad rows stay 0.0
        live = got[i, :, :, :m]
        t0, tl, f0, fl = [int(a) for a in masks[i]]
        keep = torch.ones_like(live, dtype=torch.bool)
        keep[:, :, t0:t0 + tl] = False
        keep[:, f0:f0 + fl, :] = False
        sel = want[i, :, :, :m][keep]
        assert torch.allclose(live[keep], sel, atol=2e-6)


def test_preprocessor_without_dataset_statistics_runs_no_eager_loop(b2):
    pre = b2.ASTPreprocessor(b2.PreprocessingConfig(sample_rate=44100, n_mels=128, normalize=True, frontend="kaldi_fbank"))
    assert not hasattr(pre, "_per_clip_normalize")
    clips = config1_clips(3, length=88200)
    out, nfr = pre.preprocess_batch(torch.cat(clips, 0), 44100)
    for i in range(3):
        v = out[i, :, :, :int(nfr[i])]
        assert abs(float(v.mean())) < 1e-5 and abs(float(v.std()) - 0.5) < 1e-5


def test_whole_clip_dc_removal(b2):
    g = torch.Generator().manual_seed(3)
    clips = [torch.rand(1, n, generator=g) * 2 - 1 + 0.3 for n in (44100, 50001, 61234)]
    fe = b2.FbankFrontend(orig_rates=(44100,), **b2.AST_FBANK_KWARGS)
    lens = torch.tensor([c.shape[1] for c in clips])
    offs = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)])
    flat = torch.cat([c[0] for c in clips])
    got, _ = fe(flat, 160, offsets=offs, remove_clip_mean=True)
    centred = torch.cat([(c - c.mean())[0] for c in clips])            # the AST recipe: waveform - waveform.mean()
    want, _ = fe(centred, 160, offsets=offs)
    # the means differ by float32 rounding of the reduction only (1e-4 on the quietest log-mel cells)
    assert float((got - want).abs().max()) < 5e-4 and float((got - want).abs().mean()) < 2e-6
    plain, _ = fe(flat, 160, offsets=offs)
    assert float((plain - want).abs().max()) > 1e-4                    # ... and the flag does change the features
    # dense batches take the same path
    dense = torch.stack([c[0, :44100] for c in clips])
    a, _ = fe(dense, 128, remove_clip_mean=True)
    b, _ = fe(dense - dense.mean(dim=1, keepdim=True), 128)
    assert float((a - b).abs().max()) < 5e-4


def test_mixup_more_than_65535_clips(b2):
    B, E = 70000, 96
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, E, generator=g).cuda()
    bank = torch.randn(512, E, generator=g).cuda()
    partner = torch.randint(-1, 512, (B,), generator=g).int()
    lam = torch.rand(B, generator=g)
    out = b2.mixup_batch(x, bank, b2.MixupPlan(partner, lam).to("cuda"))
    l = lam.cuda()[:, None]
    want = torch.where(partner.cuda()[:, None] >= 0, l * x + (1 - l) * bank[partner.clamp(min=0).long().cuda()], x)
    assert torch.equal(out, want)


def test_pcm16_input_is_bit_identical_to_the_float_contract(b2):
    """int16 PCM + per-clip divisor widened on the device == torchaudio.load's float32 (/32768) followed by the peak
    normalisation of scripts/prepare_esc50.py:94-101; the end-to-end call gives the same features either way."""
    g = torch.Generator().manual_seed(9)
    pcm = torch.randint(-20000, 20000, (5, 44100), generator=g, dtype=torch.int32).to(torch.int16)
    wave = pcm.to(torch.float32) / 32768.0                                # what torchaudio.load returns
    peak = wave.abs().amax(dim=1, keepdim=True)
    norm = wave / peak                                                    # prepare_esc50.py: wave / max|wave|
    fe = b2.FbankFrontend(orig_rates=(44100,), **b2.AST_FBANK_KWARGS)
    assert torch.equal(fe.pcm16_to_float(pcm.cuda()).cpu(), wave)
    div = pcm.to(torch.int32).abs().amax(dim=1).to(torch.float32)
    assert torch.equal(fe.pcm16_to_float(pcm.cuda(), div).cpu(), norm)
    a = fe.process_host(norm.pin_memory(), 128, chunk_clips=2, mean=-6.6, std=5.0)
    b = fe.process_host(pcm.pin_memory(), 128, chunk_clips=2, pcm_divisor=div, mean=-6.6, std=5.0)
    assert torch.equal(a, b)
    # ragged form
    lens = torch.tensor([1000, 7, 44100 - 3, 12345])
    offs = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)])
    flat = torch.cat([pcm[i, :int(n)] for i, n in enumerate(lens)])
    got = fe.pcm16_to_float(flat.cuda(), div[:4], offsets=offs).cpu()
    want = torch.cat([pcm[i, :int(n)].to(torch.float32) / div[i] for i, n in enumerate(lens)])
    assert torch.equal(got, want)
