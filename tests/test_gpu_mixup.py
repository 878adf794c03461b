"""GPU tests of the Mixup row: the CUDA kernel behind b200fbank_mixup against the golden
vectors made by the reference's own MixupAugmentation and against the oracle, bit for bit."""
import random

import numpy as np
import pytest
import torch

from test_mixup_host import B, C, N, golden, mixup_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b2():
    import dl_sound_classification_b200 as m
    assert torch.cuda.is_available()
    return m


def test_mixup_batch_matches_reference_golden_bit_exact(b2):
    g = golden()
    x, labels, bank, bank_labels = mixup_inputs()
    random.seed(31337)
    torch.manual_seed(31337)
    plan = b2.draw_mixup_plan(B, N, alpha=0.5, prob=0.5)
    out = b2.mixup_batch(x.cuda(), bank.cuda(), plan)
    assert torch.equal(out.cpu(), torch.from_numpy(g["out"]))
    soft = b2.mixup_labels(labels.cuda(), bank_labels.cuda(), plan, C)
    assert torch.equal(soft.cpu(), torch.from_numpy(g["soft"]))
    # in place, and on rows that are not 16-byte aligned (odd row length)
    xc = x.cuda().clone()
    assert b2.mixup_batch(xc, bank.cuda(), plan, out=xc) is xc and torch.equal(xc.cpu(), torch.from_numpy(g["out"]))
    xo, bo = x.flatten(1)[:, :637].contiguous(), bank.flatten(1)[:, :637].contiguous()
    got = b2.mixup_batch(xo.cuda(), bo.cuda(), plan).cpu()
    assert torch.equal(got, torch.from_numpy(g["out"]).flatten(1)[:, :637])


def test_mixup_augmentation_mirror_class(b2):
    from oracle import fbank_oracle as O
    x, labels, bank, bank_labels = mixup_inputs()
    random.seed(1)
    torch.manual_seed(1)
    aug = b2.MixupAugmentation(alpha=0.5, prob=1.0)
    mixed, soft = aug(x[0], bank[5], int(labels[0]), int(bank_labels[5]), C)
    torch.manual_seed(1)
    lam = float(torch.distributions.Beta(0.5, 0.5).sample())
    assert mixed.device.type == "cpu" and np.array_equal(mixed.numpy(), O.mixup(x[0].numpy(), bank[5].numpy(), lam))
    assert np.array_equal(soft.numpy(), O.mixup_soft_labels(int(labels[0]), int(bank_labels[5]), lam, C))
    s0 = x[0]
    same, one_hot = b2.MixupAugmentation(alpha=0.5, prob=0.0)(s0, bank[5], 7, 9, C)
    assert same is s0 and float(one_hot[7]) == 1.0 and float(one_hot.sum()) == 1.0   # returned untouched (preprocessing.py:950-953)


def test_mixup_full_size_properties(b2):
    """BASELINE.json config sizes: 1024 x (1, 128, 512) against a 2000-clip bank: unmixed rows are copies, lam = 1 is
    the identity, lam = 0 returns the partner, and the batch equals the per-sample oracle on a few rows."""
    from oracle import fbank_oracle as O
    Bf, Nf = 1024, 2000
    gen = torch.Generator(device="cuda").manual_seed(8)
    x = torch.randn((Bf, 1, 128, 512), generator=gen, device="cuda")
    bank = torch.randn((Nf, 1, 128, 512), generator=gen, device="cuda")
    random.seed(3)
    torch.manual_seed(3)
    plan = b2.draw_mixup_plan(Bf, Nf, alpha=0.5, prob=0.5)
    plan.lam[10], plan.partner[10] = 1.0, 17
    plan.lam[11], plan.partner[11] = 0.0, 18
    out = b2.mixup_batch(x, bank, plan)
    un = (plan.partner < 0).nonzero().flatten()
    assert 600 < len(un) < 900 and torch.equal(out[un.cuda()], x[un.cuda()])
    assert torch.equal(out[10], x[10]) and torch.equal(out[11], bank[18] + 0.0 * x[11])
    for i in (plan.partner >= 0).nonzero().flatten()[:5].tolist():
        ref = O.mixup(x[i].cpu().numpy(), bank[int(plan.partner[i])].cpu().numpy(), float(plan.lam[i]))
        assert np.array_equal(out[i].cpu().numpy(), ref), i
    with pytest.raises(IndexError):
        b2.mixup_batch(x[:2], bank[:5], b2.MixupPlan(torch.tensor([1, 9], dtype=torch.int32), torch.ones(2)))


@pytest.mark.parametrize("layout", ["btf", "bft"])
def test_mixup_fused_into_the_fbank_epilogue_is_bit_identical(b2, layout):
    """b200fbank_execute_mixup == b200fbank_execute followed by b200fbank_mixup, bit for bit: dense batch with
    normalisation, SpecAugment masks and pad rows (both layouts), partner -1 rows untouched."""
    from inputs import config1_clips
    Bq, T = 12, 320                                                        # 2 s clips -> 198 real frames + pad rows
    wav = torch.cat(config1_clips(Bq, seed=7, length=88200), 0).cuda()
    fe = b2.FbankFrontend(orig_rates=(44100,), **b2.AST_FBANK_KWARGS)
    random.seed(3)
    torch.manual_seed(3)
    masks = b2.specaugment.draw_masks(Bq, T, 128, 48, 24, variant="reference")
    bank_wav = torch.cat(config1_clips(5, seed=70, length=88200), 0).cuda()
    bank, _ = fe(bank_wav, out_frames=T, mean=-4.27, std=4.57, layout=layout)
    plan = b2.draw_mixup_plan(Bq, 5, alpha=0.5, prob=0.9)
    assert int((plan.partner >= 0).sum()) >= 2 and int((plan.partner < 0).sum()) >= 2
    plain, nfr = fe(wav, out_frames=T, masks=masks, mean=-4.27, std=4.57, layout=layout)
    want = b2.mixup_batch(plain, bank, plan)
    got, nfr2 = fe(wav, out_frames=T, masks=masks, mean=-4.27, std=4.57, layout=layout, mixup=(bank, plan))
    assert torch.equal(nfr, nfr2) and int(nfr[0]) == 198
    assert torch.equal(got, want)
    keep = plan.partner < 0
    assert torch.equal(got[keep.cuda()], plain[keep.cuda()])
    # per-clip statistics: the mix follows the normalisation pass as its own launch, same bits
    a, _ = fe(wav, out_frames=T, masks=masks, layout=layout, per_clip_norm=True)
    b_, _ = fe(wav, out_frames=T, masks=masks, layout=layout, per_clip_norm=True, mixup=(bank, plan))
    assert torch.equal(b_, b2.mixup_batch(a, bank, plan))
    with pytest.raises(ValueError):
        fe(wav, out_frames=T, mixup=(bank[:, ..., :-1].contiguous(), plan))


def test_mixup_fused_ragged_multi_rate_batch(b2):
    """The dynamic persistent launch (ragged clips, three rates) with the fused mix."""
    from inputs import us8k_small_clips
    clips, rates = us8k_small_clips(9)
    table = (22050, 44100, 48000)
    lens = torch.tensor([c.shape[1] for c in clips])
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)])
    flat = torch.cat([c[0] for c in clips]).cuda()
    rid = torch.tensor([table.index(r) for r in rates], dtype=torch.int32)
    fe = b2.FbankFrontend(orig_rates=table, **b2.AST_FBANK_KWARGS)
    T = 416
    bank = torch.randn(4, 1, 128, T, generator=torch.Generator().manual_seed(2)).cuda()
    random.seed(9)
    torch.manual_seed(9)
    plan = b2.draw_mixup_plan(9, 4, alpha=0.5, prob=1.0)
    plain, _ = fe(flat, out_frames=T, offsets=offsets, rate_ids=rid, mean=-4.27, std=4.57, layout="bft")
    got, _ = fe(flat, out_frames=T, offsets=offsets, rate_ids=rid, mean=-4.27, std=4.57, layout="bft", mixup=(bank, plan))
    assert torch.equal(got, b2.mixup_batch(plain, bank, plan))
