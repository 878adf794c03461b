"""The integration objects of SURVEY.md section 8b/8f as code: the Lightning-free DataModule mixin
(`on_after_batch_transfer`) driven by a fake trainer, and `PreprocessingCache.batch_preprocess`."""
import random
import types

import pytest
import torch

from inputs import short_clip

pytestmark = pytest.mark.gpu

CFG = dict(sample_rate=44100, n_mels=128, normalize=True, target_mean=0.0, target_std=0.5)


@pytest.fixture(scope="module")
def b2():
    import dl_sound_classification_b200 as m
    assert torch.cuda.is_available()
    return m


class HostDataModule:
    """Stands in for the reference's ESC50DataModule (src/datasets/esc50.py:353-373): only the constructor attributes."""

    def __init__(self, **kw):
        self.sample_rate, self.n_mels, self.num_classes = 44100, 128, 50
        self.time_mask, self.freq_mask, self.enable_mixup, self.mixup_alpha = True, True, True, 0.5
        self.preprocessing_config = dict(CFG)
        self.trainer = types.SimpleNamespace(training=True)
        self.__dict__.update(kw)


def test_datamodule_mixin_equals_the_per_sample_loop(b2, tmp_path):
    class DM(b2.B200DataModuleMixin, HostDataModule):
        pass

    B, N, C = 6, 44100, 50
    wav = torch.cat([short_clip(N, seed=100 + i) for i in range(B)], 0)               # (B, N) CPU
    labels = torch.tensor([3, 7, 7, 0, 49, 21])
    pre = b2.create_preprocessor("ast", dict(CFG), tmp_path)
    bank_wav = torch.cat([short_clip(N, seed=500 + i) for i in range(9)], 0)
    bank, _ = pre.preprocess_batch(bank_wav.cuda(), 44100)                            # (9, 1, 128, 276)
    bank_labels = torch.arange(9) * 5

    dm = DM()
    dm.setup_b200(base_cache_dir=tmp_path)
    dm.set_mixup_bank(bank, bank_labels.cuda())
    random.seed(5)
    torch.manual_seed(5)
    spec, soft = dm.on_after_batch_transfer((wav[:, None, :].cuda(), labels.cuda()), 0)
    assert tuple(spec.shape) == (B, 1, 128, 276) and tuple(soft.shape) == (B, C) and spec.is_cuda

    # the reference's per-sample order (ESC50Dataset._process_ast, esc50.py:246-291), with this package's per-sample mirrors
    random.seed(5)
    torch.manual_seed(5)
    mix = b2.MixupAugmentation(alpha=0.5, prob=0.5)                                   # esc50.py:51
    mixed_any = False
    for i in range(B):
        s = pre.preprocess(wav[i:i + 1], 44100)
        s = pre.apply_specaugment(s, time_mask=192, freq_mask=48)
        if random.random() > 0.5:                                                     # MixupDataset.apply_mixup, esc50.py:64-76
            lab = torch.zeros(C)
            lab[labels[i]] = 1.0
        else:
            other = random.randint(0, bank.shape[0] - 1)
            before = s
            s, lab = mix(s, bank[other].cpu(), int(labels[i]), int(bank_labels[other]), C)
            mixed_any |= s is not before
        assert torch.allclose(spec[i].cpu(), s, atol=1e-6), i
        assert torch.equal(spec[i].cpu() == 0, s == 0), i                              # masks: exactly the drawn intervals
        assert torch.allclose(soft[i].cpu(), lab, atol=1e-7), i
    assert mixed_any

    # evaluation: no augmentation, one-hot labels; spectrogram batches pass through untouched
    dm.trainer.training = False
    spec_e, soft_e = dm.on_after_batch_transfer((wav.cuda(), labels.cuda()), 0)
    assert torch.equal(soft_e.argmax(1).cpu(), labels) and float(soft_e.sum()) == B
    assert torch.allclose(spec_e[2].cpu(), pre.preprocess(wav[2:3], 44100), atol=1e-6)
    same = dm.on_after_batch_transfer((spec_e, soft_e), 0)
    assert same[0] is spec_e


def test_batch_preprocess_matches_per_clip_and_fills_the_cache(b2, tmp_path):
    files = []
    waves = {}
    for i, n in enumerate((44100, 30000, 52000, 44100)):
        p = tmp_path / f"clip{i}.pt"
        waves[p] = short_clip(n, seed=40 + i)
        torch.save({"waveform": waves[p], "label": i}, p)
        files.append(p)
    bad = tmp_path / "broken.pt"
    bad.write_bytes(b"not a torch file")
    files.insert(2, bad)
    config = b2.PreprocessingConfig(**CFG)
    cache = b2.PreprocessingCache(tmp_path / "cache")
    out = cache.batch_preprocess(files, "ast", config, num_workers=2, show_progress=False, batch_clips=3)
    good = [f for f in files if f != bad]
    assert len(out) == len(good)                                                      # the broken file is skipped
    pre = cache.setup_preprocessor("ast", config)
    for f, o in zip(good, out):
        ref = pre.preprocess(waves[f], 44100)
        assert o.device.type == "cpu" and o.shape == ref.shape
        assert torch.allclose(o, ref, atol=1e-6)
        assert torch.equal(b2.read_cache_entry(pre.cache_dir, f, config.get_hash()), o)
    # second call: every clip is a cache hit, nothing is computed
    pre.preprocess_batch = None
    again = cache.batch_preprocess(files, "ast", config, num_workers=1, show_progress=False)
    assert all(torch.equal(a, o) for a, o in zip(again, out))
