"""GPU test of the cache-format row: a directory of clips is precomputed in ragged batches on
the GPU and lands in the reference's cache layout, each entry equal to the per-clip ``preprocess`` result."""
import os
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("frontend", ["kaldi_fbank", "melspectrogram"])
def test_precompute_cache_fills_the_reference_layout(tmp_path, frontend):
    import dl_sound_classification_b200 as b2
    from dl_sound_classification_b200 import cache as CA
    cfg = dict(sample_rate=44100, n_mels=128, normalize=True, target_mean=0.0, target_std=0.5, frontend=frontend)
    if frontend == "kaldi_fbank":
        cfg.update(norm_mean=-6.6268, norm_std=5.0613, extra_rates=(22050,))
    pre = b2.create_preprocessor("ast", cfg, tmp_path / "cache")
    g = torch.Generator().manual_seed(12)
    items = []
    for i in range(23):
        rate = 22050 if (frontend == "kaldi_fbank" and i % 5 == 0) else 44100
        n = int(rate * (0.4 + 0.05 * i))
        p = tmp_path / f"clip_{i:03d}.pt"
        p.write_bytes(b"0" * (100 + i))
        items.append((p, torch.rand(1, n, generator=g) * 2 - 1, rate))
    written = b2.precompute_cache(pre, items, tmp_path / "cache", batch_clips=8)
    cdir = tmp_path / "cache" / pre.get_cache_suffix()
    assert len(written) == 23 and all(w.parent == cdir and w.name.endswith(".cache.gz") for w in written)
    assert (cdir / CA.CACHE_METADATA_NAME).exists()
    h = pre.config.get_hash()
    for path, wave, rate in items[::4]:
        got = b2.read_cache_entry(cdir, path, h)
        want = pre.preprocess(wave, rate)                       # the reference's per-clip contract: (1, n_mels, T) on CPU
        assert got.device.type == "cpu" and got.shape == want.shape and got.shape[:2] == (1, 128)
        if frontend == "kaldi_fbank":
            assert torch.equal(got, want), path                  # batch invariance of the fused kernel: bit-identical
        else:
            assert float((got - want).abs().max()) < 2e-5, path  # per-clip dB max / statistics: float64 atomics, order-dependent
    assert b2.precompute_cache(pre, items, tmp_path / "cache") == []       # everything is already there


def test_stock_config_cache_entries_are_the_reference_features(tmp_path):
    """with a stock config (hash == the reference's) the cache writer must store what the reference
    itself would have computed -- checked against the golden output of the reference's own ASTPreprocessor."""
    import numpy as np
    import dl_sound_classification_b200 as b2
    from inputs import short_clip
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_actual.npz"))
    cfg = dict(sample_rate=44100, n_mels=128, bc_mixing=False, normalize=True, target_mean=0.0, target_std=0.5)
    pre = b2.create_preprocessor("ast", cfg, tmp_path / "cache")
    assert pre.frontend_name == "melspectrogram"
    clip = tmp_path / "1-100032-A-0.pt"
    clip.write_bytes(b"x" * 64)
    w = short_clip(44100, seed=77)
    written = b2.precompute_cache(pre, [(clip, w, 44100)], tmp_path / "cache")
    assert len(written) == 1
    got = b2.read_cache_entry(tmp_path / "cache" / pre.get_cache_suffix(), clip, pre.config.get_hash())
    assert tuple(got.shape) == (1, 128, 276)
    assert np.abs(got.numpy() - g["ast_44k"]).max() < 5e-4
