"""CPU tests of the cache-format row: files written by dl_sound_classification_b200.cache
are found and decoded by the reference's own AdvancedCacheManager, and vice versa."""
import gzip
import os
import pickle
import sys
import time
from pathlib import Path

import numpy as np
import pytest
import torch

import dl_sound_classification_b200 as b2
from dl_sound_classification_b200 import cache as CA

HERE = Path(os.path.dirname(os.path.abspath(__file__)))
REF = Path("/root/reference/src/datasets")


def _fake_clip(tmp_path, name="1-100032-A-0.pt", size=1234):
    p = tmp_path / name
    p.write_bytes(b"x" * size)
    return p


def test_decodes_an_entry_written_by_the_reference(tmp_path):
    """tests/golden/cache_entry.cache.gz was written by AdvancedCacheManager._compress_and_save itself."""
    want = np.load(HERE / "golden" / "cache_entry.npz")["data"]
    clip = _fake_clip(tmp_path)
    dst = CA.cache_path(tmp_path, clip, "0123456789ab")
    dst.write_bytes((HERE / "golden" / "cache_entry.cache.gz").read_bytes())
    got = CA.read_cache_entry(tmp_path, clip, "0123456789ab")
    assert isinstance(got, torch.Tensor) and got.dtype == torch.float32 and np.array_equal(got.numpy(), want)


def test_file_name_payload_and_staleness_rules(tmp_path):
    import hashlib
    clip = _fake_clip(tmp_path)
    st = clip.stat()
    fh = hashlib.md5(f"{clip.name}_{st.st_size}_{st.st_mtime}".encode()).hexdigest()[:12]   # preprocessing.py:197-199
    assert CA.file_hash(clip) == fh
    assert CA.file_hash(tmp_path / "missing.pt") == hashlib.md5(str(tmp_path / "missing.pt").encode()).hexdigest()[:12]
    cdir = tmp_path / "ast_abc"
    cdir.mkdir()
    assert CA.cache_path(cdir, clip, "cafecafecafe") == cdir / f"1-100032-A-0_{fh}_cafecafecafe.cache.gz"
    x = torch.arange(24, dtype=torch.float32).reshape(1, 4, 6)
    assert CA.read_cache_entry(cdir, clip, "cafecafecafe") is None
    p = CA.write_cache_entry(cdir, clip, "cafecafecafe", x)
    with gzip.open(p, "rb") as f:
        assert torch.equal(pickle.loads(f.read()), x)                       # gzip(pickle(tensor)), preprocessing.py:213-217
    assert torch.equal(CA.read_cache_entry(cdir, clip, "cafecafecafe"), x)
    assert CA.read_cache_entry(cdir, clip, "000000000000") is None          # another config hash: another file
    os.utime(p, (time.time() - 100, time.time() - 100))                     # source newer than the entry: stale
    assert CA.read_cache_entry(cdir, clip, "cafecafecafe") is None


@pytest.mark.skipif(not REF.exists(), reason="needs the reference checkout (authoring container)")
def test_round_trip_with_the_reference_cache_manager(tmp_path):
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(REF.parent))
    try:
        import preprocessing as R
    finally:
        sys.path.remove(str(REF)); sys.path.remove(str(REF.parent))
    kw = dict(sample_rate=44100, n_mels=128, bc_mixing=False, normalize=True, target_mean=0.0, target_std=0.5)
    ours, theirs = b2.PreprocessingConfig(**kw), R.PreprocessingConfig(**kw)
    assert ours.get_hash() == theirs.get_hash()
    pre = b2.ASTPreprocessor(ours)                                          # no GPU needed: the plan is lazy
    assert pre.get_cache_suffix() == R.ASTPreprocessor(theirs).get_cache_suffix()
    cdir = tmp_path / pre.get_cache_suffix()
    mgr = R.AdvancedCacheManager(cdir)                                      # what BasePreprocessor.setup_cache builds (:726-727)
    clip = _fake_clip(tmp_path)
    g = torch.Generator().manual_seed(1)
    x = torch.randn((1, 128, 31), generator=g)
    p = CA.write_cache_entry(cdir, clip, ours.get_hash(), x)
    assert p == mgr._get_cache_path(clip, theirs.get_hash())
    assert mgr.is_cached(clip, theirs.get_hash())
    assert torch.equal(mgr.get_cached(clip, theirs.get_hash()), x)          # an unmodified reference run hits our entry
    clip2 = _fake_clip(tmp_path, "5-9032-A-0.pt", 99)
    mgr.save_cached(clip2, theirs.get_hash(), x * 2)
    assert torch.equal(CA.read_cache_entry(cdir, clip2, ours.get_hash()), x * 2)   # and we read theirs
    CA._update_metadata(cdir, [(clip, p)], ours.get_hash())
    assert str(clip) in R.AdvancedCacheManager(cdir).metadata["file_metadata"]     # their loader accepts our metadata file


def test_stock_config_means_the_reference_recipe():
    """A config the reference would hash identically must produce the reference's own features, so the default
    frontend of a stock config (no ``frontend`` / ``target_sample_rate`` key) is the MelSpectrogram-dB recipe;
    the kaldi recipe is opt-in and always changes the hash (its entries can never sit under the stock cache key)."""
    stock = dict(sample_rate=44100, n_mels=128, bc_mixing=False, normalize=True, target_mean=0.0, target_std=0.5)
    pre = b2.ASTPreprocessor(b2.PreprocessingConfig(**stock))
    assert pre.frontend_name == "melspectrogram"
    k1 = b2.ASTPreprocessor(b2.PreprocessingConfig(**stock, frontend="kaldi_fbank"))
    k2 = b2.ASTPreprocessor(b2.PreprocessingConfig(**stock, target_sample_rate=16000))
    assert k1.frontend_name == k2.frontend_name == "kaldi_fbank"
    assert len({pre.get_cache_suffix(), k1.get_cache_suffix(), k2.get_cache_suffix()}) == 3
