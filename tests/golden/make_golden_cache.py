"""Generate tests/golden/cache_entry.cache.gz + cache_entry.npz with the REFERENCE's own cache writer.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden_cache.py

``AdvancedCacheManager._compress_and_save`` (src/datasets/preprocessing.py:209-218) writes one entry; the tests
decode it with ``dl_sound_classification_b200.cache`` (no reference needed at test time).
"""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

HERE = Path(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference/src/datasets")
sys.path.insert(0, "/root/reference/src")
from preprocessing import AdvancedCacheManager  # noqa: E402

g = torch.Generator().manual_seed(77)
data = torch.randn((1, 128, 23), generator=g)
with tempfile.TemporaryDirectory() as d:
    mgr = AdvancedCacheManager(Path(d))
    out = Path(d) / "entry.cache.gz"
    mgr._compress_and_save(data, out)
    (HERE / "cache_entry.cache.gz").write_bytes(out.read_bytes())
np.savez_compressed(HERE / "cache_entry.npz", data=data.numpy(), meta_torch=np.array(torch.__version__))
print("written", (HERE / "cache_entry.cache.gz").stat().st_size, "bytes")
