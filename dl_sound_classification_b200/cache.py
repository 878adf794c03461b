"""Batched offline precompute into the reference's own feature-cache format.

The reference keeps a per-file disk cache because its CPU frontend is slow
(``AdvancedCacheManager``, src/datasets/preprocessing.py:150-345): one file per clip,

    <base_cache_dir>/<preprocessor.get_cache_suffix()>/<stem>_<filehash>_<confighash>.cache.gz

with ``filehash = md5(f"{name}_{st_size}_{st_mtime}")[:12]`` (:194-201), ``confighash =
PreprocessingConfig.get_hash()`` (:621-650; it embeds the python / torch / platform versions, so the
cache is only valid on the box that wrote it) and the payload ``gzip(level 6, pickle.dumps(tensor))``
(:211-218).  ``precompute_cache`` fills that directory from the GPU op, a few hundred clips per
launch, so an UNMODIFIED reference checkout -- CPU-only training included -- finds every clip already
cached (``BasePreprocessor.preprocess_with_cache`` :733-770 -> ``get_cached`` :272-293) and never runs
its torchaudio path.  Reading back (``read_cache_entry``) is provided for tests and tools; the compute
itself has no CPU path.
"""
from __future__ import annotations

import gzip
import hashlib
import json
import pickle
import time
from pathlib import Path
from typing import Iterable, List, Optional, Sequence, Tuple

import torch

CACHE_METADATA_NAME = "cache_metadata.json"     # src/datasets/preprocessing.py:166


def file_hash(file_path: Path) -> str:
    """``AdvancedCacheManager._get_file_hash`` (src/datasets/preprocessing.py:194-201)."""
    file_path = Path(file_path)
    try:
        stat = file_path.stat()
        content = f"{file_path.name}_{stat.st_size}_{stat.st_mtime}"
        return hashlib.md5(content.encode()).hexdigest()[:12]
    except OSError:
        return hashlib.md5(str(file_path).encode()).hexdigest()[:12]


def cache_path(cache_dir: Path, original_path: Path, config_hash: str) -> Path:
    """``AdvancedCacheManager._get_cache_path`` (:203-207); ``cache_dir`` = base dir / cache suffix (:726)."""
    original_path = Path(original_path)
    return Path(cache_dir) / f"{original_path.stem}_{file_hash(original_path)}_{config_hash}.cache.gz"


def write_cache_entry(cache_dir: Path, original_path: Path, config_hash: str, data: torch.Tensor) -> Path:
    """``_compress_and_save`` (:209-218): ``pickle.dumps`` of a CPU tensor, gzip level 6."""
    path = cache_path(cache_dir, original_path, config_hash)
    serialized = pickle.dumps(data.detach().cpu().contiguous().clone())
    with gzip.open(path, "wb", compresslevel=6) as f:
        f.write(serialized)
    return path


def read_cache_entry(cache_dir: Path, original_path: Path, config_hash: str) -> Optional[torch.Tensor]:
    """``get_cached`` / ``_load_and_decompress`` (:230-293) without the statistics: None on a miss, and a cache
    file older than its source is stale (:259-265)."""
    path = cache_path(cache_dir, original_path, config_hash)
    if not path.exists():
        return None
    try:
        if Path(original_path).stat().st_mtime > path.stat().st_mtime:
            return None
    except OSError:
        return None
    with gzip.open(path, "rb") as f:
        return pickle.loads(f.read())


def _update_metadata(cache_dir: Path, entries: Sequence[Tuple[Path, Path]], config_hash: str) -> None:
    """The bookkeeping of ``save_cached`` (:295-310) and ``_load_metadata`` / ``_save_metadata`` (:170-192);
    informational only -- ``get_cached`` never consults it."""
    meta_path = Path(cache_dir) / CACHE_METADATA_NAME
    meta = None
    if meta_path.exists():
        try:
            meta = json.loads(meta_path.read_text())
        except (json.JSONDecodeError, OSError):
            meta = None
    if not isinstance(meta, dict):
        meta = {"version": "1.0", "created_time": time.time(), "file_metadata": {}, "cache_stats": {}}
    fm = meta.setdefault("file_metadata", {})
    for original, cached in entries:
        fm[str(original)] = {"cache_path": str(cached), "config_hash": config_hash, "cached_time": time.time(),
                             "original_size": original.stat().st_size if original.exists() else 0}
    meta_path.write_text(json.dumps(meta, indent=2))


def precompute_cache(preprocessor, items: Iterable[Tuple[Path, torch.Tensor, int]], base_cache_dir: Path,
                     batch_clips: int = 256, skip_existing: bool = True) -> List[Path]:
    """Fill the reference's cache directory for ``preprocessor`` (an ``ASTPreprocessor`` mirror) from the GPU.

    ``items`` yields ``(original_path, waveform[1, N] or [N], sample_rate)`` -- what the reference's
    ``load_audio_bundle`` + ``preprocess_with_cache`` call chain sees per file.  Clips are grouped into ragged
    batches of ``batch_clips`` and each batch is ONE fused launch (``preprocess_batch``); every clip is cached
    with its own natural frame count, exactly what a per-clip ``preprocess`` call returns.  Returns the cache
    files written."""
    cache_dir = Path(base_cache_dir) / preprocessor.get_cache_suffix()
    cache_dir.mkdir(parents=True, exist_ok=True)
    config_hash = preprocessor.config.get_hash()
    written: List[Path] = []
    entries: List[Tuple[Path, Path]] = []
    pending: List[Tuple[Path, torch.Tensor, int]] = []

    def flush():
        if not pending:
            return
        waves = [w.reshape(-1).to(torch.float32) for _, w, _ in pending]
        rates = [int(r) for _, _, r in pending]
        lengths = [int(w.numel()) for w in waves]
        flat = torch.cat(waves)
        same = len(set(rates)) == 1
        out, nfr = preprocessor.preprocess_batch(flat, rates[0] if same else rates, lengths=lengths)
        out = out.cpu()
        nfr = nfr.cpu().tolist()
        for (path, _, _), feats, m in zip(pending, out, nfr):
            m = int(m)
            if m <= 0:
                raise AssertionError(f"{path}: clip is shorter than one analysis window")
            p = write_cache_entry(cache_dir, path, config_hash, feats[..., :m])
            written.append(p)
            entries.append((Path(path), p))
        pending.clear()

    for path, wave, rate in items:
        path = Path(path)
        if skip_existing and read_cache_entry(cache_dir, path, config_hash) is not None:
            continue
        pending.append((path, wave, int(rate)))
        if len(pending) >= batch_clips:
            flush()
    flush()
    if entries:
        _update_metadata(cache_dir, entries, config_hash)
    return written
