"""Drop-in mirror of the reference's AST preprocessing interface (src/datasets/preprocessing.py).

Same names, arguments and error behaviour as ``PreprocessingConfig`` (:612-680),
``BasePreprocessor`` (:683-792), ``ASTPreprocessor`` (:971-1113), ``resample_waveform``
(:61-76) and ``create_preprocessor`` (:1315-1346) -- computed by the fused sm_100a kernel.

New OPTIONAL config keys (free-form kwargs already flow through the reference's Hydra
configs untouched): ``frontend`` ("melspectrogram" = the reference's own recipe, the DEFAULT for a
stock config; "kaldi_fbank" = the north_star recipe, also selected by giving ``target_sample_rate``),
``target_sample_rate`` (16000), ``target_frames`` (None = the clip's own frame count),
``window_type`` ("hanning"), ``norm_mean`` / ``norm_std`` (dataset statistics, scalar or
per-bin; None + ``normalize`` = the reference's per-clip mean / unbiased std), ``extra_rates``.
The gzip/pickle disk cache is not consulted -- GPU recompute makes it moot, ``preprocess_with_cache``
simply recomputes -- but ``cache.precompute_cache`` can WRITE it in the reference's format (row N4);
the EnvNet/CNN modes are out of scope.
"""
from __future__ import annotations

import hashlib
import json
import logging
import random
from abc import ABC, abstractmethod
from functools import lru_cache
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

from . import specaugment as _sa
from .frontend import AST_FBANK_KWARGS, FbankFrontend, MelSpecFrontend, _require_cuda

logger = logging.getLogger(__name__)

AST_HOP_LENGTH = 160      # src/datasets/preprocessing.py:55-58
AST_N_FFT = 1024
AST_WIN_LENGTH = 400


class PreprocessingConfig:
    """Free-form kwargs bag + md5 hash (src/datasets/preprocessing.py:612-680)."""

    def __init__(self, **kwargs):
        self.config = kwargs
        self._hash = None
        self._validation_errors: List[str] = []

    def get_hash(self) -> str:
        if self._hash is None:
            import platform
            system_info = {"python_version": platform.python_version(), "torch_version": torch.__version__,
                           "platform": platform.platform()}

            # the reference's own conversion, quirk included (src/datasets/preprocessing.py:631-641): an omegaconf
            # DictConfig becomes a dict, but anything else that is iterable and indexable -- a plain dict too --
            # becomes the list of its items, i.e. a plain dict hashes by its KEYS only.  Reproduced literally so that
            # cache directories and cache file names agree with an unmodified reference checkout.
            def convert_omegaconf(obj):
                if hasattr(obj, "_content"):
                    return {k: convert_omegaconf(v) for k, v in obj.items()}
                elif hasattr(obj, "__iter__") and hasattr(obj, "__getitem__") and not isinstance(obj, str):
                    try:
                        return [convert_omegaconf(item) for item in obj]
                    except Exception:
                        return obj
                else:
                    return obj
            s = json.dumps({"config": convert_omegaconf(self.config), "system_info": system_info}, sort_keys=True)
            self._hash = hashlib.md5(s.encode()).hexdigest()[:12]
        return self._hash

    def validate(self) -> bool:
        self._validation_errors = []
        c = self.config
        if "sample_rate" in c and (not isinstance(c["sample_rate"], int) or c["sample_rate"] <= 0):
            self._validation_errors.append("sample_rate must be a positive integer")
        if "n_mels" in c and (not isinstance(c["n_mels"], int) or c["n_mels"] <= 0):
            self._validation_errors.append("n_mels must be a positive integer")
        if "window_length" in c and (not isinstance(c["window_length"], (int, float)) or c["window_length"] <= 0):
            self._validation_errors.append("window_length must be a positive number")
        return len(self._validation_errors) == 0

    def get_validation_errors(self) -> List[str]:
        return self._validation_errors.copy()

    def __getattr__(self, name: str) -> Any:
        cfg = self.__dict__.get("config", {})
        if name in cfg:
            return cfg[name]
        raise AttributeError(f"'{self.__class__.__name__}' object has no attribute '{name}'")


class BasePreprocessor(ABC):
    """src/datasets/preprocessing.py:683-792 without the disk cache (out of scope)."""

    def __init__(self, config: PreprocessingConfig):
        self.config = config
        self.cache_manager = None
        self._performance_stats: List[float] = []

    @abstractmethod
    def preprocess(self, waveform: torch.Tensor, sample_rate: int) -> torch.Tensor:
        ...

    @abstractmethod
    def get_cache_suffix(self) -> str:
        ...

    def multi_crop_test(self, waveform: torch.Tensor) -> List[torch.Tensor]:
        return [self.preprocess(waveform, self.config.config.get("sample_rate", 44100))]

    def setup_cache(self, base_cache_dir: Path, force_rebuild: bool = False, max_cache_size_gb: float = 5.0) -> None:
        """src/datasets/preprocessing.py:716-731.  Only the cache DIRECTORY is remembered: per-clip calls recompute on
        the GPU (faster than gunzip + unpickle), while ``cache.precompute_cache`` can fill that directory in the
        reference's own file format for consumers that still read it."""
        self.cache_manager = None
        self.cache_dir = Path(base_cache_dir) / self.get_cache_suffix()

    def preprocess_with_cache(self, waveform: torch.Tensor, sample_rate: int,
                              original_path: Optional[Path] = None) -> torch.Tensor:
        return self.preprocess(waveform, sample_rate)

    def get_cache_stats(self):
        return None

    def cleanup_cache(self, max_age_days: int = 30) -> None:
        return None

    def get_performance_stats(self) -> Dict[str, float]:
        return {}


@lru_cache(maxsize=16)
def _resampler(device_index: int, orig: int, new: int) -> FbankFrontend:
    # only the rate table matters for b200fbank_resample; the fbank options are placeholders
    return FbankFrontend(orig_rates=(orig,), device=torch.device("cuda", device_index), sample_frequency=float(new),
                         num_mel_bins=23, high_freq=0.0, low_freq=0.0 if new < 100 else 20.0)


def resample_waveform(waveform: torch.Tensor, current_rate: int, target_rate: int) -> torch.Tensor:
    """src/datasets/preprocessing.py:61-76: ``Resample(current, target)(waveform)`` or the input itself
    when the rates match.  ``waveform`` is ``(..., time)``; the result lives on the input's device."""
    if current_rate == target_rate:
        return waveform
    if not waveform.is_floating_point():
        raise TypeError(f"Expected floating point type for waveform tensor, but received {waveform.dtype}.")
    if int(current_rate) != current_rate or int(target_rate) != target_rate:
        raise Exception("Frequencies must be of integer type to ensure quality resampling computation.")
    dev = _require_cuda(waveform.device if waveform.is_cuda else None)
    fe = _resampler(dev.index, int(current_rate), int(target_rate))
    shape = waveform.shape
    flat = waveform.reshape(-1, shape[-1]).to(device=dev, dtype=torch.float32)
    out = fe.resample(flat).reshape(shape[:-1] + (-1,))
    return out.to(device=waveform.device, dtype=waveform.dtype)


class ASTPreprocessor(BasePreprocessor):
    """B200 drop-in for ``ASTPreprocessor`` (src/datasets/preprocessing.py:971-1113).

    ``preprocess(waveform[1, N], sample_rate) -> [1, n_mels, T]`` float32 on the input's device
    (the reference's per-clip contract); ``preprocess_batch`` is the batched GPU entry point a
    DataModule hook (``on_after_batch_transfer``) should call.
    """

    def __init__(self, config: PreprocessingConfig, device=None):
        super().__init__(config)
        c = config.config
        self.n_mels = c.get("n_mels", 128)
        self.sample_rate = c.get("sample_rate", 44100)
        self.target_mean = c.get("target_mean", 0.0)
        self.target_std = c.get("target_std", 0.5)
        self.normalize = c.get("normalize", True)
        # Default recipe = what the config's cache hash MEANS.  A stock reference config (no ``frontend`` /
        # ``target_sample_rate`` key) hashes exactly as the reference hashes it, so it must produce the reference's own
        # features: MelSpectrogram(1024/160 at sample_rate) + dB + per-clip normalisation, (1, 128, 1379) for 5 s.  The
        # north_star kaldi recipe is opt-in (``frontend: kaldi_fbank`` or a ``target_sample_rate``); either key changes
        # the hash, so its features can never land under the reference's stock cache key.
        self.frontend_name = c.get("frontend", "kaldi_fbank" if "target_sample_rate" in c else "melspectrogram")
        self.target_sample_rate = int(c.get("target_sample_rate", 16000))
        self.target_frames = c.get("target_frames", None)
        self.window_type = c.get("window_type", "hanning")
        self.norm_mean = c.get("norm_mean", None)
        self.norm_std = c.get("norm_std", None)
        if (self.norm_mean is None) != (self.norm_std is None):
            raise ValueError("norm_mean and norm_std must be given together")
        if self.frontend_name not in ("kaldi_fbank", "melspectrogram"):
            raise ValueError(f"Unknown frontend: {self.frontend_name}")
        self.n_fft, self.hop_length, self.win_length = AST_N_FFT, AST_HOP_LENGTH, AST_WIN_LENGTH
        rates = [int(self.sample_rate)] + [int(r) for r in c.get("extra_rates", ())]
        self.rates = tuple(dict.fromkeys(rates))
        kw = dict(AST_FBANK_KWARGS, num_mel_bins=int(self.n_mels), sample_frequency=float(self.target_sample_rate),
                  window_type=self.window_type)
        self._device = device
        self._kw = kw
        self._fe: Optional[FbankFrontend] = None
        # Mixup augmentation (src/datasets/preprocessing.py:999-1008)
        mixup_config = c.get("mixup", {}) or {}
        if mixup_config.get("enabled", False):
            from .mixup import MixupAugmentation
            self.mixup = MixupAugmentation(alpha=mixup_config.get("alpha", 0.5), prob=mixup_config.get("prob", 0.5))
        else:
            self.mixup = None

    # the plan is created lazily so that constructing the object needs no GPU (config plumbing, hashing)
    @property
    def frontend(self) -> FbankFrontend:
        if self._fe is None:
            if self.frontend_name == "melspectrogram":      # the reference's own recipe, computed at self.sample_rate
                self._fe = MelSpecFrontend(int(self.sample_rate), self.n_fft, self.hop_length, self.win_length,
                                           int(self.n_mels), 80.0, orig_rates=self.rates,
                                           device=_require_cuda(self._device))
            else:
                self._fe = FbankFrontend(orig_rates=self.rates, device=_require_cuda(self._device), **self._kw)
        return self._fe

    def get_cache_suffix(self) -> str:
        return f"ast_{self.config.get_hash()}"

    # ------------------------------------------------------------------------------------------
    def preprocess_batch(self, waveforms: torch.Tensor, sample_rate, lengths: Optional[Sequence[int]] = None,
                         masks: Optional[torch.Tensor] = None, target_frames: Optional[int] = None, mixup=None):
        """``(B, N)`` (or flat ragged + ``lengths``) -> ``((B, 1, n_mels, T), n_frames[B])`` on the GPU.
        ``sample_rate`` is an int or a per-clip sequence drawn from the plan's rate table.  ``mixup`` = ``(bank, plan)``
        (``mixup.draw_mixup_plan``): Mixup after SpecAugment as in esc50.py:267-285 -- fused into the kernel's epilogue
        for the kaldi recipe with dataset statistics, one more launch otherwise; same bits either way."""
        fe = self.frontend
        if isinstance(sample_rate, int):
            rid_list = None if fe.rate_id(sample_rate) == 0 and len(fe.orig_rates) == 1 else None
            rate_ids = None
            rid0 = fe.rate_id(sample_rate)
            B = waveforms.shape[0] if lengths is None else len(lengths)
            if rid0 != 0:
                rate_ids = torch.full((B,), rid0, dtype=torch.int32)
            del rid_list
        else:
            rate_ids = torch.tensor([fe.rate_id(int(r)) for r in sample_rate], dtype=torch.int32)
        offsets = None
        if lengths is not None:
            lens = torch.as_tensor(list(lengths), dtype=torch.int64)
            offsets = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)])
            max_len = [int(x) for x in lens]
        else:
            max_len = [int(waveforms.shape[-1])] * int(waveforms.shape[0])
        T = target_frames if target_frames is not None else self.target_frames
        if T is None:
            rids = [0] * len(max_len) if rate_ids is None else rate_ids.tolist()
            T = max(fe.num_frames(n, r) for n, r in zip(max_len, rids))
            if T <= 0:
                raise AssertionError("choose a window size {} that is [2, {}]".format(fe.plan.window_size, min(max_len)))
        if self.frontend_name == "melspectrogram":
            # src/datasets/preprocessing.py:1024-1037: dB, top_db clamp and per-clip normalisation are fused passes
            out, nfr = fe(waveforms, out_frames=int(T), offsets=offsets, rate_ids=rate_ids, masks=masks, to_db=True,
                          normalize=bool(self.normalize), target_mean=self.target_mean, target_std=self.target_std,
                          layout="bft")
            if mixup is not None:
                from .mixup import mixup_batch
                out = mixup_batch(out, mixup[0], mixup[1], out=out)
            return out, nfr
        mean = std = None
        if self.normalize and self.norm_mean is not None:
            mean, std = self.norm_mean, self.norm_std
        # without dataset statistics the reference normalises every clip with its OWN mean / unbiased std
        # (src/datasets/preprocessing.py:1030-1037): one more kernel pass on the device, masks after it
        per_clip = bool(self.normalize and self.norm_mean is None)
        return fe(waveforms, out_frames=int(T), offsets=offsets, rate_ids=rate_ids, masks=masks, mean=mean, std=std,
                  target_mean=self.target_mean, target_std=self.target_std, layout="bft", per_clip_norm=per_clip, mixup=mixup)

    def preprocess(self, waveform: torch.Tensor, sample_rate: int) -> torch.Tensor:
        if waveform.dim() == 1:
            waveform = waveform.unsqueeze(0)
        in_device = waveform.device
        out, _ = self.preprocess_batch(waveform[:1].to(torch.float32), int(sample_rate))
        out = out[0]                                            # (1, n_mels, T)
        return out if out.device == in_device else out.to(in_device)

    def multi_crop_test(self, waveform: torch.Tensor) -> List[torch.Tensor]:
        """src/datasets/preprocessing.py:1041-1073: ten evenly spaced 5 s crops (one batched launch here)."""
        total_length = waveform.shape[-1]
        n_crops = 10
        if total_length <= self.sample_rate * 5:
            return [self.preprocess(waveform, self.sample_rate)]
        crop_length = int(self.sample_rate * 5)
        starts = torch.linspace(0, total_length - crop_length, n_crops).long()
        crops = torch.stack([waveform[..., int(s):int(s) + crop_length].reshape(-1) for s in starts])
        out, _ = self.preprocess_batch(crops.to(torch.float32), int(self.sample_rate))
        out = out.to(waveform.device)
        return [out[i] for i in range(n_crops)]

    def apply_specaugment(self, spectrogram: torch.Tensor, time_mask: int = 192, freq_mask: int = 48) -> torch.Tensor:
        """src/datasets/preprocessing.py:1075-1104: clones, draws four ``random.randint`` (time first), zero-fills."""
        channels, n_mels, n_frames = spectrogram.shape
        return _sa.apply_intervals(spectrogram, _sa.reference_intervals(n_frames, n_mels, time_mask, freq_mask, random))

    def draw_specaugment_masks(self, batch: int, n_frames, time_mask: int = 192, freq_mask: int = 48) -> torch.Tensor:
        """Mask table for the fused path; same RNG consumption as ``batch`` sequential ``apply_specaugment`` calls."""
        return _sa.draw_masks(batch, n_frames, int(self.n_mels), time_mask, freq_mask, "reference", random)

    def apply_mixup(self, spec1: torch.Tensor, spec2: torch.Tensor, label1: int, label2: int, num_classes: int):
        """src/datasets/preprocessing.py:1106-1113: mix if enabled in the config, else the original with one-hot
        labels.  Batches should use ``mixup.draw_mixup_plan`` + ``mixup.mixup_batch`` (one launch) instead."""
        if self.mixup:
            return self.mixup(spec1, spec2, label1, label2, num_classes)
        labels = torch.zeros(num_classes, dtype=torch.float32)
        labels[label1] = 1.0
        return spec1, labels


B200ASTPreprocessor = ASTPreprocessor


@lru_cache(maxsize=16)
def _melspec_frontend(device_index: int, sr: int, n_mels: int, n_fft: int, hop_length: int) -> MelSpecFrontend:
    return MelSpecFrontend(sr, n_fft, hop_length, None, n_mels, 80.0, device=torch.device("cuda", device_index))


def melspectrogram(wav: torch.Tensor, sr: int = 44100, n_mels: int = 128, n_fft: int = 1_024, hop_length: int = 512,
                   log_scale: bool = True) -> torch.Tensor:
    """Mirror of src/utils/audio.py:60-84: waveform ``(1, N)`` -> ``(1, n_mels, frames)`` (dB, top_db=80, if
    ``log_scale``) via ``MelSpectrogram(win_length=n_fft, center=True)`` -- the fallback frontend of
    ``create_ast_fallback_spectrogram`` (src/datasets/preprocessing.py:79-97)."""
    dev = _require_cuda(wav.device if wav.is_cuda else None)
    fe = _melspec_frontend(dev.index, int(sr), int(n_mels), int(n_fft), int(hop_length))
    w = wav.reshape(-1, wav.shape[-1])[:1].to(torch.float32)
    T = fe.num_frames(int(w.shape[-1]))
    out, _ = fe(w, out_frames=T, to_db=bool(log_scale), normalize=False, layout="bft", return_n_frames=False)
    out = out[0]
    return out if out.device == wav.device else out.to(wav.device)


def create_ast_fallback_spectrogram(waveform: torch.Tensor, sample_rate: int, n_mels: int = 128) -> torch.Tensor:
    """src/datasets/preprocessing.py:79-97."""
    return melspectrogram(waveform, sample_rate, n_mels, AST_N_FFT, AST_HOP_LENGTH, log_scale=True)


class PreprocessingCache:
    """src/datasets/preprocessing.py:1116-1174 reduced to the dispatch it performs for mode 'ast'."""

    def __init__(self, base_cache_dir: Path, max_cache_size_gb: float = 5.0):
        self.base_cache_dir = Path(base_cache_dir)
        self.max_cache_size_gb = max_cache_size_gb
        self.preprocessors: Dict[str, BasePreprocessor] = {}

    def get_preprocessor(self, mode: str, config: PreprocessingConfig) -> BasePreprocessor:
        if not config.validate():
            raise ValueError(f"Invalid preprocessing config: {config.get_validation_errors()}")
        if mode == "ast":
            return ASTPreprocessor(config)
        if mode in ("envnet_v2", "cnn_esc50"):
            raise NotImplementedError(f"preprocessing mode {mode!r} is outside the B200 frontend's scope (AST path only)")
        raise ValueError(f"Unknown preprocessing mode: {mode}")

    def setup_preprocessor(self, mode: str, config: PreprocessingConfig, force_rebuild: bool = False) -> BasePreprocessor:
        key = f"{mode}_{config.get_hash()}"
        if key in self.preprocessors and not force_rebuild:
            return self.preprocessors[key]
        pre = self.get_preprocessor(mode, config)
        pre.setup_cache(self.base_cache_dir, force_rebuild=force_rebuild, max_cache_size_gb=self.max_cache_size_gb)
        self.preprocessors[key] = pre
        return pre

    def batch_preprocess(self, file_paths: List[Path], mode: str, config: PreprocessingConfig, num_workers: int = 4,
                         show_progress: bool = True, batch_clips: int = 256, sample_rate: int = 44100,
                         use_cache: bool = True) -> List[torch.Tensor]:
        """src/datasets/preprocessing.py:1176-1254 with the same signature and result: the preprocessed tensors
        ``[1, n_mels, T_clip]`` (CPU, float32) of the files that could be loaded, in file order, files that fail to load
        skipped with a warning.  The reference runs ``preprocess_with_cache`` per file on a thread pool; here the
        ``num_workers`` threads only LOAD the ``.pt`` bundles (``load_audio_bundle``, :100-117: ``{"waveform", "label"}``,
        44.1 kHz assumed as at :1205) while the features of ``batch_clips`` clips at a time come from ONE ragged fused
        launch.  With ``use_cache`` the reference's own cache files are read first and written for the misses, as
        ``preprocess_with_cache`` does (:733-764)."""
        from concurrent.futures import ThreadPoolExecutor
        from . import cache as _cache
        pre = self.setup_preprocessor(mode, config)
        cache_dir = getattr(pre, "cache_dir", None) if use_cache else None
        if cache_dir is not None:
            Path(cache_dir).mkdir(parents=True, exist_ok=True)
        config_hash = config.get_hash()
        paths = [Path(f) for f in file_paths]
        results: List[Optional[torch.Tensor]] = [None] * len(paths)

        def load(i: int):
            try:
                if cache_dir is not None:
                    hit = _cache.read_cache_entry(cache_dir, paths[i], config_hash)
                    if hit is not None:
                        return i, hit, True
                bundle = torch.load(paths[i], map_location="cpu")
                return i, bundle["waveform"], False
            except Exception as e:                               # the reference logs and skips (:1214-1216)
                logger.warning(f"Failed to process {paths[i]}: {e}")
                return i, None, False

        bar = None
        if show_progress:
            try:
                from tqdm import tqdm
                bar = tqdm(total=len(paths), desc=f"Preprocessing with {mode}", unit="files")
            except ImportError:
                bar = None
        pending: List[Tuple[int, torch.Tensor]] = []
        entries: List[Tuple[Path, Path]] = []

        def flush():
            if not pending:
                return
            waves = [w.reshape(-1).to(torch.float32) for _, w in pending]
            out, nfr = pre.preprocess_batch(torch.cat(waves), int(sample_rate), lengths=[int(w.numel()) for w in waves])
            out, nfr = out.cpu(), nfr.cpu().tolist()
            for (i, _), feats, m in zip(pending, out, nfr):
                if int(m) <= 0:
                    logger.warning(f"Failed to process {paths[i]}: clip is shorter than one analysis window")
                    continue
                results[i] = feats[..., :int(m)].clone()
                if cache_dir is not None:
                    entries.append((paths[i], _cache.write_cache_entry(cache_dir, paths[i], config_hash, results[i])))
            pending.clear()

        loader = ThreadPoolExecutor(max_workers=max(1, int(num_workers)))
        try:
            for i, data, hit in loader.map(load, range(len(paths))):
                if bar is not None:
                    bar.update(1)
                if data is None:
                    continue
                if hit:
                    results[i] = data
                    continue
                pending.append((i, data))
                if len(pending) >= batch_clips:
                    flush()
            flush()
        finally:
            loader.shutdown()
            if bar is not None:
                bar.close()
        if entries:
            _cache._update_metadata(cache_dir, entries, config_hash)
        return [r for r in results if r is not None]


def create_preprocessor(mode: str, config_dict: Dict[str, Any], base_cache_dir: Path, force_rebuild: bool = False,
                        max_cache_size_gb: float = 5.0) -> BasePreprocessor:
    """src/datasets/preprocessing.py:1315-1346.  Never falls back: errors propagate (the reference's caller
    swallows them into a "basic mode", src/datasets/esc50.py:149-160 -- a GPU drop-in must not)."""
    config = PreprocessingConfig(**config_dict)
    return PreprocessingCache(base_cache_dir, max_cache_size_gb).setup_preprocessor(mode, config, force_rebuild)
