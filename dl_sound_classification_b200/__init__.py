"""dl_sound_classification_b200 -- B200-native waveform -> log-mel fbank frontend.

Drop-in for the CPU torchaudio feature path of youssefg7/dl-sound-classification
(``src/datasets/preprocessing.py``), built as hand-written sm_100a CUDA behind the C ABI
in ``include/b200fbank.h``.  Importing the package loads ``lib/libb200fbank.so``; it
raises if the library has not been built -- there is no CPU fallback.
"""
from . import _capi
from .frontend import AST_FBANK_KWARGS, FbankFrontend, MelSpecFrontend, launch_count
from .kaldi import fbank
from .preprocessing import (ASTPreprocessor, B200ASTPreprocessor, BasePreprocessor, PreprocessingCache, PreprocessingConfig,
                            create_preprocessor, melspectrogram, resample_waveform)
from .datamodule import B200DataModuleMixin
from .cache import cache_path, file_hash, precompute_cache, read_cache_entry, write_cache_entry
from .mixup import MixupAugmentation, MixupPlan, draw_mixup_plan, mixup_batch, mixup_labels
from . import ops
from .patch_embed import PatchEmbed, patch_embed
from .specaugment import SpecAugment
from .stats import DatasetStats, NormStats, finalize_sums

__all__ = ["FbankFrontend", "AST_FBANK_KWARGS", "fbank", "launch_count", "_capi", "ASTPreprocessor",
           "B200ASTPreprocessor", "BasePreprocessor", "PreprocessingConfig", "create_preprocessor",
           "resample_waveform", "melspectrogram", "MelSpecFrontend", "SpecAugment", "DatasetStats", "NormStats", "finalize_sums",
           "precompute_cache", "read_cache_entry", "write_cache_entry", "cache_path", "file_hash",
           "PreprocessingCache", "B200DataModuleMixin", "ops", "PatchEmbed", "patch_embed", "MixupAugmentation", "MixupPlan", "draw_mixup_plan", "mixup_batch", "mixup_labels"]
__version__ = "0.1.0"
