"""The batched injection point (SURVEY.md section 8b hook (i)) as code: a Lightning-free mixin for the reference's
``ESC50DataModule`` (src/datasets/esc50.py:327-629).

The reference extracts features one clip at a time inside ``Dataset.__getitem__`` on DataLoader worker processes
(``ESC50Dataset._process_ast``, esc50.py:246-291: ``preprocess_with_cache`` -> ``apply_specaugment`` -> ``apply_mixup``
-> soft labels).  With this mixin the workers return RAW waveforms + integer labels (the reference dataset does exactly
that for a preprocessing mode it does not know, esc50.py:221-222), and the whole chain runs on the collated batch in the
main process right after Lightning has moved it to the GPU -- ``on_after_batch_transfer`` is the hook Lightning calls
there.  The class imports neither Lightning nor the reference, so a Hydra override can name a three-line subclass::

    class ESC50DataModuleB200(B200DataModuleMixin, ESC50DataModule):       # dataset._target_=...ESC50DataModuleB200
        pass

Random draws replay the reference's per-sample order (SpecAugment's four ``random.randint`` per clip, then Mixup's coin /
partner / coin / Beta per clip), so one seed gives the same augmentation for a batch as for the per-sample loop when the
dataset is visited in the same order.
"""
from __future__ import annotations

import random
from pathlib import Path
from typing import Any, Dict, Optional, Sequence, Tuple

import torch

from . import specaugment as _sa
from .mixup import MixupPlan, draw_mixup_plan, mixup_labels
from .preprocessing import ASTPreprocessor, create_preprocessor


class B200DataModuleMixin:
    """Attributes read from the host DataModule when present (the reference's constructor arguments, esc50.py:353-373):
    ``sample_rate``, ``num_classes``, ``time_mask`` / ``freq_mask`` (False, True = the AST defaults 192 / 48, or an int),
    ``enable_mixup``, ``mixup_alpha``, ``preprocessing_config``.  ``setup_b200`` may be called explicitly; otherwise the
    first batch builds the preprocessor from ``preprocessing_config``."""

    b200_preprocessor: Optional[ASTPreprocessor] = None
    b200_bank: Optional[torch.Tensor] = None            # un-augmented spectrograms of the training set (esc50.py:280-282)
    b200_bank_labels: Optional[torch.Tensor] = None

    def setup_b200(self, preprocessing_config: Optional[Dict[str, Any]] = None, base_cache_dir: Path = Path("/tmp/b200_cache"),
                   device=None, preprocessor: Optional[ASTPreprocessor] = None) -> ASTPreprocessor:
        if preprocessor is None:
            cfg = dict(preprocessing_config if preprocessing_config is not None else (getattr(self, "preprocessing_config", None) or {}))
            cfg.setdefault("sample_rate", int(getattr(self, "sample_rate", 44100)))
            cfg.setdefault("n_mels", int(getattr(self, "n_mels", 128)))
            preprocessor = create_preprocessor("ast", cfg, base_cache_dir)
            if device is not None:
                preprocessor._device = device
        self.b200_preprocessor = preprocessor
        return preprocessor

    def set_mixup_bank(self, spectrograms: torch.Tensor, labels: torch.Tensor) -> None:
        """``spectrograms``: ``(N, 1, n_mels, T)`` float32, the reference's ``_cached_data``; kept on the GPU."""
        self.b200_bank = spectrograms
        self.b200_bank_labels = labels

    @staticmethod
    def _mask_param(v, default: int) -> int:
        if v is True:
            return default
        return int(v) if v else 0

    def _b200_training(self) -> bool:
        tr = getattr(self, "trainer", None)
        return bool(getattr(tr, "training", False))

    def on_after_batch_transfer(self, batch, dataloader_idx: int = 0):
        """``(waveforms (B, 1, N) | (B, N), labels (B,) int | (B, C) soft)`` -> ``(spectrograms (B, 1, n_mels, T), soft
        labels (B, C))`` on the GPU.  A batch that already holds spectrograms (4-D) passes through untouched."""
        wav, labels = batch
        if not torch.is_tensor(wav) or wav.dim() == 4:
            return batch
        if self.b200_preprocessor is None:
            self.setup_b200()
        pre = self.b200_preprocessor
        if wav.dim() == 3:
            wav = wav[:, 0]
        if not wav.is_cuda:
            wav = wav.cuda(non_blocking=True)
        B = int(wav.shape[0])
        training = self._b200_training()
        num_classes = int(getattr(self, "num_classes", 50))
        sr = int(getattr(self, "sample_rate", pre.sample_rate))
        T = pre.target_frames
        if T is None:
            T = pre.frontend.num_frames(int(wav.shape[-1]), pre.frontend.rate_id(sr))
        masks = plan = None
        tm = self._mask_param(getattr(self, "time_mask", False), 192)
        fm = self._mask_param(getattr(self, "freq_mask", False), 48)
        do_sa = training and bool(tm or fm)
        do_mix = training and bool(getattr(self, "enable_mixup", False)) and self.b200_bank is not None
        if do_sa or do_mix:
            # the draws of B consecutive __getitem__ calls, in their order: per sample apply_specaugment's four randints
            # (time first, preprocessing.py:1075-1104 as called from esc50.py:267-273), then apply_mixup's coin / partner /
            # coin / Beta (esc50.py:64-76, preprocessing.py:950-958)
            rows, partner, lam = [], [], []
            for _ in range(B):
                if do_sa:
                    rows.append(_sa.reference_intervals(int(T), int(pre.n_mels), tm, fm, random))
                if do_mix:
                    one = draw_mixup_plan(1, int(self.b200_bank.shape[0]), alpha=float(getattr(self, "mixup_alpha", 0.5)), prob=0.5)
                    partner.append(one.partner)
                    lam.append(one.lam)
            if do_sa:
                masks = torch.tensor(rows, dtype=torch.int32).reshape(B, 4)
            if do_mix:
                plan = MixupPlan(torch.cat(partner), torch.cat(lam))
        mix = None if plan is None else (self.b200_bank, plan)
        spec, _ = pre.preprocess_batch(wav.to(torch.float32), sr, masks=masks, target_frames=int(T), mixup=mix)
        labels = labels.to(spec.device)
        if labels.dim() == 1:
            hard = labels.to(torch.int64)
            soft = torch.zeros((B, num_classes), dtype=torch.float32, device=spec.device)
            soft.scatter_(1, hard[:, None], 1.0)                 # create_one_hot_labels, preprocessing.py:24-36
        else:
            hard, soft = labels.argmax(1), labels.to(torch.float32)
        if plan is not None:
            soft = mixup_labels(hard, self.b200_bank_labels, plan, num_classes)
        return spec, soft
