"""SpecAugment interval replay (H10a / H10b of SURVEY.md section 8a).

The fused kernel takes mask intervals as INPUT; the random draws stay on the host and
consume the very same generators, in the very same order, as the reference:

* ``reference_intervals``  -- ``ASTPreprocessor.apply_specaugment``
  (src/datasets/preprocessing.py:1075-1104): Python ``random.randint``, time mask first.
* ``torchaudio_intervals`` -- legacy ``SpecAugment`` (src/utils/audio.py:90-103) ->
  torchaudio ``mask_along_axis`` (torchaudio/functional/functional.py:885-958):
  ``torch.rand(1)`` on the default CPU generator, two draws per axis, time then frequency.

Masks are therefore bit-exact per seed by construction.
"""
from __future__ import annotations

import random as _random
from typing import Optional, Sequence, Tuple

import torch


def reference_intervals(n_frames: int, n_mels: int, time_mask: int = 192, freq_mask: int = 48,
                        rng=_random) -> Tuple[int, int, int, int]:
    """One clip's ``(t_start, t_len, f_start, f_len)``; a length of 0 means "no mask".
    ``rng`` is the ``random`` module (default, what the reference uses) or a ``random.Random``."""
    t0 = tl = f0 = fl = 0
    if time_mask > 0 and n_frames > time_mask:
        tl = rng.randint(1, min(time_mask, n_frames // 4))
        t0 = rng.randint(0, n_frames - tl)
    if freq_mask > 0 and n_mels > freq_mask:
        fl = rng.randint(1, min(freq_mask, n_mels // 4))
        f0 = rng.randint(0, n_mels - fl)
    return t0, tl, f0, fl


def _torchaudio_axis(mask_param: int, size: int, generator: Optional[torch.Generator]) -> Tuple[int, int]:
    if mask_param < 1:                      # _get_mask_param / early return, functional.py:926-928
        return 0, 0
    value = torch.rand(1, generator=generator) * mask_param
    min_value = torch.rand(1, generator=generator) * (size - value)
    start = int(min_value.long())
    end = int(min_value.long()) + int(value.long())
    if end - start >= mask_param:
        raise ValueError("Number of columns to be masked should be less than mask_param")
    lo, hi = max(start, 0), min(end, size)
    return (lo, hi - lo) if hi > lo else (0, 0)


def torchaudio_intervals(n_frames: int, n_mels: int, time_mask: int = 80, freq_mask: int = 32,
                         generator: Optional[torch.Generator] = None) -> Tuple[int, int, int, int]:
    t0, tl = _torchaudio_axis(time_mask, n_frames, generator)
    f0, fl = _torchaudio_axis(freq_mask, n_mels, generator)
    return t0, tl, f0, fl


def draw_masks(batch: int, n_frames, n_mels: int, time_mask: int = 192, freq_mask: int = 48,
               variant: str = "reference", rng=_random, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """``(B, 4)`` int32 mask table for ``FbankFrontend.__call__(masks=...)``; clip ``i`` gets the
    intervals the reference would draw on its ``i``-th sequential call.  ``n_frames`` is an int or
    a per-clip sequence (the spectrogram's time dimension as the reference sees it)."""
    rows = []
    for i in range(batch):
        nf = int(n_frames[i]) if not isinstance(n_frames, int) else n_frames
        if variant == "reference":
            rows.append(reference_intervals(nf, n_mels, time_mask, freq_mask, rng))
        elif variant == "torchaudio":
            rows.append(torchaudio_intervals(nf, n_mels, time_mask, freq_mask, generator))
        else:
            raise ValueError(f"unknown SpecAugment variant {variant!r}")
    return torch.tensor(rows, dtype=torch.int32).reshape(batch, 4)


def apply_intervals(spec: torch.Tensor, mask: Sequence[int], value: float = 0.0) -> torch.Tensor:
    """Zero-fill ``(…, F, T)`` with one interval set (clones; the input is never mutated)."""
    t0, tl, f0, fl = (int(v) for v in mask)
    out = spec.clone()
    if tl > 0:
        out[..., :, t0:t0 + tl] = value
    if fl > 0:
        out[..., f0:f0 + fl, :] = value
    return out


class SpecAugment(torch.nn.Module):
    """Mirror of src/utils/audio.py:90-103 (time masking then frequency masking, torch.rand)."""

    def __init__(self, time_mask: int = 80, freq_mask: int = 32):
        super().__init__()
        self.time_mask_param = int(time_mask)
        self.freq_mask_param = int(freq_mask)

    def forward(self, spec: torch.Tensor) -> torch.Tensor:
        n_mels, n_frames = int(spec.shape[-2]), int(spec.shape[-1])
        return apply_intervals(spec, torchaudio_intervals(n_frames, n_mels, self.time_mask_param, self.freq_mask_param))
