"""``torch.ops.b200fbank.*``: the frontend as torch custom operators (``torch.library``).

north_star: "exposed to Python as a torch custom op with the same signature and defaults as the reference transform
and ``torchaudio.compliance.kaldi.fbank``".  Three operators, each a thin schema over the C-ABI-backed Python entry
points of this package (plans are cached per configuration and device); all of them need CUDA tensors' device to be a
B200 and have no CPU compute path:

* ``b200fbank::kaldi_fbank(waveform, blackman_coeff=0.42, channel=-1, dither=0.0, ...) -> Tensor``
  -- argument names, order and defaults of ``torchaudio/compliance/kaldi.py:514-541``;
* ``b200fbank::ast_frontend(wav, sample_rate, out_frames, mean, std, target_mean=0.0, target_std=0.5, masks=None)
  -> (features (B, 1, 128, T), n_frames (B))`` -- the batched AST recipe (``ASTPreprocessor`` + SpecAugment masks +
  normalisation, src/datasets/preprocessing.py:971-1104);
* ``b200fbank::mixup(x, bank, partner, lam) -> Tensor`` -- src/datasets/preprocessing.py:933-968.

Each has a fake (meta) implementation, so the ops trace under ``torch.compile`` / ``torch.export`` shape propagation.
"""
from __future__ import annotations

from functools import lru_cache
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import kaldi as _kaldi
from . import mixup as _mixup
from .frontend import AST_FBANK_KWARGS, FbankFrontend, _require_cuda

__all__ = ["kaldi_fbank", "ast_frontend", "mixup"]


@torch.library.custom_op("b200fbank::kaldi_fbank", mutates_args=())
def kaldi_fbank(
    waveform: Tensor,
    blackman_coeff: float = 0.42,
    channel: int = -1,
    dither: float = 0.0,
    energy_floor: float = 1.0,
    frame_length: float = 25.0,
    frame_shift: float = 10.0,
    high_freq: float = 0.0,
    htk_compat: bool = False,
    low_freq: float = 20.0,
    min_duration: float = 0.0,
    num_mel_bins: int = 23,
    preemphasis_coefficient: float = 0.97,
    raw_energy: bool = True,
    remove_dc_offset: bool = True,
    round_to_power_of_two: bool = True,
    sample_frequency: float = 16000.0,
    snip_edges: bool = True,
    subtract_mean: bool = False,
    use_energy: bool = False,
    use_log_fbank: bool = True,
    use_power: bool = True,
    vtln_high: float = -500.0,
    vtln_low: float = 100.0,
    vtln_warp: float = 1.0,
    window_type: str = "povey",
) -> Tensor:
    return _kaldi.fbank(waveform, blackman_coeff, channel, dither, energy_floor, frame_length, frame_shift, high_freq,
                        htk_compat, low_freq, min_duration, num_mel_bins, preemphasis_coefficient, raw_energy,
                        remove_dc_offset, round_to_power_of_two, sample_frequency, snip_edges, subtract_mean, use_energy,
                        use_log_fbank, use_power, vtln_high, vtln_low, vtln_warp, window_type)


@kaldi_fbank.register_fake
def _(waveform, blackman_coeff=0.42, channel=-1, dither=0.0, energy_floor=1.0, frame_length=25.0, frame_shift=10.0,
      high_freq=0.0, htk_compat=False, low_freq=20.0, min_duration=0.0, num_mel_bins=23, preemphasis_coefficient=0.97,
      raw_energy=True, remove_dc_offset=True, round_to_power_of_two=True, sample_frequency=16000.0, snip_edges=True,
      subtract_mean=False, use_energy=False, use_log_fbank=True, use_power=True, vtln_high=-500.0, vtln_low=100.0,
      vtln_warp=1.0, window_type="povey"):
    n = waveform.shape[1]
    shift, size = int(sample_frequency * frame_shift * 0.001), int(sample_frequency * frame_length * 0.001)
    m = (0 if n < size else 1 + (n - size) // shift) if snip_edges else (n + shift // 2) // shift     # kaldi.py:63-69
    return waveform.new_empty((m, num_mel_bins + int(use_energy)))


@lru_cache(maxsize=16)
def _ast_frontend(device_index: int, sample_rate: int) -> FbankFrontend:
    return FbankFrontend(orig_rates=(int(sample_rate),), device=torch.device("cuda", device_index), **AST_FBANK_KWARGS)


@torch.library.custom_op("b200fbank::ast_frontend", mutates_args=())
def ast_frontend(wav: Tensor, sample_rate: int, out_frames: int, mean: float, std: float, target_mean: float = 0.0,
                 target_std: float = 0.5, masks: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    dev = _require_cuda(wav.device if wav.is_cuda else None)
    fe = _ast_frontend(dev.index if dev.index is not None else torch.cuda.current_device(), int(sample_rate))
    out, nfr = fe(wav.to(dev), out_frames=int(out_frames), masks=masks, mean=float(mean), std=float(std),
                  target_mean=float(target_mean), target_std=float(target_std), layout="bft")
    return out, nfr


@ast_frontend.register_fake
def _(wav, sample_rate, out_frames, mean, std, target_mean=0.0, target_std=0.5, masks=None):
    B = wav.shape[0]
    return wav.new_empty((B, 1, 128, out_frames)), wav.new_empty((B,), dtype=torch.int32)


@torch.library.custom_op("b200fbank::mixup", mutates_args=())
def mixup(x: Tensor, bank: Tensor, partner: Tensor, lam: Tensor) -> Tensor:
    return _mixup.mixup_batch(x, bank, _mixup.MixupPlan(partner, lam))


@mixup.register_fake
def _(x, bank, partner, lam):
    return torch.empty_like(x)
