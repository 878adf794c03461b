"""Drop-in for ``torchaudio.compliance.kaldi.fbank`` (torchaudio/compliance/kaldi.py:514-645).

Same name, argument order, defaults and return shape; the computation runs in
the fused sm_100a kernel.  CPU tensors are copied to the current CUDA device
and the result is returned on the input's device -- there is no CPU compute
path.  ``dither != 0`` raises (it draws ``torch.randn``; RNG parity with the
CPU generator is impossible on the device).
"""
from __future__ import annotations

from functools import lru_cache

import torch
from torch import Tensor

from .frontend import FbankFrontend, _require_cuda

POVEY, HANNING, HAMMING, RECTANGULAR, BLACKMAN = "povey", "hanning", "hamming", "rectangular", "blackman"
WINDOWS = [HAMMING, HANNING, POVEY, RECTANGULAR, BLACKMAN]

__all__ = ["fbank", "WINDOWS"]


@lru_cache(maxsize=64)
def _frontend(device_index: int, key: tuple) -> FbankFrontend:
    kw = dict(key)
    sf = kw["sample_frequency"]
    return FbankFrontend(orig_rates=(int(sf),), device=torch.device("cuda", device_index), **kw)


def fbank(
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
    window_type: str = POVEY,
) -> Tensor:
    r"""Create a fbank from a raw audio signal (Kaldi ``compute-fbank-feats``), on a B200.

    Args and return value: identical to ``torchaudio.compliance.kaldi.fbank``;
    ``waveform`` is ``(c, n)`` and the result ``(m, num_mel_bins + use_energy)``.
    """
    if dither != 0.0:
        raise NotImplementedError("dither != 0.0 draws torch.randn per sample; RNG parity is impossible -- use dither=0.0")
    if waveform.dim() != 2:
        raise ValueError("waveform must be (c, n)")
    channel = max(channel, 0)
    assert channel < waveform.size(0), "Invalid channel {} for size {}".format(channel, waveform.size(0))
    in_device, in_dtype = waveform.device, waveform.dtype
    dev = _require_cuda(in_device if in_device.type == "cuda" else None)
    key = tuple(sorted(dict(
        blackman_coeff=float(blackman_coeff), energy_floor=float(energy_floor), frame_length=float(frame_length),
        frame_shift=float(frame_shift), high_freq=float(high_freq), htk_compat=bool(htk_compat),
        low_freq=float(low_freq), num_mel_bins=int(num_mel_bins),
        preemphasis_coefficient=float(preemphasis_coefficient), raw_energy=bool(raw_energy),
        remove_dc_offset=bool(remove_dc_offset), round_to_power_of_two=bool(round_to_power_of_two),
        sample_frequency=float(sample_frequency), snip_edges=bool(snip_edges), subtract_mean=bool(subtract_mean),
        use_energy=bool(use_energy), use_log_fbank=bool(use_log_fbank), use_power=bool(use_power),
        vtln_high=float(vtln_high), vtln_low=float(vtln_low), vtln_warp=float(vtln_warp),
        window_type=str(window_type)).items()))
    if float(int(sample_frequency)) != float(sample_frequency):
        raise NotImplementedError("non-integer sample_frequency")
    fe = _frontend(dev.index, key)
    wave = waveform[channel, :]
    n = int(wave.numel())
    # torchaudio/compliance/kaldi.py:142
    assert 2 <= fe.plan.window_size <= n, "choose a window size {} that is [2, {}]".format(fe.plan.window_size, n)
    if n < min_duration * sample_frequency:
        return torch.empty(0, device=in_device, dtype=in_dtype)       # kaldi.py:595-597
    m = fe.num_frames(n)
    if m == 0:
        return torch.empty((0, fe.n_cols), device=in_device, dtype=in_dtype)
    out, _ = fe(wave.reshape(1, n).to(torch.float32), out_frames=m, return_n_frames=False)
    out = out[0]
    if in_device != out.device or in_dtype != out.dtype:
        out = out.to(device=in_device, dtype=in_dtype)
    return out
