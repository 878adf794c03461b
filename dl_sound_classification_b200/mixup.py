"""On-device Mixup with host-replayed random draws.

The reference mixes one sample at a time inside ``Dataset.__getitem__``:

* ``MixupDataset.apply_mixup`` (src/datasets/esc50.py:43-76): with probability 1/2 the sample is
  left alone (``random.random() > 0.5``); otherwise a partner index is drawn from the whole
  spectrogram bank (``random.randint(0, N - 1)``) and handed to
* ``MixupAugmentation.__call__`` (src/datasets/preprocessing.py:933-968), constructed with
  ``prob=0.5`` (esc50.py:51): a second coin (``random.random() > self.prob``) may still skip the
  mix; otherwise ``lam ~ Beta(alpha, alpha)`` from ``torch.distributions`` on the default CPU
  generator, ``mixed = lam * spec1 + (1 - lam) * spec2``, soft labels ``[label1] = lam`` then
  ``[label2] = 1 - lam``.

Here the draws stay on the host and consume the very same generators in the very same order
(``draw_mixup_plan``), so a batch is mixed exactly as the per-sample loop would mix it, and the
arithmetic runs as one CUDA kernel over the batch (``b200fbank_mixup``), bit-identical to the
reference's float32 tensor expression.  There is no CPU compute path.
"""
from __future__ import annotations

import random as _random
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _capi as K


@dataclass
class MixupPlan:
    """Who is mixed with whom: ``partner[i]`` = index into the bank, or -1 (sample left alone);
    ``lam[i]`` = float32 mixing coefficient (1.0 where unmixed)."""
    partner: torch.Tensor      # int32 [B]
    lam: torch.Tensor          # float32 [B]

    def to(self, device) -> "MixupPlan":
        return MixupPlan(self.partner.to(device), self.lam.to(device))


def draw_mixup_plan(batch_size: int, bank_size: int, alpha: float = 0.5, prob: float = 0.5,
                    enable_mixup: bool = True, rng=_random) -> MixupPlan:
    """Replay of the reference's per-sample draws for ``batch_size`` consecutive ``__getitem__`` calls
    (esc50.py:64-76, preprocessing.py:950-958).  ``rng`` is the ``random`` module (what the reference
    uses) or a ``random.Random``; lambda comes from ``torch.distributions.Beta`` on the default generator."""
    partner = torch.full((batch_size,), -1, dtype=torch.int32)
    lam = torch.ones(batch_size, dtype=torch.float32)
    if not enable_mixup:
        return MixupPlan(partner, lam)
    beta = torch.distributions.Beta(alpha, alpha) if alpha > 0 else None
    for i in range(batch_size):
        if rng.random() > 0.5:                         # esc50.py:64
            continue
        other = rng.randint(0, bank_size - 1)          # esc50.py:70
        if rng.random() > prob:                        # preprocessing.py:950
            continue
        partner[i] = other
        lam[i] = beta.sample() if beta is not None else 1.0   # preprocessing.py:955-958
    return MixupPlan(partner, lam)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def mixup_batch(x: torch.Tensor, bank: torch.Tensor, plan: MixupPlan, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``out[i] = lam[i] * x[i] + (1 - lam[i]) * bank[partner[i]]`` (``x[i]`` where ``partner[i] < 0``).
    ``x``: ``(B, ...)`` float32 CUDA, ``bank``: ``(N, ...)`` with the same trailing shape; ``out`` may be ``x``."""
    if not x.is_cuda:
        raise RuntimeError("mixup_batch needs CUDA tensors: dl_sound_classification_b200 has no CPU fallback")
    if x.dtype != torch.float32 or bank.dtype != torch.float32:
        raise TypeError("mixup_batch expects float32 spectrograms")
    if tuple(x.shape[1:]) != tuple(bank.shape[1:]):
        raise ValueError(f"spectrogram shapes differ: {tuple(x.shape[1:])} vs {tuple(bank.shape[1:])}")
    x = x.contiguous()
    bank = bank.to(x.device).contiguous()
    B = x.shape[0]
    if int(plan.partner.shape[0]) != B:
        raise ValueError("plan size does not match the batch")
    partner = plan.partner.to(device=x.device, dtype=torch.int32).contiguous()
    if B and int(partner.max()) >= bank.shape[0]:
        raise IndexError("partner index outside the bank")
    lam = plan.lam.to(device=x.device, dtype=torch.float32).contiguous()
    if out is None:
        out = torch.empty_like(x)
    elif out.shape != x.shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != x.device:
        raise ValueError("out must be a contiguous float32 tensor shaped like x on the same device")
    clip = x[0].numel() if B else 1
    with torch.cuda.device(x.device):
        K.check(K.lib.b200fbank_mixup(_ptr(x), _ptr(bank), _ptr(partner), _ptr(lam), B, clip, _ptr(out),
                                      torch.cuda.current_stream().cuda_stream))
    return out


def mixup_labels(labels: torch.Tensor, bank_labels: torch.Tensor, plan: MixupPlan, num_classes: int) -> torch.Tensor:
    """Soft labels ``(B, num_classes)`` of the same plan (preprocessing.py:960-966, 39-52)."""
    dev = labels.device
    if dev.type != "cuda":
        raise RuntimeError("mixup_labels needs CUDA tensors: dl_sound_classification_b200 has no CPU fallback")
    labels = labels.to(torch.int64).contiguous()
    partner = plan.partner.to(device=dev, dtype=torch.int32).contiguous()
    plabel = bank_labels.to(device=dev, dtype=torch.int64)[partner.clamp(min=0).long()].contiguous()
    lam = plan.lam.to(device=dev, dtype=torch.float32).contiguous()
    B = labels.shape[0]
    soft = torch.empty((B, num_classes), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        K.check(K.lib.b200fbank_mixup_labels(_ptr(labels), _ptr(plabel), _ptr(partner), _ptr(lam), B, int(num_classes),
                                             _ptr(soft), torch.cuda.current_stream().cuda_stream))
    return soft


class MixupAugmentation:
    """Mirror of the reference class (src/datasets/preprocessing.py:928-968), same constructor and call
    signature; the mix itself runs on the GPU (inputs are moved there, the result comes back on the
    device of ``spec1``)."""

    def __init__(self, alpha: float = 0.5, prob: float = 1.0):
        self.alpha = alpha
        self.prob = prob

    def __call__(self, spec1: torch.Tensor, spec2: torch.Tensor, label1: int, label2: int,
                 num_classes: int) -> Tuple[torch.Tensor, torch.Tensor]:
        if _random.random() > self.prob:
            soft = torch.zeros(num_classes, dtype=torch.float32)
            soft[label1] = 1.0
            return spec1, soft
        lam = torch.distributions.Beta(self.alpha, self.alpha).sample() if self.alpha > 0 else torch.tensor(1.0)
        plan = MixupPlan(torch.zeros(1, dtype=torch.int32), lam.reshape(1).float())
        dev = spec1.device
        cuda = dev if dev.type == "cuda" else torch.device("cuda")
        mixed = mixup_batch(spec1.to(cuda)[None], spec2.to(cuda)[None], plan)[0].to(dev)
        soft = mixup_labels(torch.tensor([label1], device=cuda), torch.tensor([label2], device=cuda), plan, num_classes)[0]
        return mixed, soft.to(dev)
