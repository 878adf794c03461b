"""Seeded synthetic inputs shared by the golden generator, the tests, smoke() and bench.py.

torch's CPU generator is deterministic across machines, so fixtures never store inputs.
"""
import torch


def config1_clips(n=40, seed=1234, length=220500):
    """``torch.manual_seed(1234)`` then sequential
    ``torch.rand(1, 220500) * 2 - 1`` draws (uniform broadband noise in [-1, 1])."""
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(1, length, generator=g) * 2 - 1 for _ in range(n)]


def short_clip(length, seed=4321):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(1, length, generator=g) * 2 - 1


def us8k_small_clips(n=9, seed=808):
    """UrbanSound8K-shaped ragged clips: 1-4 s at 22.05 / 44.1 / 48 kHz."""
    g = torch.Generator().manual_seed(seed)
    rates_all = (22050, 44100, 48000)
    clips, rates = [], []
    for i in range(n):
        r = rates_all[int(torch.randint(0, 3, (1,), generator=g))]
        dur = 1.0 + 3.0 * float(torch.rand(1, generator=g))
        length = int(dur * r)
        clips.append(torch.rand(1, length, generator=g) * 2 - 1)
        rates.append(r)
    return clips, rates
