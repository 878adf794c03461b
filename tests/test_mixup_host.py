"""CPU tests of the Mixup row: the oracle against the golden vectors made by the
reference's own MixupAugmentation (tests/golden/make_golden_mixup.py), and the host-side replay of the draws."""
import os
import random

import numpy as np
import torch

from oracle import fbank_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
B, N, C, SHAPE = 24, 37, 50, (1, 16, 40)


def mixup_inputs():
    g = torch.Generator().manual_seed(2024)
    bank = torch.randn((N,) + SHAPE, generator=g) * 3 - 4
    bank_labels = torch.randint(0, C, (N,), generator=g)
    x = torch.randn((B,) + SHAPE, generator=g) * 3 - 4
    labels = torch.randint(0, C, (B,), generator=g)
    labels[3] = bank_labels[0]
    return x, labels, bank, bank_labels


def golden():
    return np.load(os.path.join(HERE, "golden", "mixup.npz"))


def test_oracle_replays_the_reference_draws_and_arithmetic():
    g = golden()
    x, labels, bank, bank_labels = mixup_inputs()
    lams = [float(v) for v, p in zip(g["lam"], g["partner"]) if p >= 0]
    partner, lam = O.mixup_plan_reference(O.PyRandom(31337), lams, B, N, prob=0.5)
    assert (partner == g["partner"]).all() and (lam == g["lam"]).all()
    assert (partner >= 0).sum() >= 5 and (partner < 0).sum() >= 5
    for i in range(B):
        if partner[i] >= 0:
            got = O.mixup(x[i].numpy(), bank[partner[i]].numpy(), lam[i])
            soft = O.mixup_soft_labels(int(labels[i]), int(bank_labels[partner[i]]), lam[i], C)
        else:
            got = x[i].numpy()
            soft = O.mixup_soft_labels(int(labels[i]), -1, 1.0, C, mixed=False)
        assert np.array_equal(got, g["out"][i]), i                   # bit-exact: three float32 roundings per element
        assert np.array_equal(soft, g["soft"][i]), i


def test_host_plan_consumes_the_generators_like_the_reference():
    import dl_sound_classification_b200 as b2
    g = golden()
    random.seed(31337)
    torch.manual_seed(31337)
    plan = b2.draw_mixup_plan(B, N, alpha=0.5, prob=0.5)
    assert plan.partner.dtype == torch.int32 and plan.lam.dtype == torch.float32
    assert (plan.partner.numpy() == g["partner"]).all()
    assert (plan.lam.numpy() == g["lam"]).all()
    # the next draws continue where the reference's would
    nxt = random.random()
    random.seed(31337)
    torch.manual_seed(31337)
    b2.draw_mixup_plan(B, N, alpha=0.5, prob=0.5)
    assert random.random() == nxt
    off = b2.draw_mixup_plan(5, N, enable_mixup=False)
    assert (off.partner == -1).all() and (off.lam == 1).all()


def test_mixup_has_no_cpu_path():
    import dl_sound_classification_b200 as b2
    import pytest
    x, labels, bank, bank_labels = mixup_inputs()
    plan = b2.MixupPlan(torch.zeros(B, dtype=torch.int32), torch.full((B,), 0.5))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b2.mixup_batch(x, bank, plan)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b2.mixup_labels(labels, bank_labels, plan, C)


def test_oracle_python_random_port_matches_cpython():
    """The oracle's MT19937 port must replay CPython's random.random() / randint() stream bit for bit."""
    for seed in (0, 1, 31337, 2 ** 40 + 7):
        random.seed(seed)
        rng = O.PyRandom(seed)
        for _ in range(50):
            assert O.py_random_float(rng) == random.random()
            assert rng.randint(0, 1999) == random.randint(0, 1999)
