"""BASELINE.json full sizes, EVERY clip against live torchaudio (the reference's own CPU path, run on the GPU box's host
cores): configs[1] = 1024 ESC-50 clips, configs[2] = 4096 ragged US8K-shaped clips at three rates.  Also measures both
float32 implementations against a float64 evaluation of the same torchaudio code ("truth") on a sample.

The measured numbers are printed in the terminal summary (tests/conftest.py) and written to
gpurun_out/parity_fullsize.json; profiles/README.md carries the table of the last run.
"""
import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest
import torch

from parity import BROADBAND_FLOOR_FRAC, LOGMEL_TOL, assert_logmel_close, logmel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = {}


@pytest.fixture(scope="module")
def b2():
    import dl_sound_classification_b200 as m
    assert torch.cuda.is_available()
    return m


@pytest.fixture(scope="module")
def ta():
    pytest.importorskip("torchaudio")
    import torchaudio.compliance.kaldi as kaldi
    import torchaudio.transforms as T
    return kaldi, T


def _pool_map(fn, items):
    n = min(16, os.cpu_count() or 1)
    old = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        with ThreadPoolExecutor(n) as ex:
            return list(ex.map(fn, items))
    finally:
        torch.set_num_threads(old)


def _save():
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_fullsize.json"), "w") as f:
            json.dump(REPORT, f, indent=1)
    except OSError:
        pass


def test_config2_all_1024_clips_vs_live_torchaudio(b2, ta):
    kaldi, T = ta
    B, N = 1024, 220500
    fe = b2.FbankFrontend(orig_rates=(44100,), **b2.AST_FBANK_KWARGS)
    gen = torch.Generator(device="cuda").manual_seed(2024)
    wav = torch.rand((B, N), generator=gen, device="cuda") * 2 - 1
    out, nfr = fe(wav, out_frames=498)
    assert bool((nfr == 498).all())
    got = out.cpu().numpy()
    host = wav.cpu()
    rs = T.Resample(44100, 16000)

    def ref_clip(i):
        return kaldi.fbank(rs(host[i:i + 1]), **b2.AST_FBANK_KWARGS).numpy()

    ref = np.stack(_pool_map(ref_clip, range(B)))
    main = assert_logmel_close(got, ref, LOGMEL_TOL, "configs[1]: all 1024 clips vs live torchaudio", BROADBAND_FLOOR_FRAC)
    _, floor, n_floor = logmel_err(got, ref)
    REPORT["config2_44100"] = {"clips": B, "max_abs": main, "near_floor_max_abs": floor, "near_floor_cells": n_floor, "cells": int(got.size)}

    # float64 evaluation of the same torchaudio code on a sample: neither float32 path may sit further from it than the bar
    rs64 = T.Resample(44100, 16000, dtype=torch.float64)
    idx = list(range(0, B, 32))

    def truth_clip(i):
        return kaldi.fbank(rs64(host[i:i + 1].double()), **b2.AST_FBANK_KWARGS).numpy()

    truth = np.stack(_pool_map(truth_clip, idx))
    ours64, _, _ = logmel_err(got[idx], truth)
    ta64, _, _ = logmel_err(ref[idx], truth)
    REPORT["config2_vs_fp64"] = {"clips": len(idx), "ours_max_abs": ours64, "torchaudio_fp32_max_abs": ta64}
    _save()
    # measured: torchaudio's own float32 path sits 1.2e-3 from its float64 evaluation on its worst cell, this kernel
    # 1.3e-3 -- the 1e-3 bar is a bar on the DISTANCE BETWEEN the two float32 paths (checked above, 7e-4), which neither
    # can hold against exact arithmetic; against truth the kernel must be no worse than the reference's own path (+25 %)
    assert ours64 <= 1.25 * ta64 + 1e-4, f"CUDA path vs float64 truth: {ours64:.2e}, torchaudio float32 vs the same: {ta64:.2e}"


def test_config3_all_4096_ragged_clips_vs_live_torchaudio(b2, ta):
    kaldi, T = ta
    B = 4096
    table = (22050, 44100, 48000)
    g = torch.Generator().manual_seed(31)
    rid = torch.randint(0, 3, (B,), generator=g)
    dur = 1.0 + 3.0 * torch.rand(B, generator=g)
    lens = (dur * torch.tensor(table)[rid]).long()
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)])
    gen = torch.Generator(device="cuda").manual_seed(32)
    flat = torch.rand(int(offsets[-1]), generator=gen, device="cuda") * 2 - 1
    fe = b2.FbankFrontend(orig_rates=table, **b2.AST_FBANK_KWARGS)
    out, nfr = fe(flat, out_frames=400, offsets=offsets, rate_ids=rid.int())
    got = out.cpu().numpy()
    nfr = nfr.cpu().numpy()
    host = flat.cpu()
    rs = [T.Resample(r, 16000) for r in table]

    def ref_clip(i):
        w = host[int(offsets[i]):int(offsets[i + 1])][None]
        return kaldi.fbank(rs[int(rid[i])](w), **b2.AST_FBANK_KWARGS).numpy()

    refs = _pool_map(ref_clip, range(B))
    for r, rate in enumerate(table):
        sel = [i for i in range(B) if int(rid[i]) == r]
        assert all(refs[i].shape[0] == nfr[i] for i in sel), f"frame counts at {rate} Hz"
        a = np.concatenate([got[i, :nfr[i]] for i in sel])
        b = np.concatenate([refs[i] for i in sel])
        main = assert_logmel_close(a, b, LOGMEL_TOL, f"configs[2]: all {len(sel)} clips at {rate} Hz vs live torchaudio", BROADBAND_FLOOR_FRAC)
        _, floor, n_floor = logmel_err(a, b)
        REPORT[f"config3_{rate}"] = {"clips": len(sel), "max_abs": main, "near_floor_max_abs": floor, "near_floor_cells": n_floor, "cells": int(a.size)}
    assert all((got[i, nfr[i]:] == 0).all() for i in range(0, B, 64))         # pad rows
    _save()
