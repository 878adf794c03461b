"""Kernel experiment driver (GPU box): time the fused kernel of the library named by $B200FBANK_LIB and report its
distance from the golden torchaudio features.

usage: B200FBANK_LIB=/root/repo/tools/build/libX.so python tools/ktime.py [--us8k] [--iters 200] [--tag NAME]
Prints one line: tag, ms per 1024-clip launch, roofline fraction, max-abs vs golden (config1) and vs live torchaudio.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

import dl_sound_classification_b200 as b2
from inputs import config1_clips, us8k_small_clips
from parity import logmel_err

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--tag", default=os.path.basename(os.environ.get("B200FBANK_LIB", "default")))
ap.add_argument("--us8k", action="store_true")
ap.add_argument("--no-parity", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
res = dict(tag=args.tag)

fe = b2.FbankFrontend(orig_rates=(44100,), device=dev, **b2.AST_FBANK_KWARGS)
if not args.no_parity:
    g = np.load(os.path.join(ROOT, "tests", "golden", "config1.npz"))
    clips = config1_clips(40)
    out, nfr = fe(torch.cat(clips, 0).to(dev), out_frames=512)
    got = out.cpu().numpy()
    worst = 0.0
    for j, i in enumerate(g["full_idx"]):
        worst = max(worst, logmel_err(got[int(i), :498], g["full"][j])[0])
    res["golden_maxabs"] = worst
    try:
        import torchaudio.compliance.kaldi as kaldi
        import torchaudio.transforms as T
        rs = T.Resample(44100, 16000)
        errs = []
        for i in range(0, 40, 3):
            ref = kaldi.fbank(rs(clips[i]), htk_compat=True, sample_frequency=16000, use_energy=False, window_type="hanning",
                              num_mel_bins=128, dither=0.0, frame_shift=10).numpy()
            errs.append(logmel_err(got[i, :498], ref)[0])
        res["live_maxabs"] = max(errs)
        clips8, rates8 = us8k_small_clips(9)
        table = (22050, 44100, 48000)
        fe8 = b2.FbankFrontend(orig_rates=table, device=dev, **b2.AST_FBANK_KWARGS)
        lens = torch.tensor([c.shape[1] for c in clips8])
        offs = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)])
        rid = torch.tensor([table.index(r) for r in rates8], dtype=torch.int32)
        o8, n8 = fe8(torch.cat([c[0] for c in clips8]).to(dev), 1024, offsets=offs, rate_ids=rid)
        o8 = o8.cpu().numpy()
        e8 = []
        for i, (c, r) in enumerate(zip(clips8, rates8)):
            ref = kaldi.fbank(T.Resample(r, 16000)(c), htk_compat=True, sample_frequency=16000, use_energy=False,
                              window_type="hanning", num_mel_bins=128, dither=0.0, frame_shift=10).numpy()
            e8.append(logmel_err(o8[i, :ref.shape[0]], ref)[0])
        res["us8k_live_maxabs"] = max(e8)
    except Exception as e:          # noqa: BLE001
        res["live_error"] = repr(e)[:120]

B = 1024
gen = torch.Generator(device=dev).manual_seed(1234)
wav = torch.rand((B, 220500), generator=gen, device=dev) * 2 - 1
out = torch.empty((B, 512, 128), device=dev)
mean, std = torch.tensor([-6.6268], device=dev), torch.tensor([5.0613], device=dev)
for _ in range(5):
    fe(wav, out_frames=512, mean=mean, std=std, out=out, return_n_frames=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.iters):
    fe(wav, out_frames=512, mean=mean, std=std, out=out, return_n_frames=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.iters
res["esc50_ms"] = ms
res["frac"] = B * (220500 * 4 + 512 * 128 * 4) / (ms * 1e-3) / 6537.6e9
res["checksum"] = float(out[0, :498].double().sum())
if args.us8k:
    import random
    B8, table = 4096, (22050, 44100, 48000)
    gg = torch.Generator().manual_seed(31)
    rid = torch.randint(0, 3, (B8,), generator=gg)
    lens = ((1.0 + 3.0 * torch.rand(B8, generator=gg)) * torch.tensor(table)[rid]).long()
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)]).to(dev)
    flat = torch.rand(int(offsets[-1]), generator=gen, device=dev) * 2 - 1
    fe8 = b2.FbankFrontend(orig_rates=table, device=dev, **b2.AST_FBANK_KWARGS)
    random.seed(77)
    masks = b2.specaugment.draw_masks(B8, 1024, 128, 192, 48).to(dev)
    rid_d = rid.int().to(dev)
    out8 = torch.empty((B8, 1024, 128), device=dev)
    f = lambda: fe8(flat, 1024, offsets=offsets, rate_ids=rid_d, masks=masks, mean=mean, std=std, out=out8, return_n_frames=False)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    res["us8k_ms"] = e0.elapsed_time(e1) / 20
print("KTIME " + json.dumps(res), flush=True)
