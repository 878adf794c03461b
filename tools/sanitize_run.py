"""Small launches of every kernel form for `compute-sanitizer` (memcheck / racecheck / synccheck):
dense persistent (static striding), dynamic persistent (ragged, three rates), one item per CTA, fused Mixup,
stats epilogue, the mel-dB kernel + per-clip pass, stand-alone mixup, patch embedding.
usage: compute-sanitizer --tool memcheck python tools/sanitize_run.py [small]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dl_sound_classification_b200 as b2  # noqa: E402

small = len(sys.argv) > 1 and sys.argv[1] == "small"
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
n_dense = 150 if small else 300                    # > #SMs: the persistent form strides over the clips
wav = torch.rand((n_dense, 22050), generator=g, device=dev) * 2 - 1          # 0.5 s clips: 2 chunks each
fe = b2.FbankFrontend(orig_rates=(44100,), device=dev, **b2.AST_FBANK_KWARGS)
out, nfr = fe(wav, out_frames=64, mean=-4.27, std=4.57)
few, _ = fe(wav[:5], out_frames=64, mean=-4.27, std=4.57)                  # one item per CTA
assert torch.equal(out[:5], few)
bank = torch.randn((7, 64, 128), generator=g, device=dev)
plan = b2.MixupPlan(torch.randint(-1, 7, (n_dense,)).int(), torch.rand(n_dense)).to(dev)
mixed, _ = fe(wav, out_frames=64, mean=-4.27, std=4.57, mixup=(bank, plan))
assert torch.equal(mixed, b2.mixup_batch(out, bank, plan))
sums = torch.zeros(257, dtype=torch.float64, device=dev)
fe.accumulate_stats(wav, sums, max_frames=64)
# ragged, three rates: dynamic persistent form
table = (22050, 44100, 48000)
B = 180 if small else 400
cg = torch.Generator().manual_seed(2)
rid = torch.randint(0, 3, (B,), generator=cg)
lens = ((0.2 + 0.5 * torch.rand(B, generator=cg)) * torch.tensor(table)[rid]).long()
offsets = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)])
flat = torch.rand(int(offsets[-1]), generator=g, device=dev) * 2 - 1
fe3 = b2.FbankFrontend(orig_rates=table, device=dev, **b2.AST_FBANK_KWARGS)
o3, n3 = fe3(flat, out_frames=80, offsets=offsets, rate_ids=rid.int(), mean=-4.27, std=4.57, layout="bft")
# mel-dB recipe + per-clip pass
fem = b2.MelSpecFrontend(44100, 1024, 160, 400, 128, 80.0, device=dev)
om, nm = fem(wav[:40], out_frames=140)
# patch embedding
w = torch.randn(192, 1, 16, 16, generator=g, device=dev) * 0.05
pe = b2.patch_embed(torch.randn(3, 1, 128, 100, generator=g, device=dev), w, torch.zeros(192, device=dev), stride=10)
torch.cuda.synchronize()
print("sanitize_run ok:", float(out.sum()), float(o3.sum()), float(om.sum()), float(sums[256]), float(pe.float().sum()), b2.launch_count())
