import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import dl_sound_classification_b200 as b2
from oracle import fbank_oracle as O
from inputs import us8k_small_clips
g = np.load("tests/golden/us8k_small.npz")
clips, rates = us8k_small_clips(9)
table = (22050, 44100, 48000)
fe = b2.FbankFrontend(orig_rates=table, **b2.AST_FBANK_KWARGS)
flat = torch.cat([c[0] for c in clips]).cuda()
lens = torch.tensor([c.shape[1] for c in clips])
offsets = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)])
rid = torch.tensor([table.index(r) for r in rates], dtype=torch.int32)
out, nfr = fe(flat, out_frames=1024, offsets=offsets, rate_ids=rid)
got = out.cpu().numpy()
for i in range(9):
    m = int(g["n_frames"][i])
    d = np.abs(got[i, :m] - g["feats"][i, :m])
    well = g["feats"][i, :m] > O.LOG_FLT_EPSILON + 3
    dd = np.where(well, d, 0)
    t, b = np.unravel_index(dd.argmax(), dd.shape)
    ref = O.kaldi_fbank(O.resample(clips[i][0].numpy(), rates[i], 16000), O.ast_fbank_options())
    print(i, rates[i], lens[i].item(), int(offsets[i]) % 4, "frames", m, "max", dd.max(), "at frame", t, "bin", b, "golden", g["feats"][i, t, b], "got", got[i, t, b], "oracle", ref[t, b],
          "n>5e-4:", int((dd > 5e-4).sum()))
# single clip 0 alone
w = clips[0][0]
o1, _ = fe(w.cuda(), out_frames=1024, offsets=torch.tensor([0, w.numel()]), rate_ids=torch.tensor([2], dtype=torch.int32))
m = int(g["n_frames"][0])
print("alone vs batch identical:", np.array_equal(o1[0, :m].cpu().numpy(), got[0, :m]))
