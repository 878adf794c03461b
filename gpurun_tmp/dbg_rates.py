import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import dl_sound_classification_b200 as b2
from oracle import fbank_oracle as O
table = (22050, 44100, 48000)
fe = b2.FbankFrontend(orig_rates=table, **b2.AST_FBANK_KWARGS)
g = torch.Generator().manual_seed(5)
for rate in table:
    for secs in (0.5, 1.3, 3.7):
        n = int(rate * secs) + 7
        w = torch.rand(1, n, generator=g) * 2 - 1
        off = torch.tensor([0, n], dtype=torch.int64)
        out, nfr = fe(w[0].cuda(), out_frames=1024, offsets=off, rate_ids=torch.tensor([table.index(rate)], dtype=torch.int32))
        ref = O.kaldi_fbank(O.resample(w[0].numpy(), rate, 16000), O.ast_fbank_options())
        m = ref.shape[0]
        got = out[0, :m].cpu().numpy()
        d = np.abs(got - ref)
        fr = d.max(1)
        bad = np.nonzero(fr > 1e-3)[0]
        print(rate, secs, "frames", m, int(nfr[0]), "max", d.max(), "bad frames", bad[:10], len(bad))
