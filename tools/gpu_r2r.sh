#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for v in base cn_256 cn_512; do echo $v; if [ $v != base ]; then export B200FBANK_LIB=$PWD/tools/build/$v.so; fi
timeout 200 python bench.py --workload melspec --steps 30 2>&1 | tail -1 | cut -c200-260
timeout 100 python - <<'PY'
import torch, sys, os
sys.path.insert(0, os.getcwd())
import dl_sound_classification_b200 as b2
fe = b2.FbankFrontend(orig_rates=(44100,), **b2.AST_FBANK_KWARGS)
wav = torch.rand((1024, 220500), device="cuda") * 2 - 1
out = torch.empty((1024, 512, 128), device="cuda")
for _ in range(5): fe(wav, 512, out=out, per_clip_norm=True, return_n_frames=False)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(50): fe(wav, 512, out=out, per_clip_norm=True, return_n_frames=False)
e1.record(); torch.cuda.synchronize()
print("per_clip_norm ms", e0.elapsed_time(e1) / 50)
PY
done
