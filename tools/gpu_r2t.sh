#!/bin/bash
# round-2: ncu captures of the pipelined patch embedding and of the fused-Mixup frontend; final launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:patch_embed_pipe -s 3 -c 1 -o gpurun_out/r02_pe_pipe_full python bench.py --workload patch_embed --steps 10 > gpurun_out/r2t_ncu1.log 2>&1; echo "ncu pe rc=$?"
cat > /tmp/mixrun.py <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
import dl_sound_classification_b200 as b2
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
wav = torch.rand((1024, 220500), generator=g, device=dev) * 2 - 1
fe = b2.FbankFrontend(orig_rates=(44100,), device=dev, **b2.AST_FBANK_KWARGS)
bank = torch.randn((2048, 512, 128), generator=g, device=dev)
gm = torch.Generator().manual_seed(5)
plan = b2.MixupPlan(torch.randint(0, 2048, (1024,), generator=gm).int(), torch.rand(1024, generator=gm)).to(dev)
out = torch.empty((1024, 512, 128), device=dev)
for _ in range(6):
    fe(wav, 512, mean=-6.6268, std=5.0613, out=out, return_n_frames=False, mixup=(bank, plan))
torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fbank_ws -s 3 -c 1 -o gpurun_out/r02_ws_mixup_full python /tmp/mixrun.py > gpurun_out/r2t_ncu2.log 2>&1; echo "ncu mix rc=$?"
ls -la gpurun_out | tail -5
