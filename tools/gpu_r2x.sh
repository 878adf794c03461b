#!/bin/bash
# round-2 final evidence: bench line, ncu launch list, full capture of the fbank kernel (dense) and of the dynamic persistent launch (application replay)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2x_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2x_ncu1.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fbank_ws -s 3 -c 1 -o gpurun_out/r02b_ws_full python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2x_ncu2.log 2>&1; echo "ncu ws rc=$?"
cat > /tmp/us8k_run.py <<'PY'
import sys, os, random, torch
sys.path.insert(0, os.getcwd())
import dl_sound_classification_b200 as b2
dev = torch.device("cuda:0")
B8, table = 4096, (22050, 44100, 48000)
g = torch.Generator().manual_seed(31)
rid = torch.randint(0, 3, (B8,), generator=g)
lens = ((1.0 + 3.0 * torch.rand(B8, generator=g)) * torch.tensor(table)[rid]).long()
offsets = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)]).to(dev)
gen = torch.Generator(device=dev).manual_seed(77)
flat = torch.rand(int(offsets[-1]), generator=gen, device=dev) * 2 - 1
fe8 = b2.FbankFrontend(orig_rates=table, device=dev, **b2.AST_FBANK_KWARGS)
random.seed(77)
masks = b2.specaugment.draw_masks(B8, 1024, 128, 192, 48).to(dev)
out8 = torch.empty((B8, 1024, 128), device=dev)
for _ in range(3):
    fe8(flat, 1024, offsets=offsets, rate_ids=rid.int().to(dev), masks=masks, mean=-6.6268, std=5.0613, out=out8, return_n_frames=False)
torch.cuda.synchronize()
print("us8k ok", float(out8.sum()))
PY
timeout 600 ncu --replay-mode application --set full --clock-control none -k regex:fbank_ws -s 2 -c 1 -o gpurun_out/r02b_ws_us8k_full python /tmp/us8k_run.py > gpurun_out/r2x_ncu3.log 2>&1; echo "ncu us8k rc=$?"; tail -3 gpurun_out/r2x_ncu3.log
ls -la gpurun_out | grep r02b
