#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r2v2_bench8.json 2> gpurun_out/r2v2_bench8.err; echo "bench8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v2_bench8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_pcm16']['value'])
for k,v in d['extra'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ('workload','points')})
PY
