#!/bin/bash
# round-2, 8 GPUs: the box's copy ceiling next to the bench's e2e at N = 8, the NCCL test, the scaling line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2v_topo.txt 2>&1
timeout 120 python tools/h2d_ceiling.py > gpurun_out/r2v_ceiling_1.json 2> gpurun_out/r2v_c1.err; cat gpurun_out/r2v_ceiling_1.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/h2d_ceiling.py > gpurun_out/r2v_ceiling_8.json 2> gpurun_out/r2v_c8.err; cat gpurun_out/r2v_ceiling_8.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/h2d_ceiling.py > gpurun_out/r2v_ceiling_2.json 2> gpurun_out/r2v_c2.err; cat gpurun_out/r2v_ceiling_2.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r2v_bench8.json 2> gpurun_out/r2v_bench8.err; echo "bench8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d.get('e2e_pcm16'))
for k,v in d['extra'].items(): print(k, {a:b for a,b in v.items() if a not in ('workload','points')})
PY
timeout 300 python -m pytest tests/test_gpu_multirank.py -x -q --timeout 200 2>&1 | tail -2
