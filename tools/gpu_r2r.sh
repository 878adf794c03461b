#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 300 python -m pytest tests/test_gpu_mixup.py tests/test_gpu_integration.py -x -q --timeout 120 2>&1 | tail -2
timeout 400 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2u_bench.json').read())
print(d['ms_per_step'], d['roofline']['frac'])
for k in ('mixup_fused','us8k'):
    v=d['extra'][k]; print(k, {a:b for a,b in v.items() if a not in ('workload','points')})
PY
