#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 300 python -m pytest tests/test_gpu_clip_norm.py tests/test_gpu_integration.py tests/test_gpu_cache.py -x -q --timeout 200 2>&1 | tail -1
timeout 400 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench.json').read())
print(d['ms_per_step'], d['roofline']['frac'])
for k in ('per_clip_norm',):
    v=d['extra'][k]; print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ('workload','points')})
PY
