#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2w_pytest.log
timeout 400 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w_bench.json').read())
print(d['ms_per_step'], d['roofline']['frac'])
for k in ('us8k','per_clip_norm','mixup_fused','melspec'):
    v=d['extra'][k]; print(k, {a:b for a,b in v.items() if a not in ('workload','points')})
print([round(p['ms_per_step'],4) for p in d['extra']['sweep']['points']])
PY
