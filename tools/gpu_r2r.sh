#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2z_pytest.log; grep -n "^E " gpurun_out/r2z_pytest.log | head -5
timeout 400 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench.json').read())
print(d['ms_per_step'], d['roofline']['frac'])
for k in ('per_clip_norm','melspec'):
    v=d['extra'][k]; print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ('workload','points')})
PY
B200FBANK_CLIPNORM=two_read timeout 200 python bench.py --workload melspec --steps 30 2>&1 | tail -1 | cut -c200-330
timeout 200 python bench.py --workload melspec --steps 30 2>&1 | tail -1 | cut -c200-330
