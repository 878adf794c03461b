#!/bin/bash
# round-2 final: full GPU suite, smoke, headline bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2y_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y_bench.json').read())
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'] if 'gpu_launches' in d else None, d['clocks'])
for k,v in d['extra'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ('workload','points')})
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-200
