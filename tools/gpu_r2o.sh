#!/bin/bash
# round-2: GPU tests (incl. full-size live-torchaudio parity, fused mixup, integration objects), sanitizer, bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2o_pytest.log
timeout 120 python tools/sanitize_run.py > gpurun_out/r2o_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/r2o_plain.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py small > gpurun_out/r2o_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2o_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_run.py small > gpurun_out/r2o_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2o_racecheck.log
timeout 400 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2o_bench.json').read())
print(d['ms_per_step'], d['roofline']['frac'])
for k,v in d['extra'].items():
    print(k, {a:b for a,b in v.items() if a not in ('workload','points')})
PY
