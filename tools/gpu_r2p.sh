#!/bin/bash
# round-2: pipelined patch embedding first light + regression check of the us8k number
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_patch_embed.py -x -q --timeout 120 > gpurun_out/r2p_pe.log 2>&1; echo "pe pytest rc=$?"; tail -15 gpurun_out/r2p_pe.log
timeout 120 python bench.py --workload patch_embed --steps 30 2>&1 | tail -1 | cut -c1-600
B200FBANK_PE=gather timeout 120 python bench.py --workload patch_embed --steps 30 2>&1 | tail -1 | cut -c1-600
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2p_pytest.log
timeout 400 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench.json').read())
print(d['ms_per_step'], d['roofline']['frac'])
for k,v in d['extra'].items():
    print(k, {a:b for a,b in v.items() if a not in ('workload','points')})
PY
