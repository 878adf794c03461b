#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 200 -k "reference_actual or cache or clip_norm or melspec" 2>&1 | tail -3
timeout 200 python bench.py --workload melspec --steps 30 2>&1 | tail -1 | cut -c150-400
for v in ms14 ms16; do echo $v; B200FBANK_LIB=$PWD/tools/build/$v.so timeout 200 python bench.py --workload melspec --steps 30 2>&1 | tail -1 | cut -c150-400; done
ncu --set full --clock-control none --import-source on -k regex:melspec_fast -s 3 -c 1 -o gpurun_out/r2m_ms_full3 python bench.py --workload melspec --steps 10 > gpurun_out/r2m_ncu.log 2>&1
