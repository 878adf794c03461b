#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:melspec_fast -s 3 -c 1 -o gpurun_out/r2m_ms_full2 python bench.py --workload melspec --steps 10 > gpurun_out/r2m_ncu.log 2>&1
ls -la gpurun_out/r2m*
