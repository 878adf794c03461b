#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2g_tests.log
tail -30 gpurun_out/r2g_tests.log
timeout 200 python bench.py --workload melspec --steps 30 2>&1 | tail -2 | tee gpurun_out/r2g_melspec.log
