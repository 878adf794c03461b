#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2s_pytest.log
timeout 120 python bench.py --workload patch_embed --steps 50 2>&1 | tail -1 | cut -c150-420
