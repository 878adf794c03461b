#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 200 -k "reference_actual or cache or clip_norm" 2>&1 | tail -25
timeout 200 python bench.py --workload melspec --steps 30 2>&1 | tail -1 | cut -c1-400
