#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_patch_embed.py tests/test_gpu_integration.py -x -q --timeout 120 2>&1 | tail -4
timeout 120 python bench.py --workload patch_embed --steps 30 2>&1 | tail -1 | cut -c150-420
for v in pp_n3; do echo $v; B200FBANK_LIB=$PWD/tools/build/$v.so timeout 120 python bench.py --workload patch_embed --steps 30 2>&1 | tail -1 | cut -c150-420; done
