#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 300 python -m pytest tests/test_gpu_patch_embed.py -x -q --timeout 120 2>&1 | tail -1
echo base; timeout 120 python bench.py --workload patch_embed --steps 50 2>&1 | tail -1 | cut -c200-300
export B200FBANK_LIB=$PWD/tools/build/pp_c8.so
timeout 300 python -m pytest tests/test_gpu_patch_embed.py -x -q --timeout 120 2>&1 | tail -1
echo c8; timeout 120 python bench.py --workload patch_embed --steps 50 2>&1 | tail -1 | cut -c200-300
