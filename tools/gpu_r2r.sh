#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 300 python -m pytest tests/test_gpu_patch_embed.py -x -q --timeout 120 2>&1 | tail -1
echo base; timeout 120 python bench.py --workload patch_embed --steps 50 2>&1 | tail -1 | cut -c200-300
for v in pp_mlp; do echo $v; B200FBANK_LIB=$PWD/tools/build/$v.so timeout 120 python bench.py --workload patch_embed --steps 50 2>&1 | tail -1 | cut -c200-300; done
