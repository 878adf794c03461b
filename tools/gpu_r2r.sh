#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
echo base; timeout 200 python bench.py --steps 100 --no-cpu-baseline --no-extra 2>&1 | tail -1 | cut -c130-170
export B200FBANK_LIB=$PWD/tools/build/ws_s0s.so
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_fullsize_parity.py -x -q --timeout 300 2>&1 | tail -8
echo scalar-stage0; timeout 200 python bench.py --steps 100 --no-cpu-baseline --no-extra 2>&1 | tail -1 | cut -c130-170
timeout 200 python bench.py --workload us8k --steps 30 2>&1 | tail -1 | cut -c150-260
