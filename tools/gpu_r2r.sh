#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
echo base; timeout 200 python bench.py --steps 100 --no-cpu-baseline --no-extra 2>&1 | tail -1 | cut -c130-170
export B200FBANK_LIB=$PWD/tools/build/ws_ep.so
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q --timeout 200 2>&1 | tail -1
echo elect; timeout 200 python bench.py --steps 100 --no-cpu-baseline --no-extra 2>&1 | tail -1 | cut -c130-170
B200FBANK_LIB=$PWD/tools/build/ws_ep_t.so timeout 120 python tools/ws_timing.py 2>&1 | tail -2
