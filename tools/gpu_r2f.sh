#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q --timeout 60 > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2f_tests.log
tail -4 gpurun_out/r2f_tests.log
timeout 300 python tools/ktime.py --us8k --tag db2 2>&1 | grep KTIME | tee -a gpurun_out/r2f_ktime.log
B200FBANK_LIB=$PWD/tools/build/v1t.so timeout 300 python tools/ws_timing2.py 2>&1 | grep WSTIMING | tee -a gpurun_out/r2f_wstiming.log
