#!/bin/bash
# round-2 GPU call D: double-buffered half-chunk input
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2d_tests.log
tail -15 gpurun_out/r2d_tests.log
timeout 300 python tools/ktime.py --us8k --tag dbuf 2>&1 | grep KTIME | tee -a gpurun_out/r2d_ktime.log
B200FBANK_LIB=$PWD/tools/build/timing.so timeout 300 python tools/ws_timing2.py 2>&1 | grep WSTIMING | tee -a gpurun_out/r2d_wstiming.log
