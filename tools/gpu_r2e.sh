#!/bin/bash
# round-2 GPU call E: lane-pair resampler (half the taps per lane), register split variants
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2e_tests.log
tail -4 gpurun_out/r2e_tests.log
for v in v1 v2 v3; do
B200FBANK_LIB=$PWD/tools/build/$v.so timeout 300 python tools/ktime.py --us8k --tag $v 2>&1 | grep KTIME | tee -a gpurun_out/r2e_ktime.log
done
for v in v1t v3t; do echo $v | tee -a gpurun_out/r2e_wstiming.log
B200FBANK_LIB=$PWD/tools/build/$v.so timeout 300 python tools/ws_timing2.py 2>&1 | grep WSTIMING | tee -a gpurun_out/r2e_wstiming.log
done
