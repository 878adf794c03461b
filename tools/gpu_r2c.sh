#!/bin/bash
# round-2 GPU call C: parity suite on HEAD, kernel timing, pipeline timing, ncu capture with source
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2c_tests.log
tail -5 gpurun_out/r2c_tests.log
python tools/ktime.py --us8k --tag head 2>&1 | grep KTIME | tee -a gpurun_out/r2c_ktime.log
B200FBANK_LIB=$PWD/tools/build/al128.so python tools/ktime.py --us8k --tag al128 2>&1 | grep KTIME | tee -a gpurun_out/r2c_ktime.log
for t in timing al128t; do echo $t | tee -a gpurun_out/r2c_wstiming.log; B200FBANK_LIB=$PWD/tools/build/$t.so python tools/ws_timing2.py 2>&1 | grep WSTIMING | tee -a gpurun_out/r2c_wstiming.log; done
python tools/ktime.py --no-parity --iters 5 > gpurun_out/r2c_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fbank_ws -s 6 -c 1 -o gpurun_out/r2c_ws_full python tools/ktime.py --no-parity --iters 5 > gpurun_out/r2c_ncu.log 2>&1
ls -la gpurun_out/r2c_*
