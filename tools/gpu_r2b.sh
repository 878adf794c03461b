#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2b_tests.log
tail -5 gpurun_out/r2b_tests.log
for lib in pk2 r64; do
  B200FBANK_LIB=$PWD/tools/build/$lib.so python tools/ktime.py --us8k --tag $lib 2>&1 | grep KTIME | tee -a gpurun_out/r2b_ktime.log
done
B200FBANK_LIB=$PWD/tools/build/r64.so python tools/ktime.py --no-parity --iters 5 > gpurun_out/r2b_plain.log 2>&1 &&
B200FBANK_LIB=$PWD/tools/build/r64.so ncu --set full --clock-control none --import-source on -k regex:fbank_ws -s 6 -c 1 -o gpurun_out/r2b_ws_full python tools/ktime.py --no-parity --iters 5 > gpurun_out/r2b_ncu.log 2>&1
