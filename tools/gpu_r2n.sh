#!/bin/bash
# round-2 evidence run: GPU tests, bench line, ncu launch list, full captures of the two tuned kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.log
timeout 600 python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2n_bench.json
timeout 200 python bench.py --workload melspec --steps 30 2>&1 | tail -1 > gpurun_out/r2n_melspec.json; cut -c1-500 gpurun_out/r2n_melspec.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2n_ncu1.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fbank_ws -s 3 -c 1 -o gpurun_out/r02_ws_full python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2n_ncu2.log 2>&1; echo "ncu ws rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:melspec_fast -s 3 -c 1 -o gpurun_out/r02_melspec_full python bench.py --workload melspec --steps 10 > gpurun_out/r2n_ncu3.log 2>&1; echo "ncu ms rc=$?"
ls -la gpurun_out/
