#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B200_BENCH_WATCHDOG=150 timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2i_bench.err | tail -1 | tee gpurun_out/r2i_bench.json
tail -40 gpurun_out/r2i_bench.err
