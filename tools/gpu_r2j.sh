#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 200 2>&1 | tail -3
B200_BENCH_WATCHDOG=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 2>gpurun_out/r2j_bench2.err | tail -1 | tee gpurun_out/r2j_bench2.json | cut -c1-200
grep -v "^$" gpurun_out/r2j_bench2.err | grep -A12 "most recent call first" | head -30
