#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2h_tests.log
tail -8 gpurun_out/r2h_tests.log
if [ "$1" = "2" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 2>gpurun_out/r2h_bench2.err | tail -1 | tee gpurun_out/r2h_bench2.json
else
  timeout 900 python bench.py --steps 100 --warmup 5 2>gpurun_out/r2h_bench.err | tail -1 | tee gpurun_out/r2h_bench.json
fi
