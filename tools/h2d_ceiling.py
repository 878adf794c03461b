"""What the BOX gives: N ranks copying plain pinned host buffers to / from their GPUs at the same time (one
cudaMemcpyAsync per buffer, no kernels) -- the ceiling of the end-to-end number of bench.py at N GPUs.
usage: python tools/h2d_ceiling.py            (one GPU)
       python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/h2d_ceiling.py"""
import json
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
H2D, D2H = 1024 * 220500 * 4, 1024 * 512 * 128 * 4          # bytes of one bench step: waveforms in, features out
h_in = torch.empty(H2D, dtype=torch.uint8).pin_memory()
h_out = torch.empty(D2H, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
d_in, d_out = torch.empty(H2D, dtype=torch.uint8, device=dev), torch.ones(D2H, dtype=torch.uint8, device=dev)
s2 = torch.cuda.Stream(dev)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def timed(fn, steps=10):
    fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def both():
    d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)


ms_in = timed(lambda: d_in.copy_(h_in, non_blocking=True))
ms_out = timed(lambda: h_out.copy_(d_out, non_blocking=True))
ms_both = timed(both)
if rank == 0:
    print(json.dumps({"n_gpus": world, "h2d_bytes": H2D, "d2h_bytes": D2H,
                      "h2d_alone_gbs_per_rank": H2D / ms_in / 1e6, "h2d_alone_gbs_total": world * H2D / ms_in / 1e6,
                      "d2h_alone_gbs_per_rank": D2H / ms_out / 1e6,
                      "both_ms_per_step": ms_both, "both_gbs_total": world * (H2D + D2H) / ms_both / 1e6,
                      "ceiling_audio_s_per_s": world * 1024 * 5.0 / (ms_both * 1e-3),
                      "cpu_affinity": len(os.sched_getaffinity(0))}))
if world > 1:
    dist.destroy_process_group()
