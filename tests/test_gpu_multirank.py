"""Multi-rank GPU test (runs when the box has >= 2 GPUs): two ranks over NCCL shard a batch, accumulate the dataset
statistics on their own GPU and all-reduce them; the result must equal the sum of the partials, a single-GPU pass over the
whole batch, and the float64 oracle (rel 1e-4, north_star)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
root = sys.argv[1]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import dl_sound_classification_b200 as b2
from dl_sound_classification_b200 import stats as ST
from inputs import config1_clips
from oracle import fbank_oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
clips = config1_clips(10, length=44100)                       # ten 1 s clips, identical on every rank
fe = b2.FbankFrontend(orig_rates=(44100,), device=dev, **b2.AST_FBANK_KWARGS)
lo, hi = ST.shard_bounds(len(clips), rank, world)
ds = b2.DatasetStats(fe, 128)
ds.update(torch.cat(clips[lo:hi], 0).to(dev))
partial = ds.sums.clone()
gathered = [torch.zeros_like(partial) for _ in range(world)]
dist.all_gather(gathered, partial)
ds.all_reduce()                                               # NCCL all-reduce of 257 float64
assert torch.equal(ds.sums, sum(gathered)), "all-reduce != sum of the partials"
whole = b2.DatasetStats(fe, 128).update(torch.cat(clips, 0).to(dev))
assert torch.allclose(ds.sums, whole.sums, rtol=1e-12, atol=0), "sharded != single pass"
if rank == 0:
    feats = [O.ast_frontend(c[0].numpy(), 44100, target_frames=128, mean=None, std=None)[0][:98] for c in clips]
    want = O.dataset_stats(feats)
    got = ds.sums.cpu().numpy()
    assert got[-1] == want[-1] == 980
    mean_w, std_w = O.stats_finalize(want)[:2]
    st_g = ds.finalize()
    scale = np.abs(mean_w) + std_w                               # rel 1e-4 of |mean| + std (tests/parity.py)
    assert (np.abs(st_g.mean_per_bin.numpy() - mean_w) / scale).max() < 1e-4
    assert (np.abs(st_g.std_per_bin.numpy() - std_w) / scale).max() < 1e-4
dist.destroy_process_group()
print("ok", rank)
"""


def test_stats_allreduce_nccl_world2(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (the driver's multi-GPU tier)")
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    port = 29600 + (os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)
