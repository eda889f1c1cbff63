"""N > 1 host logic on CPU: genes dealt to ranks, per-rank summaries gathered
into the global table with one all-gather (world_size 2, gloo).  On GPUs the
same gather runs over NCCL through misob200_comm_allgather."""
import os
import subprocess
import sys

import numpy as np

from miso_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch
import torch.distributed as dist
from miso_b200 import shard
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
rng = np.random.default_rng(5)
costs = rng.integers(100, 20000, size=37)
shards = shard.shard_genes(costs, world)
mine = shards[rank]
local = np.zeros((len(mine), 32))
local[:, 0] = mine * 1.5          # stands in for mean[0]
local[:, 24:32].view(np.int32)[:, 8] = 2 + mine % 7   # n_iso
def all_gather(flat):
    t = torch.from_numpy(np.ascontiguousarray(flat))
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return torch.cat(outs).numpy()
table = shard.gather_summaries(local, shards, rank, all_gather)
assert table.shape == (37, 32)
np.testing.assert_array_equal(table[:, 0], np.arange(37) * 1.5)
np.testing.assert_array_equal(table[:, 24:32].view(np.int32)[:, 8], 2 + np.arange(37) % 7)
dist.barrier()
print("rank", rank, "ok")
'''


def test_shard_genes_is_balanced_and_complete():
    rng = np.random.default_rng(0)
    costs = rng.integers(1, 16000, size=1000)
    for world in (1, 2, 4, 8):
        sh = shard.shard_genes(costs, world)
        allg = np.sort(np.concatenate(sh))
        np.testing.assert_array_equal(allg, np.arange(1000))
        loads = np.array([costs[s].sum() for s in sh])
        assert loads.max() <= loads.mean() * 1.02 + costs.max()


def test_gather_with_two_gloo_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29653", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o
