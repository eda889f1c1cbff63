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


def test_bench_deals_one_workload_to_the_ranks():
    """bench.py --gpus N (strong scaling, BASELINE cfg-4): every event on exactly one rank, every
    rank the same mix of isoform counts, and a rank's shard regenerates the very genes of the
    whole workload (a gene is a function of (seed, id))."""
    import bench
    from workloads import Workload
    wl = dict(bench.WORKLOADS["cfg3"], n_genes=3000, reads=20)
    whole = Workload(wl["kind"], wl["n_genes"], wl["reads"], bench.READ_LEN, *bench.PE, seed=bench.SEED)
    K = whole.n_iso()
    for world in (2, 4, 8):
        parts = [bench.shard_ids(wl, r, world, "strong")[0] for r in range(world)]
        np.testing.assert_array_equal(np.sort(np.concatenate(parts)), np.arange(3000))
        cost = np.array([sum(bench.COST_PER_READ[int(k)] for k in K[p]) for p in parts], float)
        assert cost.max() / cost.mean() < 1.01
    mine = parts[3]
    w = Workload(wl["kind"], 0, wl["reads"], bench.READ_LEN, *bench.PE, seed=bench.SEED, gene_ids=mine)
    for j in (0, len(mine) // 2, len(mine) - 1):
        a, b = w.gene(j), whole.gene(int(mine[j]))
        assert a[0] == b[0] and a[1] == b[1] and (a[2] == b[2]).all() and a[3] == b[3]
    ids, total = bench.shard_ids(wl, 1, 2, "weak")
    assert total == 6000 and ids[0] == 3000 and len(ids) == 3000


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
