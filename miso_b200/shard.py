"""Gene sharding across GPUs (one process per GPU) and the one collective.

The reference parallelises over genes with OS processes
(``/root/reference/misopy/miso.py:163-188``: gene ids chunked into
``num_processors`` batches); genes are independent, so the GPU version deals
them to ranks and gathers only fixed-size posterior summaries at the end.
``all_gather`` is injected: NCCL through the C ABI on GPUs
(``misob200_comm_allgather``), any ``torch.distributed`` backend in the CPU
tests.
"""
import numpy as np

from ._lib import SUMMARY_F64


def shard_genes(costs, world):
    """Longest-processing-time deal of genes to `world` ranks by cost
    (reads x isoforms); returns a list of index arrays, one per rank, each in
    ascending gene order.  Deterministic on every rank."""
    costs = np.asarray(costs, dtype=np.int64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(world, dtype=np.int64)
    buckets = [[] for _ in range(world)]
    for g in order:
        r = int(np.argmin(load))
        buckets[r].append(int(g))
        load[r] += int(costs[g])
    return [np.asarray(sorted(b), dtype=np.int64) for b in buckets]


def pad_summaries(local, n_max):
    """All-gather needs equal counts: pad a rank's [n, 32] records with
    status = -1 rows (int32 field 11 of the record's integer tail)."""
    out = np.zeros((n_max, SUMMARY_F64))
    out[:len(local)] = local
    if len(local) < n_max:
        ints = out[len(local):, 24:32].view(np.int32)
        ints[:, 11] = -1
    return out


def gather_summaries(local, shards, rank, all_gather):
    """local: this rank's [len(shards[rank]), 32] summary records.  Returns the
    [n_genes, 32] table in global gene order on every rank.
    all_gather(flat_f64) -> [world * len(flat)] array, rank-major."""
    world = len(shards)
    n_max = max(len(s) for s in shards)
    mine = pad_summaries(np.asarray(local, dtype=np.float64).reshape(-1, SUMMARY_F64), n_max)
    got = np.asarray(all_gather(mine.reshape(-1))).reshape(world, n_max, SUMMARY_F64)
    n = sum(len(s) for s in shards)
    table = np.zeros((n, SUMMARY_F64))
    for r, idx in enumerate(shards):
        table[idx] = got[r, :len(idx)]
    return table
