"""Parity at the benchmark's own shape: the exact cfg-2 and cfg-3 workloads of bench.py
(BASELINE.json configs[1], configs[2]: all 10k / 50k events, 5000 iterations, burn-in 500,
lag 10, one chain, default segment length and default kernel-layout rules -- no environment
overrides), run through misob200_run, and a seeded sample of the genes -- three per isoform
count -- compared with the oracle on the same stream (miso_paired.c:431-538, miso.c:827-947).
This is the one input set the timed step sees: multi-wave buckets cut into 256-step segments,
10^7 uniforms per chain."""
import os

import numpy as np
import pytest

import bench
from helpers import assert_gene_parity, oracle_gene

pytestmark = pytest.mark.gpu

ENV_KNOBS = ("MISOB200_SEG_ITERS", "MISOB200_SEG_ALWAYS", "MISOB200_CHAINS_PER_WARP", "MISOB200_QUAD_MAX_READS",
             "MISOB200_ONLY_K", "MISOB200_SERIAL", "MISOB200_CONCURRENT", "MISOB200_SMALL_K_FIRST", "MISOB200_CARVEOUT")


@pytest.fixture(scope="module")
def mb():
    import miso_b200
    if miso_b200.device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu tests must run on a B200")
    return miso_b200


@pytest.mark.parametrize("name,per_k", [("cfg3", 3), ("cfg2", 6)])
def test_benchmark_workload_matches_oracle(mb, ref_or_port, monkeypatch, name, per_k):
    for k in ENV_KNOBS:
        monkeypatch.delenv(k, raising=False)
    wl = bench.WORKLOADS[name]
    ids = np.arange(wl["n_genes"], dtype=np.uint32)
    plan, _, _ = bench.build_plan(mb, wl, ids)
    params = mb.make_params(bench.ITERS, bench.BURN, bench.LAG, bench.CHAINS, seed=bench.SEED)
    out = plan.alloc_outputs(params, pinned=True)      # as bench.py: the kernels write the posteriors to host memory themselves
    plan.run(params, out)
    assert (out["status"] == 0).all()
    assert (out["rundata"][:, 5] + out["rundata"][:, 6] == bench.ITERS).all()
    info = plan.info()
    rng = np.random.default_rng(1)
    pick = []
    for k in sorted(set(int(x) for x in info[:, 0])):
        cand = np.flatnonzero(info[:, 0] == k)
        pick += [int(g) for g in rng.choice(cand, size=min(per_k, len(cand)), replace=False)]
    # the extremes of the work list too: most and fewest drawing reads
    pick += [int(np.argmax(info[:, 2])), int(np.argmin(info[:, 2]))]
    w = mb.Workload(wl["kind"], 0, wl["reads"], bench.READ_LEN, *bench.PE, seed=bench.SEED, gene_ids=ids[pick])
    for j, g in enumerate(pick):
        want = oracle_gene(ref_or_port, w.gene(j), wl["kind"] == 1, params, gene_id=int(ids[g]), pe=bench.PE,
                           read_len=bench.READ_LEN)
        assert_gene_parity(plan.gene_result(out, g), want, tag="%s gene %d (K=%d, R2=%d)" % (name, g, info[g, 0], info[g, 2]))
    # the bench's own parity leg agrees (what BENCH's parity_checked field reports)
    res = bench.parity_leg(mb, wl, plan, out, ids, per_k=1)
    assert res["counts_bit_exact"] and res["max_abs_mean_diff"] <= 1e-3 and res["max_abs_ci_diff"] <= 1e-3
