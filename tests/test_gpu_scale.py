"""Full-size properties and the less-travelled device paths."""
import numpy as np
import pytest

from helpers import assert_gene_parity, oracle_gene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb():
    import miso_b200
    if miso_b200.device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu tests must run on a B200")
    return miso_b200


@pytest.mark.parametrize("tile_format,reads", [(0, 40000), (-1, 120000)])
def test_tile_too_big_for_shared_memory_streams_from_l2(mb, port, tile_format, reads):
    """R = 40k pairs (dense tile) / 120k pairs (class tile): the tile no longer fits a
    shared-memory slot, the kernel variant that streams it from global/L2 must make the
    same decisions."""
    w = mb.Workload(1, 3, reads, 36, 250.0, 900.0, 4.0, seed=21)
    plan = mb.Plan(tile_format=tile_format).append(w)
    params = mb.make_params(60, 10, 5, 2, seed=5)
    out = plan.run(params)
    for g in range(3):
        want = oracle_gene(port, w.gene(g), True, params, gene_id=g)
        assert_gene_parity(plan.gene_result(out, g), want, tag="big tile gene %d" % g)


def test_rerun_is_bit_reproducible_and_seed_matters(mb):
    w = mb.Workload(1, 40, 500, 36, 250.0, 900.0, 4.0, seed=2)
    plan = mb.Plan().append(w)
    a = plan.run(mb.make_params(400, 50, 5, 2, seed=9))
    b = plan.run(mb.make_params(400, 50, 5, 2, seed=9))
    c = plan.run(mb.make_params(400, 50, 5, 2, seed=10))
    np.testing.assert_array_equal(a["samples"], b["samples"])
    np.testing.assert_array_equal(a["assignment"], b["assignment"])
    assert not np.array_equal(a["samples"], c["samples"])


@pytest.mark.parametrize("chains_per_warp", [4, 1])
def test_cfg2_and_cfg3_sized_batches_have_sane_posteriors(mb, monkeypatch, chains_per_warp):
    """Size-independent properties on thousands of genes: psi on the simplex,
    every compatible read assigned to a compatible isoform, counts add up,
    accept + reject = iterations * chains, posterior mean near the simulated truth."""
    monkeypatch.setenv("MISOB200_CHAINS_PER_WARP", str(chains_per_warp))
    for kind, G, R in ((0, 3000, 1000), (1, 3000, 2000)):
        w = mb.Workload(kind, G, R, 36, 250.0, 900.0, 4.0, seed=31)
        plan = mb.Plan(keep_match=False).append(w)
        params = mb.make_params(1500, 300, 10, 1, seed=3)
        out = plan.run(params)
        summ = plan.summarize()
        info = plan.info()
        assert (out["status"] == 0).all()
        assert (out["rundata"][:, 5] + out["rundata"][:, 6] == 1500).all()
        err = []
        off = 0
        for g in range(G):
            K = int(info[g, 0])
            r = plan.gene_result(out, g)
            s = r["samples"]
            assert s.shape == (K, 120) and np.isfinite(s).all() and (s > 0).all()
            np.testing.assert_allclose(s.sum(axis=0), 1.0, atol=1e-12)
            a = r["assignment"]
            assert a.max() < K and a.min() >= -1
            d = mb.decode_summary(summ[g])
            assert d["assigned_counts"].sum() == (a >= 0).sum()
            if g < 300:
                err.append(np.abs(s.mean(axis=1) - w.truth(g, K)).max())
        assert np.median(err) < 0.05, np.median(err)


def test_two_sample_bayes_factor_on_device(mb):
    """BASELINE config 5 in small: two samples of the same events (different reads), Delta psi
    and Bayes factor per event on the device vs the numpy/scipy statement of
    hypothesis_test.py."""
    from miso_b200 import miso_format
    wa = mb.Workload(1, 60, 600, 36, 250.0, 900.0, 4.0, seed=41)
    wb = mb.Workload(1, 60, 600, 36, 250.0, 900.0, 4.0, seed=41)      # same genes/psi ...
    wc = mb.Workload(1, 60, 600, 36, 250.0, 900.0, 4.0, seed=42)      # ... vs different psi
    params = mb.make_params(1200, 200, 5, 2, seed=8)
    pa, pb, pc = (mb.Plan().append(w) for w in (wa, wb, wc))
    oa, ob, oc = pa.run(params), pb.run(mb.make_params(1200, 200, 5, 2, seed=9)), pc.run(params)
    # genes of seed 41 and 42 differ in K: compare only a vs b (same structure)
    cmp_ab = pa.compare(pb)
    for g in range(60):
        ra, rb = pa.gene_result(oa, g), pb.gene_result(ob, g)
        want = miso_format.bayes_factor(ra["samples"].T, rb["samples"].T)
        K = ra["samples"].shape[0]
        np.testing.assert_allclose(cmp_ab[g, :K], want, rtol=1e-9)
        np.testing.assert_allclose(cmp_ab[g, 8:8 + K], ra["samples"].mean(axis=1) - rb["samples"].mean(axis=1),
                                   atol=1e-12)
    assert (pa.info()[:, 0] != pc.info()[:, 0]).any()
    with pytest.raises(mb.InternalError, match="different number of isoforms"):
        pa.compare(pc)


def test_segment_hand_over_at_scale(mb, monkeypatch):
    """Thousands of chains handed from warp to warp through the ready queues, several waves per
    bucket and buckets of less than a wave alike: outputs identical to the uncut run, every
    chain complete.  (tools/seg_verify.py is the same check at the full cfg-3 size.)"""
    w = mb.Workload(1, 12000, 600, 36, 250.0, 900.0, 4.0, seed=77)
    plan = mb.Plan().append(w)
    params = mb.make_params(1200, 200, 10, 1, seed=6)
    res = {}
    for name, env in (("uncut", {"MISOB200_SEG_ITERS": "100000000"}),
                      ("cut", {"MISOB200_SEG_ITERS": "64", "MISOB200_SEG_ALWAYS": "1"}),
                      ("default", {"MISOB200_SEG_ITERS": "64"})):
        monkeypatch.delenv("MISOB200_SEG_ALWAYS", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        out = plan.run(params)
        res[name] = {k: np.array(out[k], copy=True) for k in ("samples", "loglik", "assignment", "rundata")}
        assert (res[name]["rundata"][:, 5] + res[name]["rundata"][:, 6] == 1200).all(), name
    for name in ("cut", "default"):
        for k in res["uncut"]:
            np.testing.assert_array_equal(res[name][k], res["uncut"][k], err_msg="%s %s" % (name, k))
