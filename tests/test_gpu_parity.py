"""GPU parity: the sm_100a chain kernel, called through the C ABI, against the
oracle on the same seeded inputs (bit-exact assignments / accept counts,
posterior samples and scores to fp64 rounding)."""
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


@pytest.mark.parametrize("kind,n_genes,reads", [(0, 24, 300), (1, 32, 400)])
def test_chain_matches_oracle(mb, port, kind, n_genes, reads):
    w = mb.Workload(kind, n_genes, reads, 36, 250.0, 900.0, 4.0, seed=11, first_gene_id=100)
    plan = mb.Plan().append(w)
    params = mb.make_params(n_iters=600, burn_in=100, lag=5, n_chains=2, seed=77)
    out = plan.run(params)
    assert out["launches"] >= 1
    for g in range(n_genes):
        want = oracle_gene(port, w.gene(g), kind == 1, params, gene_id=100 + g)
        assert_gene_parity(plan.gene_result(out, g), want, tag="kind %d gene %d" % (kind, g))


def test_summary_matches_numpy(mb):
    w = mb.Workload(1, 16, 300, 36, 250.0, 900.0, 4.0, seed=5)
    plan = mb.Plan().append(w)
    params = mb.make_params(n_iters=700, burn_in=100, lag=3, n_chains=3, seed=1)
    out = plan.run(params)
    summ = plan.summarize()
    for g in range(16):
        r = plan.gene_result(out, g)
        s = mb.decode_summary(summ[g])
        smp = r["samples"]
        n = smp.shape[1]
        lo = int(np.floor(0.025 * n + 0.5)) - 1
        hi = int(np.floor(0.975 * n + 0.5)) - 1
        srt = np.sort(smp, axis=1)
        np.testing.assert_allclose(s["mean"], smp.mean(axis=1), rtol=1e-12)
        np.testing.assert_array_equal(s["ci_low"], srt[:, lo])
        np.testing.assert_array_equal(s["ci_high"], srt[:, hi])
        cnt = np.bincount(r["assignment"][r["assignment"] >= 0], minlength=s["n_iso"])
        np.testing.assert_array_equal(s["assigned_counts"], cnt)
        assert s["accepted"] == r["rundata"][5] and s["rejected"] == r["rundata"][6]


def test_wide_insert_model_uses_16_bit_codes(mb, port):
    """sd = 50 -> 401 fragment lengths: codes no longer fit a byte; the 16-bit tile variant
    of the kernel must make the same decisions."""
    w = mb.Workload(1, 20, 500, 36, 300.0, 2500.0, 4.0, seed=13)
    plan = mb.Plan().append(w)
    fp, fs = plan.fragment_table()
    assert len(fp) > 255
    params = mb.make_params(500, 100, 5, 2, seed=21)
    out = plan.run(params)
    for g in range(20):
        want = oracle_gene(port, w.gene(g), True, params, gene_id=g, pe=(300.0, 2500.0, 4.0))
        assert_gene_parity(plan.gene_result(out, g), want, tag="wide gene %d" % g)
