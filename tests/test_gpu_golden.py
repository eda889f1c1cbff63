"""GPU vs the committed golden vectors (outputs of the UNMODIFIED reference,
tests/golden/make_golden.py), through the reference-facing Python entry points
pysplicing.MISO / MISOPaired, the batched plan API, and the .miso writer."""
import numpy as np
import pytest

from golden_util import STREAMS, load_cases

pytestmark = pytest.mark.gpu
CASES = [c for v in STREAMS for c in load_cases(v)]       # both stream versions (Philox4x32-10 / -7)
IDS = ["v%d-%s" % (c.stream, c.name) for c in CASES]


@pytest.fixture(autouse=True)
def stream_of_the_case(request):
    """Select the product's random-stream version the golden case was generated with."""
    import miso_b200
    case = request.node.callspec.params.get("case") if hasattr(request.node, "callspec") else None
    was = miso_b200.stream_version()
    if case is not None:
        miso_b200.stream_version(case.stream)
    yield
    miso_b200.stream_version(was)


@pytest.fixture(scope="module")
def mb():
    import miso_b200
    if miso_b200.device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu tests must run on a B200")
    return miso_b200


def check(case, samples, loglik, assignment, acc, rej):
    np.testing.assert_array_equal(np.asarray(assignment), case.assignment)
    assert (acc, rej) == (case.accepted, case.rejected)
    np.testing.assert_allclose(np.asarray(samples), case.samples, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(np.asarray(loglik), case.loglik, rtol=1e-9, atol=1e-9)
    # the contract of BASELINE.json: posterior mean and 95% CI within 1e-3
    from miso_b200.miso_format import credible_interval_indices
    lo, hi = credible_interval_indices(case.samples.shape[1])
    for k in range(case.samples.shape[0]):
        a, b = np.sort(np.asarray(samples)[k]), np.sort(case.samples[k])
        assert abs(a.mean() - b.mean()) < 1e-3 and abs(a[lo] - b[lo]) < 1e-3 and abs(a[hi] - b[hi]) < 1e-3


@pytest.fixture(params=[4, 1], ids=["quad", "single"])
def chains_per_warp(request, monkeypatch):
    """four gene-chains per warp (quad_kernel.cuh, what big batches run) / one (chain_kernel.cuh)"""
    monkeypatch.setenv("MISOB200_CHAINS_PER_WARP", str(request.param))
    return request.param


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_pysplicing_entry_points_match_reference_golden(mb, case, chains_per_warp):
    import pysplicing
    gene = pysplicing.createGene(case.exons, case.isoforms)
    pos, cig = tuple(int(p) for p in case.pos), tuple(case.cig)
    common = (case.n_iters, case.burn_in, case.lag, None, case.overhang, case.n_chains, case.start, 0)
    # batch API with the golden gene id (stream = (seed, gene_id, chain))
    g = mb.Gene(case.exons, case.isoforms)
    rb = mb.ReadBatch([g], [case.pos], [case.cig], case.read_len, case.overhang, bool(case.paired), *case.pe,
                      gene_ids=[case.gene_id])
    plan = mb.Plan().append(rb)
    params = mb.make_params(case.n_iters, case.burn_in, case.lag, case.n_chains, start=case.start, seed=case.seed)
    out = plan.run(params)
    r = plan.gene_result(out, 0)
    assert r["status"] == 0
    check(case, r["samples"], r["loglik"], r["assignment"], int(r["rundata"][5]), int(r["rundata"][6]))
    t, c = plan.classes(0)
    np.testing.assert_array_equal(t, case.class_templates.T)
    np.testing.assert_array_equal(c, case.class_counts)
    if case.gene_id == 0:       # pysplicing.MISO numbers its single gene 0: same stream as the golden run
        if case.paired:
            res = pysplicing.MISOPaired(gene, 0, pos, cig, case.read_len, *case.pe, *common, seed=case.seed)
        else:
            res = pysplicing.MISO(gene, 0, pos, cig, case.read_len, *common, 0, seed=case.seed)
        assert isinstance(res, tuple) and len(res) == 6 and all(isinstance(x, tuple) for x in res)
        assert len(res[0]) == len(case.isoforms) and isinstance(res[0][0][0], float)
        check(case, res[0], res[1], res[4], res[5][4], res[5][5])
        assert res[5][:4] == (len(case.isoforms), case.n_iters, case.burn_in, case.lag)
        np.testing.assert_array_equal(np.asarray(res[2]), case.class_templates.T)


def test_sampler_front_end_writes_reference_format(mb, tmp_path):
    from miso_b200 import sampler, miso_format
    case = [c for c in load_cases() if c.name == "cfg1_default"][0]      # the default stream's golden
    parts = [sampler.Part("up", *case.exons[0]), sampler.Part("se", *case.exons[1]), sampler.Part("dn", *case.exons[2])]
    gene = sampler.GeneModel("ev1", parts, [["up", "se", "dn"], ["up", "dn"]], chrom="chr10", strand="+")
    prm = sampler.get_single_end_sampler_params(2, 36)
    smp = sampler.MISOSampler(prm, paired_end=False, seed=case.seed)
    out = smp.run_sampler(case.n_iters, (case.pos - 1, case.cig), gene, None, prm, str(tmp_path / "chr10" / "ev1"),
                          num_chains=case.n_chains, burn_in=case.burn_in, lag=case.lag)
    samples, header, scores, _, _, counts = miso_format.load_samples(out)
    np.testing.assert_allclose(samples, np.round(case.samples.T, 4), atol=1.1e-4)
    np.testing.assert_allclose(scores, case.loglik, atol=6e-3)
    assert header["iters"] == "5000" and header["burn_in"] == "500" and header["lag"] == "10"
    acc = 100.0 * case.accepted / (case.accepted + case.rejected)
    assert header["percent_accept"] == "%.2f" % acc
    want_counts = ",".join("%s:%d" % (str(tuple(int(v) for v in t)).replace(" ", ""), int(n))
                           for t, n in zip(case.class_templates.T, case.class_counts))
    assert counts == want_counts
    # second call: the file exists -> skipped (miso_sampler.py:233-238)
    assert smp.run_sampler(case.n_iters, (case.pos - 1, case.cig), gene, None, prm, str(tmp_path / "chr10" / "ev1"),
                           num_chains=6, burn_in=500, lag=10) is None
