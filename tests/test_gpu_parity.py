"""GPU parity: the sm_100a chain kernel, called through the C ABI, against the
oracle on the same seeded inputs (bit-exact assignments / accept counts,
posterior samples and scores to fp64 rounding)."""
import numpy as np
import pytest

from helpers import assert_gene_parity, oracle_gene, simulate_pairs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb():
    import miso_b200
    if miso_b200.device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu tests must run on a B200")
    return miso_b200


# Every device layout must reproduce the oracle:
#   "quad"   class tiles (integer thresholds per weight class, class_pass.cuh), four gene-chains
#            per warp (quad_kernel.cuh) -- what a full-size batch runs
#   "single" class tiles, one gene-chain per warp (chain_kernel.cuh) -- small batches
#   "dense"  dense tiles (per-read fp64 weights, dense_pass.cuh) -- genes with too many classes
LAYOUTS = ["quad", "single", "dense"]


@pytest.fixture(params=[4, 1], ids=["quad", "single"])
def chains_per_warp(request, monkeypatch):
    monkeypatch.setenv("MISOB200_CHAINS_PER_WARP", str(request.param))
    return request.param


@pytest.fixture(params=LAYOUTS)
def tile_format(request, monkeypatch):
    monkeypatch.setenv("MISOB200_CHAINS_PER_WARP", "4" if request.param == "quad" else "1")
    # chains are handed from warp to warp in segments (ChainState, chain_kernel.cuh): an odd short
    # segment makes every chain resume many times, at every phase of the lag counter
    monkeypatch.setenv("MISOB200_SEG_ITERS", "37")
    monkeypatch.setenv("MISOB200_SEG_ALWAYS", "1")      # (by default only buckets of more than one wave are cut)
    return 0 if request.param == "dense" else -1


@pytest.mark.parametrize("kind,n_genes,reads", [(0, 24, 300), (1, 32, 400)])
def test_chain_matches_oracle(mb, port, kind, n_genes, reads, tile_format):
    w = mb.Workload(kind, n_genes, reads, 36, 250.0, 900.0, 4.0, seed=11, first_gene_id=100)
    plan = mb.Plan(tile_format=tile_format).append(w)
    assert (plan.tile_info()[:, 0] == (1 if tile_format else 0)).all()
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


def test_wide_insert_model_uses_16_bit_codes(mb, port, tile_format):
    """sd = 50 -> 401 fragment lengths: codes no longer fit a byte; the 16-bit tile variant
    of the kernel must make the same decisions."""
    w = mb.Workload(1, 20, 500, 36, 300.0, 2500.0, 4.0, seed=13)
    plan = mb.Plan(tile_format=tile_format).append(w)
    fp, fs = plan.fragment_table()
    assert len(fp) > 255
    params = mb.make_params(500, 100, 5, 2, seed=21)
    out = plan.run(params)
    for g in range(20):
        want = oracle_gene(port, w.gene(g), True, params, gene_id=g, pe=(300.0, 2500.0, 4.0))
        assert_gene_parity(plan.gene_result(out, g), want, tag="wide gene %d" % g)


def _cassette_batch(mb, n_genes, n_pairs, exon_len, two_cassettes, seed):
    """Events whose alternative exons are SHORTER than the insert-length window: a pair
    flanking a cassette exon is compatible with the inclusion and the exclusion isoform at
    two different fragment lengths, so its weight vector is not a 0/1 pattern and every
    distinct fragment length is a weight class of its own."""
    rng = np.random.default_rng(seed)
    genes, poss, cigs, raw = [], [], [], []
    for g in range(n_genes):
        if two_cassettes:
            exons = ((1, 300), (401, 400 + exon_len), (601, 600 + exon_len + 7), (801, 1100))
            isoforms = ((0, 1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 3))
        else:
            exons = ((1, 300), (401, 400 + exon_len), (601, 900))
            isoforms = ((0, 1, 2), (0, 2))
        psi = rng.dirichlet(np.ones(len(isoforms)))
        pos, cig = simulate_pairs(exons, isoforms, psi, n_pairs, 36, 250.0, 30.0, 4.0, rng)
        genes.append(mb.Gene(exons, isoforms))
        poss.append(pos)
        cigs.append(cig)
        raw.append((exons, isoforms, pos, cig))
    rb = mb.ReadBatch(genes, poss, cigs, 36, 1, True, 250.0, 900.0, 4.0)
    return rb, raw


@pytest.mark.parametrize("two_cassettes,want_format", [(False, 1), (True, 0)])
def test_reads_with_two_fragment_lengths(mb, port, two_cassettes, want_format, chains_per_warp):
    """Non-uniform weight classes.  One short cassette exon: up to ~200 classes, still a class
    tile.  Two: more than 254 classes, the plan must fall back to the dense tile by itself."""
    rb, raw = _cassette_batch(mb, 10, 1500 if two_cassettes else 700, 61, two_cassettes, seed=5)
    plan = mb.Plan().append(rb)
    ti = plan.tile_info()
    assert (ti[:, 0] == want_format).all(), ti
    if want_format == 1:
        assert ti[:, 1].max() > 60          # genuinely many classes, most of them non-uniform
    params = mb.make_params(n_iters=400, burn_in=100, lag=5, n_chains=2, seed=3)
    out = plan.run(params)
    for g, wl_gene in enumerate(raw):
        want = oracle_gene(port, wl_gene, True, params, gene_id=g)
        assert_gene_parity(plan.gene_result(out, g), want, tag="cassette gene %d" % g)


@pytest.mark.parametrize("kind", [0, 1])
def test_start_random_and_uniform(mb, port, kind, chains_per_warp):
    """MISO_START_RANDOM: Dirichlet(1) start from K gamma(1,1) = -log(uniform) draws
    (src/miso.c:388-404, :309-326) and MISO_START_UNIFORM (:372-386)."""
    w = mb.Workload(kind, 16, 300, 36, 250.0, 900.0, 4.0, seed=17, first_gene_id=40)
    plan = mb.Plan().append(w)
    for start in (mb.MISO_START_RANDOM, mb.MISO_START_UNIFORM):
        params = mb.make_params(n_iters=400, burn_in=100, lag=5, n_chains=2, start=start, seed=9)
        out = plan.run(params)
        for g in range(16):
            want = oracle_gene(port, w.gene(g), kind == 1, params, gene_id=40 + g)
            assert_gene_parity(plan.gene_result(out, g), want, tag="start %d kind %d gene %d" % (start, kind, g))


@pytest.mark.parametrize("kind", [0, 1])
def test_results_do_not_depend_on_segment_length(mb, kind, chains_per_warp, monkeypatch):
    """Cutting the chains into segments (resume from ChainState) must not change a single bit."""
    w = mb.Workload(kind, 37, 400, 36, 250.0, 900.0, 4.0, seed=23)
    plan = mb.Plan().append(w)
    params = mb.make_params(n_iters=500, burn_in=100, lag=7, n_chains=3, seed=4)
    ref = None
    monkeypatch.setenv("MISOB200_SEG_ALWAYS", "1")          # (by default only buckets of more than one wave are cut)
    for seg in ("1000000", "256", "50", "7", "1"):
        monkeypatch.setenv("MISOB200_SEG_ITERS", seg)
        out = plan.run(params)
        got = {k: np.array(out[k], copy=True) for k in ("samples", "loglik", "assignment", "rundata")}
        if ref is None:
            ref = got
            assert (got["rundata"][:, 5] + got["rundata"][:, 6] == 1500).all()
        else:
            for k in ref:
                np.testing.assert_array_equal(got[k], ref[k], err_msg="seg %s %s" % (seg, k))


def test_degenerate_genes_in_one_batch(mb, port, chains_per_warp):
    """One batch mixing degenerate inputs: no reads at all (the chain runs on the prior; defined by
    the port only -- the reference C aborts on an empty read set, error.c:121, and misopy never
    passes one, miso_sampler.py:229), reads that fit no isoform (assignment -1, miso.c:65), a
    gene whose reads all hit the same two isoforms (a single class), a bad CIGAR (status EINVAL,
    the other genes unaffected), shorter-than-read-length alignments (all-zero column,
    solve.c:55), an ordinary gene, and non-flat hyperparameters."""
    ex = ((1, 100), (201, 300), (401, 500))
    iso = ((0, 1, 2), (0, 2), (0, 1))
    g = mb.Gene(ex, iso)
    rng = np.random.default_rng(3)
    inc2 = [int(p) for p in rng.integers(210, 260, 40)]          # inside exon 1: isoforms 0 and 2
    only0 = [95] * 25                                           # junction exon 0 -> 1 (6M100N27M): isoforms 0 and 2
    nowhere = [120] * 10                                        # intron
    ordinary_pos = inc2 + [int(p) for p in rng.integers(1, 60, 50)] + [int(p) for p in rng.integers(405, 460, 30)]
    cases = [
        ([], []),
        (nowhere, ["33M"] * len(nowhere)),
        (only0, ["6M100N27M"] * len(only0)),
        ([10], ["33Q"]),
        (inc2, ["30M"] * len(inc2)),
        (ordinary_pos, ["33M"] * len(ordinary_pos)),
        (ordinary_pos, ["33M"] * len(ordinary_pos)),
    ]
    hyper = [None] * 6 + [(0.5, 2.0, 1.5)]
    rb = mb.ReadBatch([g] * len(cases), [c[0] for c in cases], [c[1] for c in cases], 33,
                      hyper=[h if h is not None else (1.0, 1.0, 1.0) for h in hyper])
    plan = mb.Plan().append(rb)
    params = mb.make_params(n_iters=300, burn_in=50, lag=5, n_chains=2, seed=12)
    out = plan.run(params)
    assert [int(s) for s in out["status"]] == [0, 0, 0, 4, 0, 0, 0]
    for i, (pos, cig) in enumerate(cases):
        r = plan.gene_result(out, i)
        if i == 3:
            assert (r["assignment"] == -1).all()
            continue
        want = oracle_gene(port, (ex, iso, np.asarray(pos, np.int32), cig), False, params, gene_id=i,
                           read_len=33, hyper=hyper[i])
        assert_gene_parity(r, want, tag="degenerate case %d" % i)
    assert (plan.gene_result(out, 1)["assignment"] == -1).all()
    assert (plan.gene_result(out, 4)["assignment"] == -1).all()


def test_read_score_lookup_checked_and_unchecked(mb, port, monkeypatch):
    """The read-score passes look -log(lp) up in a table; genes whose lp range the host proved
    inside the table skip the range test (GeneDesc.lp_safe, class_pass.cuh MODE 2).  (a) isoforms
    shorter than the longest insert length: lp can leave the table, the checked path must run;
    (b) an ordinary batch forced through the checked path."""
    rng = np.random.default_rng(9)
    exons = ((1, 100), (201, 260), (401, 500))
    isoforms = ((0, 1, 2), (0, 2))
    genes, poss, cigs, raw = [], [], [], []
    for g in range(12):
        psi = rng.dirichlet(np.ones(2))
        pos, cig = simulate_pairs(exons, isoforms, psi, 400, 36, 150.0, 20.0, 4.0, rng)
        genes.append(mb.Gene(exons, isoforms)); poss.append(pos); cigs.append(cig)
        raw.append((exons, isoforms, pos, cig))
    rb = mb.ReadBatch(genes, poss, cigs, 36, 1, True, 150.0, 400.0, 4.0)
    plan = mb.Plan().append(rb)
    fp, fs = plan.fragment_table()
    assert fs + len(fp) - 1 > 200          # inserts longer than the short isoform exist in the model
    params = mb.make_params(n_iters=400, burn_in=50, lag=5, n_chains=2, seed=31)
    out = plan.run(params)
    for g, wl_gene in enumerate(raw):
        want = oracle_gene(port, wl_gene, True, params, gene_id=g, pe=(150.0, 400.0, 4.0))
        assert_gene_parity(plan.gene_result(out, g), want, tag="short isoform gene %d" % g)
    monkeypatch.setenv("MISOB200_NO_LP_SAFE", "1")
    w = mb.Workload(1, 16, 400, 36, 250.0, 900.0, 4.0, seed=12)
    plan = mb.Plan().append(w)
    out = plan.run(params)
    for g in range(16):
        want = oracle_gene(port, w.gene(g), True, params, gene_id=g)
        assert_gene_parity(plan.gene_result(out, g), want, tag="checked lookup gene %d" % g)


@pytest.mark.parametrize("kind", [0, 1])
def test_literal_proposal_scores(mb, port, kind, chains_per_warp, monkeypatch):
    """The proposal densities are evaluated in log space (proposal_scores, chain_kernel.cuh); the
    reference's literal operation sequence (mvplogisnorm, miso.c:97-122) is kept for values a chain
    does not normally see (a psi of 0, NaN, underflow).  MISOB200_LITERAL_SCORES=1 sends EVERY proposal
    down that route: the same decisions as the oracle, and as the log-space route."""
    w = mb.Workload(kind, 20, 300, 36, 250.0, 900.0, 4.0, seed=13, first_gene_id=40)
    plan = mb.Plan().append(w)
    params = mb.make_params(n_iters=500, burn_in=100, lag=5, n_chains=2, seed=5)
    fast = plan.run(params)
    fast = {k: np.array(v, copy=True) for k, v in fast.items() if isinstance(v, np.ndarray)}
    monkeypatch.setenv("MISOB200_LITERAL_SCORES", "1")
    lit = plan.run(params)
    for g in range(20):
        want = oracle_gene(port, w.gene(g), kind == 1, params, gene_id=40 + g)
        assert_gene_parity(plan.gene_result(lit, g), want, tag="literal route, kind %d gene %d" % (kind, g))
    np.testing.assert_array_equal(lit["assignment"], fast["assignment"])
    np.testing.assert_array_equal(lit["rundata"], fast["rundata"])
    np.testing.assert_allclose(lit["samples"], fast["samples"], rtol=1e-9, atol=0)
