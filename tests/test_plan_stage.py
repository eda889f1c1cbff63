"""Host setup stage of the product (miso_b200/csrc/plan.cpp, through the C ABI)
against the oracle: compatibility codes, the draw order (ties included -- it
decides which read gets which uniform), read classes, the insert-length table.
Integer / index work: bit-exact.  No GPU needed."""
import numpy as np
import pytest

import miso_b200 as mb
from golden_util import load_cases

CASES = load_cases()


def plan_of(case):
    g = mb.Gene(case.exons, case.isoforms)
    rb = mb.ReadBatch([g], [case.pos], [case.cig], case.read_len, case.overhang, bool(case.paired), *case.pe)
    return mb.Plan(keep_match=True).append(rb)


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_setup_matches_oracle_on_golden_inputs(port, case):
    plan = plan_of(case)
    K, R, R2, ncls, status = plan.info()[0]
    assert status == 0 and K == len(case.isoforms)
    codes, order = plan.match(0)
    if case.paired:
        want = port.match_pe(case.exons, case.isoforms, case.pos, case.cig, case.read_len, *case.pe, case.overhang)
        fp, fs = plan.fragment_table()
        wfp, wfs = port.fragment_table(case.pe[0], case.pe[1], case.pe[2], case.read_len)
        np.testing.assert_array_equal(fp, wfp)
        assert fs == wfs
        np.testing.assert_array_equal(np.where(codes > 0, codes - 1 + fs, -1), want["fraglen"])
        np.testing.assert_array_equal(np.where(codes > 0, fp[np.maximum(codes - 1, 0)], 0.0), want["match"])
        ct, cc = want["bin_class_templates"], want["bin_class_counts"]
    else:
        want = port.match_se(case.exons, case.isoforms, case.pos, case.cig, case.read_len, case.overhang)
        np.testing.assert_array_equal(codes.astype(float), want["match"])
        ct, cc = want["class_templates"], want["class_counts"]
    np.testing.assert_array_equal(order, want["order"])
    t, c = plan.classes(0)
    np.testing.assert_array_equal(t, ct.T)
    np.testing.assert_array_equal(c, cc)
    # reads that draw = reads with >= 2 compatible isoforms (miso.c:65-68)
    assert R2 == int(((codes != 0).sum(axis=0) >= 2).sum())
    # golden class tables too (they came from the unmodified reference)
    np.testing.assert_array_equal(t, case.class_templates.T)
    np.testing.assert_array_equal(c, case.class_counts)


@pytest.mark.parametrize("kind,n_genes,reads", [(0, 10, 250), (1, 10, 250), (1, 48, 700)])
def test_setup_matches_oracle_on_synthetic_workload(port, kind, n_genes, reads):
    """(the larger paired-end case has every isoform count 2..8 several times and thousands of tied
    columns: the draw order comes from integer sort keys here, plan.cpp, and from the reference's
    qsort over the probability columns in the oracle)"""
    w = mb.Workload(kind, n_genes, reads, 36, 250.0, 900.0, 4.0, seed=3)
    plan = mb.Plan(keep_match=True).append(w)
    for g in range(n_genes):
        ex, iso, pos, cig = w.gene(g)
        codes, order = plan.match(g)
        if kind:
            want = port.match_pe(ex, iso, pos, cig, 36, 250.0, 900.0, 4.0)
            _, fs = plan.fragment_table()
            np.testing.assert_array_equal(np.where(codes > 0, codes - 1 + fs, -1), want["fraglen"])
        else:
            want = port.match_se(ex, iso, pos, cig, 36)
            np.testing.assert_array_equal(codes.astype(float), want["match"])
        np.testing.assert_array_equal(order, want["order"])


def test_edge_cases_and_status_codes():
    g3 = mb.Gene(((1, 100), (201, 300), (401, 500)), ((0, 1), (0, 2), (0, 1, 2)))
    # empty read set: a valid plan entry with nothing to draw
    p = mb.Plan().append(mb.ReadBatch([g3], [[]], [[]], 33))
    assert p.info()[0].tolist() == [3, 0, 0, 0, 0]
    # bad CIGAR -> EINVAL for that gene only (solve.c:295-298 aborts the call in the reference)
    p = mb.Plan().append(mb.ReadBatch([g3, g3], [[10], [10]], [["33Q"], ["33M"]], 33))
    assert p.info()[:, 4].tolist() == [4, 0]
    # S/H inside the alignment (solve.c:244-247)
    p = mb.Plan().append(mb.ReadBatch([g3], [[10]], [["10M2S21M"]], 33))
    assert p.info()[0, 4] == 4
    # one isoform: rejected; nine isoforms: not implemented on chip
    g1 = mb.Gene(((1, 100),), ((0,),))
    assert mb.Plan().append(mb.ReadBatch([g1], [[1]], [["33M"]], 33)).info()[0, 4] == 4
    ex9 = tuple((1 + 200 * i, 100 + 200 * i) for i in range(10))
    g9 = mb.Gene(ex9, tuple((0, i + 1) for i in range(9)))
    assert mb.Plan().append(mb.ReadBatch([g9], [[1]], [["33M"]], 33)).info()[0, 4] == 12
    # overhang >= read_len / 2 (miso.c:691-694)
    assert mb.Plan().append(mb.ReadBatch([g3], [[1]], [["33M"]], 33, overhang=16)).info()[0, 4] == 4
    # one plan = one library
    p = mb.Plan().append(mb.ReadBatch([g3], [[1]], [["33M"]], 33))
    with pytest.raises(mb.InternalError):
        p.append(mb.ReadBatch([g3], [[1]], [["36M"]], 36))
    # odd trailing mate of a paired batch is ignored (solve.c:187)
    p = mb.Plan().append(mb.ReadBatch([g3], [[10, 60, 20]], [["33M"] * 3], 33, paired=True, frag_mean=80.0,
                                      frag_var=100.0, num_devs=4.0))
    assert p.info()[0, 1] == 1


def test_tile_layout_and_sizes():
    w = mb.Workload(1, 6, 500, 36, 250.0, 900.0, 4.0, seed=9)
    plan = mb.Plan(tile_format=0).append(w)
    G, n_reads, tile_bytes = plan.size()
    info = plan.info()
    assert G == 6 and n_reads == 6 * 500
    want = sum(((int(r2) + 3 + 127) // 128 * 128 + 16) * (int(k) + 1) for k, _, r2, _, _ in info)
    assert tile_bytes == want and tile_bytes % 16 == 0
    assert (plan.tile_info()[:, 0] == 0).all()
    # class tiles (the default): an id row and a uniform-code row whatever K is, plus
    # 16 + 4 bytes per weight class
    cplan = mb.Plan().append(w)
    ti = cplan.tile_info()
    assert (ti[:, 0] == 1).all() and (ti[:, 1] >= 1).all() and (ti[:, 1] <= 254).all()
    for (k, _, r2, _, _), (_, ncls, tb) in zip(info, ti):
        padded = (int(r2) + 3 + 127) // 128 * 128
        assert tb == 2 * (padded + 16) + 16 * ncls + (4 * (ncls + 1) + 15) // 16 * 16
    assert cplan.size()[2] == int(ti[:, 2].sum()) < tile_bytes
    p = mb.make_params(1000, 100, 9, 3)
    ns, nl, na = plan.output_sizes(p)
    assert nl == 6 * 3 * 100 and na == n_reads and ns == int((info[:, 0] * 300).sum())


def test_wide_insert_model_plan(port):
    w = mb.Workload(1, 6, 300, 36, 300.0, 2500.0, 4.0, seed=13)
    plan = mb.Plan(keep_match=True, tile_format=0).append(w)
    fp, fs = plan.fragment_table()
    wfp, wfs = port.fragment_table(300.0, 2500.0, 4.0, 36)
    np.testing.assert_array_equal(fp, wfp)
    assert fs == wfs and len(fp) == 401
    info = plan.info()
    assert (info[:, 4] == 0).all()
    for g in range(6):
        ex, iso, pos, cig = w.gene(g)
        codes, order = plan.match(g)
        want = port.match_pe(ex, iso, pos, cig, 36, 300.0, 2500.0, 4.0)
        np.testing.assert_array_equal(np.where(codes > 0, codes - 1 + fs, -1), want["fraglen"])
        np.testing.assert_array_equal(order, want["order"])
    _, _, tile_bytes = plan.size()
    want_bytes = sum(((int(r2) + 3 + 127) // 128 * 128 * 2 + 16) * int(k) + (int(r2) + 3 + 127) // 128 * 128 + 16
                     for k, _, r2, _, _ in info)
    assert tile_bytes == want_bytes
    # an insert model too wide even for 16-bit tiles in shared memory is refused, not mis-run
    big = mb.Plan().append(mb.Workload(1, 1, 10, 36, 3000.0, 360000.0, 4.0, seed=1))
    assert big.info()[0, 4] == 12


ODD_CIGARS = ["33M", "10M5I23M", "3S30M", "30M3S", "2H31M", "10M2D21M", "6M100N27M", "33=", "33X", " 33M", "+33M",
              "033M", "20M", "40M", "1M100N32M", "6M100N20M100N7M", "3S3M100N27M", "0M33M"]


def test_odd_cigars_match_the_reference(ref_or_port):
    """CIGAR corner cases of splicing_parse_cigar (solve.c:220-306) -- insertions, clips, deletions,
    strtol's leading blank / sign / zeros, too short, too long, zero-length blocks -- through the shared
    matching code (csrc/match_core.hpp) against the oracle."""
    ex = ((1, 100), (201, 300), (401, 500))
    iso = ((0, 1, 2), (0, 2), (0, 1))
    pos, cig = [], []
    for c in ODD_CIGARS:
        for p in (1, 68, 95, 98, 210, 268, 295, 405):
            pos.append(p)
            cig.append(c)
    g = mb.Gene(ex, iso)
    for overhang in (1, 4):
        plan = mb.Plan(keep_match=True).append(mb.ReadBatch([g], [pos], [cig], 33, overhang=overhang))
        assert plan.info()[0, 4] == 0
        codes, order = plan.match(0)
        want = ref_or_port.match_se(ex, iso, np.asarray(pos, np.int32), cig, 33, overhang)
        np.testing.assert_array_equal(codes.astype(float), want["match"])
        np.testing.assert_array_equal(order, want["order"])
    # paired: the same strings as mates
    ppos, pcig = [], []
    for i in range(0, len(pos) - 1, 2):
        ppos += [pos[i], pos[i] + 40]
        pcig += [cig[i], cig[i + 1]]
    plan = mb.Plan(keep_match=True).append(mb.ReadBatch([g], [ppos], [pcig], 33, paired=True, frag_mean=80.0,
                                                        frag_var=400.0, num_devs=4.0))
    codes, order = plan.match(0)
    fp, fs = plan.fragment_table()
    want = ref_or_port.match_pe(ex, iso, np.asarray(ppos, np.int32), pcig, 33, 80.0, 400.0, 4.0, 1)
    np.testing.assert_array_equal(np.where(codes > 0, codes - 1 + fs, -1), want["fraglen"])
    np.testing.assert_array_equal(order, want["order"])


def test_synthetic_workload_has_the_read_mix_of_the_reference_simulator(ref):
    """bench.py's generator (workloads/synth.cpp) against the reference's own paired-end simulator
    (splicing_simulate_paired_reads, simulator.c:221-442, through oracle/_ref) on the same genes and
    the same true psi: what sets the GPU's cost per gene -- the number of reads that draw (R2) and the
    number of weight classes -- must agree within sampling noise."""
    import miso_b200 as mb
    from workloads import Workload
    w = Workload(1, 60, 2000, 36, 250.0, 900.0, 4.0, seed=20260925)
    ours = mb.Plan().append(w)
    info = ours.info()
    genes, poss, cigs = [], [], []
    for g in range(60):
        ex, isos, _, _ = w.gene(g)
        K = len(isos)
        pos, cig, _ = ref.simulate_pe(ex, isos, w.truth(g, K), 2000, 36, 250.0, 900.0, 4.0, seed=1000 + g)
        genes.append(mb.Gene(ex, isos))
        poss.append(pos)
        cigs.append(cig)
    theirs = mb.Plan().append(mb.ReadBatch(genes, poss, cigs, 36, 1, True, 250.0, 900.0, 4.0)).info()
    assert (theirs[:, 0] == info[:, 0]).all() and (theirs[:, 1] == 2000).all() and (info[:, 1] == 2000).all()
    for k in sorted(set(int(x) for x in info[:, 0])):
        m = info[:, 0] == k
        a, b = info[m, 2].mean(), theirs[m, 2].mean()
        assert abs(a - b) <= 0.05 * max(a, b) + 6, (k, a, b)                       # drawing reads
        assert abs(info[m, 3].mean() - theirs[m, 3].mean()) <= 1.5, k              # read classes (0/1 patterns)
    # gene by gene: same psi, same structure -> R2 within binomial noise of each other
    d = np.abs(info[:, 2] - theirs[:, 2])
    assert np.median(d) <= 40 and d.max() <= 160, (np.median(d), d.max())
