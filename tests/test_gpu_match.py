"""Read <-> isoform compatibility and the draw-order sort on the device (csrc/match.cu: match_kernel,
order_kernel) against the host plan stage and the oracle: the same integer
codes, draw order (the unstable quicksort's tie order included), classes, tiles and -- run through the
chain kernels -- the same posteriors."""
import numpy as np
import pytest

from golden_util import load_cases
from test_plan_stage import ODD_CIGARS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb():
    import miso_b200
    if miso_b200.device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu tests must run on a B200")
    return miso_b200


def same_plan(mb, a, b, n_genes, run=True):
    np.testing.assert_array_equal(a.info(), b.info())
    np.testing.assert_array_equal(a.tile_info(), b.tile_info())
    assert a.size() == b.size()
    for g in range(n_genes):
        if a.info()[g, 4] != 0:
            continue
        ca, oa = a.match(g)
        cb, ob = b.match(g)
        np.testing.assert_array_equal(ca, cb, err_msg="codes gene %d" % g)
        np.testing.assert_array_equal(oa, ob, err_msg="order gene %d" % g)
    if run:
        params = mb.make_params(200, 40, 5, 2, seed=3)
        ra, rb = a.run(params), b.run(params)
        for k in ("samples", "loglik", "assignment", "rundata", "status"):
            np.testing.assert_array_equal(ra[k], rb[k], err_msg=k)


@pytest.mark.parametrize("kind,n_genes,reads", [(0, 300, 700), (1, 300, 900)])
def test_device_matching_equals_host_on_synthetic_batches(mb, kind, n_genes, reads):
    w = mb.Workload(kind, n_genes, reads, 36, 250.0, 900.0, 4.0, seed=19)
    host = mb.Plan(keep_match=True).append(w)
    dev = mb.Plan(keep_match=True).append(w, match_device=0)
    same_plan(mb, host, dev, n_genes)
    k_ms, h_ms, d_ms, b_in, b_out = mb.Plan.last_match_stats()
    info = host.info()
    # one byte per code (at most 256 codes in this insert model) + four bytes of draw order per read
    assert k_ms > 0 and b_in > 0 and b_out == int((info[:, 0] * info[:, 1]).sum()) + 4 * int(info[:, 1].sum())


def test_device_order_with_wide_keys_and_long_genes(mb, monkeypatch):
    """sd = 70: 561 fragment lengths -> 16-bit codes and more than 255 distinct probabilities, i.e. 128-bit
    sort keys; a gene with 9000 pairs does not fit the sort kernel's shared-memory arrays and is ordered
    by the host (same code, bm_sort.hpp); and the host-sort switch gives the same plan."""
    w = mb.Workload(1, 40, 700, 36, 300.0, 4900.0, 4.0, seed=23)
    host = mb.Plan(keep_match=True).append(w)
    dev = mb.Plan(keep_match=True).append(w, match_device=0)
    assert len(host.fragment_table()[0]) > 510
    same_plan(mb, host, dev, 40)
    big = mb.Workload(1, 3, 9000, 36, 250.0, 900.0, 4.0, seed=29)
    same_plan(mb, mb.Plan(keep_match=True).append(big), mb.Plan(keep_match=True).append(big, match_device=0), 3, run=False)
    monkeypatch.setenv("MISOB200_HOST_SORT", "1")
    same_plan(mb, host, mb.Plan(keep_match=True).append(w, match_device=0), 40, run=False)


def test_device_matching_on_odd_and_bad_cigars(mb, port):
    ex = ((1, 100), (201, 300), (401, 500))
    iso = ((0, 1, 2), (0, 2), (0, 1))
    pos, cig = [], []
    for c in ODD_CIGARS:
        for p in (1, 68, 95, 98, 210, 268, 295, 405):
            pos.append(p)
            cig.append(c)
    g = mb.Gene(ex, iso)
    bad = ([10, 20], ["33M", "12Q21M"])
    mid_clip = ([10], ["10M2S21M"])
    empty = ([], [])
    genes = [g, g, g, g, g]
    poss = [pos, bad[0], pos[::-1], mid_clip[0], empty[0]]
    cigs = [cig, bad[1], cig[::-1], mid_clip[1], empty[1]]
    for paired in (False, True):
        kw = dict(paired=True, frag_mean=80.0, frag_var=400.0, num_devs=4.0) if paired else {}
        for overhang in (1, 4):
            rb = mb.ReadBatch(genes, poss, cigs, 33, overhang=overhang, **kw)
            host = mb.Plan(keep_match=True).append(rb)
            dev = mb.Plan(keep_match=True).append(rb, match_device=0)
            # (paired: the lone read of gene 3 is an odd trailing mate, never parsed, solve.c:187)
            assert host.info()[:, 4].tolist() == ([0, 4, 0, 0, 0] if paired else [0, 4, 0, 4, 0])
            same_plan(mb, host, dev, 5, run=False)
    # and against the oracle directly
    dev = mb.Plan(keep_match=True).append(mb.ReadBatch([g], [pos], [cig], 33), match_device=0)
    codes, order = dev.match(0)
    want = port.match_se(ex, iso, np.asarray(pos, np.int32), cig, 33, 1)
    np.testing.assert_array_equal(codes.astype(float), want["match"])
    np.testing.assert_array_equal(order, want["order"])


def test_device_matching_on_the_golden_inputs(mb):
    for case in load_cases():
        g = mb.Gene(case.exons, case.isoforms)
        rb = mb.ReadBatch([g], [case.pos], [case.cig], case.read_len, case.overhang, bool(case.paired), *case.pe,
                          gene_ids=[case.gene_id])
        host = mb.Plan(keep_match=True).append(rb)
        dev = mb.Plan(keep_match=True).append(rb, match_device=0)
        same_plan(mb, host, dev, 1, run=False)
        t, c = dev.classes(0)
        np.testing.assert_array_equal(t, case.class_templates.T)
        np.testing.assert_array_equal(c, case.class_counts)


def test_device_matching_needs_a_valid_device(mb):
    w = mb.Workload(0, 4, 50, 36, 250.0, 900.0, 4.0, seed=1)
    with pytest.raises(mb.InternalError):
        mb.Plan().append(w, match_device=99)
