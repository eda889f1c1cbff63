"""`.miso` writer / parser / credible intervals (miso_b200/miso_format.py) --
the on-disk contract of misopy/miso_sampler.py:376-466 and
misopy/credible_intervals.py:4-55."""
import os

import numpy as np
import pytest

from miso_b200 import miso_format as mf


def test_header_and_body_format(tmp_path):
    rng = np.random.default_rng(0)
    psi = rng.dirichlet(np.ones(2), size=30)
    sc = -1000 - rng.random(30) * 10
    ass = np.array([0, 0, 1, -1, 1, 1, 0])
    h = mf.format_header([["A", "B", "C"], ["A", "C"]], [("A", 183), ("B", 154), ("C", 3336)], 5000, 500, 10,
                         23.456, "drift", [(0.0, 1.0), (1.0, 0.0), (1.0, 1.0)], [5.0, 7.0, 700.0], ass, "chr10",
                         "+", [98481349, 98481349], [98488777, 98488777])
    assert h.startswith("#isoforms=['A_B_C','A_C']\texon_lens=('A',183),('B',154),('C',3336)\titers=5000\t"
                        "burn_in=500\tlag=10\tpercent_accept=23.46\tproposal_type=drift\t"
                        "counts=(0,1):5,(1,0):7,(1,1):700\tassigned_counts=0:3,1:3\tchrom=chr10\tstrand=+\t")
    assert h.endswith("mRNA_starts=98481349,98481349\tmRNA_ends=98488777,98488777\n")
    path = str(tmp_path / "ev.miso")
    mf.write_miso(path, h, psi, sc)
    lines = open(path).read().splitlines()
    assert lines[1] == "sampled_psi\tlog_score" and len(lines) == 32
    assert lines[2] == "%.4f,%.4f\t%.2f" % (psi[0, 0], psi[0, 1], sc[0])
    samples, header, scores, smap, smap_score, counts = mf.load_samples(path)
    np.testing.assert_allclose(samples, np.round(psi, 4), atol=1e-12)
    np.testing.assert_allclose(scores, np.round(sc, 2), atol=1e-9)
    assert header["iters"] == "5000" and counts == "(0,1):5,(1,0):7,(1,1):700"
    assert smap == [float(v) for v in max(ln.split("\t")[0] for ln in lines[2:]).split(",")]


def test_header_without_chrom():
    h = mf.format_header(["iso1", "iso2"], [("e", 10)], 10, 1, 1, 50.0, "drift", [], [], np.array([-1, -1]))
    assert "chrom=NA\tstrand=NA" in h and "assigned_counts=\t" in h


def test_credible_interval_indices_follow_numpy_round():
    # credible_intervals.py:45-49 with `from numpy import *`: half-to-even on the fp64 product
    for n in (20, 60, 100, 450, 900, 2700, 1001):
        alpha = 1 - 0.95
        lo, hi = mf.credible_interval_indices(n)
        assert lo == int(np.round((alpha / 2) * n)) - 1 and hi == int(np.round((1 - alpha / 2) * n)) - 1
    s = np.linspace(0, 1, 2700)[::-1].copy()
    lo, hi = mf.compute_credible_intervals(np.stack([s, 1 - s], axis=1))
    assert (lo, hi) == (np.sort(s)[67], np.sort(s)[2631])      # 0.025*2700 = 67.5(+eps) -> 68
    out = mf.format_credible_intervals("ev", np.stack([s, 1 - s], axis=1))
    assert out[0] == "ev" and out[1] == "0.50"
    out3 = mf.format_credible_intervals("ev", np.random.default_rng(1).dirichlet(np.ones(3), size=500))
    assert len(out3) == 4 and out3[1].count(",") == 2


def test_count_isoform_assignments():
    assert mf.count_isoform_assignments(np.array([0, 2, 2, -1])) == [(0, 1), (1, 0), (2, 2)]
    assert mf.count_isoform_assignments(np.array([-1, -1])) == []   # range(max+1) is empty, reads_utils.py:42-45


def test_bayes_factor_matches_scipy_kde():
    """misopy's gaussian_kde_covfact(delta, 0.3) evaluated at 0 (hypothesis_test.py:41-57,168-169)."""
    from scipy import stats

    class kde_covfact(stats.gaussian_kde):
        def covariance_factor(self):
            return 0.3

    rng = np.random.default_rng(3)
    s1 = rng.dirichlet([4, 2, 6], size=450)
    s2 = rng.dirichlet([5, 2, 5], size=450)
    got = mf.bayes_factor(s1, s2)
    for k in range(3):
        want = 1.0 / kde_covfact(s1[:, k] - s2[:, k]).evaluate([0])[0]
        assert abs(got[k] - want) <= 1e-9 * want
    # peaked on the null
    assert mf.bayes_factor(s1, s1 + 1e-4)[0] == 0.0
    # far apart: capped
    a = np.stack([np.full(100, 0.9) + rng.normal(0, 1e-3, 100), np.full(100, 0.1)], axis=1)
    b = np.stack([np.full(100, 0.1) + rng.normal(0, 1e-3, 100), np.full(100, 0.9)], axis=1)
    assert mf.bayes_factor(a, b)[0] == 1e12
    line = mf.format_bf_line("ev", s1[:, :2], s2[:, :2], got, {"isoforms": "['a','b']"}, {})
    assert line.count("\t") == len(mf.BF_HEADER) - 1


def test_summarize_and_compare_directories(tmp_path):
    """summarize_miso / compare_miso over sample directories (samples_utils.py:263-329,
    hypothesis_test.py:182-345)."""
    from miso_b200 import postprocess as pp
    rng = np.random.default_rng(5)
    for label, shift in (("ctrl", 0.0), ("kd", 0.5)):
        for ev, K in (("evA", 2), ("evB", 3)):
            psi = rng.dirichlet(np.ones(K) * 200, size=300)
            psi[:, 0] = np.clip(psi[:, 0] + shift, 0, 1)
            psi /= psi.sum(axis=1, keepdims=True)
            header = mf.format_header([["a%d" % k] for k in range(K)], [("a%d" % k, 100) for k in range(K)], 3500, 500,
                                      10, 40.0, "drift", tuple((1.0,) * K for _ in range(1)), (300.0,),
                                      np.zeros(300, int), "chr1", "+", [10] * K, [900] * K)
            d = tmp_path / label / "chr1"
            d.mkdir(parents=True, exist_ok=True)
            mf.write_miso(str(d / (ev + ".miso")), header, psi, np.zeros(300))
    (tmp_path / "kd" / "chr1" / "evC.miso").write_text((tmp_path / "kd" / "chr1" / "evA.miso").read_text())
    n = pp.summarize_sampler_results(str(tmp_path / "ctrl"), str(tmp_path / "out" / "ctrl.miso_summary"))
    assert n == 2
    lines = (tmp_path / "out" / "ctrl.miso_summary").read_text().splitlines()
    assert lines[0].split("\t") == pp.SUMMARY_HEADER
    fa = lines[1].split("\t")
    assert fa[0] == "evA" and fa[4] == "'a0','a1'" and fa[7:9] == ["chr1", "+"]
    smp = mf.load_samples(str(tmp_path / "ctrl" / "chr1" / "evA.miso"))[0]
    assert fa[1] == "%.2f" % smp.mean(axis=0)[0]
    assert len(lines[2].split("\t")[1].split(",")) == 3              # multi-isoform: comma-separated means
    path, m = pp.output_samples_comparison(str(tmp_path / "ctrl"), str(tmp_path / "kd"), str(tmp_path / "cmp"))
    assert m == 2 and path.endswith(os.path.join("ctrl_vs_kd", "bayes-factors", "ctrl_vs_kd.miso_bf"))
    rows = [ln.split("\t") for ln in open(path).read().splitlines()]
    assert rows[0] == mf.BF_HEADER and [r[0] for r in rows[1:]] == ["evA", "evB"]
    s2 = mf.load_samples(str(tmp_path / "kd" / "chr1" / "evA.miso"))[0]
    assert float(rows[1][8]) == float("%.2f" % mf.bayes_factor(smp, s2)[0]) and float(rows[1][8]) > 100
    assert abs(float(rows[1][7]) - (smp.mean(axis=0)[0] - s2.mean(axis=0)[0])) < 0.011


def test_cli_summarize_and_compare(tmp_path, capsys):
    from miso_b200 import run_miso
    rng = np.random.default_rng(2)
    for label in ("s1", "s2"):
        psi = rng.dirichlet(np.ones(2) * 50, size=100)
        header = mf.format_header([["a"], ["b"]], [("a", 10), ("b", 20)], 1500, 500, 10, 50.0, "drift", ((1.0, 1.0),),
                                  (100.0,), np.zeros(100, int), "chr2", "-", [1, 1], [9, 9])
        d = tmp_path / label / "chr2"
        d.mkdir(parents=True)
        mf.write_miso(str(d / "ev.miso"), header, psi, np.zeros(100))
    run_miso.main(["--summarize-samples", str(tmp_path / "s1"), str(tmp_path / "out")])
    assert (tmp_path / "out" / "summary" / "s1.miso_summary").read_text().count("\n") == 2
    run_miso.main(["--compare-samples", str(tmp_path / "s1"), str(tmp_path / "s2"), str(tmp_path / "out")])
    assert (tmp_path / "out" / "s1_vs_s2" / "bayes-factors" / "s1_vs_s2.miso_bf").is_file()
    assert "1 events" in capsys.readouterr().out


# ---- batched writer (miso_b200/csrc/writer.cpp) -------------------------------------------------

def test_fixed_point_formatting_equals_python_percent_operator():
    """The writer's "%.4f" / "%.2f": round-half-even on the exact binary value, as CPython's "%"
    (miso_sampler.py:461-463), incl. exact ties (multiples of 1/32, x.125), signs, non-finite."""
    import ctypes as C
    from miso_b200._lib import lib
    buf = C.create_string_buffer(64)
    rng = np.random.default_rng(3)
    vals = list(rng.random(20000)) + list(-rng.random(5000) * 1e5) + [k / 32 for k in range(-64, 64)]
    vals += [(2 * m + 1) * 625 / 20000.0 for m in range(500)] + [(2 * m + 1) / 8.0 for m in range(200)]
    vals += [0.0, -0.0, 1.0, 0.99995, 0.00005, 1e-300, -1e-9, 123456.785, float("inf"), float("-inf"), float("nan"), 1e15]
    for v in vals:
        for d, fmt in ((2, "%.2f"), (4, "%.4f")):
            n = lib.misob200_format_fixed(float(v), d, buf)
            assert buf.value.decode() == fmt % v and n == len(fmt % v), (v, d)


def test_batched_writer_equals_the_python_writer(tmp_path):
    """misob200_plan_write_miso on a whole plan = format_header + write_miso gene by gene, byte for byte
    (plan and layout are host-side; the output buffers are filled with random posteriors here)."""
    import miso_b200 as mb
    w = mb.Workload(1, 40, 120, 36, 250.0, 900.0, 4.0, seed=4)
    plan = mb.Plan().append(w)
    params = mb.make_params(300, 60, 7, 2, seed=1)
    G = plan.size()[0]
    info = plan.info()
    ns, nl, na = plan.output_sizes(params)
    rng = np.random.default_rng(9)
    out = dict(samples=rng.random(ns), loglik=-rng.random(nl) * 5000, params=params,
               assignment=np.zeros(na, np.int32), rundata=np.zeros((G, 9), np.int32), status=np.zeros(G, np.int32))
    so, lo, ao = plan.offsets(params)
    for g in range(G):
        K, R = int(info[g, 0]), int(info[g, 1])
        out["assignment"][ao[g]:ao[g] + R] = rng.integers(-1, K, size=R)
        out["rundata"][g, 5:7] = (rng.integers(1, 300), rng.integers(0, 300))
    out["assignment"][ao[3]:ao[3] + int(info[3, 1])] = -1           # no compatible read: no file (miso_sampler.py:352-354)
    metas, paths, pre, suf = [], [], [], []
    for g in range(G):
        K = int(info[g, 0])
        descs = [["e%d" % e for e in range(K + 1) if e != k or k == 0] for k in range(K)]
        meta = dict(isoform_descs=descs, exon_lens=[("e%d" % e, 200) for e in range(K + 1)], chrom="chr%d" % (g % 3),
                    strand="+-"[g % 2], mRNA_starts=[1] * K, mRNA_ends=[400 * K + 200] * K)
        metas.append(meta)
        a, b = mf.header_static_parts(**meta)
        pre.append(a)
        suf.append(b)
        paths.append(None if g == 5 else str(tmp_path / ("g%d.miso" % g)))
    nf, nb = plan.write_miso(out, paths, pre, suf, n_threads=3)
    assert nf == G - 2 and not os.path.exists(str(tmp_path / "g3.miso")) and not os.path.exists(str(tmp_path / "g5.miso"))
    total = 0
    for g in range(G):
        if g in (3, 5):
            continue
        r = plan.gene_result(out, g)
        templ, counts = plan.classes(g)
        acc, rej = int(r["rundata"][5]), int(r["rundata"][6])
        m = metas[g]
        header = mf.format_header(m["isoform_descs"], m["exon_lens"], 300, 60, 7, float(acc) / (acc + rej) * 100, "drift",
                                  templ, counts, r["assignment"], m["chrom"], m["strand"], m["mRNA_starts"], m["mRNA_ends"])
        ref = str(tmp_path / "ref.miso")
        mf.write_miso(ref, header, r["samples"].T, r["loglik"])
        want = open(ref, "rb").read()
        got = open(paths[g], "rb").read()
        assert got == want, "gene %d" % g
        total += len(got)
    assert nb == total


REF_MISO = "/root/reference/misopy/sashimi_plot/test-data/miso-data"


@pytest.mark.skipif(not os.path.isdir(REF_MISO), reason="the reference tree is not mounted on this box")
def test_reference_held_miso_files_round_trip(tmp_path):
    """SURVEY.md section 8c fixture 4: the four .miso files the reference ships.  The parser reads them, the
    summary numbers equal an independent numpy computation on the raw text, and both writers re-emit
    header and psi column byte for byte (the reference wrote them with the same format strings)."""
    import glob
    import ctypes as C
    from miso_b200._lib import lib, ptr
    files = sorted(glob.glob(os.path.join(REF_MISO, "*", "chr17", "*.miso")))
    assert len(files) == 4
    for path in files:
        raw = open(path).read()
        lines = raw.splitlines()
        samples, header, scores, smap, smap_score, counts = mf.load_samples(path)
        body = [ln.split("\t") for ln in lines[2:]]
        psi = np.array([[float(v) for v in b[0].split(",")] for b in body])
        np.testing.assert_array_equal(samples, psi)
        np.testing.assert_array_equal(scores, np.array([float(b[1]) for b in body]))
        assert set(header) >= {"isoforms", "exon_lens", "iters", "burn_in", "lag", "percent_accept", "proposal_type",
                               "counts", "assigned_counts"}
        assert counts == header["counts"] and samples.shape[1] == 2
        # summary fields (credible_intervals.py:4-28): mean and order statistics
        n = len(psi)
        srt = np.sort(psi[:, 0])
        lo, hi = int(np.round(0.025 * n)) - 1, int(np.round(0.975 * n)) - 1
        assert mf.format_credible_intervals("ev", samples) == ["ev", "%.2f" % psi[:, 0].mean(), "%.2f" % srt[lo], "%.2f" % srt[hi]]
        # python writer: header, column line and the psi column byte for byte; these legacy files carry
        # the log score with four decimals, the current writer (miso_sampler.py:461-463) prints two
        def same_but_score_decimals(text):
            got = text.splitlines()
            assert got[:2] == lines[:2] and len(got) == len(lines)
            for a, b in zip(got[2:], lines[2:]):
                (pa, sa), (pb, sb) = a.split("\t"), b.split("\t")
                assert pa == pb and sa == "%.2f" % float(sb)
        out = str(tmp_path / "py.miso")
        mf.write_miso(out, lines[0] + "\n", samples, scores)
        same_but_score_decimals(open(out).read())
        # batched C++ writer, same bytes
        out2 = str(tmp_path / "cc.miso")
        s = np.ascontiguousarray(samples, np.float64)
        sc = np.ascontiguousarray(scores, np.float64)
        off = np.zeros(1, np.int64)
        k = np.array([2], np.int32)
        pa = (C.c_char_p * 1)(out2.encode())
        hd = (C.c_char_p * 1)((lines[0] + "\n").encode())
        nb = C.c_int64()
        assert lib.misob200_write_miso_files(1, pa, hd, ptr(s), ptr(off), ptr(sc), ptr(off), ptr(k), n, 1, C.addressof(nb)) == 0
        same_but_score_decimals(open(out2).read())
        assert open(out2).read() == open(out).read() and nb.value == os.path.getsize(out)
