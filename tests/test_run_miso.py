"""The Python-3 caller of the hot path (miso_b200/run_miso.py, mirror of misopy/run_miso.py
compute_gene_psi + misopy/sam_utils.py) on a fixture written from the golden config-1 case: the
SE event of SE.mm9.gff:45081 in the reference's own GFF layout and its reads as SAM text, plus
decoys the front end must drop."""
import os

import numpy as np
import pytest

from golden_util import load_cases

EVENT = "ENSMUSG00000019943.chr10:98481349:98481531:+@chr10:98481912:98482065:+@chr10:98485442:98488777:+"


def write_fixture(tmp_path, case, paired_decoys=True):
    ex = case.exons
    gff = tmp_path / "SE.gff"
    rows = [("gene", ex[0][0], ex[-1][1], "ID=%s;Name=%s" % (EVENT, EVENT)),
            ("mRNA", ex[0][0], ex[-1][1], "ID=%s.A;Parent=%s" % (EVENT, EVENT)),
            ("mRNA", ex[0][0], ex[-1][1], "ID=%s.B;Parent=%s" % (EVENT, EVENT)),
            # exons of .A deliberately out of order: a transcript's exons are sorted by start
            ("exon", ex[2][0], ex[2][1], "ID=%s.A.dn;Parent=%s.A" % (EVENT, EVENT)),
            ("exon", ex[0][0], ex[0][1], "ID=%s.A.up;Parent=%s.A" % (EVENT, EVENT)),
            ("exon", ex[1][0], ex[1][1], "ID=%s.A.se;Parent=%s.A" % (EVENT, EVENT)),
            ("exon", ex[0][0], ex[0][1], "ID=%s.B.up;Parent=%s.B" % (EVENT, EVENT)),
            ("exon", ex[2][0], ex[2][1], "ID=%s.B.dn;Parent=%s.B" % (EVENT, EVENT)),
            # a second, single-isoform gene and one without reads
            ("gene", 1000, 2000, "ID=solo;Name=solo"), ("mRNA", 1000, 2000, "ID=solo.A;Parent=solo"),
            ("exon", 1000, 2000, "ID=solo.A.e;Parent=solo.A"),
            ("gene", 5000000, 5001000, "ID=empty;Name=empty"), ("mRNA", 5000000, 5001000, "ID=empty.A;Parent=empty"),
            ("mRNA", 5000000, 5001000, "ID=empty.B;Parent=empty"),
            ("exon", 5000000, 5000400, "ID=empty.A.e;Parent=empty.A"), ("exon", 5000600, 5001000, "ID=empty.B.e;Parent=empty.B")]
    with open(gff, "w") as f:
        f.write("##gff-version 3\n")
        for typ, s, e, attr in rows:
            f.write("chr10\tSE\t%s\t%d\t%d\t.\t+\t.\t%s\n" % (typ, s, e, attr))
    sam = tmp_path / "reads.sam"
    seq, qual = "A" * 36, "I" * 36
    with open(sam, "w") as f:
        f.write("@HD\tVN:1.0\tSO:coordinate\n@SQ\tSN:10\tLN:129993255\n")
        # decoys: before the gene, other chromosome, unmapped, wrong read length, no CIGAR
        f.write("d1\t0\t10\t%d\t255\t36M\t*\t0\t0\t%s\t%s\n" % (ex[0][0] - 5000, seq, qual))
        f.write("d2\t0\t11\t%d\t255\t36M\t*\t0\t0\t%s\t%s\n" % (ex[0][0] + 10, seq, qual))
        f.write("d3\t4\t10\t%d\t255\t36M\t*\t0\t0\t%s\t%s\n" % (ex[0][0] + 10, seq, qual))
        for i, (p, c) in enumerate(zip(case.pos, case.cig)):
            f.write("r%d\t%d\t10\t%d\t255\t%s\t*\t0\t0\t%s\t%s\n" % (i, 16 * (i % 2), int(p), c, seq, qual))
            if i == 100:
                f.write("d4\t0\t10\t%d\t255\t30M\t*\t0\t0\t%s\t%s\n" % (int(p), "A" * 30, "I" * 30))
                f.write("d5\t0\t10\t%d\t255\t*\t*\t0\t0\t%s\t%s\n" % (int(p), seq, qual))
        f.write("d6\t0\t10\t%d\t255\t36M\t*\t0\t0\t%s\t%s\n" % (ex[-1][1] + 5000, seq, qual))
    return str(gff), str(sam)


def cfg1():
    return [c for c in load_cases() if c.name == "cfg1_default"][0]


def test_front_end_extracts_the_golden_reads(tmp_path):
    from miso_b200 import run_miso as rm
    case = cfg1()
    gff, sam = write_fixture(tmp_path, case)
    genes = rm.load_gff_genes(gff)
    assert list(genes) == [EVENT, "solo", "empty"]
    gene = rm.make_gene_from_gff_records(EVENT, genes[EVENT])
    assert [[lab.split(".")[-1] for lab in iso.desc] for iso in gene.isoforms] == [["up", "se", "dn"], ["up", "dn"]]
    assert [(p.start, p.end) for p in gene.isoforms[0].parts] == list(case.exons)
    assert (gene.chrom, gene.strand) == ("chr10", "+")
    assert rm.get_inclusive_txn_bounds(genes[EVENT]) == (case.exons[0][0], case.exons[-1][1])
    reads = rm.load_sam(sam)
    raw = rm.fetch_reads_in_gene(reads, "chr10", *rm.get_inclusive_txn_bounds(genes[EVENT]))    # SAM says "10"
    (pos, cig), n = rm.sam_parse_reads(raw, given_read_len=36)
    assert n == len(case.pos)
    np.testing.assert_array_equal(np.asarray(pos) + 1, case.pos)           # 0-based here, +1 in the sampler
    assert list(cig) == list(case.cig)
    # strand rule: fr-firststrand keeps the reads on the gene's strand only
    (pos_f, _), n_f = rm.sam_parse_reads(raw, strand_rule="fr-firststrand", target_strand="+", given_read_len=36)
    assert n_f == (len(case.pos) + 1) // 2


def test_pairing_rules():
    from miso_b200 import run_miso as rm
    R = rm.SamRead
    reads = [R("pair_a_x/1", 99, "1", 100, "36M", 36, 136), R("pair_a_x/2", 147, "1", 300, "36M", 36, 336),      # proper pair
             R("pair_b_x/1", 99, "1", 110, "36M", 36, 146),                                               # mate missing
             R("pair_c_x/1", 65, "1", 120, "36M", 36, 156), R("pair_c_x/2", 129, "1", 320, "36M", 36, 356),      # same strand
             R("pair_d_x/1", 99 | 0x200, "1", 130, "36M", 36, 166), R("pair_d_x/2", 147, "1", 330, "36M", 36, 366),   # QC fail
             R("pair_e_x/1", 99, "1", 140, "30M", 30, 170), R("pair_e_x/2", 147, "1", 340, "36M", 36, 376),      # wrong length
             R("pair_f_x/1", 83, "1", 350, "36M", 36, 386), R("pair_f_x/2", 163, "1", 150, "36M", 36, 186)]      # read1 reverse
    (pos, cig), n = rm.sam_parse_reads(reads, paired_end=True, given_read_len=36)
    assert n == 2 and pos == (100, 300, 350, 150)
    (pos, cig), n = rm.sam_parse_reads(reads, paired_end=True, strand_rule="fr-firststrand", target_strand="+",
                                       given_read_len=36)
    # fr-firststrand puts the reverse /1 read second (sam_utils.py:236-247); pair a then fails the '+' rule? no:
    # a: read1 forward -> kept; f: swapped to (f/2 forward, f/1 reverse) -> first mate '+' -> kept
    assert n == 2 and pos == (100, 300, 150, 350)
    assert rm.strip_mate_id("x#1") == "" and rm.strip_mate_id("name/2") == "nam"        # three characters, as in the reference


@pytest.mark.gpu
def test_compute_gene_psi_reproduces_the_golden_posterior(tmp_path):
    import miso_b200
    if miso_b200.device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu tests must run on a B200")
    from miso_b200 import run_miso as rm, miso_format
    case = cfg1()
    gff, sam = write_fixture(tmp_path, case)
    out_dir = str(tmp_path / "out")
    res = rm.compute_gene_psi([EVENT, "solo", "empty", "nope"], gff, sam, out_dir, 36, seed=case.seed)
    assert res["solo"].startswith("skipped") and res["empty"].startswith("skipped") and res["nope"].startswith("skipped")
    path = res[EVENT]
    assert path == os.path.join(out_dir, "chr10", EVENT + ".miso") and os.path.isfile(path)
    samples, header, scores, _, _, counts = miso_format.load_samples(path)
    np.testing.assert_allclose(samples, np.round(case.samples.T, 4), atol=1.1e-4)
    np.testing.assert_allclose(scores, case.loglik, atol=6e-3)
    assert header["iters"] == "5000" and header["burn_in"] == "500" and header["lag"] == "10"
    assert header["chrom"] == "chr10" and header["strand"] == "+"
    # a second call finds the file and skips the gene (miso_sampler.py:233-238)
    assert rm.compute_gene_psi([EVENT], gff, sam, out_dir, 36, seed=case.seed)[EVENT] == "skipped: output exists"


def test_paired_end_insert_model_is_truncated_like_the_reference(monkeypatch):
    """misopy/run_miso.py:81-83: mean_frag_len = int(mean), frag_variance = int(sd) ** 2 -- miso.py
    forwards the two numbers as %.1f, so fractional values are normal."""
    from miso_b200 import run_miso, batch
    seen = {}

    class Stop(Exception):
        pass

    def fake_read_batch(gs, poss, cigs, read_len, overhang, paired, mean, var, devs):
        seen.update(mean=mean, var=var, devs=devs)
        raise Stop()

    monkeypatch.setattr(batch, "ReadBatch", fake_read_batch)
    gff = os.path.join(os.path.dirname(__file__), "golden", "_pe_trunc.gff")
    sam = os.path.join(os.path.dirname(__file__), "golden", "_pe_trunc.sam")
    try:
        with open(gff, "w") as f:
            f.write("chr1\tt\tgene\t1\t1000\t.\t+\t.\tID=g1\n")
            for t, exons in (("t1", ((1, 100), (301, 400), (601, 700))), ("t2", ((1, 100), (601, 700)))):
                f.write("chr1\tt\tmRNA\t1\t700\t.\t+\t.\tID=%s;Parent=g1\n" % t)
                for a, b in exons:
                    f.write("chr1\tt\texon\t%d\t%d\t.\t+\t.\tParent=%s\n" % (a, b, t))
        with open(sam, "w") as f:
            for i in range(30):
                f.write("r%d\t99\tchr1\t%d\t255\t36M\t=\t%d\t0\t%s\t*\n" % (i, 5 + i, 330 + i, "A" * 36))
                f.write("r%d\t147\tchr1\t%d\t255\t36M\t=\t%d\t0\t%s\t*\n" % (i, 330 + i, 5 + i, "A" * 36))
        with pytest.raises(Stop):
            run_miso.compute_gene_psi(["g1"], gff, sam, os.path.join(os.path.dirname(gff), "_out"), 36,
                                      paired_end=(251.7, 15.9), settings={"filter_reads": False})
    finally:
        for p in (gff, sam):
            if os.path.exists(p):
                os.remove(p)
    assert seen == dict(mean=251.0, var=225.0, devs=4.0)


def test_exons_without_id_and_shared_labels():
    """Exons without an ID attribute are named parent@start@end@strand (gff_utils.py:370-376) and an
    isoform is the first part carrying each of its labels, in its own order (Gene.py:305-321)."""
    from miso_b200.sampler import GeneModel, Part
    parts = [Part("a", 1, 10), Part("b", 20, 30), Part("c", 40, 50), Part("a", 1, 10), Part("c", 40, 50)]
    g = GeneModel("g", parts, [["a", "b", "c"], ["a", "c"]])
    assert [len(i.parts) for i in g.isoforms] == [3, 2]
    assert [p.label for p in g.isoforms[1].parts] == ["a", "c"]


def test_settings_file(tmp_path):
    """misopy/settings.py: ConfigParser file, sections ignored, literals evaluated, defaults and checks of
    get_sampler_params / get_min_event_reads / get_strand_param."""
    from miso_b200 import run_miso as rm
    ref = "/root/reference/misopy/settings/miso_settings.txt"
    if os.path.isfile(ref):
        st = rm.load_settings(ref)
        assert st == dict(rm.DEFAULT_SETTINGS)          # the shipped file IS the defaults
    p = tmp_path / "s.txt"
    p.write_text("[data]\nfilter_results = True\nmin_event_reads = 5\nstrand = fr-firststrand\n"
                 "[cluster]\ncluster_command = long\n[sampler]\nburn_in = 100\nlag = 2\nnum_iters = 1000\n")
    st = rm.load_settings(str(p))
    assert (st["burn_in"], st["lag"], st["num_iters"], st["num_chains"]) == (100, 2, 1000, 6)
    assert st["min_event_reads"] == 5 and st["strand_rule"] == "fr-firststrand" and st["filter_reads"] is True
    p.write_text("[sampler]\nburn_in = 100\nlag = 2\n")
    with pytest.raises(ValueError, match="num_iters"):
        rm.load_settings(str(p))
    p.write_text("[data]\nstrand = sideways\n[sampler]\nburn_in = 1\nlag = 1\nnum_iters = 10\n")
    with pytest.raises(ValueError, match="Invalid strand"):
        rm.load_settings(str(p))
    with pytest.raises(FileNotFoundError):
        rm.load_settings(str(tmp_path / "missing.txt"))
