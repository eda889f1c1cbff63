"""BAM / BAI input without pysam (miso_b200/bam.py): the alignment side of compute_gene_psi
(misopy/sam_utils.py:143-181 opens a pysam.Samfile and fetches the reads of a gene's span).

Checked here, without a GPU: the BGZF / BAM / BAI decoding against files this module wrote itself
(records spanning block boundaries, indexed fetch == scan == the SAM text loader), and -- where the
reference tree is mounted -- against the four BAM + BAI files the reference ships
(misopy/sashimi_plot/test-data/bam-data, written by samtools)."""
import glob
import os
import random

import pytest

from test_run_miso import cfg1, write_fixture

REF_BAM = "/root/reference/misopy/sashimi_plot/test-data/bam-data"


def sorted_reads(sam):
    """The reads of a load_sam() result as one coordinate-ordered list + the reference table."""
    refs = [(name, 200000000) for name in sam]
    reads = []
    for name in sam:
        reads += sorted(sam[name], key=lambda r: r.pos)
    return refs, reads


def test_bam_written_here_reads_back_like_the_sam_text(tmp_path):
    from miso_b200 import bam, run_miso as rm
    case = cfg1()
    gff, sam_path = write_fixture(tmp_path, case)
    sam = rm.load_sam(sam_path)
    refs, reads = sorted_reads(sam)
    path = str(tmp_path / "reads.bam")
    # 700-byte blocks: most records straddle a BGZF block boundary
    bam.write_bam(path, refs, reads, header_text="@HD\tVN:1.0\tSO:coordinate\n", block_bytes=700, index_path=path + ".bai")
    with bam.BamFile(path) as b:
        assert b.has_index and b.references == [name for name, _ in refs]
        assert b.header_text.startswith("@HD")
        assert list(b) == reads                       # every field of every record, file order
        genes = rm.load_gff_genes(gff)
        from test_run_miso import EVENT
        lo, hi = rm.get_inclusive_txn_bounds(genes[EVENT])
        via_sam = rm.fetch_reads_in_gene(sam, "chr10", lo, hi)
        via_bam = rm.fetch_reads_in_gene(b, "chr10", lo, hi)          # "chr10" -> "10": the prefix fallback
        assert sorted(via_bam) == sorted(via_sam) and len(via_bam) > 700
        assert rm.sam_parse_reads(via_bam, given_read_len=36)[1] == len(case.pos)
        assert rm.fetch_reads_in_gene(b, "chrNope", lo, hi) == []
    # the same file without its index: fetch scans
    os.remove(path + ".bai")
    with bam.BamFile(path) as b:
        assert not b.has_index
        assert sorted(rm.fetch_reads_in_gene(b, "chr10", lo, hi)) == sorted(via_sam)
    assert isinstance(rm.load_alignments(path), bam.BamFile) and isinstance(rm.load_alignments(sam_path), dict)


def test_indexed_fetch_equals_scan_on_random_regions(tmp_path):
    from miso_b200 import bam
    rng = random.Random(7)
    R = bam.SamRead
    refs = [("chrA", 3000000), ("chrB", 500000), ("chrEmpty", 1000)]
    reads = []
    for name, length in refs[:2]:
        rs = []
        for i in range(1500):
            pos = rng.randrange(0, length - 5000)
            kind = rng.random()
            if kind < 0.6:
                cig, span = "36M", 36
            elif kind < 0.9:
                n = rng.choice((80, 900, 20000, 140000))          # spliced: long spans land in higher bins
                cig, span = "20M%dN16M" % n, 36 + n
            elif kind < 0.95:
                cig, span = "5S31M", 31
            else:
                cig, span = None, 0                                # placed, unaligned: covers one base
            rs.append(R("%s_r%d" % (name, i), rng.choice((0, 16, 99, 147, 4)), name, pos, cig, 36, pos + span))
        reads += sorted(rs, key=lambda r: r.pos)
    path = str(tmp_path / "rand.bam")
    bam.write_bam(path, refs, reads, block_bytes=4096, index_path=path + ".bai")
    indexed, scan = bam.BamFile(path), bam.BamFile(path, index=os.devnull + ".none")
    scan._index = None
    assert indexed.has_index and not scan.has_index
    for _ in range(300):
        name, length = refs[rng.randrange(2)]
        a = rng.randrange(0, length)
        b = a + rng.choice((1, 50, 3000, 20000, 400000))
        want = [r for r in reads if r.rname == name and r.pos < b and max(r.aend, r.pos + 1) > a]
        assert indexed.fetch(name, a, b) == want
        assert scan.fetch(name, a, b) == want
    assert indexed.fetch("chrEmpty", 0, 1000) == [] and indexed.fetch("chrA", 10, 10) == []
    with pytest.raises(ValueError):
        indexed.fetch("chrZ", 0, 10)
    indexed.close(); scan.close()


def test_bins():
    from miso_b200 import bam
    rng = random.Random(3)
    for _ in range(2000):
        beg = rng.randrange(0, 1 << 29)
        end = min((1 << 29), beg + rng.choice((1, 30, 20000, 1 << 15, 1 << 21, 1 << 27)))
        assert bam._reg2bin(beg, end) in bam.reg2bins(beg, end)
        # every sub-interval's bin is among the bins searched for the whole interval
        mid = rng.randrange(beg, end)
        assert bam._reg2bin(mid, mid + 1) in bam.reg2bins(beg, end)
    assert bam.reg2bins(0, 1) == [0, 1, 9, 73, 585, 4681]


def test_not_a_bam(tmp_path):
    from miso_b200 import bam
    p = tmp_path / "x.bam"
    p.write_bytes(b"\x1f\x8b\x08\x00" + b"\x00" * 40)          # gzip, but no BGZF extra field
    with pytest.raises(bam.BamError):
        bam.BamFile(str(p))
    good = str(tmp_path / "g.bam")
    bam.write_bam(good, [("c", 100)], [bam.SamRead("r", 0, "c", 5, "10M", 10, 15)])
    blob = open(good, "rb").read()
    (tmp_path / "cut.bam").write_bytes(blob[:len(blob) // 2])
    with pytest.raises((bam.BamError, Exception)):
        list(bam.BamFile(str(tmp_path / "cut.bam")))


@pytest.mark.skipif(not os.path.isdir(REF_BAM), reason="the reference tree is not mounted on this box")
def test_reference_held_bam_files():
    """The BAM + BAI files the reference ships (samtools output): every record decodes, files are
    coordinate-sorted, and the indexed fetch returns exactly what a scan of the file returns."""
    from miso_b200 import bam, run_miso as rm
    files = sorted(glob.glob(os.path.join(REF_BAM, "*.sorted.bam")))
    assert len(files) == 4
    rng = random.Random(11)
    total = 0
    for path in files:
        with bam.BamFile(path) as b:
            assert b.has_index and b.references and all(l > 0 for l in b.lengths)
            reads = list(b)
            total += len(reads)
            assert reads, path
            by_ref = {}
            for r in reads:
                by_ref.setdefault(r.rname, []).append(r)
                assert r.rlen > 0 and r.flag >= 0
                if r.cigar is not None:
                    assert rm._reference_span(r.cigar) == r.aend - r.pos
            for name, rs in by_ref.items():
                assert [r.pos for r in rs] == sorted(r.pos for r in rs)
                lo, hi = rs[0].pos, max(r.aend for r in rs)
                assert b.fetch(name, lo, hi) == rs
                for _ in range(40):
                    a = rng.randrange(max(0, lo - 500), hi + 500)
                    z = a + rng.choice((1, 40, 500, 5000))
                    want = [r for r in rs if r.pos < z and max(r.aend, r.pos + 1) > a]
                    assert b.fetch(name, a, z) == want
                # and through the front end's fetch (the "chr" prefix is only ever stripped, sam_utils.py:161-168)
                assert rm.fetch_reads_in_gene(b, name, lo, hi) == [r for r in rs if not (r.flag & 4)]
                if not name.startswith("chr"):
                    assert rm.fetch_reads_in_gene(b, "chr" + name, lo, hi) == [r for r in rs if not (r.flag & 4)]
    assert total > 100
