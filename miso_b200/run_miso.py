"""``compute_gene_psi`` in Python 3: annotation + alignments -> ``.miso`` files on the B200.

The caller of the hot path in the reference is ``misopy/run_miso.py:34-202``
(``compute_gene_psi``): load the gene from the indexed GFF, fetch the BAM reads in
its span, pair / strand-filter them (``misopy/sam_utils.py:210-456``), call
``MISOSampler.run_sampler`` once per gene.  That module is Python 2 and needs
``pysam``; this one restates the same steps for what this image can read --
**GFF3 text, and SAM text or BAM (+ BAI) alignments** (``miso_b200/bam.py``: BGZF, BAM
records and the BAI index read directly, no pysam) -- and, instead of one sampler call
per gene, puts all genes of the call into ONE plan, so the device sees a batch
(SURVEY.md section 8f-4).

Mirrored rules (file:line of the reference):
  * gene construction from gene -> mRNA/transcript -> exon records, transcripts in
    file order, exons of a transcript sorted by start, exon label = its ID
    (``misopy/Gene.py:920-1015``); the most inclusive mRNA bounds are the fetch region
    (``misopy/gff_utils.py:955-984``)
  * ``chr`` prefix fallback when the alignment file names chromosomes differently
    (``sam_utils.py:161-168``); region fetch = alignments overlapping [start, end)
  * single-end: every read with a CIGAR and ``rlen == read_len`` (``sam_utils.py:427-447``)
  * paired-end: mates grouped by name without ``/1 /2 #1 #2`` (``:199-208``), QC-fail /
    unmapped / mate-unmapped / unpaired reads dropped, exactly two mates, opposite
    strands, both with a CIGAR and the given length, mates consecutive (``:210-425``)
  * strand rules ``fr-unstranded`` (default) and ``fr-firststrand`` (``:303-351``)
  * skip rules of ``compute_gene_psi`` (all isoforms shorter than the read,
    ``run_miso.py:110-113``; fewer than ``min_event_reads`` reads, ``:134-141``) and of
    ``run_sampler`` (no reads, one isoform, output exists, every read unassignable,
    ``miso_sampler.py:229-277,352-354``)
  * defaults of ``misopy/settings/miso_settings.txt``: 5000 iterations, burn-in 500,
    lag 10, 6 chains, at least 20 reads per event; ``load_settings`` reads such a file
    (``misopy/settings.py:19-145``)
  * output ``<output_dir>/<chrom>/<gene id>.miso`` in the format of
    ``miso_sampler.py:376-466`` (``miso_b200/miso_format.py``)
Pairs are passed in order of first appearance of the read name (the reference iterates a
Python-2 dict, i.e. in no defined order).
"""
import os
from collections import OrderedDict, namedtuple

import numpy as np

from . import batch as _batch
from .bam import BamFile, SamRead
from .miso_format import format_header, write_miso
from .sampler import GeneModel, Part

DEFAULT_SETTINGS = dict(num_iters=5000, burn_in=500, lag=10, num_chains=6, min_event_reads=20,
                        filter_reads=True, strand_rule="fr-unstranded")

GffRecord = namedtuple("GffRecord", "seqid source type start end strand id parent")


# ---- settings file ------------------------------------------------------------------------
def _try_eval(text):
    """``parse_csv.tryEval`` (``misopy/parse_csv.py:88-92``) without ``eval``: a Python literal, else the string."""
    import ast
    try:
        return ast.literal_eval(text)
    except (ValueError, SyntaxError):
        return text


def load_settings(path):
    """A MISO settings file (``misopy/settings/miso_settings.txt``) as the ``settings`` dict of
    ``compute_gene_psi``.  ``misopy/settings.py:19-60``: ConfigParser format, section headers ignored,
    values evaluated as literals except in ``[cluster]``.  What the path reads (``run_miso.py:68-97``,
    ``settings.py:62-145``): ``burn_in``, ``lag``, ``num_iters`` (required), ``num_chains`` (default 6),
    ``min_event_reads`` (default 20), ``strand`` (default fr-unstranded, validated), ``filter_reads``
    (default True).  ``num_processors`` and the cluster options have no meaning here (one batch on the GPU)."""
    import configparser
    if not os.path.isfile(path):
        raise FileNotFoundError("Error: Settings file %s does not exist." % path)
    cfg = configparser.ConfigParser()
    cfg.read(path)
    raw = {}
    for section in cfg.sections():
        for option in cfg.options(section):
            v = cfg.get(section, option)
            raw[option] = str(v) if section == "cluster" else _try_eval(v)
    for name in ("burn_in", "lag", "num_iters"):
        if name not in raw:
            raise ValueError("Error: need %s parameter to be set in settings file." % name)
    st = dict(DEFAULT_SETTINGS)
    for name in ("burn_in", "lag", "num_iters", "num_chains", "min_event_reads", "filter_reads"):
        if name in raw:
            st[name] = raw[name]
    if "strand" in raw:
        if raw["strand"] not in ("fr-unstranded", "fr-firststrand", "fr-secondstrand"):
            raise ValueError("Error: Invalid strand parameter %s" % raw["strand"])
        st["strand_rule"] = raw["strand"]
    return st


# ---- GFF3 ---------------------------------------------------------------------------------
def _attrs(field):
    out = {}
    for kv in field.strip().split(";"):
        if "=" in kv:
            k, v = kv.split("=", 1)
            out[k.strip()] = v.strip()
    return out


def load_gff_genes(path):
    """gene id -> {'record', 'mRNAs': {mRNA id -> {'record', 'exons': {exon id -> record}}}}, file order."""
    recs = []
    with open(path) as f:
        for line in f:
            if not line.strip() or line.startswith("#"):
                continue
            c = line.rstrip("\n").split("\t")
            if len(c) < 9:
                continue
            a = _attrs(c[8])
            # multi-valued attributes: the first value counts (gff_utils.py:463-470 get_value);
            # an exon without ID is named parent@start@end@strand (gff_utils.py:370-376)
            rid = a["ID"].split(",")[0].rstrip() if "ID" in a else None
            parent = a["Parent"].split(",")[0].rstrip() if "Parent" in a else None
            if c[2] == "exon" and rid is None:
                rid = "%s@%s@%s@%s" % (parent or "", c[3], c[4], c[6])
            recs.append(GffRecord(c[0], c[1], c[2], int(c[3]), int(c[4]), c[6], rid, parent))
    genes = OrderedDict()
    for r in recs:
        if r.type == "gene":
            genes[r.id] = dict(record=r, mRNAs=OrderedDict())
    mrna_gene = {}
    for r in recs:
        if r.type in ("mRNA", "transcript") and r.parent in genes:
            genes[r.parent]["mRNAs"][r.id] = dict(record=r, exons=OrderedDict())
            mrna_gene[r.id] = r.parent
    for r in recs:
        if r.type == "exon" and r.parent in mrna_gene:
            genes[mrna_gene[r.parent]]["mRNAs"][r.parent]["exons"][r.id] = r
    return genes


def make_gene_from_gff_records(gene_id, hierarchy):
    """``misopy/Gene.py:920-1015``."""
    parts, isoform_desc = [], []
    chrom, strand = None, "NA"
    if not hierarchy["mRNAs"]:
        raise ValueError("Error: %s has no transcripts..." % gene_id)
    for tid, info in hierarchy["mRNAs"].items():
        chrom, strand = info["record"].seqid, info["record"].strand
        if not info["exons"]:
            continue
        exons = sorted((Part(e.id, e.start, e.end) for e in info["exons"].values()), key=lambda p: p.start)
        parts.extend(exons)
        isoform_desc.append([p.label for p in exons])
    return GeneModel(gene_id, parts, isoform_desc, chrom=chrom, strand=strand)


def get_inclusive_txn_bounds(hierarchy):
    """``misopy/gff_utils.py:955-984``."""
    starts = [m["record"].start for m in hierarchy["mRNAs"].values()]
    ends = [m["record"].end for m in hierarchy["mRNAs"].values()]
    return min(starts), max(ends)


# ---- SAM ----------------------------------------------------------------------------------
def _reference_span(cigar):
    n, num = 0, ""
    for ch in cigar:
        if ch.isdigit():
            num += ch
        else:
            if ch in "MDN=X":
                n += int(num or 0)
            num = ""
    return n


def load_sam(path):
    """All alignment lines of a SAM text file; per reference name in file order."""
    by_ref = OrderedDict()
    with open(path) as f:
        for line in f:
            if line.startswith("@"):
                continue
            c = line.rstrip("\n").split("\t")
            if len(c) < 11:
                continue
            cigar = None if c[5] == "*" else c[5]
            pos = int(c[3]) - 1                                      # pysam's 0-based pos
            aend = pos + (_reference_span(cigar) if cigar else 0)
            rlen = len(c[9]) if c[9] != "*" else 0
            by_ref.setdefault(c[2], []).append(SamRead(c[0], int(c[1]), c[2], pos, cigar, rlen, aend))
    return by_ref


def load_alignments(path):
    """``sam_utils.load_bam_reads`` (``sam_utils.py:143-152``): a BAM file (by its magic bytes; indexed
    when ``<path>.bai`` exists) or SAM text."""
    with open(path, "rb") as f:
        magic = f.read(4)
    if magic[:2] == b"\x1f\x8b":
        return BamFile(path)
    return load_sam(path)


def fetch_reads_in_gene(sam, chrom, start, end):
    """``sam_utils.py:155-181``: region fetch with the ``chr`` prefix fallback.  ``sam``: what
    ``load_sam`` or ``load_alignments`` returned."""
    names = sam.references if isinstance(sam, BamFile) else sam
    if chrom not in names:
        parts = chrom.split("chr")
        chrom = parts[0] if len(parts) <= 1 else parts[1]
    if isinstance(sam, BamFile):
        try:
            reads = sam.fetch(chrom, start, end)
        except ValueError:                      # "Cannot fetch reads in region" (sam_utils.py:171-174)
            reads = []
        return [r for r in reads if not (r.flag & 4)]
    return [r for r in sam.get(chrom, ()) if r.pos < end and r.aend > start and not (r.flag & 4)]


def flag_to_strand(flag):
    return "-" if flag & 16 else "+"


def strip_mate_id(name):
    if name.endswith(("/1", "/2", "#1", "#2")):
        name = name[0:-3]                  # sic: the reference drops three characters (sam_utils.py:207)
    return name


def pair_sam_reads(reads, strand_rule=None):
    """``sam_utils.py:210-291``."""
    paired = OrderedDict()
    for r in reads:
        name = strip_mate_id(r.qname)
        if (r.flag & 0x200) or (r.flag & 0x4) or (r.flag & 0x8) or not (r.flag & 0x1):
            continue
        paired.setdefault(name, []).append(r)
        if len(paired[name]) == 2 and strand_rule == "fr-firststrand":
            first = paired[name][0]
            if (first.flag & 0x40) and (first.flag & 0x10):
                paired[name] = paired[name][::-1]
            first = paired[name][0]
            if (first.flag & 0x80) and (first.flag & 0x10):
                paired[name] = paired[name][::-1]
    out = OrderedDict()
    for name, rs in paired.items():
        if len(rs) != 2:
            continue
        if flag_to_strand(rs[0].flag) == flag_to_strand(rs[1].flag):
            continue
        out[name] = rs
    return out


def read_matches_strand(read, target_strand, strand_rule, paired_end=None):
    """``sam_utils.py:303-351``."""
    if strand_rule == "fr-unstranded":
        return True
    if strand_rule == "fr-secondstrand":
        raise Exception("fr-secondstrand currently unsupported.")
    if strand_rule != "fr-firststrand":
        raise Exception("Unknown strandedness rule.")
    if paired_end:            # (the reference tests `is not None`; its callers pass None for single-end)
        r1, r2 = read
        if target_strand == "+":
            return flag_to_strand(r1.flag) == "+"
        if target_strand == "-":
            return flag_to_strand(r2.flag) == "-"
        return None
    return flag_to_strand(read.flag) == target_strand


def sam_parse_reads(reads, paired_end=False, strand_rule=None, target_strand=None, given_read_len=None):
    """``sam_utils.py:354-456``: ((positions 0-based, CIGARs), number of reads / pairs)."""
    pos, cig, n = [], [], 0
    check = not (strand_rule is None or strand_rule == "fr-unstranded" or target_strand is None)
    if paired_end:
        for name, (r1, r2) in pair_sam_reads(reads, strand_rule=strand_rule).items():
            if check and not read_matches_strand((r1, r2), target_strand, strand_rule, paired_end=paired_end):
                continue
            if r1.cigar is None or r2.cigar is None:
                continue
            if given_read_len is not None and (r1.rlen != given_read_len or r2.rlen != given_read_len):
                continue
            pos += [r1.pos, r2.pos]
            cig += [r1.cigar, r2.cigar]
            n += 1
    else:
        for r in reads:
            if r.cigar is None:
                continue
            if given_read_len is not None and r.rlen != given_read_len:
                continue
            if check and not read_matches_strand(r, target_strand, strand_rule, paired_end=paired_end):
                continue
            pos.append(r.pos)
            cig.append(r.cigar)
            n += 1
    return (tuple(pos), tuple(cig)), n


# ---- the driver ---------------------------------------------------------------------------
def compute_gene_psi(gene_ids, gff_filename, sam_filename, output_dir, read_len, overhang_len=1,
                     paired_end=None, settings=None, seed=None, device=0, verbose=False):
    """``misopy/run_miso.py:34-202`` for a set of genes, as one batch on the device.

    ``paired_end``: ``None`` or ``(mean_frag_len, frag_sd)``.  Returns ``{gene id: path of the
    .miso file written, or a string saying why the gene was skipped}``.
    """
    st = dict(DEFAULT_SETTINGS)
    st.update(settings or {})
    genes = load_gff_genes(gff_filename)
    sam = load_alignments(sam_filename)
    os.makedirs(output_dir, exist_ok=True)
    result, todo = OrderedDict(), []
    try:
        for gid in gene_ids:
            if gid not in genes:
                result[gid] = "skipped: not in the GFF"
                continue
            gene = make_gene_from_gff_records(gid, genes[gid])
            iso_lens = [sum(p.end - p.start + 1 for p in iso.parts) for iso in gene.isoforms]
            if all(l < read_len for l in iso_lens):                                   # run_miso.py:110-113
                result[gid] = "skipped: all isoforms shorter than the read"
                continue
            tx_start, tx_end = get_inclusive_txn_bounds(genes[gid])
            raw = fetch_reads_in_gene(sam, gene.chrom, tx_start, tx_end)
            reads, n_raw = sam_parse_reads(raw, paired_end=bool(paired_end), strand_rule=st["strand_rule"],
                                           target_strand=gene.strand, given_read_len=read_len)
            if st["filter_reads"] and n_raw < st["min_event_reads"]:                  # run_miso.py:134-141
                result[gid] = "skipped: only %d reads in gene (needed >= %d)" % (n_raw, st["min_event_reads"])
                continue
            out_file = os.path.join(output_dir, gene.chrom, gid) + ".miso"
            if len(reads[0]) == 0:                                                    # miso_sampler.py:229-231
                result[gid] = "skipped: no reads"
            elif os.path.isfile(os.path.normpath(out_file)):                          # :233-238
                result[gid] = "skipped: output exists"
            elif len(gene.isoforms) == 1:                                             # :272-277
                result[gid] = "skipped: one isoform"
            else:
                todo.append((gid, gene, reads, out_file))
    finally:
        if isinstance(sam, BamFile):
            sam.close()
    if not todo:
        return result

    # one plan for all genes of the call (the reference loops run_sampler per gene)
    gs, poss, cigs = [], [], []
    for gid, gene, reads, _ in todo:
        exons = tuple((p.start, p.end) for p in gene.parts)                       # py2c_gene.py:10-21
        isoforms = tuple(tuple(gene.parts.index(p) for p in iso.parts) for iso in gene.isoforms)
        gs.append(_batch.Gene(exons, isoforms))
        poss.append([int(p) + 1 for p in reads[0]])                               # miso_sampler.py:284
        cigs.append(list(reads[1]))
    # the reference truncates both numbers: mean_frag_len = int(paired_end[0]),
    # frag_variance = int(paired_end[1]) ** 2 (misopy/run_miso.py:81-83)
    pe = (float(int(paired_end[0])), float(int(paired_end[1]) ** 2), 4.0) if paired_end else (0.0, 0.0, 0.0)
    rb = _batch.ReadBatch(gs, poss, cigs, int(read_len), int(overhang_len), bool(paired_end), *pe)
    plan = _batch.Plan().append(rb)
    try:
        from .pysplicing_api import _seed
        params = _batch.make_params(st["num_iters"], st["burn_in"], st["lag"], st["num_chains"], device=device,
                                    seed=_seed(seed))
        out = plan.run(params)
        for i, (gid, gene, reads, out_file) in enumerate(todo):
            r = plan.gene_result(out, i)
            if r["status"] != 0:
                result[gid] = "failed: status %d" % r["status"]
                continue
            if np.all(r["assignment"] == -1):                                     # miso_sampler.py:352-354
                result[gid] = "skipped: no read is compatible with an isoform"
                continue
            templ, counts = plan.classes(i)
            acc, rej = int(r["rundata"][5]), int(r["rundata"][6])
            header = format_header([iso.desc for iso in gene.isoforms],
                                   [(p.label, p.end - p.start + 1) for p in gene.parts], st["num_iters"],
                                   st["burn_in"], st["lag"], float(acc) / (acc + rej) * 100, "drift",
                                   tuple(tuple(float(v) for v in row) for row in templ),
                                   tuple(float(v) for v in counts), r["assignment"], gene.chrom, gene.strand,
                                   [iso.genomic_start for iso in gene.isoforms],
                                   [iso.genomic_end for iso in gene.isoforms])
            os.makedirs(os.path.dirname(os.path.abspath(out_file)), exist_ok=True)
            write_miso(out_file, header, r["samples"].T, r["loglik"])
            result[gid] = out_file
            if verbose:
                print("%s: %d reads -> %s" % (gid, len(reads[0]), out_file))
    finally:
        plan.close()
    return result


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(
        description="MISO on the B200: GFF3 + SAM / BAM -> .miso files (miso --run), .miso_summary "
                    "(summarize_miso --summarize-samples), .miso_bf (compare_miso --compare-samples)")
    ap.add_argument("--compute-gene-psi", nargs=4, metavar=("GENE_IDS", "GFF", "ALIGNMENTS", "OUTPUT_DIR"),
                    help="comma-separated gene ids (or 'all'), GFF3 annotation, SAM text or BAM (+ .bai) alignments, "
                         "output directory")
    ap.add_argument("--summarize-samples", nargs=2, metavar=("SAMPLES_DIR", "OUTPUT_DIR"),
                    help="write OUTPUT_DIR/summary/<name of SAMPLES_DIR>.miso_summary (summarize_miso.py:24-47)")
    ap.add_argument("--compare-samples", nargs=3, metavar=("SAMPLES_DIR_1", "SAMPLES_DIR_2", "OUTPUT_DIR"),
                    help="Bayes factors of every event both samples have (compare_miso.py:40-95)")
    ap.add_argument("--read-len", type=int)
    ap.add_argument("--overhang-len", type=int, default=1)
    ap.add_argument("--paired-end", nargs=2, type=float, metavar=("MEAN", "SD"), default=None)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--settings-filename", default=None,
                    help="MISO settings file (misopy/settings/miso_settings.txt format); default: its shipped values")
    a = ap.parse_args(argv)
    if a.summarize_samples:
        from .postprocess import summarize_sampler_results
        samples_dir, out = a.summarize_samples
        name = os.path.basename(os.path.normpath(samples_dir))
        path = os.path.join(out, "summary", "%s.miso_summary" % name)
        print("%s\t%d events" % (path, summarize_sampler_results(samples_dir, path)))
    elif a.compare_samples:
        from .postprocess import output_samples_comparison
        path, n = output_samples_comparison(*a.compare_samples)
        print("%s\t%d events" % (path, n))
    elif a.compute_gene_psi:
        if a.read_len is None:
            ap.error("--compute-gene-psi needs --read-len")
        ids, gff, sam, out = a.compute_gene_psi
        gene_ids = list(load_gff_genes(gff)) if ids == "all" else ids.split(",")
        res = compute_gene_psi(gene_ids, gff, sam, out, a.read_len, a.overhang_len, a.paired_end,
                               settings=load_settings(a.settings_filename) if a.settings_filename else None,
                               seed=a.seed, verbose=True)
        for gid, r in res.items():
            print("%s\t%s" % (gid, r))
    else:
        ap.error("one of --compute-gene-psi, --summarize-samples, --compare-samples is required")


if __name__ == "__main__":
    main()
