"""The `.miso` posterior file and its summaries, in Python 3.

Restates the reference's text formats so that outputs of this framework drop
into the downstream tools unchanged:
  * writer  : ``misopy/miso_sampler.py:376-466`` (``output_miso_results``)
  * parser  : ``misopy/samples_utils.py:130-156`` (``load_samples``)
  * summary : ``misopy/credible_intervals.py:4-55`` and
              ``misopy/samples_utils.py:263-329`` (``.miso_summary`` fields)
All paths relative to /root/reference.
"""
import numpy as np


def count_isoform_assignments(assignments):
    """``misopy/reads_utils.py:38-47``: (isoform, count) for 0..max assigned."""
    a = np.asarray(assignments)
    if a.size == 0:
        return []
    n = int(a.max())
    return [(k, int((a == k).sum())) for k in range(n + 1)]


def header_static_parts(isoform_descs, exon_lens, chrom=None, strand=None, mRNA_starts=(), mRNA_ends=()):
    """The two pieces of a .miso header line that depend on the annotation only
    (``miso_sampler.py:385-403,430-454``): ("#isoforms=..\texon_lens=..\t", "\tchrom=..\n").  The
    batched writer (``misob200_plan_write_miso``) fills in the run-dependent fields between them."""
    if len(isoform_descs) and isinstance(isoform_descs[0], (list, tuple)):
        str_isoforms = "[" + ",".join("'" + "_".join(d) + "'" for d in isoform_descs) + "]"
    else:
        str_isoforms = "[" + ",".join("'" + d + "'" for d in isoform_descs) + "]"
    exon_lens_s = ",".join("('%s',%d)" % (label, length) for label, length in exon_lens)
    prefix = "#isoforms=%s\texon_lens=%s\t" % (str_isoforms, exon_lens_s)
    suffix = "\tchrom=%s\tstrand=%s\tmRNA_starts=%s\tmRNA_ends=%s\n" % (
        "NA" if chrom is None else chrom, "NA" if strand is None else strand,
        ",".join(str(s) for s in mRNA_starts), ",".join(str(e) for e in mRNA_ends))
    return prefix, suffix


def format_header(isoform_descs, exon_lens, iters, burn_in, lag, percent_accept, proposal_type,
                  read_classes, read_class_counts, assignments, chrom=None, strand=None,
                  mRNA_starts=(), mRNA_ends=()):
    """Header line of a .miso file (``miso_sampler.py:385-454``)."""
    prefix, suffix = header_static_parts(isoform_descs, exon_lens, chrom, strand, mRNA_starts, mRNA_ends)
    counts = []
    for cls, cnt in zip(read_classes, read_class_counts):
        counts.append("%s:%s" % (str(tuple(int(c) for c in cls)).replace(" ", ""), int(cnt)))
    assigned = ",".join("%d:%d" % c for c in count_isoform_assignments(assignments))
    return (prefix + "iters=%d\tburn_in=%d\tlag=%d\tpercent_accept=%.2f\tproposal_type=%s\t"
            "counts=%s\tassigned_counts=%s" % (iters, burn_in, lag, percent_accept, proposal_type,
                                              ",".join(counts), assigned) + suffix)


def write_miso(path, header, psi_vectors, log_scores):
    """Body of a .miso file: ``%.4f`` psi, ``%.2f`` score (``miso_sampler.py:456-465``)."""
    with open(path, "w") as out:
        out.write(header)
        out.write("sampled_psi\tlog_score\n")
        for psi, sc in zip(psi_vectors, log_scores):
            out.write("%s\t%.2f\n" % (",".join("%.4f" % p for p in psi), sc))


def parse_header(line):
    """``#k=v\\tk=v...`` -> dict of raw strings."""
    fields = {}
    for item in line.lstrip("#").rstrip("\n").split("\t"):
        if "=" in item:
            k, v = item.split("=", 1)
            fields[k] = v
    return fields


def load_samples(path_or_lines):
    """(samples [n x K], header dict, log_scores, sampled MAP, its score, counts string)."""
    if isinstance(path_or_lines, str):
        with open(path_or_lines) as f:
            lines = f.read().splitlines()
    else:
        lines = [ln.rstrip("\n") for ln in path_or_lines]
    header = parse_header(lines[0])
    cols = lines[1].split("\t")
    i_psi, i_sc = cols.index("sampled_psi"), cols.index("log_score")
    raw, scores = [], []
    for ln in lines[2:]:
        if not ln:
            continue
        f = ln.split("\t")
        raw.append(f[i_psi])
        scores.append(float(f[i_sc]))
    samples = np.asarray([[float(v) for v in r.split(",")] for r in raw])
    # the reference takes max() over the psi *strings* (samples_utils.py:141-142)
    m = max(raw)
    k = raw.index(m)
    return samples, header, np.asarray(scores), [float(v) for v in m.split(",")], scores[k], header.get("counts")


def credible_interval_indices(n, confidence_level=0.95):
    """Order-statistic indices of ``compute_credible_intervals``
    (``credible_intervals.py:31-55``).  ``round`` there is numpy's
    (``from numpy import *``): half-to-even on the fp64 product."""
    alpha = 1 - confidence_level
    lo = int(np.round((alpha / 2) * n)) - 1
    hi = int(np.round((1 - alpha / 2) * n)) - 1
    return lo, hi


def compute_credible_intervals(samples, confidence_level=0.95):
    s = np.asarray(samples, dtype=float)
    if s.ndim == 2:
        s = s[:, 0]
    lo, hi = credible_interval_indices(len(s), confidence_level)
    s = np.sort(s)
    return [s[lo], s[hi]]


def compute_multi_iso_credible_intervals(samples, confidence_level=0.95):
    s = np.asarray(samples, dtype=float)
    return [compute_credible_intervals(s[:, k], confidence_level) for k in range(s.shape[1])]


def format_credible_intervals(event_name, samples, confidence_level=0.95):
    """Fields of one .miso_summary line (``credible_intervals.py:4-28``)."""
    s = np.asarray(samples, dtype=float)
    if s.shape[1] > 2:
        ci = compute_multi_iso_credible_intervals(s, confidence_level)
        return [event_name, ",".join("%.2f" % v for v in s.mean(axis=0)),
                ",".join("%.2f" % c[0] for c in ci), ",".join("%.2f" % c[1] for c in ci)]
    ci = compute_credible_intervals(s, confidence_level)
    return [event_name, "%.2f" % s.mean(axis=0)[0], "%.2f" % ci[0], "%.2f" % ci[1]]


def isoforms_field(header):
    """``get_isoforms_from_header`` (``samples_utils.py:177-189``): the isoforms= value without
    its brackets."""
    v = header.get("isoforms", "")
    return v[1:-1] if v.startswith("[") and v.endswith("]") else v


# ---- two-sample comparison (compare_miso) -----------------------------------------

BF_HEADER = ["event_name", "sample1_posterior_mean", "sample1_ci_low", "sample1_ci_high",
             "sample2_posterior_mean", "sample2_ci_low", "sample2_ci_high", "diff", "bayes_factor",
             "isoforms", "sample1_counts", "sample1_assigned_counts", "sample2_counts",
             "sample2_assigned_counts", "chrom", "strand", "mRNA_starts", "mRNA_ends"]


def bayes_factor(samples1, samples2, smoothing_param=0.3, max_bf=1e12):
    """Per-isoform Bayes factor of delta psi != 0, as ``compute_delta_densities`` +
    ``compute_bayes_factor`` (``misopy/hypothesis_test.py:89-179,348-380``): Gaussian KDE of
    the paired differences with covariance = var(ddof=1) * smoothing_param**2, evaluated
    at 0; a posterior peaked on the null (mean |delta| <= 0.009 or constant) gives 0."""
    s1, s2 = np.asarray(samples1, float), np.asarray(samples2, float)
    out = []
    for k in range(s1.shape[1]):
        d = s1[:, k] - s2[:, k]
        if np.mean(np.abs(d)) <= 0.009 or np.all(d - d[0] == 0):
            out.append(0.0)
            continue
        cov = np.var(d, ddof=1) * smoothing_param ** 2
        kde0 = np.sum(np.exp(-0.5 * d * d / cov)) / (len(d) * np.sqrt(2 * np.pi * cov))
        out.append(max_bf if kde0 == 0 else min(1.0 / kde0, max_bf))
    return out


def format_bf_line(event_name, samples1, samples2, bf, header1, header2):
    """One line of a ``.miso_bf`` file (``hypothesis_test.py:255-338``); header1/2 are the
    parsed .miso headers of the two samples."""
    from decimal import Decimal
    s1, s2 = np.asarray(samples1, float), np.asarray(samples2, float)
    c1, c2 = format_credible_intervals(event_name, s1), format_credible_intervals(event_name, s2)
    m1, m2 = s1.mean(axis=0), s2.mean(axis=0)
    if s1.shape[1] == 2:
        q1 = Decimal(str(m1[0])).quantize(Decimal("0.01"))
        q2 = Decimal(str(m2[0])).quantize(Decimal("0.01"))
        mean1, mean2, diff, bfs = str(q1), str(q2), "%.2f" % (q1 - q2), "%.2f" % bf[0]
    else:
        mean1, mean2 = c1[1], c2[1]
        diff = ",".join("%.2f" % v for v in (m1 - m2))
        bfs = ",".join("%.2f" % max(v, 0) for v in bf)
    return "\t".join([event_name, mean1, c1[2], c1[3], mean2, c2[2], c2[3], diff, bfs,
                      isoforms_field(header1), header1.get("counts", ""),
                      header1.get("assigned_counts", ""), header2.get("counts", ""),
                      header2.get("assigned_counts", ""), header1.get("chrom", "NA"),
                      header1.get("strand", "NA"), header1.get("mRNA_starts", ""),
                      header1.get("mRNA_ends", "")]) + "\n"
