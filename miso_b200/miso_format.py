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


def format_header(isoform_descs, exon_lens, iters, burn_in, lag, percent_accept, proposal_type,
                  read_classes, read_class_counts, assignments, chrom=None, strand=None,
                  mRNA_starts=(), mRNA_ends=()):
    """Header line of a .miso file (``miso_sampler.py:385-454``)."""
    if len(isoform_descs) and isinstance(isoform_descs[0], (list, tuple)):
        str_isoforms = "[" + ",".join("'" + "_".join(d) + "'" for d in isoform_descs) + "]"
    else:
        str_isoforms = "[" + ",".join("'" + d + "'" for d in isoform_descs) + "]"
    exon_lens_s = ",".join("('%s',%d)" % (label, length) for label, length in exon_lens)
    counts = []
    for cls, cnt in zip(read_classes, read_class_counts):
        counts.append("%s:%s" % (str(tuple(int(c) for c in cls)).replace(" ", ""), int(cnt)))
    assigned = ",".join("%d:%d" % c for c in count_isoform_assignments(assignments))
    return ("#isoforms=%s\texon_lens=%s\titers=%d\tburn_in=%d\tlag=%d\t"
            "percent_accept=%.2f\tproposal_type=%s\t"
            "counts=%s\tassigned_counts=%s\tchrom=%s\tstrand=%s\tmRNA_starts=%s\tmRNA_ends=%s\n"
            % (str_isoforms, exon_lens_s, iters, burn_in, lag, percent_accept, proposal_type,
               ",".join(counts), assigned, "NA" if chrom is None else chrom,
               "NA" if strand is None else strand,
               ",".join(str(s) for s in mRNA_starts), ",".join(str(e) for e in mRNA_ends)))


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
