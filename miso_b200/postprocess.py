"""Directory-level post-processing of ``.miso`` files: ``summarize_miso`` and ``compare_miso``.

Python-3 restatements of ``misopy/samples_utils.py:263-329`` (``summarize_sampler_results`` ->
``.miso_summary``) and ``misopy/hypothesis_test.py:182-345`` (``output_samples_comparison`` ->
``<s1>_vs_<s2>/bayes-factors/<s1>_vs_<s2>.miso_bf``) for plain (uncompressed) sample directories
``<dir>/<chrom>/<event>.miso``.  The per-line formatting lives in ``miso_format.py``; the device
versions of the same numbers are ``Plan.summarize`` / ``Plan.compare``.
"""
import os

import numpy as np

from . import miso_format as mf

SUMMARY_HEADER = ["event_name", "miso_posterior_mean", "ci_low", "ci_high", "isoforms", "counts",
                  "assigned_counts", "chrom", "strand", "mRNA_starts", "mRNA_ends"]


def miso_files(samples_dir):
    """event name -> path, for every ``*.miso`` below ``samples_dir`` (sorted by event name)."""
    found = {}
    for root, _, files in os.walk(samples_dir):
        for f in files:
            if f.endswith(".miso"):
                found[f[:-len(".miso")]] = os.path.join(root, f)
    return dict(sorted(found.items()))


def summarize_sampler_results(samples_dir, summary_filename):
    """One ``.miso_summary`` line per event; returns the number of events summarised."""
    n = 0
    os.makedirs(os.path.dirname(os.path.abspath(summary_filename)), exist_ok=True)
    with open(summary_filename, "w") as out:
        out.write("\t".join(SUMMARY_HEADER) + "\n")
        for event, path in miso_files(samples_dir).items():
            try:
                samples, header, _, _, _, counts = mf.load_samples(path)
            except (ValueError, IndexError):
                continue                                   # unparsable file: skipped, as the reference does
            if samples.ndim < 2:
                continue
            fields = mf.format_credible_intervals(event, samples)
            fields += [mf.isoforms_field(header), header.get("counts", ""), header.get("assigned_counts", ""),
                       header.get("chrom", "NA"), header.get("strand", "NA"), header.get("mRNA_starts", "NA"),
                       header.get("mRNA_ends", "NA")]
            out.write("\t".join(fields) + "\n")
            n += 1
    return n


def output_samples_comparison(sample1_dir, sample2_dir, output_dir, sample_labels=None):
    """``compare_miso --compare-samples``: Bayes factor and delta psi of every event both samples have.
    Returns (path of the ``.miso_bf`` file, number of events compared)."""
    if sample_labels is None:
        l1 = os.path.basename(os.path.normpath(sample1_dir))
        l2 = os.path.basename(os.path.normpath(sample2_dir))
    else:
        l1, l2 = sample_labels
    bf_dir = os.path.join(output_dir, "%s_vs_%s" % (l1, l2), "bayes-factors")
    os.makedirs(bf_dir, exist_ok=True)
    out_name = os.path.join(bf_dir, "%s_vs_%s.miso_bf" % (l1, l2))
    f1, f2 = miso_files(sample1_dir), miso_files(sample2_dir)
    n = 0
    with open(out_name, "w") as out:
        out.write("\t".join(mf.BF_HEADER) + "\n")
        for event, p1 in f1.items():
            if event not in f2:
                continue
            s1, h1 = mf.load_samples(p1)[:2]
            s2, h2 = mf.load_samples(f2[event])[:2]
            if s1.shape != s2.shape:
                continue
            bf = mf.bayes_factor(s1, s2)
            out.write(mf.format_bf_line(event, s1, s2, bf, h1, h2))
            n += 1
    return out_name, n
