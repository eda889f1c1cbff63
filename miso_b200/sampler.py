"""``MISOSampler`` in Python 3: the front end of the hot path.

Mirror of ``/root/reference/misopy/miso_sampler.py:169-466`` (``run_sampler`` +
``output_miso_results``) and ``misopy/py2c_gene.py:4-23``: same arguments, same
skip rules, same ``.miso`` file.  The reference module is Python 2 and cannot
be imported by this interpreter; this mirror keeps its call into
``pysplicing.MISO`` / ``MISOPaired`` byte for byte so either module can drive
the device.
"""
import os
from collections import namedtuple

import numpy as np

from . import pysplicing_api as pysplicing
from .miso_format import format_header, write_miso

Part = namedtuple("Part", "label start end")
Isoform = namedtuple("Isoform", "desc parts genomic_start genomic_end")


class GeneModel:
    """Just the fields of misopy.Gene the sampler reads (``miso_sampler.py:384-441``)."""

    def __init__(self, label, parts, isoforms, chrom=None, strand=None):
        self.label, self.parts, self.chrom, self.strand = label, list(parts), chrom, strand
        self.isoforms = []
        for desc in isoforms:
            # the first part carrying each label, in the isoform's own order
            # (misopy/Gene.py:305-321 create_isoforms / get_part_by_label)
            ps = [next(p for p in self.parts if p.label == lab) for lab in desc]
            self.isoforms.append(Isoform(list(desc), ps, min(p.start for p in ps), max(p.end for p in ps)))


def py2c_gene(py_gene):
    """``misopy/py2c_gene.py:4-23``."""
    exons = tuple((p.start, p.end) for p in py_gene.parts)
    isoforms = tuple(tuple(py_gene.parts.index(p) for p in iso.parts) for iso in py_gene.isoforms)
    return pysplicing.createGene(exons, isoforms)


def get_single_end_sampler_params(num_isoforms, read_len, overhang_len=1):
    return {"read_len": read_len, "overhang_len": overhang_len, "uniform_proposal": False,
            "sigma_proposal": np.eye(num_isoforms - 1) * 0.05}


def get_paired_end_sampler_params(num_isoforms, mean_frag_len, frag_variance, read_len, overhang_len=1):
    p = get_single_end_sampler_params(num_isoforms, read_len, overhang_len)
    p.update(mean_frag_len=mean_frag_len, frag_variance=frag_variance)
    return p


class MISOSampler:
    def __init__(self, params, paired_end=False, log_dir=None, seed=None, device=0):
        self.params, self.paired_end, self.seed, self.device = params, paired_end, seed, device
        if paired_end:
            if "mean_frag_len" not in params or "frag_variance" not in params:
                raise Exception("Must set mean_frag_len and frag_variance when "
                                "running in sampler on paired-end data.")
            self.mean_frag_len, self.frag_variance = params["mean_frag_len"], params["frag_variance"]

    def run_sampler(self, num_iters, reads, gene, hyperparameters, params, output_file, num_chains=6,
                    burn_in=1000, lag=2, prior_params=None, algorithm=pysplicing.MISO_ALGO_CLASSES,
                    start_cond=pysplicing.MISO_START_AUTO, stop_cond=pysplicing.MISO_STOP_FIXEDNO,
                    verbose=True):
        num_isoforms = len(gene.isoforms)
        if prior_params is None:
            prior_params = (1.0,) * num_isoforms
        read_positions, read_cigars = reads[0], reads[1]
        if len(read_positions) == 0:                       # miso_sampler.py:229-231
            return None
        output_file = output_file + ".miso"
        if os.path.isfile(os.path.normpath(output_file)):  # :233-238
            return None
        if num_isoforms == 1:                              # :272-277
            return None
        proposal_type = "unif" if params["uniform_proposal"] else "drift"
        c_gene = py2c_gene(gene)
        read_positions = tuple(int(r) + 1 for r in read_positions)   # :284
        read_cigars = tuple(read_cigars)
        if self.paired_end:
            res = pysplicing.MISOPaired(c_gene, 0, read_positions, read_cigars, int(self.params["read_len"]),
                                        float(self.mean_frag_len), float(self.frag_variance), 4.0,
                                        int(num_iters), int(burn_in), int(lag), tuple(prior_params),
                                        int(self.params["overhang_len"]), int(num_chains), start_cond,
                                        stop_cond, seed=self.seed, device=self.device)
        else:
            res = pysplicing.MISO(c_gene, 0, read_positions, read_cigars, int(self.params["read_len"]),
                                  int(num_iters), int(burn_in), int(lag), tuple(prior_params),
                                  int(self.params["overhang_len"]), int(num_chains), start_cond, stop_cond,
                                  pysplicing.MISO_ALGO_REASSIGN, seed=self.seed, device=self.device)
        psi_vectors = np.transpose(np.array(res[0]))
        kept_log_scores = np.array(res[1])
        assignments = np.array(res[4])
        if np.all(assignments == -1):                      # :352-354
            return None
        acc, rej = res[5][4], res[5][5]
        percent_acceptance = float(acc) / (acc + rej) * 100
        header = format_header([iso.desc for iso in gene.isoforms],
                               [(p.label, p.end - p.start + 1) for p in gene.parts], num_iters, burn_in, lag,
                               percent_acceptance, proposal_type, res[2], res[3], assignments, gene.chrom,
                               gene.strand, [iso.genomic_start for iso in gene.isoforms],
                               [iso.genomic_end for iso in gene.isoforms])
        os.makedirs(os.path.dirname(os.path.abspath(output_file)), exist_ok=True)
        write_miso(output_file, header, psi_vectors, kept_log_scores)
        return output_file
