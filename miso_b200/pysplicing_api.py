"""The ``pysplicing`` entry points of the reference, on the B200.

Same names, positional order, defaults, argument checks and 6-tuple return as
the CPython-2 extension (``/root/reference/pysplicing/src/pysplicing.c:41-131``
``MISO``, ``:152-244`` ``MISOPaired``, ``:246-278`` ``createGene``; constants
``pysplicing/pysplicing/__init__.py:2-13``), so ``misopy.miso_sampler`` can call
them unchanged (``misopy/miso_sampler.py:292-322``).  One gene per call is a
batch of one on the device; use ``miso_b200.Plan`` for throughput.

Not on the sampler path and not provided: readGFF, writeGFF, simulateReads,
simulatePairedReads, assignmentMatrix, solveIsoGene, geneComplexity, ...
(they raise NotImplementedError).
"""
import os
import random

import numpy as np

from . import _lib
from ._lib import InternalError
from .batch import (Gene, Plan, ReadBatch, make_params, MISO_START_AUTO, MISO_START_UNIFORM,  # noqa: F401
                    MISO_START_RANDOM, MISO_START_GIVEN, MISO_START_LINEAR, MISO_STOP_FIXEDNO,
                    MISO_STOP_CONVERGENT_MEAN, MISO_ALGO_REASSIGN, MISO_ALGO_MARGINAL,
                    MISO_ALGO_CLASSES)


def createGene(exons, isoforms, id="insilicogene", seqid="seq1", source="protein_coding", strand=2):
    return Gene(exons, isoforms, id, seqid, source, strand)


def _need_tuple(x):
    if not isinstance(x, tuple):
        raise TypeError("Need a tuple")            # pyconvert.c:7-10
    return x


def _seed(seed):
    if seed is not None:
        return int(seed)
    env = os.environ.get("MISO_B200_SEED")
    if env is not None:
        return int(env)
    # the reference draws from Python's `random` module (pyrandom.c:106-141):
    # derive the stream key from it, so random.seed(n) makes a run repeatable
    return random.getrandbits(64)


def _run(gene, geneNo, positions, cigars, read_len, iters, burn, lag, hyperp, overhang, chains, start,
         stop, algo, paired, pe, seed, device):
    if not isinstance(gene, Gene):
        raise TypeError("gene must come from createGene")
    if geneNo != 0:
        raise InternalError("Invalid gene id")     # gff.c:588-590
    _need_tuple(positions)
    _need_tuple(cigars)
    K = gene.n_iso
    if hyperp is None:
        hyperp = (1.0,) * K                         # pysplicing.c:88-93
    else:
        _need_tuple(hyperp)
        if len(hyperp) != K:
            raise InternalError("Invalid hyperparameter vector length")   # miso.c:698-701
    if chains < 1:
        raise InternalError("Number of chains must be at least one.")     # miso.c:703-706
    oh = 1 if overhang == 0 else overhang
    if oh < 1 or oh >= read_len // 2:
        raise InternalError("Overhang length invalid. Must be between 0 and readLength/2")
    params = make_params(iters, burn, lag, chains, start, stop, algo, device=device, seed=_seed(seed))
    batch = ReadBatch([gene], [positions], [cigars], read_len, overhang, paired,
                      *(pe if paired else (0.0, 0.0, 0.0)), hyper=[hyperp])
    plan = Plan().append(batch)
    try:
        K_, R, _, ncls, status = (int(v) for v in plan.info()[0])
        if status != 0:
            _lib.check(status)
        out = plan.run(params)
        res = plan.gene_result(out, 0)
        templ, counts = plan.classes(0)
    finally:
        plan.close()
    samples = tuple(tuple(float(v) for v in row) for row in res["samples"])
    loglik = tuple(float(v) for v in res["loglik"])
    class_templates = tuple(tuple(float(v) for v in row) for row in templ)
    class_counts = tuple(float(v) for v in counts)
    assignment = tuple(int(v) for v in res["assignment"])
    rd = res["rundata"]
    rundata = (int(rd[0]), int(rd[1]), int(rd[3]), int(rd[4]), int(rd[5]), int(rd[6]))   # pyconvert.c:174-183
    return samples, loglik, class_templates, class_counts, assignment, rundata


def MISO(gene, geneNo, positions, cigars, readLength, noIterations=5000, noBurnIn=500, noLag=10,
         hyperp=None, overhang=1, no_chains=6, start=MISO_START_AUTO, stop=MISO_STOP_FIXEDNO,
         algo=MISO_ALGO_REASSIGN, seed=None, device=0):
    return _run(gene, geneNo, positions, cigars, int(readLength), int(noIterations), int(noBurnIn),
                int(noLag), hyperp, int(overhang), int(no_chains), int(start), int(stop), int(algo),
                False, None, seed, device)


def MISOPaired(gene, geneNo, positions, cigars, readLength, normalMean, normalVar, numDevs,
               noIterations=5000, noBurnIn=500, noLag=10, hyperp=None, overhang=1, no_chains=6,
               start=MISO_START_AUTO, stop=MISO_STOP_FIXEDNO, seed=None, device=0):
    return _run(gene, geneNo, positions, cigars, int(readLength), int(noIterations), int(noBurnIn),
                int(noLag), hyperp, int(overhang), int(no_chains), int(start), int(stop),
                MISO_ALGO_REASSIGN, True, (float(normalMean), float(normalVar), float(numDevs)),
                seed, device)


def noIso(gene):
    return gene.n_iso


def isoLength(gene):
    return tuple(sum(gene.exons[i][1] - gene.exons[i][0] + 1 for i in iso) for iso in gene.isoforms)


def _off_path(name):
    def f(*a, **k):
        raise NotImplementedError(
            "pysplicing.%s is not on the sampler path this framework implements" % name)
    f.__name__ = name
    return f


for _n in ("readGFF", "writeGFF", "simulateReads", "simulatePairedReads", "assignmentMatrix",
           "solveIsoGene", "geneComplexity", "noGenes", "i_fromGFF", "toGFF"):
    globals()[_n] = _off_path(_n)
