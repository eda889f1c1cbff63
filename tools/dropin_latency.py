"""Latency of the drop-in entry point: N consecutive pysplicing.MISOPaired / MISO calls, one gene each
(the way misopy/run_miso.py drives the sampler, miso_sampler.py:292-322), default run parameters
(5000 iterations, burn-in 500, lag 10, 6 chains).  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import pysplicing
from workloads import Workload


def measure(n_calls=100, reads=2000, kind=1, seed=5):
    w = Workload(kind, n_calls, reads, 36, 250.0, 900.0, 4.0, seed=seed)
    genes = []
    for g in range(n_calls):
        ex, isos, pos, cig = w.gene(g)
        genes.append((pysplicing.createGene(ex, isos), tuple(int(p) for p in pos), tuple(cig)))
    ms = []
    for g, (gene, pos, cig) in enumerate(genes):
        t0 = time.perf_counter()
        if kind == 1:
            r = pysplicing.MISOPaired(gene, 0, pos, cig, 36, 250.0, 900.0, 4.0, seed=g)
        else:
            r = pysplicing.MISO(gene, 0, pos, cig, 36, seed=g)
        ms.append((time.perf_counter() - t0) * 1e3)
        assert len(r) == 6 and len(r[0][0]) == 6 * 450
    ms = np.array(ms)
    return {"calls": n_calls, "reads_per_gene": reads, "paired": bool(kind), "first_call_ms": float(ms[0]),
            "mean_ms_after_first": float(ms[1:].mean()), "median_ms": float(np.median(ms)), "max_ms_after_first": float(ms[1:].max()),
            "note": "5000 iterations x 6 chains per call; the reference C takes ~0.3 s per chain on one host core"}


if __name__ == "__main__":
    print(json.dumps(measure(int(sys.argv[1]) if len(sys.argv) > 1 else 100)))
