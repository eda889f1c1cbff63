"""Loader of tests/golden/golden_v<stream>.npz (written by tests/golden/make_golden.py): v1 = the
reference driven by Philox4x32-10, v2 = by Philox4x32-7 (the product's default stream)."""
import os

import numpy as np

DEFAULT_STREAM = 1 if os.environ.get("MISOB200_STREAM") == "1" else 2
STREAMS = (1, 2)


def golden_path(version):
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v%d.npz" % version)


class Case:
    pass


def load_cases(version=DEFAULT_STREAM):
    z = np.load(golden_path(version))
    names = sorted({k.split("/")[0] for k in z.files})
    cases = []
    for name in names:
        c = Case()
        c.name = name
        c.stream = version
        g = lambda f: z[name + "/" + f]   # noqa: E731
        c.exons = tuple((int(a), int(b)) for a, b in g("exons"))
        iso, cur = [], []
        for v in g("iso_flat"):
            if v < 0:
                iso.append(tuple(cur)); cur = []
            else:
                cur.append(int(v))
        c.isoforms = tuple(iso)
        c.pos = g("pos")
        c.cig = str(g("cig")).split("\n") if len(c.pos) else []
        m = g("meta")
        (c.paired, c.read_len, c.overhang, c.n_iters, c.burn_in, c.lag, c.n_chains, c.seed, c.start,
         c.gene_id) = (int(x) for x in m)
        c.pe = tuple(float(x) for x in g("pe"))
        c.samples, c.loglik, c.assignment = g("samples"), g("loglik"), g("assignment")
        c.accepted, c.rejected = (int(x) for x in g("accrej"))
        c.class_templates, c.class_counts = g("class_templates"), g("class_counts")
        cases.append(c)
    return cases


class Params:
    def __init__(self, c):
        self.n_iters, self.burn_in, self.lag, self.n_chains, self.seed, self.start = (
            c.n_iters, c.burn_in, c.lag, c.n_chains, c.seed, c.start)
