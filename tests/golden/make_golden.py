"""Generate the committed golden vectors under tests/golden/ (run in the build
container only: needs /root/reference for the SAM fixture and oracle/_ref, the
UNMODIFIED reference C core, as the source of truth).

    python tests/golden/make_golden.py [1|2]      (stream version; default: both)

Every case stores its inputs (gene structure, read positions, CIGARs) and the
outputs of the reference driven by the miso-b200 stream (oracle/philox_ref.h; golden_v1.npz:
stream v1 = Philox4x32-10, golden_v2.npz: stream v2 = Philox4x32-7, the product's default)
with the framework's chain convention: chain c of gene g owns stream
(seed, g, c); the oracle is called once per chain with noChains=1 and the
columns are interleaved s*C + c (tests/helpers.py:oracle_gene).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import refdriver  # noqa: E402
from helpers import oracle_gene  # noqa: E402


class P:  # minimal params carrier
    def __init__(self, n_iters, burn_in, lag, n_chains, seed, start=0):
        self.n_iters, self.burn_in, self.lag, self.n_chains, self.seed, self.start = (
            n_iters, burn_in, lag, n_chains, seed, start)


def cfg1_reads():
    """BASELINE config 1: the SE event of misopy/gff-events/mm9/SE.mm9.gff:45081 and
    the reads of misopy/test-data/sam-data/c2c12.Atp2b1.sam that start in its span,
    treated as single-end (SAM POS is 1-based = what misopy passes after its +1,
    miso_sampler.py:284)."""
    exons = ((98481349, 98481531), (98481912, 98482065), (98485442, 98488777))
    isoforms = ((0, 1, 2), (0, 2))
    pos, cig = [], []
    sam = "/root/reference/misopy/test-data/sam-data/c2c12.Atp2b1.sam"
    for line in open(sam):
        if line.startswith("@"):
            continue
        f = line.split("\t")
        if f[2] != "10":
            continue
        p = int(f[3])
        if exons[0][0] <= p <= exons[-1][1]:
            pos.append(p)
            cig.append(f[5])
    return exons, isoforms, np.asarray(pos, np.int32), cig


def skip_gene(K):
    ex = tuple((1 + 400 * i, 200 + 400 * i) for i in range(K + 1))
    iso = (tuple(range(K + 1)),) + tuple(tuple(j for j in range(K + 1) if j != k) for k in range(1, K))
    return ex, iso


def main(version):
    ref = refdriver.RefOracle(stream=version)
    assert ref.stream == version
    cases = {}

    ex, iso, pos, cig = cfg1_reads()
    print("cfg-1: %d reads" % len(pos))
    for name, prm in (("cfg1_default", P(5000, 500, 10, 6, 1)), ("cfg1_long", P(7500, 2500, 10, 1, 2))):
        cases[name] = dict(kind="se", exons=ex, isoforms=iso, pos=pos, cig=cig, read_len=36, overhang=1,
                           params=prm, gene_id=0)

    rng = np.random.default_rng(7)
    # synthetic SE / PE genes from the reference's own simulators
    for K, R, tag in ((2, 300, "se_k2"), (3, 500, "se_k3"), (5, 400, "se_k5")):
        ex, iso = skip_gene(K)
        psi = rng.dirichlet(np.ones(K))
        pos, cig, _ = ref.simulate_se(ex, iso, psi, R, 36, seed=K)
        cases[tag] = dict(kind="se", exons=ex, isoforms=iso, pos=pos, cig=cig, read_len=36, overhang=1,
                          params=P(800, 100, 7, 2, 11), gene_id=K)
    for K, R, tag in ((2, 300, "pe_k2"), (4, 600, "pe_k4"), (8, 500, "pe_k8")):
        ex, iso = skip_gene(K)
        psi = rng.dirichlet(np.ones(K))
        pos, cig, _ = ref.simulate_pe(ex, iso, psi, R, 36, 250.0, 900.0, 4.0, seed=K)
        cases[tag] = dict(kind="pe", exons=ex, isoforms=iso, pos=pos, cig=cig, read_len=36, overhang=1,
                          pe=(250.0, 900.0, 4.0), params=P(800, 100, 7, 2, 13), gene_id=100 + K)
    # edge cases
    ex, iso = skip_gene(3)
    pos, cig, _ = ref.simulate_se(ex, iso, (0.3, 0.3, 0.4), 200, 36, seed=99)
    cig2 = list(cig)
    for i in range(0, 200, 3):
        cig2[i] = "30M"                       # shorter than read_len -> incompatible (solve.c:55)
    for i in range(1, 200, 11):
        cig2[i] = "2S30M1I4M"                 # clipping + insertion (solve.c:271-294)
    cases["se_ragged"] = dict(kind="se", exons=ex, isoforms=iso, pos=pos, cig=cig2, read_len=36, overhang=3,
                              params=P(600, 60, 5, 3, 17), gene_id=7)
    cases["se_all_incompatible"] = dict(kind="se", exons=ex, isoforms=iso, pos=pos[:50],
                                        cig=["20M"] * 50, read_len=36, overhang=1,
                                        params=P(300, 50, 5, 1, 19), gene_id=8)
    cases["se_uniform_start"] = dict(kind="se", exons=ex, isoforms=iso, pos=pos, cig=cig, read_len=36,
                                     overhang=1, params=P(500, 100, 4, 2, 23, start=1), gene_id=9)
    cases["se_lag_not_dividing"] = dict(kind="se", exons=ex, isoforms=iso, pos=pos, cig=cig, read_len=36,
                                        overhang=1, params=P(1000, 100, 7, 3, 29), gene_id=10)

    out = {}
    for name, c in cases.items():
        prm = c["params"]
        paired = c["kind"] == "pe"
        want = oracle_gene(ref, (c["exons"], c["isoforms"], c["pos"], c["cig"]), paired, prm, c["gene_id"],
                           pe=c.get("pe", (0, 0, 0)), read_len=c["read_len"], overhang=c["overhang"])
        S = (prm.n_iters - prm.burn_in) // prm.lag
        print(name, "K", len(c["isoforms"]), "reads", len(c["pos"]), "mean psi", want["samples"].mean(axis=1))
        pre = name + "/"
        out[pre + "exons"] = np.asarray(c["exons"], np.int32)
        out[pre + "iso_flat"] = refdriver.flatten_gene(c["exons"], c["isoforms"])[1]
        out[pre + "pos"] = np.asarray(c["pos"], np.int32)
        out[pre + "cig"] = np.asarray("\n".join(c["cig"]))
        out[pre + "meta"] = np.asarray([paired, c["read_len"], c["overhang"], prm.n_iters, prm.burn_in, prm.lag,
                                        prm.n_chains, prm.seed, prm.start, c["gene_id"]], np.int64)
        out[pre + "pe"] = np.asarray(c.get("pe", (0.0, 0.0, 0.0)), np.float64)
        out[pre + "samples"] = want["samples"][:, :prm.n_chains * S]
        out[pre + "loglik"] = want["loglik"][:prm.n_chains * S]
        out[pre + "assignment"] = want["assignment"]
        out[pre + "accrej"] = np.asarray([want["accepted"], want["rejected"]], np.int64)
        out[pre + "class_templates"] = want["class_templates"]
        out[pre + "class_counts"] = want["class_counts"]
    path = os.path.join(HERE, "golden_v%d.npz" % version)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    for v in ([int(sys.argv[1])] if len(sys.argv) > 1 else [1, 2]):
        main(v)
