"""Shared helpers for the parity tests."""
import numpy as np


def oracle_gene(oracle, wl_gene, paired, params, gene_id, pe=(250.0, 900.0, 4.0), read_len=36,
                overhang=1, hyper=None):
    """Run the oracle the way the framework defines multi-chain runs: every
    chain owns the stream (seed, gene_id, chain), so the oracle is called with
    noChains=1 per chain and the columns are interleaved s*C + c afterwards
    ("same seed" definition)."""
    ex, isos, pos, cig = wl_gene
    C = params.n_chains
    outs = []
    for c in range(C):
        kw = dict(iters=params.n_iters, burn=params.burn_in, lag=params.lag, hyper=hyper,
                  overhang=overhang, chains=1, start=params.start, seed=params.seed,
                  gene_id=gene_id, chain_id=c)
        if paired:
            outs.append(oracle.miso_pe(ex, isos, pos, cig, read_len, pe[0], pe[1], pe[2], **kw))
        else:
            outs.append(oracle.miso_se(ex, isos, pos, cig, read_len, **kw))
    S = (params.n_iters - params.burn_in) // params.lag
    K = len(isos)
    smp = np.zeros((K, C * S))
    ll = np.zeros(C * S)
    for c, o in enumerate(outs):
        smp[:, c::C] = o["samples"][:, :S]
        ll[c::C] = o["loglik"][:S]
    return dict(samples=smp, loglik=ll, assignment=outs[0]["assignment"],
                accepted=sum(int(o["rundata"][5]) for o in outs),
                rejected=sum(int(o["rundata"][6]) for o in outs),
                class_templates=outs[0]["class_templates"], class_counts=outs[0]["class_counts"])


def assert_gene_parity(got, want, tag=""):
    """Decisions identical (assignments, accept counts), values to fp64 noise."""
    assert got["status"] == 0, tag
    np.testing.assert_array_equal(got["assignment"], want["assignment"], err_msg=tag)
    assert int(got["rundata"][5]) == want["accepted"], tag
    assert int(got["rundata"][6]) == want["rejected"], tag
    np.testing.assert_allclose(got["samples"], want["samples"], rtol=1e-9, atol=1e-12, err_msg=tag)
    np.testing.assert_allclose(got["loglik"], want["loglik"], rtol=1e-9, atol=1e-9, err_msg=tag)


def simulate_pairs(exons, isoforms, psi, n_pairs, read_len, mean, sd, num_devs, rng):
    """Small paired-end read simulator for the tests (numpy): isoform ~ psi weighted by
    length, fragment length ~ discretised N(mean, sd^2) within num_devs, uniform start;
    mates given as 1-based genomic positions + CIGARs (M/N blocks), mate order 1, 2."""
    iso_exons = [[exons[e] for e in iso] for iso in isoforms]
    iso_len = np.array([sum(b - a + 1 for a, b in ex) for ex in iso_exons])
    w = np.asarray(psi, float) * iso_len
    w /= w.sum()

    def to_genome(ex, ipos, length):
        """isoform coordinate (1-based) + length -> genomic start, CIGAR"""
        out, left, start = [], length, None
        off = ipos - 1
        for i, (a, b) in enumerate(ex):
            el = b - a + 1
            if off >= el:
                off -= el
                continue
            if start is None:
                start = a + off
            take = min(left, el - off)
            out.append("%dM" % take)
            left -= take
            off = 0
            if left == 0:
                break
            out.append("%dN" % (ex[i + 1][0] - b - 1))
        return start, "".join(out)

    lo, hi = max(int(mean - num_devs * sd), read_len), int(mean + num_devs * sd)
    pos, cig = [], []
    while len(pos) < 2 * n_pairs:
        k = rng.choice(len(isoforms), p=w)
        f = int(round(rng.normal(mean, sd)))
        if f < lo or f > hi or f > iso_len[k]:
            continue
        s = int(rng.integers(1, iso_len[k] - f + 2))
        for ip in (s, s + f - read_len):
            g, c = to_genome(iso_exons[k], ip, read_len)
            pos.append(g)
            cig.append(c)
    return np.asarray(pos, np.int32), cig
