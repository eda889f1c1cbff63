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
