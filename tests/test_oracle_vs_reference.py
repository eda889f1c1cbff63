"""Pins the restatement to the reference itself: oracle/miso_oracle.c against
oracle/_ref (the unmodified reference C core, built from /root/reference by
oracle/Makefile) on seeded random genes -- every output bit-identical.
Skipped where oracle/_ref was never built."""
import numpy as np
import pytest


def _gene(rng, K, kind):
    if kind == 0:
        ex = [(1 + 400 * i, 200 + 400 * i) for i in range(K + 1)]
        iso = [tuple(range(K + 1))] + [tuple(j for j in range(K + 1) if j != k) for k in range(1, K)]
        return tuple(ex), tuple(iso)
    n = int(rng.integers(4, 8))
    s, ex = 1, []
    for _ in range(n):
        ln = int(rng.integers(60, 300))
        ex.append((s, s + ln - 1))
        s += ln + int(rng.integers(50, 400))
    iso = []
    for _ in range(1000):
        if len(iso) == K:
            break
        c = tuple(sorted(rng.choice(n, int(rng.integers(2, n + 1)), replace=False).tolist()))
        if c not in iso:
            iso.append(c)
    assert len(iso) == K
    return tuple(ex), tuple(iso)


@pytest.mark.parametrize("t", range(12))
def test_port_equals_reference(ref, port, t):
    rng = np.random.default_rng(1000 + t)
    K = int(rng.integers(2, 9))
    ex, iso = _gene(rng, K, t % 2)
    psi = rng.dirichlet(np.ones(K))
    R, rl, oh, C = int(rng.integers(1, 300)), int(rng.integers(20, 50)), int(rng.integers(1, 5)), int(rng.integers(1, 4))
    if t % 3 == 0:
        pos, cig, _ = ref.simulate_se(ex, iso, psi, R, rl, seed=t)
        cig = list(cig)
        for i in range(0, R, 17):
            cig[i] = "%dM" % (rl - 3)
        for i in range(5, R, 23):
            cig[i] = "2S%dM1I2M" % (rl - 4)
        a, b = ref.match_se(ex, iso, pos, cig, rl, oh), port.match_se(ex, iso, pos, cig, rl, oh)
        kw = dict(overhang=oh, chains=C, seed=t, gene_id=t)
        ra, rb = ref.miso_se(ex, iso, pos, cig, rl, 300, 50, 5, **kw), port.miso_se(ex, iso, pos, cig, rl, 300, 50, 5, **kw)
    else:
        mean, sd = float(rng.integers(80, 300)), float(rng.integers(5, 40))
        pos, cig, _ = ref.simulate_pe(ex, iso, psi, R, rl, mean, sd * sd, 4.0, seed=t)
        a = ref.match_pe(ex, iso, pos, cig, rl, mean, sd * sd, 4.0, oh)
        b = port.match_pe(ex, iso, pos, cig, rl, mean, sd * sd, 4.0, oh)
        kw = dict(overhang=oh, chains=C, seed=t, gene_id=t)
        ra = ref.miso_pe(ex, iso, pos, cig, rl, mean, sd * sd, 4.0, 300, 50, 5, **kw)
        rb = port.miso_pe(ex, iso, pos, cig, rl, mean, sd * sd, 4.0, 300, 50, 5, **kw)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    for k in ra:
        np.testing.assert_array_equal(ra[k], rb[k], err_msg=k)   # NaN-free by construction here


@pytest.mark.parametrize("t", range(6))
def test_start_random_port_equals_reference(ref, port, t):
    """MISO_START_RANDOM (Dirichlet start via gamma draws, src/miso.c:388-404, :309-326).
    The reference reads `sigma` uninitialised on this branch; the harness pre-loads that
    stack slot with SIGMA (oracle/ref_harness.c refh_paint_stack) -- agreement bit for bit
    with the port, which sets sigma = SIGMA, shows the pre-load took."""
    rng = np.random.default_rng(7000 + t)
    K = int(rng.integers(2, 9))
    ex, iso = _gene(rng, K, 0)
    psi = rng.dirichlet(np.ones(K))
    R, rl, C = int(rng.integers(20, 300)), 36, int(rng.integers(1, 4))
    kw = dict(overhang=1, chains=C, seed=t, gene_id=t, start=2)
    if t % 2 == 0:
        pos, cig, _ = ref.simulate_se(ex, iso, psi, R, rl, seed=t)
        ra, rb = ref.miso_se(ex, iso, pos, cig, rl, 300, 50, 5, **kw), port.miso_se(ex, iso, pos, cig, rl, 300, 50, 5, **kw)
    else:
        pos, cig, _ = ref.simulate_pe(ex, iso, psi, R, rl, 250.0, 900.0, 4.0, seed=t)
        ra = ref.miso_pe(ex, iso, pos, cig, rl, 250.0, 900.0, 4.0, 300, 50, 5, **kw)
        rb = port.miso_pe(ex, iso, pos, cig, rl, 250.0, 900.0, 4.0, 300, 50, 5, **kw)
    for k in ra:
        np.testing.assert_array_equal(ra[k], rb[k], err_msg=k)
    # and the start differs from AUTO's (the branch really ran)
    kw["start"] = 0
    rc = (port.miso_se(ex, iso, pos, cig, rl, 300, 50, 5, **kw) if t % 2 == 0 else
          port.miso_pe(ex, iso, pos, cig, rl, 250.0, 900.0, 4.0, 300, 50, 5, **kw))
    assert not np.array_equal(rb["samples"], rc["samples"])


@pytest.mark.parametrize("t", range(6))
def test_port_equals_reference_with_hyperparameters_and_uniform_start(ref, port, t):
    """Non-flat Dirichlet hyperparameters (ldirichlet, src/miso.c:165-182, enters the MH ratio) and
    MISO_START_UNIFORM (src/miso.c:372-386): port and unmodified reference bit for bit."""
    rng = np.random.default_rng(9000 + t)
    K = int(rng.integers(2, 9))
    ex, iso = _gene(rng, K, t % 2)
    psi = rng.dirichlet(np.ones(K))
    hyper = tuple(float(h) for h in rng.uniform(0.3, 3.0, K))
    R, rl, C = int(rng.integers(30, 250)), 36, int(rng.integers(1, 3))
    kw = dict(overhang=int(rng.integers(1, 4)), chains=C, seed=50 + t, gene_id=t, start=t % 2, hyper=hyper)
    if t % 3 == 0:
        pos, cig, _ = ref.simulate_se(ex, iso, psi, R, rl, seed=t)
        ra, rb = ref.miso_se(ex, iso, pos, cig, rl, 250, 50, 4, **kw), port.miso_se(ex, iso, pos, cig, rl, 250, 50, 4, **kw)
    else:
        pos, cig, _ = ref.simulate_pe(ex, iso, psi, R, rl, 200.0, 625.0, 4.0, seed=t)
        ra = ref.miso_pe(ex, iso, pos, cig, rl, 200.0, 625.0, 4.0, 250, 50, 4, **kw)
        rb = port.miso_pe(ex, iso, pos, cig, rl, 200.0, 625.0, 4.0, 250, 50, 4, **kw)
    for k in ra:
        np.testing.assert_array_equal(ra[k], rb[k], err_msg=k)
