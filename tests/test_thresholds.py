"""The integer-threshold rule of miso_b200/csrc/class_pass.cuh against the reference's
fp64 choice rule (src/miso.c:59-83, src/miso_paired.c:11-22,45-78), in numpy.

class_pass.cuh claims: for a weight class, `reference test k is true  <=>  w > floor(tau_k)`
for EVERY 32-bit word w, provided tau_k = C_k * (2^32 / C_{K-1}) - 0.5 is further than
2^-15 from an integer (otherwise the kernel declines the pass and runs the literal rule).
Here the exact integer threshold of the reference's arithmetic is found by bisection over w
and compared with floor(tau) -- on random psi and, mainly, on psi nudged ulp by ulp so that
tau lands next to an integer, where the claim could fail.  numpy's multiply/add/divide are
the same correctly rounded IEEE operations the kernel uses (it is built with -fmad=false).
CPU only."""
import numpy as np
import pytest

MARGIN = 2.0 ** -15          # kThrMargin in class_pass.cuh
TWO32 = 4294967296.0


def fragment_table(mean=250.0, sd=30.0, devs=4.0, read_len=36):
    fs, fe = max(int(mean - sd * devs), read_len), int(mean + sd * devs)
    x = (np.arange(fs, fe + 1) - mean) / sd
    p = 0.398942280401432677939946059934 * np.exp(-0.5 * x * x) / sd
    return np.concatenate([[0.0], p / p.sum()])


PTAB = fragment_table()


def cumsums(psi, wts):
    """S = S + psi_k * w_k, k ascending (CUMSUM): [N, K] running sums."""
    C = np.empty_like(psi)
    S = np.zeros(psi.shape[0])
    for k in range(psi.shape[1]):
        S = S + psi[:, k] * wts[:, k]
        C[:, k] = S
    return C


def reference_test(w, C, k, two):
    """Is test k of the reference true for the word w?  rnd = u * C_{K-1};
    two compatible isoforms: !(rnd < C_k), otherwise rnd > C_k (miso.c:71,78)."""
    u = (w.astype(np.float64) + 0.5) * 2.0 ** -32
    rnd = u * C[:, -1]
    return np.where(two, rnd >= C[:, k], rnd > C[:, k])


def exact_threshold(C, k, two):
    """Smallest word for which the reference's test k is true (2^32 if none), by bisection."""
    n = C.shape[0]
    lo = np.zeros(n, np.int64)                # invariant: test false below lo ... true from hi on
    hi = np.full(n, 1 << 32, np.int64)
    for _ in range(34):
        mid = (lo + hi) >> 1
        t = reference_test(np.minimum(mid, (1 << 32) - 1), C, k, two) & (mid < (1 << 32))
        hi = np.where(t, mid, hi)
        lo = np.where(t, lo, np.minimum(mid + 1, hi))
    return hi


def kernel_threshold(Cmine, k):
    """thr_update: tau, floor(tau) saturated to u32, and the trust test."""
    inv = TWO32 / Cmine[:, -1]
    tau = Cmine[:, k] * inv - 0.5
    tk = np.clip(np.trunc(tau), 0.0, TWO32 - 1.0)
    fr = tau - tk
    good = (fr > MARGIN) & (fr < 1.0 - MARGIN)
    return tau, tk.astype(np.int64), good


def draw_classes(rng, n, K, uniform):
    """n random (psi, compatibility pattern, codes): at least two compatible isoforms."""
    conc = rng.choice([0.05, 0.3, 1.0, 8.0], size=(n, 1))
    psi = rng.gamma(np.broadcast_to(conc, (n, K))) + 1e-300
    psi /= psi.sum(axis=1, keepdims=True)
    psi[:, K - 1] = 1.0 - psi[:, : K - 1].sum(axis=1)            # as logit_inv leaves it (miso.c:467)
    psi = np.abs(psi) + 1e-30
    mask = rng.random((n, K)) < 0.6
    for i in np.where(mask.sum(axis=1) < 2)[0]:
        mask[i, rng.choice(K, 2, replace=False)] = True
    if uniform:
        codes = np.where(mask, rng.integers(1, len(PTAB), size=(n, 1)), 0)
    else:
        codes = np.where(mask, rng.integers(1, len(PTAB), size=(n, K)), 0)
    return psi, mask, codes


@pytest.mark.parametrize("uniform", [True, False])
@pytest.mark.parametrize("K", [2, 3, 5, 8])
def test_integer_threshold_equals_reference_rule(K, uniform):
    rng = np.random.default_rng(100 * K + uniform)
    n = 6000
    psi, mask, codes = draw_classes(rng, n, K, uniform)
    two = mask.sum(axis=1) == 2
    # the kernel's weights: 1.0 for a uniform-code class, ptab[code] otherwise
    w_mine = np.where(mask, 1.0, 0.0) if uniform else PTAB[codes]
    w_ref = PTAB[codes]
    worst = 0.0
    checked = declined = 0
    for k in range(K - 1):
        first = mask.argmax(axis=1)
        live = k >= first                      # earlier tests are constant true (GeneDesc.g_always)
        # --- nudge psi_first so that tau_k sits next to an integer --------------------
        p = psi.copy()
        for _ in range(3):                     # Newton on rho_k = P_k / P
            C = cumsums(p, w_mine)
            Pk, P = C[:, k], C[:, -1]
            tau = Pk / P * TWO32 - 0.5
            target = np.round(tau) + rng.choice([-1.0, 1.0], n) * MARGIN * rng.uniform(0.0, 3.0, n)
            wf = w_mine[np.arange(n), first]
            with np.errstate(divide="ignore", invalid="ignore"):
                d = (target - tau) / TWO32 * P * P / np.maximum(P - Pk, 1e-300) / wf
            ok = live & (P > Pk) & np.isfinite(d) & (np.abs(d) < 0.5 * p[np.arange(n), first])
            p[np.arange(n), first] += np.where(ok, d, 0.0)
        # a cloud of ulp-sized perturbations around it
        reps = 12
        P2 = np.repeat(p, reps, axis=0)
        jit = rng.integers(-400, 401, size=n * reps)
        idx = np.repeat(first, reps)
        rows = np.arange(n * reps)
        P2[rows, idx] = P2[rows, idx] * (1.0 + jit * 2.0 ** -52)
        Wm, Wr = np.repeat(w_mine, reps, axis=0), np.repeat(w_ref, reps, axis=0)
        tw = np.repeat(two, reps)
        lv = np.repeat(live, reps)
        Cm, Cr = cumsums(P2, Wm), cumsums(P2, Wr)
        tau, tk, good = kernel_threshold(Cm, k)
        exact = exact_threshold(Cr, k, tw)     # smallest w with the test true
        # kernel rule: test true <=> w > tk, i.e. smallest true word = tk + 1 (2^32 when tk saturates)
        mine = np.where(tk >= (1 << 32) - 1, 1 << 32, tk + 1)
        trusted = good & lv
        bad = trusted & (mine != exact)
        assert not bad.any(), (K, k, uniform, P2[bad][:3], tau[bad][:3], exact[bad][:3])
        checked += int(trusted.sum())
        declined += int((lv & ~good).sum())
        # how far from an integer does a disagreement ever occur?  (the margin must cover it)
        # (tau < 0 -- a compatible isoform with psi below 2^-33 -- is declined outright: fr < 0)
        dis = lv & (mine != exact) & (tau >= 0.0)
        if dis.any():
            dist = np.abs(tau[dis] - np.round(tau[dis]))
            worst = max(worst, float(dist.max()))
    assert checked > 10000 and declined > 1000       # the cloud really straddles the margin
    assert worst < MARGIN / 1.5, worst


def test_random_psi_is_rarely_declined():
    """Away from the nudged cases the trust test almost never fires (about 6e-5 per threshold)."""
    rng = np.random.default_rng(7)
    psi, mask, codes = draw_classes(rng, 400000, 5, True)
    psi = np.maximum(psi, 1e-8)                # (a share below 2^-33 is declined by design, see above)
    C = cumsums(psi, np.where(mask, 1.0, 0.0))
    first = mask.argmax(axis=1)
    frac = []
    for k in range(4):
        _, _, good = kernel_threshold(C, k)
        live = (k >= first) & (C[:, k] < C[:, -1])
        frac.append((live & ~good).sum() / max(live.sum(), 1))
    assert max(frac) < 4e-4, frac
