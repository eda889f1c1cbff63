"""oracle/refdriver.py -- TEST INFRASTRUCTURE ONLY (never imported by miso_b200).

ctypes driver for oracle/_ref/libsplicing_ref.so: the UNMODIFIED reference C
core (/root/reference/pysplicing/src/*.c on the sampler path) + ref_harness.c.
Used by tests/ to pin the plain-C restatement (oracle/miso_oracle.c), by
tests/golden/make_golden.py to generate the committed fixtures, and by
bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libsplicing_ref.so")
PORT_SO = os.path.join(HERE, "libmiso_oracle.so")

_i32p = C.POINTER(C.c_int)
_f64p = C.POINTER(C.c_double)


def available():
    return os.path.isfile(REF_SO)


def _ip(a):
    return a.ctypes.data_as(_i32p) if a is not None else None


def _dp(a):
    return a.ctypes.data_as(_f64p) if a is not None else None


def flatten_gene(exons, isoforms):
    """(exons, isoforms) as pysplicing.createGene takes them
    (pysplicing/src/pyconvert.c:59-93) -> flat int arrays."""
    ex = np.asarray([c for e in exons for c in e], dtype=np.int32)
    iso = []
    for i in isoforms:
        iso.extend(int(x) for x in i)
        iso.append(-1)
    return ex, np.asarray(iso, dtype=np.int32)


def _cigar_array(cigars):
    arr = (C.c_char_p * len(cigars))()
    arr[:] = [c.encode() if isinstance(c, str) else c for c in cigars]
    return arr


class _Prefixed:
    """lib.refh_xxx / lib.mo_xxx behind one attribute namespace."""

    def __init__(self, lib, prefix):
        self._lib, self._prefix = lib, prefix

    def __getattr__(self, name):
        assert name.startswith("refh_")
        return getattr(self._lib, self._prefix + name[5:])


class RefOracle:
    """kind == "reference": the unmodified reference C (oracle/_ref);
    PortOracle below shares every method except the simulators."""
    kind = "reference"

    def __init__(self, path=REF_SO, prefix="refh_", stream=None):
        self.lib = _Prefixed(C.CDLL(path), prefix)
        self.lib.refh_init()
        # the Philox stream the oracle is driven by (rng_mode 0): version 2 = 7 rounds (the
        # product's default), 1 = 10 rounds; MISOB200_STREAM follows the product's switch
        if stream is None:
            stream = 1 if os.environ.get("MISOB200_STREAM") == "1" else 2
        self.set_stream(stream)

    def set_stream(self, version):
        self.stream = int(self.lib.refh_set_stream(int(version)))
        return self.stream

    # -- setup stage ---------------------------------------------------
    def gene_info(self, exons, isoforms):
        ex, iso = flatten_gene(exons, isoforms)
        n = C.c_int()
        il = np.zeros(len(isoforms), np.int32)
        ne = np.zeros(len(isoforms), np.int32)
        r = self.lib.refh_gene_info(len(exons), _ip(ex), len(iso), _ip(iso),
                                    C.byref(n), _ip(il), _ip(ne))
        assert r == 0
        return n.value, il, ne

    def match_se(self, exons, isoforms, pos, cigars, read_len, overhang=1):
        ex, iso = flatten_gene(exons, isoforms)
        K, R = len(isoforms), len(pos)
        pos = np.ascontiguousarray(pos, np.int32)
        match = np.zeros((R, K), np.float64)      # column-major K x R
        order = np.zeros(R, np.int32)
        ct = np.zeros((max(R, 1), K), np.float64)
        cc = np.zeros(max(R, 1), np.float64)
        ncls = C.c_int()
        r = self.lib.refh_match_se(len(exons), _ip(ex), len(iso), _ip(iso),
                                   R, _ip(pos), _cigar_array(cigars),
                                   read_len, overhang, _dp(match), _ip(order),
                                   C.byref(ncls), _dp(ct), _dp(cc))
        if r != 0:
            raise RuntimeError("reference returned error %d" % r)
        n = ncls.value
        return dict(match=match.T.copy(), order=order,
                    class_templates=ct[:n].T.copy(), class_counts=cc[:n].copy())

    def fragment_table(self, mean, var, num_devs, read_len):
        cap = 1 << 16
        prob = np.zeros(cap)
        st, il = C.c_int(), C.c_int()
        r = self.lib.refh_fragment_table(C.c_double(mean), C.c_double(var),
                                         C.c_double(num_devs), read_len, cap,
                                         _dp(prob), C.byref(st), C.byref(il))
        assert r == 0
        return prob[:il.value].copy(), st.value

    def match_pe(self, exons, isoforms, pos, cigars, read_len, mean, var,
                 num_devs, overhang=1):
        ex, iso = flatten_gene(exons, isoforms)
        K, n = len(isoforms), len(pos)
        R = n // 2
        pos = np.ascontiguousarray(pos, np.int32)
        match = np.zeros((n, K), np.float64)      # harness copies K x R only
        fl = np.zeros((max(R, 1), K), np.int32)
        order = np.zeros(max(R, 1), np.int32)
        ct = np.zeros((max(R, 1), K), np.float64)
        cc = np.zeros(max(R, 1), np.float64)
        ncls = C.c_int()
        r = self.lib.refh_match_pe(len(exons), _ip(ex), len(iso), _ip(iso),
                                   n, _ip(pos), _cigar_array(cigars),
                                   read_len, overhang, C.c_double(mean),
                                   C.c_double(var), C.c_double(num_devs),
                                   _dp(match), _ip(fl), _ip(order),
                                   C.byref(ncls), _dp(ct), _dp(cc))
        if r != 0:
            raise RuntimeError("reference returned error %d" % r)
        m = ncls.value
        return dict(match=match[:R].T.copy(), fraglen=fl[:R].T.copy(),
                    order=order[:R], bin_class_templates=ct[:m].T.copy(),
                    bin_class_counts=cc[:m].copy())

    # -- sampler -------------------------------------------------------
    def _miso(self, paired, exons, isoforms, pos, cigars, read_len, iters,
              burn, lag, hyper, overhang, chains, start, stop, rng_mode, seed,
              gene_id, chain_id, pe=None):
        ex, iso = flatten_gene(exons, isoforms)
        K, n = len(isoforms), len(pos)
        R = n // 2 if paired else n
        pos = np.ascontiguousarray(pos, np.int32)
        if hyper is None:
            hyper = np.ones(K)
        hyper = np.ascontiguousarray(hyper, np.float64)
        ns = chains * (iters - burn) // lag
        samples = np.zeros((max(ns, 1), K), np.float64)
        ll = np.zeros(max(ns, 1), np.float64)
        ct = np.zeros((max(R, 1), K), np.float64)
        cc = np.zeros(max(R, 1), np.float64)
        ass = np.zeros(max(R, 1), np.int32)
        rd = np.zeros(9, np.int32)
        ncls = C.c_int()
        common = (len(exons), _ip(ex), len(iso), _ip(iso), n, _ip(pos),
                  _cigar_array(cigars), read_len, overhang)
        tail = (chains, iters, burn, lag, _dp(hyper), start, stop, rng_mode,
                C.c_uint64(seed), C.c_uint32(gene_id), C.c_uint32(chain_id),
                _dp(samples), _dp(ll), C.byref(ncls), _dp(ct), _dp(cc),
                _ip(ass), _ip(rd))
        if paired:
            mean, var, nd = pe
            r = self.lib.refh_miso_pe(*common, C.c_double(mean),
                                      C.c_double(var), C.c_double(nd), *tail)
        else:
            r = self.lib.refh_miso_se(*common, *tail)
        if r != 0:
            raise RuntimeError("reference returned error %d" % r)
        nu, nn = C.c_uint64(), C.c_uint64()
        self.lib.refh_rng_counts(C.byref(nu), C.byref(nn))
        m = ncls.value
        return dict(samples=samples[:ns].T.copy(), loglik=ll[:ns].copy(),
                    class_templates=ct[:m].T.copy(),
                    class_counts=cc[:m].copy(), assignment=ass[:R].copy(),
                    rundata=rd.copy(), n_unif=nu.value, n_norm=nn.value)

    def miso_se(self, exons, isoforms, pos, cigars, read_len, iters=5000,
                burn=500, lag=10, hyper=None, overhang=1, chains=1, start=0,
                stop=0, rng_mode=0, seed=0, gene_id=0, chain_id=0):
        return self._miso(False, exons, isoforms, pos, cigars, read_len, iters,
                          burn, lag, hyper, overhang, chains, start, stop,
                          rng_mode, seed, gene_id, chain_id)

    def miso_pe(self, exons, isoforms, pos, cigars, read_len, mean, var,
                num_devs, iters=5000, burn=500, lag=10, hyper=None,
                overhang=1, chains=1, start=0, stop=0, rng_mode=0, seed=0,
                gene_id=0, chain_id=0):
        return self._miso(True, exons, isoforms, pos, cigars, read_len, iters,
                          burn, lag, hyper, overhang, chains, start, stop,
                          rng_mode, seed, gene_id, chain_id,
                          pe=(mean, var, num_devs))

    # -- simulators ------------------------------------------------------
    def _unpack(self, buf, off, n):
        raw = buf.raw
        return [raw[off[i]:off[i + 1] - 1].decode() for i in range(n)]

    def simulate_se(self, exons, isoforms, expression, noreads, read_len,
                    seed=1, rng_mode=1):
        ex, iso = flatten_gene(exons, isoforms)
        expr = np.ascontiguousarray(expression, np.float64)
        isoout = np.zeros(noreads, np.int32)
        pos = np.zeros(noreads, np.int32)
        cap = 64 * noreads + 64
        buf = C.create_string_buffer(cap)
        off = np.zeros(noreads + 1, np.int32)
        r = self.lib.refh_simulate_se(len(exons), _ip(ex), len(iso), _ip(iso),
                                      _dp(expr), noreads, read_len, rng_mode,
                                      C.c_uint64(seed), _ip(isoout), _ip(pos),
                                      buf, cap, _ip(off))
        if r != 0:
            raise RuntimeError("reference returned error %d" % r)
        return pos, self._unpack(buf, off, noreads), isoout

    def simulate_pe(self, exons, isoforms, expression, nopairs, read_len,
                    mean, var, num_devs, seed=1, rng_mode=1):
        ex, iso = flatten_gene(exons, isoforms)
        expr = np.ascontiguousarray(expression, np.float64)
        n = 2 * nopairs
        isoout = np.zeros(n, np.int32)
        pos = np.zeros(n, np.int32)
        cap = 64 * n + 64
        buf = C.create_string_buffer(cap)
        off = np.zeros(n + 1, np.int32)
        r = self.lib.refh_simulate_pe(len(exons), _ip(ex), len(iso), _ip(iso),
                                      _dp(expr), nopairs, read_len,
                                      C.c_double(mean), C.c_double(var),
                                      C.c_double(num_devs), rng_mode,
                                      C.c_uint64(seed), _ip(isoout), _ip(pos),
                                      buf, cap, _ip(off))
        if r != 0:
            raise RuntimeError("reference returned error %d" % r)
        return pos, self._unpack(buf, off, n), isoout


class PortOracle(RefOracle):
    """kind == "port": oracle/miso_oracle.c, the plain-C restatement."""
    kind = "port"

    def __init__(self, path=PORT_SO, stream=None):
        RefOracle.__init__(self, path, "mo_", stream)

    def simulate_se(self, *a, **k):
        raise NotImplementedError("simulators exist only in oracle/_ref")

    simulate_pe = simulate_se


def port_available():
    return os.path.isfile(PORT_SO)
