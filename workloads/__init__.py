"""Synthetic workloads of BASELINE.json's configs (bench / test infrastructure).

ctypes binding of ``workloads/libmiso_synth.so`` (``include/miso_synth.h``,
``workloads/synth.cpp``).  Kept outside the product package on purpose: both
arms of ``bench.py`` draw the same genes from here, and the reference arm loads
nothing of ``libmiso_b200.so``.  A ``Workload`` exposes a ``misob200_reads_t``
view (``.struct``) that ``miso_b200.Plan.append`` takes as is, and per-gene
Python tuples (``.gene(g)``) in the form the oracle drivers take.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmiso_synth.so")


class Reads(C.Structure):
    """misob200_reads_t (include/miso_b200.h); same layout as miso_b200._lib.Reads."""
    _fields_ = [
        ("n_genes", C.c_int32),
        ("iso_off", C.c_void_p), ("exon_off", C.c_void_p),
        ("exon_start", C.c_void_p), ("exon_end", C.c_void_p),
        ("read_off", C.c_void_p), ("position", C.c_void_p),
        ("cigar_off", C.c_void_p), ("cigar", C.c_void_p),
        ("hyper", C.c_void_p), ("gene_id", C.c_void_p),
        ("read_len", C.c_int32), ("overhang", C.c_int32), ("paired", C.c_int32),
        ("frag_mean", C.c_double), ("frag_var", C.c_double), ("num_devs", C.c_double),
    ]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _load():
    if not os.path.isfile(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.misob200_workload_last_error.restype = C.c_char_p
    lib.misob200_workload_create.argtypes = [
        C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double,
        C.c_uint64, C.c_uint32, C.c_int, C.POINTER(vp)]
    lib.misob200_workload_create_ids.argtypes = [
        C.c_int, C.c_int32, vp, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double,
        C.c_uint64, C.c_uint32, C.c_int, C.POINTER(vp)]
    lib.misob200_workload_n_iso.argtypes = [vp, vp]
    lib.misob200_workload_view.argtypes = [vp, C.POINTER(Reads)]
    lib.misob200_workload_truth.argtypes = [vp, C.c_int32, vp]
    lib.misob200_workload_destroy.argtypes = [vp]
    return lib


lib = _load()

EXPORTS = ["misob200_workload_create", "misob200_workload_create_ids", "misob200_workload_n_iso",
           "misob200_workload_last_error", "misob200_workload_view", "misob200_workload_truth",
           "misob200_workload_destroy"]


def _check(rc):
    if rc:
        raise ValueError("workload error %d: %s" % (rc, (lib.misob200_workload_last_error() or b"").decode()))


class Workload:
    """kind 0: K = 2 skipped-exon single-end events (cfg-2); kind 1: K ~ U{2..8}
    paired-end events with a discretised normal insert model (cfg-3/4/5).  A gene is a
    function of (seed, gene id) only, so any shard of a workload can be rebuilt by itself:
    pass ``gene_ids`` for an explicit list, else ids are first_gene_id .. +n_genes-1.
    ``sample`` > 0 gives another sample of the same events (same structures, its own psi and reads)."""

    def __init__(self, kind, n_genes, reads_per_gene, read_len=36, frag_mean=250.0,
                 frag_var=900.0, num_devs=4.0, seed=1, first_gene_id=0, n_threads=0, gene_ids=None, sample=0):
        self.h = C.c_void_p()
        if sample and gene_ids is None:
            gene_ids = np.arange(first_gene_id, first_gene_id + n_genes)
        if gene_ids is not None:
            self._ids = np.ascontiguousarray(gene_ids, np.uint32)
            n_genes = len(self._ids)
            _check(lib.misob200_workload_create_ids(kind, n_genes, self._ids.ctypes.data, reads_per_gene, read_len,
                                                    frag_mean, frag_var, num_devs, seed, int(sample), n_threads,
                                                    C.byref(self.h)))
        else:
            _check(lib.misob200_workload_create(kind, n_genes, reads_per_gene, read_len,
                                                frag_mean, frag_var, num_devs, seed,
                                                first_gene_id, n_threads, C.byref(self.h)))
        self.struct = Reads()
        _check(lib.misob200_workload_view(self.h, C.byref(self.struct)))
        self.n_genes = n_genes

    def _arr(self, addr, n, dtype):
        dtype = np.dtype(dtype)
        if int(n) == 0:
            return np.zeros(0, dtype)
        buf = (C.c_char * (int(n) * dtype.itemsize)).from_address(addr)
        return np.frombuffer(buf, dtype=dtype, count=int(n))

    def n_iso(self):
        out = np.zeros(max(self.n_genes, 1), np.int32)
        _check(lib.misob200_workload_n_iso(self.h, out.ctypes.data))
        return out[:self.n_genes]

    def gene_ids(self):
        return self._arr(self.struct.gene_id, self.n_genes, np.uint32).copy()

    def gene(self, g):
        """(exons, isoforms, positions, cigars) of gene g, the form the oracle drivers take."""
        s = self.struct
        iso_off = self._arr(s.iso_off, s.n_genes + 1, np.int32)
        n_iso = int(iso_off[-1])
        exon_off = self._arr(s.exon_off, n_iso + 1, np.int32)
        n_ex = int(exon_off[-1])
        xs = self._arr(s.exon_start, n_ex, np.int32)
        xe = self._arr(s.exon_end, n_ex, np.int32)
        read_off = self._arr(s.read_off, s.n_genes + 1, np.int64)
        n_reads = int(read_off[-1])
        pos = self._arr(s.position, n_reads, np.int32)
        cig_off = self._arr(s.cigar_off, n_reads + 1, np.int64)
        exons, isoforms, seen = [], [], {}
        for k in range(iso_off[g], iso_off[g + 1]):
            iso = []
            for e in range(exon_off[k], exon_off[k + 1]):
                key = (int(xs[e]), int(xe[e]))
                if key not in seen:
                    seen[key] = len(exons)
                    exons.append(key)
                iso.append(seen[key])
            isoforms.append(tuple(iso))
        r0, r1 = int(read_off[g]), int(read_off[g + 1])
        if r1 > r0:
            b0, b1 = int(cig_off[r0]), int(cig_off[r1])
            raw = C.string_at(s.cigar + b0, b1 - b0)
            cig = raw[:-1].decode().split("\0")
        else:
            cig = []
        return tuple(exons), tuple(isoforms), pos[r0:r1].copy(), cig

    def truth(self, g, K):
        out = np.zeros(K)
        _check(lib.misob200_workload_truth(self.h, g, out.ctypes.data))
        return out

    def close(self):
        if self.h:
            lib.misob200_workload_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        self.close()
