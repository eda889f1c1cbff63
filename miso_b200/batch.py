"""Batched front end: many genes per device launch.

The reference feeds its C core one gene per call
(``/root/reference/misopy/run_miso.py:198-202`` ->
``misopy/miso_sampler.py:292-322``); a GPU needs thousands of gene-chains in
flight, so the native unit here is a *plan* holding a batch of genes.
``pysplicing.MISO`` / ``MISOPaired`` (``miso_b200/pysplicing_api.py``) are
one-gene batches over the same code.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Params, Reads, check, lib, ptr

MISO_START_AUTO, MISO_START_UNIFORM, MISO_START_RANDOM, MISO_START_GIVEN, MISO_START_LINEAR = range(5)
MISO_STOP_FIXEDNO, MISO_STOP_CONVERGENT_MEAN = 0, 1
MISO_ALGO_REASSIGN, MISO_ALGO_MARGINAL, MISO_ALGO_CLASSES = 0, 1, 2


class Gene:
    """What ``createGene`` returns: the exon list of every isoform
    (``pysplicing/src/pysplicing.c:246-278``, ``src/simulator.c:9-66``)."""

    def __init__(self, exons, isoforms, id="insilicogene", seqid="seq1",
                 source="protein_coding", strand=2):
        self.exons = tuple((int(s), int(e)) for s, e in exons)
        self.isoforms = tuple(tuple(int(i) for i in iso) for iso in isoforms)
        for iso in self.isoforms:
            for i in iso:
                if i < 0 or i >= len(self.exons):
                    raise _lib.InternalError("createGene: exon index out of range")
        self.id, self.seqid, self.source, self.strand = id, seqid, source, strand

    @property
    def n_iso(self):
        return len(self.isoforms)


def make_params(n_iters=5000, burn_in=500, lag=10, n_chains=6, start=0, stop=0,
                algo=0, device=0, seed=0):
    if start not in (MISO_START_AUTO, MISO_START_UNIFORM, MISO_START_RANDOM):
        # GIVEN: pysplicing.MISO passes start_psi=0 (pysplicing.c:99), unusable there too;
        # LINEAR: NNLS deconvolution (solve.c:308-536), off the sampler path
        raise NotImplementedError(
            "start=%r: MISO_START_AUTO / UNIFORM / RANDOM run on the device "
            "(misopy always passes AUTO, miso_sampler.py:210)" % (start,))
    if stop != MISO_STOP_FIXEDNO:
        raise NotImplementedError("stop=%r: only MISO_STOP_FIXEDNO" % (stop,))
    if algo != MISO_ALGO_REASSIGN:
        raise NotImplementedError(
            "algo=%r: only MISO_ALGO_REASSIGN (misopy forces it, "
            "miso_sampler.py:322)" % (algo,))
    return Params(int(n_iters), int(burn_in), int(lag), int(n_chains), int(start),
                  int(stop), int(algo), int(device), int(seed) & (2 ** 64 - 1))


class ReadBatch:
    """Flat arrays behind a ``misob200_reads_t``; keeps them alive."""

    def __init__(self, genes, positions, cigars, read_len, overhang=1, paired=False,
                 frag_mean=0.0, frag_var=0.0, num_devs=0.0, hyper=None, gene_ids=None):
        iso_off, exon_off, xs, xe = [0], [0], [], []
        for g in genes:
            for iso in g.isoforms:
                for i in iso:
                    xs.append(g.exons[i][0])
                    xe.append(g.exons[i][1])
                exon_off.append(len(xs))
            iso_off.append(iso_off[-1] + g.n_iso)
        read_off, pos, cig_off, blob = [0], [], [0], bytearray()
        for p, c in zip(positions, cigars):
            if len(p) != len(c):
                raise _lib.InternalError("positions and cigars differ in length")
            pos.extend(int(x) for x in p)
            for s in c:
                b = s if isinstance(s, bytes) else str(s).encode()
                blob += b + b"\0"
                cig_off.append(len(blob))
            read_off.append(len(pos))
        self.a = dict(
            iso_off=np.asarray(iso_off, np.int32), exon_off=np.asarray(exon_off, np.int32),
            exon_start=np.asarray(xs, np.int32), exon_end=np.asarray(xe, np.int32),
            read_off=np.asarray(read_off, np.int64), position=np.asarray(pos, np.int32),
            cigar_off=np.asarray(cig_off, np.int64),
            cigar=np.frombuffer(bytes(blob) + b"\0", np.uint8).copy())
        if hyper is not None:
            flat = [float(h) for hs in hyper for h in hs]
            if len(flat) != iso_off[-1]:
                raise _lib.InternalError("Invalid hyperparameter vector length")
            self.a["hyper"] = np.asarray(flat, np.float64)
        if gene_ids is not None:
            self.a["gene_id"] = np.asarray(gene_ids, np.uint32)
        r = Reads()
        r.n_genes = len(genes)
        for k, v in self.a.items():
            setattr(r, k, ptr(v))
        r.read_len, r.overhang, r.paired = int(read_len), int(overhang), int(bool(paired))
        r.frag_mean, r.frag_var, r.num_devs = float(frag_mean), float(frag_var), float(num_devs)
        self.struct = r


class Plan:
    def __init__(self, keep_match=False, tile_format=-1):
        """tile_format: -1 class tiles where a gene allows it (default), 0 dense tiles only."""
        self.h = C.c_void_p()
        check(lib.misob200_plan_create(C.byref(self.h)))
        if keep_match:
            check(lib.misob200_plan_keep_match(self.h, 1))
        if tile_format != -1:
            check(lib.misob200_plan_tile_format(self.h, tile_format))
        self._info = None

    def append(self, reads, n_threads=0, match_device=None):
        """Setup stage for a batch (a-12 ... a-18).  ``match_device``: GPU ordinal that computes
        the read <-> isoform compatibility (csrc/match.cu) instead of the host threads."""
        struct = reads.struct if hasattr(reads, "struct") else reads      # any misob200_reads_t mirror
        if match_device is None:
            check(lib.misob200_plan_append(self.h, C.addressof(struct), n_threads))
        else:
            check(lib.misob200_plan_append_device(self.h, C.addressof(struct), n_threads, int(match_device)))
        self._info = None
        return self

    def append_begin(self, reads, match_device=0):
        """First half of a device append: enqueue the batch's GPU work (copies, match_kernel,
        order_kernel) and return at once.  ``reads`` must stay alive until ``append_finish``."""
        struct = reads.struct if hasattr(reads, "struct") else reads
        pending = C.c_void_p()
        check(lib.misob200_plan_append_device_begin(self.h, C.addressof(struct), int(match_device), C.byref(pending)))
        self._pending = (pending, reads)
        return self

    def append_finish(self, n_threads=0):
        """Second half: wait for the GPU work, then classes and tiles on the host threads."""
        pending, _ = self._pending
        self._pending = None
        check(lib.misob200_plan_append_device_finish(self.h, pending, n_threads))
        self._info = None
        return self

    @staticmethod
    def last_match_stats():
        """(kernel ms, H2D ms, D2H ms, bytes in, bytes out) of the last device matching."""
        k, h, d = C.c_double(), C.c_double(), C.c_double()
        bi, bo = C.c_int64(), C.c_int64()
        check(lib.misob200_last_match_stats(C.addressof(k), C.addressof(h), C.addressof(d), C.addressof(bi),
                                            C.addressof(bo)))
        return k.value, h.value, d.value, bi.value, bo.value

    def close(self):
        if self.h:
            if getattr(self, "_pending", None):      # a device append was begun and abandoned: release its stage
                lib.misob200_plan_append_device_finish(self.h, self._pending[0], 0)
                self._pending = None
            lib.misob200_plan_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        self.close()

    # ---- introspection ---------------------------------------------------
    def size(self):
        g, r, t = C.c_int32(), C.c_int64(), C.c_int64()
        check(lib.misob200_plan_size(self.h, C.addressof(g), C.addressof(r), C.addressof(t)))
        return g.value, r.value, t.value

    def tile_info(self):
        """int32 array [n_genes, 3]: tile format (0 dense, 1 class), weight classes, tile bytes."""
        G = self.size()[0]
        out = np.zeros((G, 3), np.int32)
        v = [C.c_int32() for _ in range(3)]
        for g in range(G):
            check(lib.misob200_plan_gene_tile(self.h, g, *[C.addressof(x) for x in v]))
            out[g] = [x.value for x in v]
        return out

    def info(self):
        """int32 array [n_genes, 5]: K, R, R2, n_classes, status."""
        if self._info is None:
            G = self.size()[0]
            out = np.zeros((max(G, 1), 5), np.int32)
            check(lib.misob200_plan_info_all(self.h, ptr(out)))
            self._info = out[:G]
        return self._info

    def offsets(self, params):
        """int64 arrays (sample_off, loglik_off, assign_off), one entry per gene (elements)."""
        G = self.size()[0]
        a, b, c = (np.zeros(max(G, 1), np.int64) for _ in range(3))
        check(lib.misob200_plan_offsets_all(self.h, C.byref(params), ptr(a), ptr(b), ptr(c)))
        return a[:G], b[:G], c[:G]

    def classes(self, g):
        K, _, _, ncls, _ = self.info()[g]
        t = np.zeros((max(ncls, 1), K))
        c = np.zeros(max(ncls, 1))
        check(lib.misob200_plan_gene_classes(self.h, g, ptr(t), ptr(c)))
        return t[:ncls], c[:ncls]

    def match(self, g):
        K, R, _, _, _ = self.info()[g]
        codes = np.zeros((max(R, 1), K), np.int32)
        order = np.zeros(max(R, 1), np.int32)
        check(lib.misob200_plan_gene_match(self.h, g, ptr(codes), ptr(order)))
        return codes[:R].T.copy(), order[:R]

    def fragment_table(self):
        n, s = C.c_int32(), C.c_int32()
        check(lib.misob200_plan_fragment_table(self.h, 0, None, C.addressof(s), C.addressof(n)))
        p = np.zeros(max(n.value, 1))
        check(lib.misob200_plan_fragment_table(self.h, n.value, ptr(p), None, None))
        return p[:n.value], s.value

    # ---- execution ---------------------------------------------------------
    def output_sizes(self, params):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib.misob200_plan_output_sizes(self.h, C.byref(params), C.addressof(a),
                                             C.addressof(b), C.addressof(c)))
        return a.value, b.value, c.value

    def alloc_outputs(self, params, pinned=True):
        ns, nl, na = self.output_sizes(params)
        G = self.size()[0]
        mk = _lib.pinned_empty if pinned else (lambda n, d: np.empty(int(n), d))
        return dict(samples=mk(ns, np.float64), loglik=mk(nl, np.float64),
                    assignment=mk(na, np.int32), rundata=np.zeros((G, 9), np.int32),
                    status=np.zeros(G, np.int32))

    def run(self, params, out=None):
        """H2D + kernels + D2H through the public C entry point."""
        if out is None:
            out = self.alloc_outputs(params, pinned=False)
        timing = np.zeros(4)
        launches = C.c_int32()
        check(lib.misob200_run(self.h, C.byref(params), ptr(out["samples"]), ptr(out["loglik"]),
                               ptr(out["assignment"]), ptr(out["rundata"]), ptr(out["status"]),
                               ptr(timing), C.addressof(launches)))
        out["timing_ms"] = timing
        out["launches"] = launches.value
        out["params"] = params
        return out

    def upload(self, params):
        check(lib.misob200_upload(self.h, C.byref(params)))
        self._params = params

    def run_resident(self):
        ms, n = C.c_double(), C.c_int32()
        check(lib.misob200_run_resident(self.h, C.addressof(ms), C.addressof(n)))
        return ms.value, n.value

    def download(self, out=None):
        if out is None:
            out = self.alloc_outputs(self._params, pinned=False)
        check(lib.misob200_download(self.h, ptr(out["samples"]), ptr(out["loglik"]),
                                    ptr(out["assignment"]), ptr(out["rundata"]), ptr(out["status"])))
        out["params"] = self._params
        return out

    def summarize(self):
        G = self.size()[0]
        s = np.zeros((max(G, 1), _lib.SUMMARY_F64))
        check(lib.misob200_summarize(self.h, ptr(s)))
        return s[:G]

    def compare(self, other):
        """Bayes factors of self (sample 1) vs other (sample 2), [n_genes, 32]:
        bf[8], mean1 - mean2 [8], mean|delta| [8], KDE(0) [8]."""
        G = self.size()[0]
        out = np.zeros((max(G, 1), 32))
        check(lib.misob200_compare(self.h, other.h, ptr(out)))
        return out[:G]

    def write_miso(self, out, paths, prefixes, suffixes, n_threads=0):
        """Batched ``.miso`` writer (``misob200_plan_write_miso``): one file per gene from the buffers
        ``run`` filled.  paths[g] None skips a gene; prefixes / suffixes from
        ``miso_format.header_static_parts``.  Returns (files written, bytes written)."""
        G = self.size()[0]
        if not (len(paths) == len(prefixes) == len(suffixes) == G):
            raise _lib.InternalError("write_miso: one path / prefix / suffix per gene")

        def carr(strs):
            a = (C.c_char_p * max(G, 1))()
            for i, t in enumerate(strs):
                a[i] = None if t is None else (t if isinstance(t, bytes) else str(t).encode())
            return a
        pa, pr, su = carr(paths), carr(prefixes), carr(suffixes)
        nf, nb = C.c_int64(), C.c_int64()
        check(lib.misob200_plan_write_miso(self.h, C.byref(out["params"]), pa, pr, su, ptr(out["samples"]),
                                           ptr(out["loglik"]), ptr(out["assignment"]), ptr(out["rundata"]),
                                           n_threads, C.addressof(nf), C.addressof(nb)))
        return nf.value, nb.value

    def bucket_timing(self):
        ms = np.zeros(9)
        check(lib.misob200_bucket_timing(self.h, ptr(ms)))
        return ms

    def transfer_bytes(self):
        a, b = C.c_int64(), C.c_int64()
        check(lib.misob200_transfer_bytes(self.h, C.addressof(a), C.addressof(b)))
        return a.value, b.value

    def release_device(self):
        lib.misob200_release_device(self.h)

    # ---- unpacking -----------------------------------------------------------
    def gene_result(self, out, g):
        """Per-gene view in the reference's shapes: samples K x (C*S)."""
        p = out["params"]
        info = self.info()
        K, R = int(info[g, 0]), int(info[g, 1])
        so, lo, ao = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib.misob200_plan_offsets(self.h, C.byref(p), g, C.addressof(so),
                                        C.addressof(lo), C.addressof(ao)))
        n = p.n_chains * ((p.n_iters - p.burn_in) // p.lag)
        smp = out["samples"][so.value:so.value + K * n].reshape(n, K).T
        return dict(samples=smp, loglik=out["loglik"][lo.value:lo.value + n],
                    assignment=out["assignment"][ao.value:ao.value + R],
                    rundata=out["rundata"][g], status=int(out["status"][g]))


def decode_summary(rec):
    """One 256-byte summary record -> dict (include/miso_b200.h)."""
    ints = rec[24:32].view(np.int32)
    K = int(ints[8])
    return dict(mean=rec[0:K].copy(), ci_low=rec[8:8 + K].copy(), ci_high=rec[16:16 + K].copy(),
                assigned_counts=ints[0:K].copy(), n_iso=K, accepted=int(ints[9]),
                rejected=int(ints[10]), status=int(ints[11]))
