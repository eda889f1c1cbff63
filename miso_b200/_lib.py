"""ctypes binding of libmiso_b200.so (include/miso_b200.h).

This is the reference-side stub a maintainer would add in place of the
CPython-2 extension (``/root/reference/pysplicing/src/pysplicing.c:659-710``):
plain pointers and sizes, no framework types.  There is no CPU path: if the
shared library is missing the import fails loudly, and every device call fails
with ``InternalError`` when no sm_100 GPU is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MISOB200_LIB: development only (A/B runs of differently tuned builds on the GPU box)
LIB_PATH = os.environ.get("MISOB200_LIB") or os.path.join(_HERE, "libmiso_b200.so")

SUCCESS, FAILURE, ENOMEM, EINVAL, UNIMPLEMENTED, ECUDA, ENCCL = 0, 1, 2, 4, 12, 100, 101
SUMMARY_F64 = 32
MAX_ISO = 8


class InternalError(Exception):
    """pysplicing.InternalError (``pysplicing/src/pyerror.c:26-45``)."""


class Reads(C.Structure):
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


class Params(C.Structure):
    _fields_ = [
        ("n_iters", C.c_int32), ("burn_in", C.c_int32), ("lag", C.c_int32),
        ("n_chains", C.c_int32), ("start", C.c_int32), ("stop", C.c_int32),
        ("algo", C.c_int32), ("device", C.c_int32), ("seed", C.c_uint64),
    ]


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "miso_b200: %s is missing -- build it with `python -c 'import "
            "__graft_entry__ as g; g.build()'` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.misob200_last_error.restype = C.c_char_p
    lib.misob200_host_alloc.restype = C.c_void_p
    lib.misob200_host_alloc.argtypes = [C.c_int64]
    lib.misob200_host_free.argtypes = [C.c_void_p]
    vp = C.c_void_p
    lib.misob200_plan_create.argtypes = [C.POINTER(vp)]
    lib.misob200_plan_destroy.argtypes = [vp]
    lib.misob200_plan_keep_match.argtypes = [vp, C.c_int]
    lib.misob200_plan_tile_format.argtypes = [vp, C.c_int]
    lib.misob200_plan_gene_tile.argtypes = [vp, C.c_int32, vp, vp, vp]
    lib.misob200_plan_append.argtypes = [vp, vp, C.c_int]
    lib.misob200_plan_append_device.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.misob200_plan_append_device_begin.argtypes = [vp, vp, C.c_int, C.POINTER(vp)]
    lib.misob200_plan_append_device_finish.argtypes = [vp, vp, C.c_int]
    lib.misob200_last_match_stats.argtypes = [vp, vp, vp, vp, vp]
    lib.misob200_plan_size.argtypes = [vp, vp, vp, vp]
    lib.misob200_plan_gene_info.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp]
    lib.misob200_plan_info_all.argtypes = [vp, vp]
    lib.misob200_plan_offsets_all.argtypes = [vp, C.POINTER(Params), vp, vp, vp]
    lib.misob200_write_miso_files.argtypes = [C.c_int32, vp, vp, vp, vp, vp, vp, vp, C.c_int32, C.c_int, vp]
    lib.misob200_plan_write_miso.argtypes = [vp, C.POINTER(Params), vp, vp, vp, vp, vp, vp, vp, C.c_int, vp, vp]
    lib.misob200_format_fixed.argtypes = [C.c_double, C.c_int, vp]
    lib.misob200_plan_gene_classes.argtypes = [vp, C.c_int32, vp, vp]
    lib.misob200_plan_gene_match.argtypes = [vp, C.c_int32, vp, vp]
    lib.misob200_plan_fragment_table.argtypes = [vp, C.c_int32, vp, vp, vp]
    lib.misob200_plan_offsets.argtypes = [vp, C.POINTER(Params), C.c_int32, vp, vp, vp]
    lib.misob200_plan_output_sizes.argtypes = [vp, C.POINTER(Params), vp, vp, vp]
    lib.misob200_run.argtypes = [vp, C.POINTER(Params), vp, vp, vp, vp, vp, vp, vp]
    lib.misob200_upload.argtypes = [vp, C.POINTER(Params)]
    lib.misob200_run_resident.argtypes = [vp, vp, vp]
    lib.misob200_download.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.misob200_release_device.argtypes = [vp]
    lib.misob200_summarize.argtypes = [vp, vp]
    lib.misob200_compare.argtypes = [vp, vp, vp]
    lib.misob200_bucket_timing.argtypes = [vp, vp]
    lib.misob200_transfer_bytes.argtypes = [vp, vp, vp]
    lib.misob200_comm_unique_id.argtypes = [vp]
    lib.misob200_comm_init.argtypes = [vp, C.c_int, C.c_int]
    lib.misob200_comm_allgather.argtypes = [vp, C.c_int64, vp]
    lib.misob200_comm_allgather_summaries.argtypes = [vp, C.c_int64, vp]
    lib.misob200_comm_allgather_compare.argtypes = [vp, vp, C.c_int64, vp]
    lib.misob200_comm_barrier_max.argtypes = [vp]
    return lib


lib = _load()

EXPORTS = [
    "misob200_version", "misob200_last_error", "misob200_init", "misob200_shutdown",
    "misob200_device_count", "misob200_plan_create", "misob200_plan_destroy",
    "misob200_plan_append", "misob200_plan_append_device", "misob200_last_match_stats",
    "misob200_plan_append_device_begin", "misob200_plan_append_device_finish",
    "misob200_plan_keep_match", "misob200_plan_size",
    "misob200_plan_tile_format", "misob200_plan_gene_tile",
    "misob200_plan_gene_info", "misob200_plan_gene_classes", "misob200_plan_gene_match",
    "misob200_plan_fragment_table", "misob200_plan_offsets", "misob200_plan_output_sizes",
    "misob200_run", "misob200_upload", "misob200_run_resident", "misob200_download",
    "misob200_release_device", "misob200_summarize", "misob200_bucket_timing", "misob200_compare",
    "misob200_transfer_bytes", "misob200_comm_unique_id",
    "misob200_comm_init", "misob200_comm_allgather", "misob200_comm_allgather_summaries",
    "misob200_comm_allgather_compare", "misob200_host_threads", "misob200_stream_version",
    "misob200_plan_info_all", "misob200_plan_offsets_all", "misob200_write_miso_files", "misob200_format_fixed", "misob200_plan_write_miso", "misob200_comm_barrier_max",
    "misob200_comm_destroy", "misob200_host_alloc", "misob200_host_free",
]


def check(rc):
    if rc == 0:
        return
    msg = (lib.misob200_last_error() or b"").decode()
    if rc == ENOMEM:
        raise MemoryError(msg)
    if rc == UNIMPLEMENTED:
        raise NotImplementedError(msg)
    raise InternalError("miso_b200 error %d: %s" % (rc, msg))


def ptr(a):
    return None if a is None else a.ctypes.data


def stream_version(set=0):
    """Random-stream version in force (2: Philox4x32-7, default; 1: Philox4x32-10); set=1|2 selects."""
    return int(lib.misob200_stream_version(int(set)))


def device_count():
    n = C.c_int(0)
    lib.misob200_device_count(C.byref(n))
    return n.value


def pinned_empty(n, dtype):
    """numpy array over page-locked memory (falls back to pageable)."""
    dtype = np.dtype(dtype)
    nbytes = max(int(n), 1) * dtype.itemsize
    p = lib.misob200_host_alloc(nbytes)
    if not p:
        return np.empty(int(n), dtype)
    buf = (C.c_char * nbytes).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(n))
    _PINNED[arr.ctypes.data] = p
    return arr


_PINNED = {}


def pinned_free(arr):
    p = _PINNED.pop(arr.ctypes.data, None)
    if p:
        lib.misob200_host_free(p)
