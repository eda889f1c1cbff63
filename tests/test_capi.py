"""The C-ABI library: loads, exports every symbol include/miso_b200.h declares,
struct layouts agree with the ctypes mirrors, and the device entry points fail
loudly (never silently fall back) on a box without a GPU."""
import ctypes as C
import os
import re

import pytest

import miso_b200 as mb
from miso_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "miso_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(misob200_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(_lib.lib, s), s
    assert set(_lib.EXPORTS) <= set(syms) | {"misob200_plan_keep_match"}
    assert set(syms) == set(_lib.EXPORTS)


def test_workload_library_exports_its_header():
    """workloads/libmiso_synth.so (bench / test inputs, outside the product) against include/miso_synth.h."""
    import workloads
    src = open(os.path.join(ROOT, "include", "miso_synth.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    syms = sorted(set(re.findall(r"\b(misob200_\w+)\s*\(", src)))
    assert set(syms) == set(workloads.EXPORTS)
    for s in syms:
        assert hasattr(workloads.lib, s), s
    assert C.sizeof(workloads.Reads) == C.sizeof(_lib.Reads)
    # the product library holds no generator
    assert not hasattr(_lib.lib, "misob200_workload_create")


def test_struct_layouts():
    assert C.sizeof(_lib.Params) == 40
    assert C.sizeof(_lib.Reads) == 4 + 4 + 10 * 8 + 3 * 4 + 4 + 3 * 8
    assert _lib.lib.misob200_version() >= 100


def test_no_oracle_or_torch_in_the_product():
    """The product path must not route through the oracle or a framework."""
    pkg = os.path.join(ROOT, "miso_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "refdriver" not in txt and "miso_oracle" not in txt and "libsplicing_ref" not in txt, f
                assert not re.search(r"^\s*(import|from)\s+torch", txt, flags=re.M), f


@pytest.mark.skipif(mb.device_count() > 0, reason="this check is for boxes without a GPU")
def test_device_calls_fail_loudly_without_gpu():
    g = mb.Gene(((1, 100), (201, 300), (401, 500)), ((0, 1), (0, 2), (0, 1, 2)))
    plan = mb.Plan().append(mb.ReadBatch([g], [[10, 20]], [["33M", "33M"]], 33))
    with pytest.raises(mb.InternalError, match="no CUDA device"):
        plan.run(mb.make_params(100, 10, 5, 1))
    import pysplicing
    with pytest.raises(pysplicing.InternalError):
        pysplicing.MISO(pysplicing.createGene(g.exons, g.isoforms), 0, (10, 20), ("33M", "33M"), 33, 100, 10, 5)


def test_argument_checks_mirror_the_reference():
    import pysplicing
    g = pysplicing.createGene(((1, 100), (201, 300), (401, 500)), ((0, 1), (0, 2), (0, 1, 2)))
    with pytest.raises(TypeError, match="Need a tuple"):            # pyconvert.c:7-10
        pysplicing.MISO(g, 0, [1, 2], ("33M", "33M"), 33)
    with pytest.raises(pysplicing.InternalError, match="hyperparameter"):   # miso.c:698-701
        pysplicing.MISO(g, 0, (1, 2), ("33M", "33M"), 33, 100, 10, 5, (1.0, 1.0))
    with pytest.raises(pysplicing.InternalError, match="Overhang"):  # miso.c:691-694
        pysplicing.MISO(g, 0, (1, 2), ("33M", "33M"), 33, 100, 10, 5, None, 20)
    with pytest.raises(pysplicing.InternalError, match="chains"):    # miso.c:703-706
        pysplicing.MISO(g, 0, (1, 2), ("33M", "33M"), 33, 100, 10, 5, None, 1, 0)
    with pytest.raises(NotImplementedError):                          # ALGO_CLASSES is off-path
        pysplicing.MISO(g, 0, (1, 2), ("33M", "33M"), 33, 100, 10, 5, None, 1, 1, 0, 0, pysplicing.MISO_ALGO_CLASSES)
    with pytest.raises(NotImplementedError):
        pysplicing.simulateReads(g, 0, (0.2, 0.3, 0.5), 10, 33)
    assert (pysplicing.MISO_START_AUTO, pysplicing.MISO_START_LINEAR, pysplicing.MISO_STOP_CONVERGENT_MEAN,
            pysplicing.MISO_ALGO_CLASSES) == (0, 4, 1, 2)


def test_device_matching_fails_loudly_without_gpu():
    """misob200_plan_append_device has no host fallback: without a GPU it is an error, the plan stays empty."""
    if mb.device_count() > 0:
        pytest.skip("a GPU is present")
    w = mb.Workload(0, 4, 50, 36, 250.0, 900.0, 4.0, seed=1)
    p = mb.Plan()
    with pytest.raises(mb.InternalError, match="no CUDA device"):
        p.append(w, match_device=0)
    assert p.size()[0] == 0
