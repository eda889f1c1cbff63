"""miso_b200 -- B200-native per-gene MCMC PSI sampler behind the pysplicing API.

Host code is Python over a ctypes C ABI (include/miso_b200.h) over hand-written
sm_100a CUDA (miso_b200/csrc/).  No CPU fallback: importing this package needs
libmiso_b200.so, running anything needs a B200.
"""
from ._lib import InternalError, LIB_PATH, device_count, stream_version  # noqa: F401
from .batch import (Gene, Plan, ReadBatch, make_params, decode_summary,  # noqa: F401
                    MISO_START_AUTO, MISO_START_UNIFORM, MISO_START_RANDOM,
                    MISO_START_GIVEN, MISO_START_LINEAR, MISO_STOP_FIXEDNO,
                    MISO_STOP_CONVERGENT_MEAN, MISO_ALGO_REASSIGN, MISO_ALGO_MARGINAL,
                    MISO_ALGO_CLASSES)


def __getattr__(name):
    # Workload (synthetic bench / test inputs) lives outside the product, in workloads/
    if name == "Workload":
        import os
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        if root not in sys.path:
            sys.path.insert(0, root)
        from workloads import Workload
        return Workload
    raise AttributeError(name)
