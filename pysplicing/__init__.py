"""Drop-in ``pysplicing`` module: ``import pysplicing`` resolves here when the
repository root is on sys.path, with the names the reference's extension module
exports (``/root/reference/pysplicing/pysplicing/__init__.py``)."""
from miso_b200.pysplicing_api import *  # noqa: F401,F403
from miso_b200.pysplicing_api import (InternalError, createGene, MISO, MISOPaired, noIso,  # noqa: F401
                                      isoLength)
