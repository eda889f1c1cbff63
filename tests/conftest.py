import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def port():
    """oracle/libmiso_oracle.so, the plain-C restatement (built on demand: gcc only)."""
    import subprocess
    import refdriver
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])   # no-op when up to date
    return refdriver.PortOracle()


@pytest.fixture(scope="session")
def ref():
    """oracle/_ref/libsplicing_ref.so, the unmodified reference C (skipped if never built)."""
    import refdriver
    if not refdriver.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return refdriver.RefOracle()


@pytest.fixture(scope="session")
def ref_or_port(request):
    """The unmodified reference where it was built (this container), else the pinned port."""
    import refdriver
    if refdriver.available():
        return refdriver.RefOracle()
    return request.getfixturevalue("port")
