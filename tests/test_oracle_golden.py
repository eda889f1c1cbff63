"""The plain-C restatement (oracle/miso_oracle.c) against the golden vectors the
UNMODIFIED reference produced (tests/golden/make_golden.py): decisions exact,
values to the last bits (same arithmetic order; 1e-12 leaves room for a
different libm build)."""
import numpy as np
import pytest

from golden_util import Params, STREAMS, load_cases
from helpers import oracle_gene

CASES = [c for v in STREAMS for c in load_cases(v)]       # both stream versions (Philox4x32-10 / -7)


@pytest.mark.parametrize("case", CASES, ids=["v%d-%s" % (c.stream, c.name) for c in CASES])
def test_port_reproduces_reference_golden(port, case, request):
    was = port.stream
    request.addfinalizer(lambda: port.set_stream(was))
    port.set_stream(case.stream)
    got = oracle_gene(port, (case.exons, case.isoforms, case.pos, case.cig), bool(case.paired), Params(case),
                      case.gene_id, pe=case.pe, read_len=case.read_len, overhang=case.overhang)
    S = (case.n_iters - case.burn_in) // case.lag
    n = case.n_chains * S
    np.testing.assert_array_equal(got["assignment"], case.assignment)
    assert (got["accepted"], got["rejected"]) == (case.accepted, case.rejected)
    np.testing.assert_allclose(got["samples"][:, :n], case.samples, rtol=1e-12, atol=0)
    np.testing.assert_allclose(got["loglik"][:n], case.loglik, rtol=1e-12, atol=0)
    np.testing.assert_array_equal(got["class_templates"], case.class_templates)
    np.testing.assert_array_equal(got["class_counts"], case.class_counts)


def test_golden_covers_baseline_config_1():
    c = [c for c in CASES if c.name == "cfg1_default"][-1]
    assert len(c.pos) > 700 and len(c.isoforms) == 2 and c.n_chains == 6
    # the skipped exon of this event is mostly excluded in the C2C12 sample
    assert 0.03 < c.samples[0].mean() < 0.12
