"""miso_b200.pipeline.run_pipelined: batches planned by background threads (host setup, or the device
setup in two phases on the library's stages) while the GPU runs the previous batch -- the posteriors are
those of one plan holding all the genes, bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb():
    import miso_b200
    if miso_b200.device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu tests must run on a B200")
    return miso_b200


@pytest.mark.parametrize("match_device", [None, 0], ids=["host-setup", "device-setup"])
def test_pipelined_batches_equal_one_plan(mb, match_device):
    from miso_b200.pipeline import run_pipelined
    ids = np.arange(700, 700 + 5 * 120, dtype=np.uint32)
    params = mb.make_params(400, 80, 4, 2, seed=12)
    whole = mb.Workload(1, 0, 300, 36, 250.0, 900.0, 4.0, seed=3, gene_ids=ids)
    big = mb.Plan().append(whole)
    big_out = big.run(params)
    batches = [mb.Workload(1, 0, 300, 36, 250.0, 900.0, 4.0, seed=3, gene_ids=ids[a:a + 120]) for a in range(0, 600, 120)]
    seen, stats = [], {}
    res = run_pipelined(batches, params, match_device=match_device, on_result=lambda i, p, o: seen.append(i), stats=stats)
    assert seen == [0, 1, 2, 3, 4] and len(res) == 5 and len(stats["run"]) == 5 and len(stats["plan"]) == 5
    for b, (plan, out) in enumerate(res):
        assert (out["status"] == 0).all()
        for j in range(0, 120, 17):
            got, want = plan.gene_result(out, j), big.gene_result(big_out, b * 120 + j)
            for k in ("samples", "loglik", "assignment"):
                np.testing.assert_array_equal(got[k], want[k], err_msg="batch %d gene %d %s" % (b, j, k))
            np.testing.assert_array_equal(got["rundata"], want["rundata"])


def test_begun_append_can_be_abandoned(mb):
    """A plan closed between append_begin and append_finish gives its stage back (four in a row would
    block forever otherwise: the library has three stages)."""
    w = mb.Workload(0, 30, 100, 36, 250.0, 900.0, 4.0, seed=5)
    for _ in range(5):
        p = mb.Plan().append_begin(w, 0)
        p.close()
    p = mb.Plan().append_begin(w, 0).append_finish()
    assert p.size()[0] == 30
