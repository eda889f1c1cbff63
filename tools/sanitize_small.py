"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck): device setup (match_kernel,
order_kernel), chain + quad kernels with chains cut into short segments and handed over through the
ready queues, helper grids, zero-copy outputs, summary and comparison kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MISOB200_SEG_ITERS", "16")
os.environ.setdefault("MISOB200_SEG_ALWAYS", "1")
import numpy as np
import miso_b200 as mb
from workloads import Workload

params = mb.make_params(80, 16, 4, 2, seed=3)
plans = []
for smp in (0, 1):
    w = Workload(1, 96, 160, 36, 250., 900., 4., seed=7, sample=smp)
    p = mb.Plan().append(w, match_device=0)
    out = p.alloc_outputs(params, pinned=True)
    p.run(params, out)
    assert (out["status"] == 0).all() and np.isfinite(out["samples"]).all()
    plans.append(p)
s = plans[0].summarize()
c = plans[0].compare(plans[1])
w = Workload(0, 64, 120, 36, seed=9)
p = mb.Plan().append(w, match_device=0)
o = p.run(mb.make_params(60, 10, 5, 1, seed=1))
print("ok", s.shape, c.shape, o["launches"])
