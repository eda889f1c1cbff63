"""Small run of the three auxiliary kernels for ncu captures: summary_kernel (posterior mean / 95% CI /
assigned counts), compare_kernel (two-sample Bayes factors) and match_kernel (read <-> isoform matching
on the device).  python tools/profile_aux.py [genes]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import miso_b200 as mb
from workloads import Workload

G = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
params = mb.make_params(5000, 500, 10, 1, seed=1)
plans = []
for smp in (0, 1):
    w = Workload(1, G, 2000, 36, 250., 900., 4., seed=1, sample=smp)
    p = mb.Plan().append(w, match_device=0 if smp == 0 else None)      # sample 0: matching on the GPU
    if smp == 0:
        print("match stats (kernel ms, h2d ms, d2h ms, bytes in, bytes out):", mb.Plan.last_match_stats())
    w.close()
    p.upload(mb.make_params(300, 30, 1, 1, seed=1))      # 270 recorded samples per chain, short chains
    p.run_resident()
    plans.append(p)
s = plans[0].summarize()
c = plans[0].compare(plans[1])
print("summaries", s.shape, "mean psi_0 %.4f" % s[:, 0].mean(), "compare", c.shape, "median BF %.3g" % np.median(c[:, 0]))
