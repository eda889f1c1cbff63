"""One-chain-per-warp vs four-chains-per-warp with the machine full either way: 28 000 events of ONE isoform count
(7 000 four-chain units = 3 waves of 2 368 warps).  python tools/quad_fair.py [K ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, miso_b200 as mb
from workloads import Workload

Ks = [int(a) for a in sys.argv[1:]] or [3, 5, 7]
wl = bench.WORKLOADS["cfg3"]
allk = Workload(1, 300000, 0, 36, 250., 900., 4., seed=bench.SEED).n_iso()
params = mb.make_params(bench.ITERS, bench.BURN, bench.LAG, 1, seed=1)
for K in Ks:
    ids = np.flatnonzero(allk == K)[:28000].astype(np.uint32)
    plan, _, _ = bench.build_plan(mb, wl, ids)
    plan.upload(params)
    res = []
    for name, lim in (("one chain per warp", "0"), ("four chains per warp", "100000")):
        os.environ["MISOB200_QUAD_MAX_READS"] = lim
        plan.run_resident()
        ms = min(plan.run_resident()[0] for _ in range(2))
        res.append("%s %.1f ms" % (name, ms))
    print("K = %d, %d events, mean drawing reads %.0f: %s" % (K, len(ids), plan.info()[:, 2].mean(), ";  ".join(res)), flush=True)
    plan.close()
