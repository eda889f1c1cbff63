"""Where does the first run of a fresh plan spend its time?  python tools/fresh_plan_probe.py [events per plan] [plans]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MISOB200_RUN_DEBUG"] = "1"
import numpy as np
import bench, miso_b200 as mb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
m = int(sys.argv[2]) if len(sys.argv) > 2 else 3
params = mb.make_params(bench.ITERS, bench.BURN, bench.LAG, 1, seed=1)
wl = bench.WORKLOADS["cfg3"]
plans = [bench.build_plan(mb, wl, np.arange(i * n, (i + 1) * n, dtype=np.uint32))[0] for i in range(m)]
outs = [p.alloc_outputs(params, pinned=True) for p in plans]
for rep in range(2):
    for p, o in zip(plans, outs):
        t0 = time.perf_counter()
        p.run(params, o)
        print("rep %d: plan.run %.1f ms, timing_ms %s" % (rep, (time.perf_counter() - t0) * 1e3, np.round(o["timing_ms"], 1)), flush=True)
