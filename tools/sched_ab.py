"""A/B of the launch policies on one GPU: python tools/sched_ab.py [genes ...]
For each batch size (50000 = the whole cfg-3 workload, 6250 = a rank's shard at 8 GPUs ...) the
resident step and the e2e step (pinned outputs) under MISOB200_SCHED=serial and the default."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import bench
import miso_b200 as mb

sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [50000, 25000, 12500, 6250]
MODES = [a for a in sys.argv[1:] if not a.isdigit()] or ["serial", "balanced"]
wl = bench.WORKLOADS["cfg3"]
params = mb.make_params(bench.ITERS, bench.BURN, bench.LAG, bench.CHAINS, seed=bench.SEED)
base = None
for n in sizes:
    ids = bench.shard_ids(dict(wl, n_genes=50000), 0, 50000 // n, "strong")[0] if n < 50000 else np.arange(n, dtype=np.uint32)
    plan, _, _ = bench.build_plan(mb, wl, ids)
    out = plan.alloc_outputs(params, pinned=True)
    for mode in MODES:
        os.environ["MISOB200_SCHED"] = mode.split("+")[0]
        os.environ.pop("MISOB200_CLUSTER", None)
        if "+c" in mode:
            os.environ["MISOB200_CLUSTER"] = mode.split("+c")[1]
        plan.upload(params)
        plan.run_resident()
        ms = [plan.run_resident()[0] for _ in range(6)]
        bt = plan.bucket_timing()
        plan.run(params, out)
        t0 = time.perf_counter()
        for _ in range(2):
            plan.run(params, out)
        e2e = (time.perf_counter() - t0) / 2 * 1e3
        os.environ["MISOB200_NO_ZEROCOPY"] = "1"
        plan.run(params, out)
        t0 = time.perf_counter()
        plan.run(params, out)
        e2e_copy = (time.perf_counter() - t0) * 1e3
        del os.environ["MISOB200_NO_ZEROCOPY"]
        r = min(ms)
        if n == 50000 and mode == "serial":
            base = r
        print("genes %6d  %-11s resident %7.2f ms (%s)  e2e %7.2f ms  e2e with D2H copies %7.2f ms  launches %d%s  buckets %s" % (
            len(ids), mode, r, " ".join("%.1f" % x for x in ms), e2e, e2e_copy, out["launches"],
            ("  strong-scaling efficiency vs 50000/serial: %.3f" % (base * len(ids) / 50000 / r)) if base else "",
            " ".join("%d:%.1f" % (k, bt[k]) for k in range(2, 9))), flush=True)
    plan.close()
