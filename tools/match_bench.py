"""Setup stage, host threads vs GPU (match_kernel + order_kernel): python tools/match_bench.py [events] [chunk]
Plans the cfg-3 workload in chunks both ways and prints the time of each stage."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
import miso_b200 as mb
from workloads import Workload

G = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 12500
wl = bench.WORKLOADS["cfg3"]
ws = [Workload(1, 0, wl["reads"], 36, 250., 900., 4., seed=bench.SEED, gene_ids=np.arange(a, min(a + chunk, G), dtype=np.uint32))
      for a in range(0, G, chunk)]
mb.Plan().append(Workload(1, 64, 200, 36, 250., 900., 4., seed=1), match_device=0)      # context + buffers up
for mode in ("host", "device", "device, host sort"):
    os.environ.pop("MISOB200_HOST_SORT", None)
    if mode.endswith("host sort"):
        os.environ["MISOB200_HOST_SORT"] = "1"
    tot, stats = 0.0, np.zeros(5)
    for w in ws:
        t0 = time.perf_counter()
        p = mb.Plan().append(w, match_device=None if mode == "host" else 0)
        tot += time.perf_counter() - t0
        if mode != "host":
            stats += np.array(mb.Plan.last_match_stats())
        p.close()
    line = "%-18s plan stage of %d events in %d chunk(s): %.3f s" % (mode, G, len(ws), tot)
    if mode != "host":
        line += "  (GPU kernels %.1f ms, H2D %.1f ms of %.2f GB, D2H %.1f ms of %.2f GB)" % (
            stats[0], stats[1], stats[3] / 1e9, stats[2], stats[4] / 1e9)
    print(line, flush=True)
