"""Fixed overhead of a bucket run: short chains with and without segment hand-over (development, GPU box).
  MISOB200_LIB=... python tools/seg_overhead_probe.py [G] [K,K,...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miso_b200 as mb
G = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8]
w = mb.Workload(1, G, 2000, 36, 250., 900., 4., seed=1)
plan = mb.Plan().append(w)
os.environ["MISOB200_SCHED"] = "serial"
for k in ks:
    os.environ["MISOB200_ONLY_K"] = str(k)
    for iters in (300, 600, 1500):
        plan.upload(mb.make_params(iters, iters // 10, 10, 1, seed=1))
        for seg in ("256", "100000"):
            os.environ["MISOB200_SEG_ITERS"] = seg
            print("K %d iters %4d seg %6s" % (k, iters, seg), flush=True)
            sys.stderr.flush()
            ms = min(plan.run_resident()[0] for _ in range(2))
            print("   -> %.2f ms" % ms, flush=True)
