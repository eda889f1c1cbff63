"""Development probe: one K bucket, several segment lengths (GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miso_b200 as mb
G = int(sys.argv[1]); ks = [int(x) for x in sys.argv[2].split(",")]; segs = sys.argv[3].split(",")
w = mb.Workload(1, G, 2000, 36, 250., 900., 4., seed=1)
plan = mb.Plan().append(w)
plan.upload(mb.make_params(5000, 500, 10, 1, seed=1))
for k in ks:
    os.environ["MISOB200_ONLY_K"] = str(k)
    for seg in segs:
        os.environ["MISOB200_SEG_ITERS"] = seg
        best = min(plan.run_resident()[0] for _ in range(2))
        print("K", k, "seg", seg, "%.1f ms" % best, flush=True)
