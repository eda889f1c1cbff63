import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miso_b200 as mb
w = mb.Workload(1, 50000, 2000, 36, 250., 900., 4., seed=1)
plan = mb.Plan().append(w)
plan.upload(mb.make_params(5000, 500, 10, 1, seed=1))
os.environ["MISOB200_ONLY_K"] = "5"
for seg in ("100000", "1251"):
    os.environ["MISOB200_SEG_ITERS"] = seg
    print(seg, plan.run_resident()[0], flush=True)
