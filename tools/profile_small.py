"""Small resident run for ncu captures: one wave of PE genes, short chains."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miso_b200 as mb
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 600
w = mb.Workload(1, G, 2000, 36, 250., 900., 4., seed=1)
plan = mb.Plan().append(w)
plan.upload(mb.make_params(iters, iters // 10, 10, 1, seed=1))
ms, nl = plan.run_resident()
print("G", G, "iters", iters, "kernel ms", ms, "launches", nl, "iters/s %.3g" % (G * iters / ms * 1e3))
print("bucket ms", plan.bucket_timing())
