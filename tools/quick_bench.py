"""Quick resident-kernel timing of both workloads and both tile formats (GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miso_b200 as mb
G3 = int(sys.argv[1]) if len(sys.argv) > 1 else 16000
fmts = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [-1, 0]
for kind, G, R in ((0, 10000, 1000), (1, G3, 2000)):
    w = mb.Workload(kind, G, R, 36, 250., 900., 4., seed=1)
    for fmt in fmts:
        t = time.time(); plan = mb.Plan(tile_format=fmt).append(w); t2 = time.time()
        params = mb.make_params(5000, 500, 10, 1, seed=1)
        plan.upload(params)
        for rep in range(2):
            ms, nl = plan.run_resident()
        print('kind', kind, 'fmt', fmt, 'G', G, 'plan %.2fs tiles %.1f MB kernel %.1f ms launches %d -> %.4g iters/s'
              % (t2 - t, plan.size()[2] / 1e6, ms, nl, G * 5000 / (ms / 1e3)), flush=True)
        print('   bucket ms', [round(x, 1) for x in plan.bucket_timing()], flush=True)
        plan.close()
