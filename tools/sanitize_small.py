"""Small run of every device layout for compute-sanitizer (memcheck / racecheck) on the GPU box:
  compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miso_b200 as mb
for cpw in ("4", "1"):
    os.environ["MISOB200_CHAINS_PER_WARP"] = cpw
    for kind, fmt, wide in ((0, -1, False), (1, -1, False), (1, 0, False), (1, -1, True)):
        sd2 = 2500.0 if wide else 900.0
        w = mb.Workload(kind, 23, 300, 36, 300.0 if wide else 250.0, sd2, 4.0, seed=3)
        plan = mb.Plan(tile_format=fmt).append(w)
        for start in (0, 2):
            out = plan.run(mb.make_params(60, 10, 5, 2, start=start, seed=5))
        plan.summarize()
        plan.close()
        print("ok cpw", cpw, "kind", kind, "fmt", fmt, "wide", wide, flush=True)
