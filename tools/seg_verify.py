"""Full-size check that cutting chains into segments changes nothing (GPU box):
the cfg-3 sized batch unsplit vs cut everywhere (MISOB200_SEG_ALWAYS) -- samples, scores,
assignments and accept counts must be identical bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import miso_b200 as mb
G = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
for kind, g, r in ((1, G, 2000), (0, 10000, 1000)):
    w = mb.Workload(kind, g, r, 36, 250., 900., 4., seed=1)
    plan = mb.Plan().append(w)
    params = mb.make_params(iters, iters // 10, 10, 1, seed=1)
    res = {}
    for name, env in (("unsplit", {"MISOB200_SEG_ITERS": "100000000"}),
                      ("cut-everywhere", {"MISOB200_SEG_ITERS": "256", "MISOB200_SEG_ALWAYS": "1"}),
                      ("default", {})):
        for k in ("MISOB200_SEG_ITERS", "MISOB200_SEG_ALWAYS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        out = plan.run(params)
        res[name] = {k: np.array(out[k], copy=True) for k in ("samples", "loglik", "assignment", "rundata")}
        rd = res[name]["rundata"]
        print(kind, name, "kernel ms %.1f" % out["kernel_ms"] if "kernel_ms" in out else "", "acc+rej ok:",
              bool((rd[:, 5] + rd[:, 6] == iters).all()), flush=True)
    for name in ("cut-everywhere", "default"):
        for k in res["unsplit"]:
            same = np.array_equal(res["unsplit"][k], res[name][k])
            print("   ", name, k, "identical" if same else "DIFFERENT", flush=True)
            if not same:
                sys.exit(1)
print("segments verified")
