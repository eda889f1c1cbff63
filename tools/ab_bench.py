"""A/B timing of differently built libraries on the GPU box (development only).

  python tools/ab_bench.py [n_genes] lib_a.so lib_b.so ...

For every library: the cfg-3 shaped workload, each isoform-count bucket timed by itself
(MISOB200_ONLY_K) and then the whole step.  Each library runs in its own process."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import miso_b200 as mb
    G = int(sys.argv[2])
    kind = int(sys.argv[3])
    w = mb.Workload(kind, G, 2000 if kind else 1000, 36, 250., 900., 4., seed=1)
    plan = mb.Plan().append(w)
    plan.upload(mb.make_params(5000, 500, 10, 1, seed=1))
    out = []
    ks = [0] if (not kind or os.environ.get('AB_ALL_ONLY')) else [2, 3, 4, 5, 6, 7, 8, 0]
    for k in ks:
        if k:
            os.environ["MISOB200_ONLY_K"] = str(k)
        else:
            os.environ.pop("MISOB200_ONLY_K", None)
        best = min(plan.run_resident()[0] for _ in range(2 if k else 3))
        out.append("%s %.1f" % ("K%d" % k if k else "all", best))
    if kind and not os.environ.get('AB_ALL_ONLY'):
        out.append("buckets " + " ".join("%.0f" % x for x in plan.bucket_timing() if x > 0))
    print("kind %d G %d: " % (kind, G) + "  ".join(out) + "  -> %.4g it/s" % (G * 5000 / (best / 1e3)), flush=True)
    plan.close()
    sys.exit(0)
args = sys.argv[1:]
G = int(args.pop(0)) if args and args[0].isdigit() else 50000
for lib in args or [os.path.join(ROOT, "miso_b200", "libmiso_b200.so")]:
    env = dict(os.environ, MISOB200_LIB=os.path.abspath(lib))
    print("==", lib, flush=True)
    for kind, g in ((1, G), (0, 10000)):
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(g), str(kind)], env=env)
