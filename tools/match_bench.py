"""Device vs host read<->isoform matching on the cfg-3 sized batch (GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miso_b200 as mb
G = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
for kind, g, r in ((1, G, 2000), (0, 10000, 1000)):
    w = mb.Workload(kind, g, r, 36, 250., 900., 4., seed=1)
    t = time.time(); ph = mb.Plan().append(w); th = time.time() - t
    for rep in range(2):
        t = time.time(); pd = mb.Plan().append(w, match_device=0); td = time.time() - t
        k, h, d, bi, bo = mb.Plan.last_match_stats()
        print("kind %d G %d: host plan %.2f s | device-matched plan %.2f s (kernel %.2f ms, H2D %.1f ms, D2H %.1f ms; "
              "in %.2f GB out %.2f GB -> kernel %.0f GB/s)" % (kind, g, th, td, k, h, d, bi / 1e9, bo / 1e9, (bi + bo) / k / 1e6), flush=True)
        assert pd.size() == ph.size()
        pd.close()
    ph.close()
