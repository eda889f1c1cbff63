"""How many host cores does this box really give us?  Throughput of the reference C path vs worker count."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(f, open(f).read().strip())
    except Exception as e: print(f, "n/a")
os.system("lscpu | egrep 'Model name|Socket|Core|Thread|MHz' | head -8")
if __name__ == "__main__":
    wl = dict(bench.WORKLOADS["cfg3"])
    for cores in (8, 16, 32, 64, 128):
        t = time.time()
        r = bench.cpu_arm(wl, 2, cores=cores)
        print(cores, "workers: %.0f it/s total, %.0f per worker, busy %.1fs, call %.1fs" % (r["value"], r["value"] / cores, r["seconds"], time.time() - t), flush=True)
