#!/usr/bin/env python
"""bench.py -- gene-chain iterations/sec of the per-gene MCMC PSI sampler.

Metric (BASELINE.json): gene-chain iterations/sec at 1/2/4/8 B200 vs the
reference pysplicing C path on host CPU.  A "step" is one pass of the hot path
over one batch: the cfg-3 workload of BASELINE.json configs[2] -- 50k mixed
events (2-8 isoforms), 2k paired-end reads each with the insert-length model,
5000 iterations (burn-in 500, lag 10, 1 chain), synthetic (workloads/synth.cpp).

N > 1 (BASELINE cfg-4): THE SAME 50k-event workload is dealt to the N ranks
(longest-processing-time first, miso_b200/shard.py), every rank plans and runs
only its shard -- no data-path collective -- and the step ends with one NCCL
all-gather of the 256-byte per-gene posterior summaries, device to device
("scaling": "strong").  `--scaling weak` gives every rank a full-size shard of
its own instead (secondary number).  `--workload cfg5`: two samples of the 50k
events (both on the rank that owns the event) + Bayes factors on the device,
all-gather of the comparison records.

  value : whole-job iterations/s with the packed tiles already resident in HBM
          (misob200_run_resident), CUDA-event time, max over ranks
  e2e   : the same metric through the public C entry point misob200_run with
          host buffers: pinned tiles H2D + kernels + D2H of posteriors + the
          host epilogue, every step, plus summary kernel and all-gather
  roofline / cpu_baseline / clocks / parity_checked: see DESIGN.md "measurement"

`--impl reference` times the reference's own CPU implementation (oracle/_ref,
the unmodified C core; falls back to the plain-C port) on all host cores on a
bounded sample of the same workload, a different slice every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gene-chain iterations/sec"
UNIT = "iterations/s"

WORKLOADS = {
    # name: kind (0 SE K=2, 1 PE mixed K), events, reads per event, samples per event
    "cfg3": dict(kind=1, n_genes=50000, reads=2000, samples=1,
                 label="cfg-3: 50k mixed events (K 2-8), 2k PE reads each, insert N(250,30^2) +-4sd, read_len 36"),
    "cfg2": dict(kind=0, n_genes=10000, reads=1000, samples=1,
                 label="cfg-2: 10k 2-isoform SE events, 1k SE reads each, read_len 36"),
    "cfg5": dict(kind=1, n_genes=50000, reads=2000, samples=2,
                 label="cfg-5: two samples of the 50k mixed events of cfg-3 (different psi) + Bayes-factor comparison"),
}
ITERS, BURN, LAG, CHAINS = 5000, 500, 10, 1
PE = (250.0, 900.0, 4.0)
READ_LEN = 36
SEED = 20260925
# relative cost of one gene-chain per isoform count (measured bucket times per gene on a B200,
# BENCH_r01: 24 / 71 / 92 / 100 / 118 / 138 / 160 ms per ~7.1k genes), times reads: the LPT key
COST_PER_READ = {2: 24, 3: 71, 4: 92, 5: 100, 6: 118, 7: 138, 8: 160}


def algorithmic_bytes(info, S):
    """bytes per gene-chain = 4R(K+2) + 8S(K+1) + 16K + 64."""
    K = info[:, 0].astype("int64")
    R = info[:, 1].astype("int64")
    return int((4 * R * (K + 2) + 8 * S * (K + 1) + 16 * K + 64).sum())


def measured_traffic(workload, n_genes):
    """dram__bytes_read + dram__bytes_write of one step from the committed ncu capture
    (profiles/traffic.json), only when it was taken at this workload size."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload)
        if t and t["events"] == n_genes:
            return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
    except Exception:
        pass
    return None


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# CPU arm (reference implementation on host cores).  Imports workloads/ and oracle/
# only: nothing of libmiso_b200.so is mapped by these processes.

_ORACLE = None


def _cpu_init(kind_ref):
    global _ORACLE
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refdriver
    _ORACLE = refdriver.RefOracle() if kind_ref == "reference" else refdriver.PortOracle()


def _cpu_gene(args):
    """One gene through the oracle; returns (iterations, seconds)."""
    gene, gid, paired = args
    ex, isos, pos, cig = gene
    kw = dict(iters=ITERS, burn=BURN, lag=LAG, chains=CHAINS, seed=SEED, gene_id=gid,
              rng_mode=1 if _ORACLE.kind == "reference" else 0)      # reference: its own MT19937 (fastest)
    t0 = time.perf_counter()
    if paired:
        _ORACLE.miso_pe(ex, isos, pos, cig, READ_LEN, PE[0], PE[1], PE[2], **kw)
    else:
        _ORACLE.miso_se(ex, isos, pos, cig, READ_LEN, **kw)
    return ITERS * CHAINS, time.perf_counter() - t0


def usable_cores():
    """Host cores this container may actually use: the scheduler affinity capped by the
    cgroup CPU quota (the GPU boxes show 128 logical CPUs but cpu.max = 16 cores' worth;
    more workers than that only thrash -- tools/cpu_probe.py)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period) + 0.5)))
    except Exception:
        pass
    return n


class CpuArm:
    """Reference C path on all usable host cores: a persistent pool of worker processes
    that pull genes one at a time (longest first), so a step's time is not the imbalance of
    a static split.  value = iterations / wall time of the step."""

    def __init__(self, wl):
        import multiprocessing as mp
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import refdriver
        self.kind = "reference" if refdriver.available() else "port"
        if self.kind == "port" and not refdriver.port_available():
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
        self.wl = wl
        self.cores = usable_cores()
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_cpu_init, initargs=(self.kind,))
        self.pool.map(_noop, range(self.cores * 4))          # workers up, library loaded

    def step(self, first_gene, n):
        """genes first_gene .. first_gene + n - 1 of the workload"""
        from workloads import Workload
        wl = self.wl
        w = Workload(wl["kind"], n, wl["reads"], READ_LEN, PE[0], PE[1], PE[2], seed=SEED, first_gene_id=first_gene)
        K = w.n_iso()
        order = sorted(range(n), key=lambda g: -int(K[g]))
        jobs = [(w.gene(g), first_gene + g, wl["kind"] == 1) for g in order]
        w.close()
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_gene, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        return dict(iters=sum(r[0] for r in res), wall=wall, busy=sum(r[1] for r in res), genes=n)

    def describe(self, steps):
        genes = sum(s["genes"] for s in steps)
        wall = sum(s["wall"] for s in steps)
        busy = sum(s["busy"] for s in steps)
        return ("%d genes of the workload (%.1f%% of its %d events; a different slice of %d genes every step), %d "
                "iterations each; %d persistent worker processes = all usable host cores (%d logical CPUs visible, "
                "cgroup quota applied) pulling one gene at a time, longest first; %s; wall %.1fs, summed worker "
                "time %.1fs (pool efficiency %.2f)" % (
                    genes, 100.0 * genes / self.wl["n_genes"], self.wl["n_genes"], steps[0]["genes"], ITERS,
                    self.cores, os.cpu_count() or 0,
                    "reference's own MT19937 stream" if self.kind == "reference" else "Philox stream",
                    wall, busy, busy / (wall * self.cores)))

    def close(self):
        self.pool.close()
        self.pool.join()


def _noop(_):
    return 0


def cpu_baseline_leg(wl, genes_per_core):
    arm = CpuArm(wl)
    s = arm.step(0, arm.cores * genes_per_core)
    out = dict(value=s["iters"] / s["wall"], unit=UNIT, cores=arm.cores, kind=arm.kind, sample=arm.describe([s]))
    arm.close()
    return out


# ---------------------------------------------------------------------------

def shard_ids(wl, rank, world, scaling):
    """Gene ids of this rank.  strong: the one workload dealt LPT by cost; weak: a private full-size shard."""
    import numpy as np
    n = wl["n_genes"]
    if world == 1:
        return np.arange(n, dtype=np.uint32), n
    if scaling == "weak":
        return np.arange(rank * n, (rank + 1) * n, dtype=np.uint32), n * world
    from workloads import Workload
    from miso_b200.shard import shard_genes
    w = Workload(wl["kind"], n, 0, READ_LEN, PE[0], PE[1], PE[2], seed=SEED)      # structures only
    K = w.n_iso()
    w.close()
    cost = np.array([COST_PER_READ.get(int(k), 160) for k in K], np.int64) * wl["reads"]
    return shard_genes(cost, world)[rank].astype(np.uint32), n


def build_plan(mb, wl, ids, sample=0, chunk=5000, match_device=None):
    from workloads import Workload
    plan = mb.Plan()
    t_gen = t_plan = 0.0
    for i in range(0, len(ids), chunk):
        t0 = time.perf_counter()
        w = Workload(wl["kind"], 0, wl["reads"], READ_LEN, PE[0], PE[1], PE[2], seed=SEED,
                     gene_ids=ids[i:i + chunk], sample=sample)
        t1 = time.perf_counter()
        plan.append(w, match_device=match_device)
        t2 = time.perf_counter()
        w.close()
        t_gen += t1 - t0
        t_plan += t2 - t1
    return plan, t_gen, t_plan


def parity_leg(mb, wl, plan, out, ids, per_k=2):
    """Untimed: the unmodified reference (oracle/_ref; else the pinned port), driven by the same
    Philox stream, on a seeded sample of the very genes that were timed -- `per_k` per isoform
    count.  Counts bit-exact, posterior mean / 95% CI within 1e-3 (BASELINE.json north_star)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refdriver
    from workloads import Workload
    oracle = refdriver.RefOracle() if refdriver.available() else refdriver.PortOracle()
    info = plan.info()
    rng = np.random.default_rng(SEED)
    pick = []
    for k in sorted(set(int(x) for x in info[:, 0])):
        cand = np.flatnonzero((info[:, 0] == k) & (info[:, 4] == 0))
        pick += [int(g) for g in rng.choice(cand, size=min(per_k, len(cand)), replace=False)]
    w = Workload(wl["kind"], 0, wl["reads"], READ_LEN, PE[0], PE[1], PE[2], seed=SEED, gene_ids=ids[pick])
    worst_mean = worst_ci = 0.0
    exact = True
    for j, g in enumerate(pick):
        ex, isos, pos, cig = w.gene(j)
        kw = dict(iters=ITERS, burn=BURN, lag=LAG, chains=1, seed=SEED, gene_id=int(ids[g]), rng_mode=0)
        if wl["kind"] == 1:
            want = oracle.miso_pe(ex, isos, pos, cig, READ_LEN, PE[0], PE[1], PE[2], **kw)
        else:
            want = oracle.miso_se(ex, isos, pos, cig, READ_LEN, **kw)
        got = plan.gene_result(out, g)
        S = (ITERS - BURN) // LAG
        exact = exact and bool((got["assignment"] == want["assignment"]).all()) \
            and int(got["rundata"][5]) == int(want["rundata"][5]) and int(got["rundata"][6]) == int(want["rundata"][6])
        a, b = got["samples"][:, :S], want["samples"][:, :S]
        worst_mean = max(worst_mean, float(np.abs(a.mean(axis=1) - b.mean(axis=1)).max()))
        lo, hi = int(np.floor(0.025 * S + 0.5)) - 1, int(np.floor(0.975 * S + 0.5)) - 1
        sa, sb = np.sort(a, axis=1), np.sort(b, axis=1)
        worst_ci = max(worst_ci, float(np.abs(sa[:, [lo, hi]] - sb[:, [lo, hi]]).max()))
    w.close()
    res = {"genes": len(pick), "oracle": oracle.kind, "counts_bit_exact": exact,
           "max_abs_mean_diff": worst_mean, "max_abs_ci_diff": worst_ci,
           "what": "per-read assignments and accept/reject counts equal; posterior mean and 95%% CI bounds vs the "
                   "oracle on the same stream, %d genes per isoform count of the timed plan" % per_k}
    if not exact or worst_mean > 1e-3 or worst_ci > 1e-3:
        raise SystemExit("bench.py: PARITY FAILURE on the timed plan: %s" % json.dumps(res))
    return res


def setup_leg(mb, wl, ids, params, match_device, big_plan, big_out, barrier):
    """Reads in -> posteriors out (untimed by the step clock, timed by itself): the rank's events in a few
    batches through miso_b200.pipeline.run_pipelined -- plan stage of batch i+1 on the host threads
    while the GPU runs batch i.  Generation of the synthetic reads and the allocation of the pinned
    output buffers are outside the clock, and so is one warm-up pass; results are checked against the
    one-plan run."""
    import numpy as np
    from miso_b200._lib import pinned_empty
    from miso_b200.pipeline import run_pipelined
    from workloads import Workload
    n_chunks = max(1, min(4, len(ids) // 6000))
    bounds = np.linspace(0, len(ids), n_chunks + 1).astype(int)
    ws, outs = [], []
    S = (ITERS - BURN) // LAG * CHAINS
    for a, b in zip(bounds[:-1], bounds[1:]):
        w = Workload(wl["kind"], 0, wl["reads"], READ_LEN, PE[0], PE[1], PE[2], seed=SEED, gene_ids=ids[a:b])
        K = w.n_iso().astype(np.int64)
        ws.append(w)
        outs.append(dict(samples=pinned_empty(int(K.sum()) * S, np.float64), loglik=pinned_empty(len(K) * S, np.float64),
                         assignment=pinned_empty(len(K) * wl["reads"], np.int32),
                         rundata=np.zeros((len(K), 9), np.int32), status=np.zeros(len(K), np.int32)))
    # one untimed pass first (like every other timing here: warm-up, then the clock): device and pinned
    # buffers of the library's pools exist afterwards
    for plan, _ in run_pipelined(ws, params, outputs=outs, match_device=match_device):
        plan.close()
    barrier(0.0)
    t0 = time.perf_counter()
    stats = {}
    res = run_pipelined(ws, params, outputs=outs, match_device=match_device, stats=stats)
    wall = barrier(time.perf_counter() - t0)
    for (plan, out), a in zip(res, bounds[:-1]):       # same posteriors as the one-plan run
        got, want = plan.gene_result(out, 0), big_plan.gene_result(big_out, int(a))
        assert np.array_equal(got["samples"], want["samples"]) and np.array_equal(got["assignment"], want["assignment"])
    for (plan, _), w in zip(res, ws):
        plan.close()
        w.close()
    return wall, n_chunks, {k: [round(x, 4) for x in v] for k, v in stats.items()}


def writer_leg(mb, plan, out, ids):
    """Untimed by the step clock, timed by itself: every event of the plan to `<dir>/<chrom>/<event>.miso`
    through the batched writer (misob200_plan_write_miso; the reference writes file by file from
    Python, miso_sampler.py:456-465).  Files go to shared memory when the box has it, so the
    number is the formatter + the file system calls, not a disk."""
    import shutil
    import tempfile
    from miso_b200 import miso_format as mf
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    root = tempfile.mkdtemp(prefix="misob200_bench_", dir=base)
    try:
        info = plan.info()
        G = info.shape[0]
        chroms = ["chr%d" % (1 + c) for c in range(20)]
        for c in chroms:
            os.mkdir(os.path.join(root, c))
        parts = {}
        for k in sorted(set(int(x) for x in info[:, 0])):     # the synthetic events of one K share a structure
            descs = [["E%d" % e for e in range(k + 1) if e != i or i == 0] for i in range(k)]
            parts[k] = (descs, [("E%d" % e, 200) for e in range(k + 1)], [1] * k, [400 * k + 200] * k)
        t0 = time.perf_counter()
        paths, pre, suf = [], [], []
        for g in range(G):
            k = int(info[g, 0])
            c = chroms[g % 20]
            a, b = mf.header_static_parts(parts[k][0], parts[k][1], c, "+", parts[k][2], parts[k][3])
            paths.append(os.path.join(root, c, "event%07d.miso" % int(ids[g])))
            pre.append(a)
            suf.append(b)
        t1 = time.perf_counter()
        nf, nb = plan.write_miso(out, paths, pre, suf)
        t2 = time.perf_counter()
        return {"files": int(nf), "bytes": int(nb), "seconds": t2 - t1, "files_per_s": nf / (t2 - t1),
                "MB_per_s": nb / (t2 - t1) / 1e6, "header_static_parts_s": t1 - t0,
                "threads": int(mb._lib.lib.misob200_host_threads()), "dir": "shared memory" if base else "tmp dir",
                "what": "one .miso file per event of rank 0's plan (header + %d posterior samples) from the pinned "
                        "output buffers" % ((ITERS - BURN) // LAG * CHAINS)}
    finally:
        shutil.rmtree(root, ignore_errors=True)


_JSON_OUT = None


def keep_stdout_for_json():
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints
    its version banner on fd 1): keep the real stdout aside for the JSON line and point fd 1
    at stderr for everything else."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def main():
    keep_stdout_for_json()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = the one workload sharded over the ranks (cfg-4), weak = one full-size shard per rank")
    ap.add_argument("--genes", type=int, default=0, help="override the number of events (debugging only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-writer", action="store_true")
    ap.add_argument("--cpu-genes-per-core", type=int, default=40)
    ap.add_argument("--ref-genes-per-core-per-step", type=int, default=10)
    ap.add_argument("--setup", default="device", choices=["device", "host"],
                    help="e2e_with_setup leg: matching + draw-order sort on the GPU (default) or the whole plan stage on host threads")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = dict(WORKLOADS[args.workload])
    if args.genes:
        wl["n_genes"] = args.genes
    S = (ITERS - BURN) // LAG
    scaling = args.scaling if world > 1 else "strong"
    total_events = wl["n_genes"] * (world if scaling == "weak" else 1)
    config = {"workload": wl["label"] + "; %d iterations, burn-in %d, lag %d, %d chain; %d events in total"
              % (ITERS, BURN, LAG, CHAINS, total_events),
              "events_total": total_events, "reads_per_event": wl["reads"], "samples_per_event": wl["samples"],
              "parallelism": ("1 GPU" if world == 1 else
                              "dp%d: the workload's events dealt LPT to %d ranks, one process per GPU, one NCCL "
                              "all-gather of the summaries per step" % (world, world) if scaling == "strong" else
                              "dp%d: one full-size shard per rank (weak scaling), one NCCL all-gather per step" % world),
              "l2_note": "packed inputs (tens to hundreds of MB per GPU) are read once per step and the outputs "
                         "(>= 140 MB per GPU) exceed the 126 MB L2; no flush needed"}

    # ------------------------------------------------------------------ CPU arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        arm = CpuArm(wl)
        per_step = arm.cores * args.ref_genes_per_core_per_step
        steps = []
        for i in range(args.warmup + args.steps):
            s = arm.step((i * per_step) % max(wl["n_genes"] - per_step, 1), per_step)
            if i >= args.warmup:
                steps.append(s)
        v = sum(s["iters"] for s in steps) / sum(s["wall"] for s in steps)
        ms = 1e3 * sum(s["wall"] for s in steps) / len(steps)
        sample = arm.describe(steps)
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": sample},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        arm.close()
        emit(line)
        return 0

    # ------------------------------------------------------------------ GPU arm
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_baseline_leg(wl, args.cpu_genes_per_core)      # before CUDA is touched in this process

    import numpy as np
    import miso_b200 as mb
    from miso_b200._lib import lib, check, ptr, pinned_empty
    import ctypes as C

    if mb.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device -- miso_b200 has no CPU path")
    device = local_rank % mb.device_count()

    if world > 1:
        import datetime
        import torch.distributed as dist     # rendezvous only (unique-id broadcast); no tensors
        dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(minutes=30))
        check(lib.misob200_init(device))
        idbuf = (C.c_char * 128)()
        if rank == 0:
            check(lib.misob200_comm_unique_id(idbuf))
        obj = [bytes(idbuf.raw)]
        dist.broadcast_object_list(obj, src=0)
        check(lib.misob200_comm_init(obj[0], world, rank))

    def barrier_max(x):
        if world == 1:
            return x
        v = np.array([x], np.float64)
        check(lib.misob200_comm_barrier_max(ptr(v)))
        return float(v[0])

    def barrier_sum_i64(x):
        if world == 1:
            return int(x)
        t = [None] * world
        dist.all_gather_object(t, int(x))
        return int(sum(t))

    ids, _ = shard_ids(wl, rank, world, scaling)
    md = device if args.setup == "device" else None
    plans, t_gen, t_plan = [], 0.0, 0.0
    for smp in range(wl["samples"]):
        p, tg, tp = build_plan(mb, wl, ids, sample=smp)
        plans.append(p)
        t_gen += tg
        t_plan += tp
    params = mb.make_params(ITERS, BURN, LAG, CHAINS, seed=SEED, device=device)
    info = plans[0].info()
    G = info.shape[0]
    ok = sum(int((p.info()[:, 4] == 0).sum()) for p in plans)
    iters_local = ok * CHAINS * ITERS
    iters_total = barrier_sum_i64(iters_local)
    n_pad = int(barrier_max(float(G)))
    outs = [p.alloc_outputs(params, pinned=True) for p in plans]
    gathered = pinned_empty(world * n_pad * 32, np.float64) if world > 1 else None
    extra_launches = 0

    def e2e_step():
        nl = 0
        for p, o in zip(plans, outs):
            p.run(params, o)
            nl += o["launches"]
        if wl["samples"] == 2:
            if world > 1:
                check(lib.misob200_comm_allgather_compare(plans[0].h, plans[1].h, n_pad, ptr(gathered)))
            else:
                plans[0].compare(plans[1])
            return nl + 1
        if world > 1:
            check(lib.misob200_comm_allgather_summaries(plans[0].h, n_pad, ptr(gathered)))
        else:
            plans[0].summarize()
        return nl + 1

    def resident_step():
        ms = nl = 0
        for p in plans:
            a, b = p.run_resident()
            ms += a
            nl += b
        return ms, nl

    # warm-up: both paths
    for p in plans:
        p.upload(params)
    for _ in range(args.warmup):
        resident_step()
    clocks = ClockSampler(device)
    barrier_max(0.0)
    clocks.start()
    t0 = time.perf_counter()
    dev_ms, launches, bucket = 0.0, 0, np.zeros(9)
    for _ in range(args.steps):
        ms, nl = resident_step()
        dev_ms += ms
        launches += nl
        for p in plans:
            bucket += p.bucket_timing()
    wall_res = time.perf_counter() - t0
    clk = clocks.stop()
    dev_ms = barrier_max(dev_ms)
    wall_res = barrier_max(wall_res)

    # e2e through misob200_run
    for _ in range(min(args.warmup, 2)):
        e2e_step()
    barrier_max(0.0)
    t0 = time.perf_counter()
    e2e_launches = 0
    for _ in range(args.steps):
        e2e_launches += e2e_step()
    wall_e2e = barrier_max(time.perf_counter() - t0)
    h2d = d2h = 0
    for p in plans:
        a, b = p.transfer_bytes()
        h2d += a
        d2h += b
    if world > 1:
        d2h += gathered.nbytes
    timing = outs[0]["timing_ms"].copy()
    e2e_ms = 1e3 * wall_e2e / args.steps

    setup_wall, setup_chunks, setup_stats = setup_leg(mb, wl, ids, params, md, plans[0], outs[0], barrier_max) \
        if wl["samples"] == 1 else (t_plan + e2e_ms / 1e3, 1, None)

    # what was timed, checked (untimed): the reference on a seeded sample of this very plan
    parity = None
    if rank == 0 and not args.no_parity:
        parity = parity_leg(mb, wl, plans[0], outs[0], ids)
    writer = None
    if rank == 0 and not args.no_writer:
        writer = writer_leg(mb, plans[0], outs[0], ids)
    if world > 1 and wl["samples"] == 1:
        # the gathered table holds every rank's records: all events accounted for
        tab = np.asarray(gathered).reshape(world, n_pad, 32)
        status = tab[:, :, 24:32].copy().view(np.int32)[:, :, 11]
        assert int((status == 0).sum()) * CHAINS * ITERS == iters_total, "all-gather lost records"

    if rank == 0:
        ms_per_step = dev_ms / args.steps
        value = iters_total / (ms_per_step / 1e3)
        e2e_val = iters_total / (wall_e2e / args.steps)
        alg = algorithmic_bytes(info[info[:, 4] == 0], S) * wl["samples"]
        peak, how = hbm_peak()
        achieved = alg / (ms_per_step / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "clocks": clk,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms,
                    "last_step_ms": {"h2d": timing[0], "kernels": timing[1], "d2h": timing[2]}},
            "e2e_with_setup": {"value": iters_total / setup_wall, "unit": UNIT, "seconds": setup_wall,
                               "what": "reads in -> posteriors out, max over ranks: every rank's events in %d batch(es) "
                                       "through miso_b200.pipeline.run_pipelined -- the plan stage (matching, draw order, "
                                       "classes, tile packing; %s) of batch i+1 overlaps the GPU run of batch i"
                                       % (setup_chunks, "matching and draw-order sort on the GPU, classes and tiles on host threads"
                                          if args.setup == "device" else "host threads"),
                               "unpipelined_host_setup_seconds": t_plan + e2e_ms / 1e3,
                               "per_batch_seconds": setup_stats,
                               "plan_stage_s": t_plan, "host_threads": int(lib.misob200_host_threads())},
            "gpu_launches": int(launches),
            "gpu_launches_e2e": int(e2e_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic(args.workload, G) if world == 1 else None,
                         "peak_source": how,
                         "algorithmic_bytes_per_step": alg, "note": "rank 0's shard; algorithmic bytes and launch "
                         "time are per rank" if world > 1 else "whole workload",
                         "kernel": "chain_kernel<K> / quad_kernel<K>, one launch per isoform-count bucket K = 2..8",
                         "bucket_ms_per_step": {str(k): bucket[k] / args.steps for k in range(2, 9) if bucket[k] > 0}},
            "cpu_baseline": cpu,
            "parity_checked": parity,
            "writer": writer,
            "events_per_s_e2e": total_events / (e2e_ms / 1e3),
            "setup_seconds": {"synthetic_generation": t_gen, "host_plan_stage": t_plan,
                              "note": "the timed plan is built once, on host threads, outside every clock"},
            "wall_ms_per_resident_step": 1e3 * wall_res / args.steps,
        }
        emit(line)
    if world > 1:
        lib.misob200_comm_destroy()
    return 0


if __name__ == "__main__":
    sys.exit(main())
