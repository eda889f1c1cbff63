#!/usr/bin/env python
"""bench.py -- gene-chain iterations/sec of the per-gene MCMC PSI sampler.

Metric (BASELINE.json): gene-chain iterations/sec at 1/2/4/8 B200 vs the
reference pysplicing C path on host CPU.  A "step" is one pass of the hot path
over one batch: the cfg-3 workload of BASELINE.json configs[2] -- 50k mixed
events (2-8 isoforms), 2k paired-end reads each with the insert-length model,
5000 iterations (burn-in 500, lag 10, 1 chain) -- generated synthetically inside
the library (miso_b200/csrc/synth.cpp).  With N > 1 every rank owns a full-size
shard of its own (weak scaling; genes are independent, no data-path collective)
and the per-gene posterior summaries are all-gathered once per step over NCCL.

  value : whole-job iterations/s with the packed tiles already resident in HBM
          (misob200_run_resident), CUDA-event time, max over ranks
  e2e   : the same metric through the public C entry point misob200_run with
          host buffers: pinned tiles H2D + kernels + D2H of posteriors + the
          host epilogue, every step, plus the summary all-gather
  roofline / cpu_baseline / clocks: see DESIGN.md "measurement"

`--impl reference` times the reference's own CPU implementation (oracle/_ref,
the unmodified C core; falls back to the plain-C port) on all host cores on a
bounded sample of the same workload.
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
    # name: (kind, n_genes, reads_per_gene, paired)
    "cfg3": dict(kind=1, n_genes=50000, reads=2000, label="cfg-3: 50k mixed events (K 2-8), 2k PE reads each, "
                 "insert N(250,30^2) +-4sd, read_len 36"),
    "cfg2": dict(kind=0, n_genes=10000, reads=1000, label="cfg-2: 10k 2-isoform SE events, 1k SE reads each, read_len 36"),
}
ITERS, BURN, LAG, CHAINS = 5000, 500, 10, 1
PE = (250.0, 900.0, 4.0)
READ_LEN = 36
SEED = 20260925


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
# CPU arm (reference implementation on host cores)

def _cpu_worker(args):
    """One process: run the oracle on a slice of genes, return (iterations, seconds)."""
    kind_ref, genes, paired, mt = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refdriver
    orc = refdriver.RefOracle() if kind_ref == "reference" else refdriver.PortOracle()
    t0 = time.perf_counter()
    n = 0
    for (ex, isos, pos, cig, gid) in genes:
        kw = dict(iters=ITERS, burn=BURN, lag=LAG, chains=CHAINS, seed=SEED, gene_id=gid,
                  rng_mode=1 if (mt and kind_ref == "reference") else 0)
        if paired:
            orc.miso_pe(ex, isos, pos, cig, READ_LEN, PE[0], PE[1], PE[2], **kw)
        else:
            orc.miso_se(ex, isos, pos, cig, READ_LEN, **kw)
        n += ITERS * CHAINS
    return n, time.perf_counter() - t0


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


def cpu_arm(wl, genes_per_core, cores=None):
    """Reference C path on `cores` processes over the first genes of the workload."""
    import concurrent.futures as cf
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refdriver
    import miso_b200 as mb
    kind_ref = "reference" if refdriver.available() else "port"
    if kind_ref == "port" and not refdriver.port_available():
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    cores = cores or usable_cores()
    n = cores * genes_per_core
    w = mb.Workload(wl["kind"], n, wl["reads"], READ_LEN, PE[0], PE[1], PE[2], seed=SEED, first_gene_id=0)
    slices = [[] for _ in range(cores)]
    for g in range(n):
        ex, isos, pos, cig = w.gene(g)
        slices[g % cores].append((ex, isos, pos, cig, g))
    w.close()
    paired = wl["kind"] == 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with cf.ProcessPoolExecutor(max_workers=cores, mp_context=ctx) as ex:
        res = list(ex.map(_cpu_worker, [(kind_ref, s, paired, True) for s in slices]))
    wall = time.perf_counter() - t0
    busy = max(r[1] for r in res)
    iters = sum(r[0] for r in res)
    return dict(value=iters / busy, unit=UNIT, cores=cores, kind=kind_ref,
                sample="%d genes of the workload (%d per core), %d iterations each, %d worker processes = all "
                       "usable host cores (%d logical CPUs visible, cgroup quota applied) in parallel, %s; busy "
                       "%.1fs wall %.1fs" % (
                           n, genes_per_core, ITERS, cores, os.cpu_count() or 0,
                           "reference's own MT19937 stream" if kind_ref == "reference" else "Philox stream",
                           busy, wall),
                seconds=busy)


# ---------------------------------------------------------------------------

def build_plan(mb, wl, n_genes, first_gene_id, chunk=5000):
    plan = mb.Plan()
    t_gen = t_plan = 0.0
    done = 0
    while done < n_genes:
        n = min(chunk, n_genes - done)
        t0 = time.perf_counter()
        w = mb.Workload(wl["kind"], n, wl["reads"], READ_LEN, PE[0], PE[1], PE[2], seed=SEED,
                        first_gene_id=first_gene_id + done)
        t1 = time.perf_counter()
        plan.append(w)
        t2 = time.perf_counter()
        w.close()
        t_gen += t1 - t0
        t_plan += t2 - t1
        done += n
    return plan, t_gen, t_plan


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
    ap.add_argument("--genes", type=int, default=0, help="override events per GPU (debugging only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-genes-per-core", type=int, default=40)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = dict(WORKLOADS[args.workload])
    if args.genes:
        wl["n_genes"] = args.genes
    S = (ITERS - BURN) // LAG
    config = {"workload": wl["label"] + "; %d iterations, burn-in %d, lag %d, %d chain; %d events per GPU"
              % (ITERS, BURN, LAG, CHAINS, wl["n_genes"]),
              "events_per_gpu": wl["n_genes"], "reads_per_event": wl["reads"],
              "parallelism": "genes sharded over %d GPU(s), one process per GPU" % world,
              "l2_note": "packed inputs (hundreds of MB) exceed the 126 MB L2; no flush needed"}

    # ------------------------------------------------------------------ CPU arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        per_step = max(1, args.cpu_genes_per_core // 8)
        vals = []
        base = None
        for i in range(args.warmup + args.steps):
            base = cpu_arm(wl, per_step)
            if i >= args.warmup:
                vals.append(base)
        v = sum(b["value"] for b in vals) / len(vals)
        ms = 1e3 * sum(b["seconds"] for b in vals) / len(vals)
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": base["cores"], "kind": base["kind"],
                                 "sample": base["sample"]},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # ------------------------------------------------------------------ GPU arm
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_arm(wl, args.cpu_genes_per_core)      # before CUDA is touched in this process

    import numpy as np
    import miso_b200 as mb
    from miso_b200._lib import lib, check, ptr
    import ctypes as C

    if mb.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device -- miso_b200 has no CPU path")
    device = local_rank % mb.device_count()

    if world > 1:
        import torch.distributed as dist     # rendezvous only (unique-id broadcast); no tensors
        dist.init_process_group("gloo", rank=rank, world_size=world)
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

    plan, t_gen, t_plan = build_plan(mb, wl, wl["n_genes"], first_gene_id=rank * wl["n_genes"])
    params = mb.make_params(ITERS, BURN, LAG, CHAINS, seed=SEED, device=device)
    info = plan.info()
    G = info.shape[0]
    ok = int((info[:, 4] == 0).sum())
    iters_per_step = ok * CHAINS * ITERS
    out = plan.alloc_outputs(params, pinned=True)
    summ_all = np.zeros((world * G, 32)) if world > 1 else None

    def e2e_step():
        plan.run(params, out)
        s = plan.summarize()
        if world > 1:
            check(lib.misob200_comm_allgather(ptr(s), s.size, ptr(summ_all)))
        return out["launches"]

    def resident_step():
        return plan.run_resident()

    # warm-up: both paths
    plan.upload(params)
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
        bucket += plan.bucket_timing()
    wall_res = time.perf_counter() - t0
    clk = clocks.stop()
    dev_ms = barrier_max(dev_ms)
    wall_res = barrier_max(wall_res)

    # e2e through misob200_run
    for _ in range(min(args.warmup, 2)):
        e2e_step()
    barrier_max(0.0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    wall_e2e = barrier_max(time.perf_counter() - t0)
    h2d, d2h = plan.transfer_bytes()
    timing = out["timing_ms"].copy()

    # sanity of what was computed (not timed): posterior means vs simulation truth on a few genes
    r0 = plan.gene_result(out, 0)
    assert r0["status"] == 0 and np.isfinite(r0["samples"]).all()

    if rank == 0:
        total_iters = iters_per_step * world
        ms_per_step = dev_ms / args.steps
        value = total_iters / (ms_per_step / 1e3)
        e2e_val = total_iters / (wall_e2e / args.steps)
        alg = algorithmic_bytes(info[info[:, 4] == 0], S)
        peak, how = hbm_peak()
        achieved = alg / (ms_per_step / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "clocks": clk,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * wall_e2e / args.steps,
                    "last_step_ms": {"h2d": timing[0], "kernels": timing[1], "d2h": timing[2]}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic(args.workload, G), "peak_source": how,
                         "algorithmic_bytes_per_step": alg,
                         "kernel": "chain_kernel<K> / quad_kernel<K>, one launch per isoform-count bucket K = 2..8 (7 launches per step)",
                         "bucket_ms_per_step": {str(k): bucket[k] / args.steps for k in range(2, 9) if bucket[k] > 0}},
            "cpu_baseline": ({k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None),
            "setup_seconds": {"synthetic_generation": t_gen, "host_plan_stage": t_plan},
            "wall_ms_per_resident_step": 1e3 * wall_res / args.steps,
        }
        emit(line)
    if world > 1:
        lib.misob200_comm_destroy()
    return 0


if __name__ == "__main__":
    sys.exit(main())
