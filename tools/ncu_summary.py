"""Summarise an .ncu-rep (read here, no GPU): per-kernel key metrics + opcode mix + stall reasons."""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
idx = {k: hdr.index(k) for k in keys if k in hdr}
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print("=" * 100)
    for k, i in idx.items():
        print("%-70s %s %s" % (k, r[i], units[i]))
    st = sorted(((float(r[i].replace(",", "")), h) for i, h in stall if r[i]), reverse=True)[:8]
    for v, h in st:
        print("   stall %-60s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
if len(sys.argv) > 2:
    n = int(sys.argv[2])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(n), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    print("#" * 100); print(rows[0][:2])
    hdr = rows[1]; iS = hdr.index("Source"); iE = hdr.index("Instructions Executed"); iSm = hdr.index("# Samples")
    byop, samp, tot = collections.Counter(), collections.Counter(), 0
    for r in rows[2:]:
        try: e = int(r[iE])
        except Exception: continue
        t = r[iS].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        byop[op] += e; tot += e
        try: samp[op] += int(r[iSm])
        except Exception: pass
    ts = sum(samp.values())
    print("total warp-instructions", tot, "static SASS", len(rows) - 2)
    for op, c in byop.most_common(24):
        print("%-10s %14d %5.1f%%   stall-samples %5.1f%%" % (op, c, 100 * c / tot, 100 * samp[op] / max(ts, 1)))
