"""Per-source-line view of one kernel launch of an .ncu-rep (read here, no GPU).

  python tools/ncu_lines.py report.ncu-rep LAUNCH_INDEX GENE_ITERATIONS [top]

Needs a capture taken with --import-source on from a -lineinfo build.  Prints, per source file
and for the heaviest source lines, warp-instructions per gene-chain iteration
(GENE_ITERATIONS = gene-chains of the launch x iterations, e.g. 2286 * 200) and the share of
warp-stall samples, i.e. where the instructions are and where the time goes."""
import collections, csv, io, subprocess, sys
rep, idx, n_iter = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, agg, smp, fn = None, {}, {}, None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        fn = fn or r[1]
        continue
    if len(r) < 8 or r[0] in ("Line No", ""):
        continue
    try:
        e, ln, s = int(r[7]), int(r[0]), int(r[6])
    except ValueError:
        continue
    key = (cur, ln, r[1].strip()[:96])
    agg[key] = agg.get(key, 0) + e
    smp[key] = smp.get(key, 0) + s
ts, te = sum(smp.values()) or 1, sum(agg.values()) or 1
print(fn)
print("warp-instructions with source attribution per gene-chain iteration: %.0f" % (te / n_iter))
byf_s, byf_e = collections.Counter(), collections.Counter()
for k in agg:
    byf_s[k[0]] += smp[k]
    byf_e[k[0]] += agg[k]
for f, e in byf_e.most_common():
    print("  %-28s %7.1f instr/iter  %5.1f%% of instr  %5.1f%% of stall samples" % (f, e / n_iter, 100 * e / te, 100 * byf_s[f] / ts))
print("heaviest lines (by stall samples):")
for k, v in sorted(smp.items(), key=lambda x: -x[1])[:top]:
    print("  %7.1f instr/iter %5.1f%% smp  %-18s %4d  %s" % (agg[k] / n_iter, 100 * v / ts, k[0], k[1], k[2]))
