#!/bin/bash
# Round-2 ncu evidence, run on the GPU box (one GPU).  Summaries are made THERE (gpurun brings
# back at most 64 MiB) and the reports are deleted afterwards.
set -u
O=gpurun_out
mkdir -p $O
# 1. launch list of a bench run (serialised, cold cache: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_v2_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-writer > $O/r2_v2_launches_bench.out 2>&1
# 2. full set on the chain kernels, each bucket alone on the whole machine (serial policy), one full wave per K
MISOB200_SCHED=serial ncu --set full --clock-control none --import-source on -k regex:"chain_kernel|quad_kernel" -c 7 \
    -o $O/r2_chain_full python tools/profile_small.py 40000 300 > $O/r2_chain_full.out 2>&1
python tools/ncu_summary.py $O/r2_chain_full.ncu-rep > $O/r2_v2_ncu_full_summary.txt 2>&1
# genes per K of `profile_small.py 40000` (seed 1): K=8 5684, K=5 5638, K=2 5746 (launch order K = 8 .. 2)
python tools/ncu_lines.py $O/r2_chain_full.ncu-rep 0 $((5684 * 300)) 45 > $O/r2_v2_single_K8_lines.txt 2>&1
python tools/ncu_lines.py $O/r2_chain_full.ncu-rep 3 $((5638 * 300)) 45 > $O/r2_v2_single_K5_lines.txt 2>&1
python tools/ncu_lines.py $O/r2_chain_full.ncu-rep 6 $((5746 * 300)) 45 > $O/r2_v2_quad_K2_lines.txt 2>&1
rm -f $O/r2_chain_full.ncu-rep
# 3. the auxiliary kernels
ncu --set full --clock-control none --import-source on -k regex:"summary_kernel|compare_kernel|match_kernel" -c 3 \
    -o $O/r2_aux_full python tools/profile_aux.py 20000 > $O/r2_aux_full.out 2>&1
python tools/ncu_summary.py $O/r2_aux_full.ncu-rep > $O/r2_v2_aux_kernels_ncu_summary.txt 2>&1
ncu -i $O/r2_aux_full.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
keys=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','l1tex__t_bytes.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print({k:r[h.index(k)] for k in keys if k in h})
" >> $O/r2_v2_aux_kernels_ncu_summary.txt 2>&1
rm -f $O/r2_aux_full.ncu-rep
# 4. DRAM traffic of a full-size step (-> profiles/traffic.json)
MISOB200_SCHED=serial ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/r2_v2_dram_traffic_full_cfg3.csv python tools/profile_small.py 50000 5000 > $O/r2_dram_traffic.out 2>&1
ls -la $O
