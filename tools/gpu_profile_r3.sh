#!/bin/bash
# ncu evidence of the round-2 final kernels (log-domain proposal scores, gather / fold reductions),
# run on the GPU box (one GPU).  Summaries are made THERE (gpurun brings back at most 64 MiB).
set -u
O=gpurun_out
mkdir -p $O
# 1. launch list of a bench run (serialised, cold cache: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r3_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-writer > $O/r3_launches_bench.out 2>&1
# 2. full set on the chain kernels, each bucket alone on the whole machine (serial policy)
MISOB200_SCHED=serial ncu --set full --clock-control none --import-source on -k regex:"chain_kernel|quad_kernel" -c 7 \
    -o $O/r3_chain_full python tools/profile_small.py 40000 300 > $O/r3_chain_full.out 2>&1
python tools/ncu_summary.py $O/r3_chain_full.ncu-rep > $O/r3_ncu_full_summary.txt 2>&1
# genes per K of `profile_small.py 40000` (seed 1): K=8 5684, K=5 5638, K=2 5746 (launch order K = 8 .. 2)
python tools/ncu_lines.py $O/r3_chain_full.ncu-rep 0 $((5684 * 300)) 45 > $O/r3_single_K8_lines.txt 2>&1
python tools/ncu_lines.py $O/r3_chain_full.ncu-rep 3 $((5638 * 300)) 45 > $O/r3_K5_lines.txt 2>&1
python tools/ncu_lines.py $O/r3_chain_full.ncu-rep 6 $((5746 * 300)) 45 > $O/r3_quad_K2_lines.txt 2>&1
rm -f $O/r3_chain_full.ncu-rep
ls -la $O | tail -12
