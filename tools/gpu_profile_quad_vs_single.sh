#!/bin/bash
# One K bucket under both layouts, ncu --set full, summarised on the box: bash tools/gpu_profile_quad_vs_single.sh [K]
K=${1:-5}
O=gpurun_out
mkdir -p $O
for mode in single quad; do
  if [ $mode = quad ]; then export MISOB200_QUAD_MAX_READS=100000; else export MISOB200_QUAD_MAX_READS=0; fi
  MISOB200_ONLY_K=$K MISOB200_SCHED=serial ncu --set full --clock-control none --import-source on -k regex:"chain_kernel|quad_kernel" -c 1 \
      -o $O/r2_k${K}_$mode python tools/profile_small.py 40000 300 > $O/r2_k${K}_$mode.out 2>&1
  python tools/ncu_summary.py $O/r2_k${K}_$mode.ncu-rep 0 > $O/r2_k${K}_${mode}_summary.txt 2>&1
  python tools/ncu_lines.py $O/r2_k${K}_$mode.ncu-rep 0 $((5638 * 300)) 60 > $O/r2_k${K}_${mode}_lines.txt 2>&1
  rm -f $O/r2_k${K}_$mode.ncu-rep
done
