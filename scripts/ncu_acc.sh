#!/bin/bash
# ncu --set full of the forward tile kernels at the bench workload (first launches after the sizing pass), summary + line view
# usage (on the GPU box): bash scripts/ncu_acc.sh <tag> [kernel regex] [nseg]
TAG=${1:-r2}; KRE=${2:-k_acc_tiles}; NSEG=${3:-10000000}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 2 -c 2 -o gpurun_out/${TAG} -f python scripts/prof_bench_size.py $NSEG > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/${TAG}_raw.csv > gpurun_out/${TAG}_summary.txt 2>&1
NCU_LINES_TOP=100000 python scripts/ncu_lines.py gpurun_out/${TAG}.ncu-rep $KRE ${4:-accumulate_sorted} ${5:-k_acc_tilesILi1ELi4E} > gpurun_out/${TAG}_lines.txt 2>&1
python scripts/ncu_regions.py gpurun_out/${TAG}_lines.txt larnd-sim-jax_b200/csrc/${4:-accumulate_sorted}.cu > gpurun_out/${TAG}_regions.txt 2>&1
