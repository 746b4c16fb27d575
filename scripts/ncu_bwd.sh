#!/bin/bash
# ncu --set full of the backward tile kernels at the bench workload; usage (GPU box): bash scripts/ncu_bwd.sh <tag> [filter]
TAG=${1:-r2bwd}; NSEG=${3:-10000000}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_bwd_tiles -s 0 -c 2 -o gpurun_out/${TAG} -f python scripts/prof_bench_size.py $NSEG bwd > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/${TAG}_raw.csv > gpurun_out/${TAG}_summary.txt 2>&1
NCU_LINES_TOP=100000 python scripts/ncu_lines.py gpurun_out/${TAG}.ncu-rep k_bwd_tiles accumulate_bwd_sorted ${2:-k_bwd_tilesILi4ELi4E} > gpurun_out/${TAG}_lines.txt 2>&1
python scripts/ncu_regions.py gpurun_out/${TAG}_lines.txt larnd-sim-jax_b200/csrc/accumulate_bwd_sorted.cu > gpurun_out/${TAG}_regions.txt 2>&1
