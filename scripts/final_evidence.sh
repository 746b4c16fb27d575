#!/bin/bash
# End-of-round evidence on one B200 (run through gpurun): GPU test suite, smoke, bench lines (ours + reference arm), the ncu
# launch list of the bench command and one `ncu --set full` capture per variant of the dominant kernels.  Outputs land in
# gpurun_out/; the summaries under profiles/ are made from them with scripts/ncu_summary.py, ncu_traffic.py, ncu_lines.py.
set -x
cd ${GRAFT_REPO_ROOT:-.}
tag=${1:-r1e}
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt; tail -3 gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -1 gpurun_out/smoke.txt
python bench.py > gpurun_out/${tag}_bench_10M.json 2> gpurun_out/${tag}_bench_10M.err; tail -c 600 gpurun_out/${tag}_bench_10M.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_10M.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
if [ "$2" = "full" ]; then
  ncu --set full --clock-control none --import-source on -k regex:k_acc_tiles -c 2 -f -o gpurun_out/acc_${tag} python scripts/prof_bench_size.py 10000000 > gpurun_out/ncu_acc.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_bwd_tiles -c 2 -f -o gpurun_out/bwd_${tag} python scripts/prof_bench_size.py 10000000 bwd > gpurun_out/ncu_bwd.log 2>&1
fi
ls -la gpurun_out | tail -12
