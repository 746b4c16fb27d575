#!/bin/bash
# End-of-round evidence on one B200 (run through gpurun); same steps as final_evidence_r2.sh for the state at the end of the
# third session of round 2.  Outputs land in gpurun_out/; the summaries under profiles/ are copied / derived from them.
set -x
cd ${GRAFT_REPO_ROOT:-.}
tag=${1:-r3_final}
ncu --set full --clock-control none --import-source on -k regex:"k_acc_tiles|k_bwd_tiles|k_fee_forward|k_prepare" -s 4 -c 5 -o gpurun_out/${tag}_kernels -f python scripts/prof_bench_size.py 10000000 steps > gpurun_out/${tag}_kernels_ncu.log 2>&1
ncu -i gpurun_out/${tag}_kernels.ncu-rep --page raw --csv > gpurun_out/${tag}_kernels_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/${tag}_kernels_raw.csv > gpurun_out/${tag}_kernels_ncu_summary.txt 2>&1
python scripts/ncu_traffic.py gpurun_out/${tag}_kernels_raw.csv > gpurun_out/r2_traffic_10M.json
cp gpurun_out/r2_traffic_10M.json profiles/r2_traffic_10M.json
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/${tag}_pytest_gpu.txt; tail -2 gpurun_out/${tag}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.txt 2>&1; tail -1 gpurun_out/${tag}_smoke.txt
python bench.py > gpurun_out/${tag}_bench_10M.json 2> gpurun_out/${tag}_bench_10M.err; tail -c 400 gpurun_out/${tag}_bench_10M.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2>/dev/null; tail -c 300 gpurun_out/${tag}_bench_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_10M.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/b_ncu.log 2>&1
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_sanitizer_racecheck_smoke.txt 2>&1; echo racecheck rc=$?; tail -3 gpurun_out/${tag}_sanitizer_racecheck_smoke.txt
ls -la gpurun_out | tail -8
