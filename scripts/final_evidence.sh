set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt; tail -3 gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -1 gpurun_out/smoke.txt
python bench.py > gpurun_out/r1c_bench_10M.json 2> gpurun_out/r1c_bench_10M.err; tail -c 600 gpurun_out/r1c_bench_10M.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1c_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_launches_10M.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_acc_tiles -c 2 -f -o gpurun_out/acc_r1d python scripts/prof_bench_size.py 10000000 > gpurun_out/ncu_acc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bwd_tiles -c 2 -f -o gpurun_out/bwd_r1d python scripts/prof_bench_size.py 10000000 bwd > gpurun_out/ncu_bwd.log 2>&1
ls -la gpurun_out | tail -12
