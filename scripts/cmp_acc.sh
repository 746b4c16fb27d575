#!/bin/bash
# Compare the accumulate kernels (LARND_ACC_IMPL = sorted / chunk) on the bench workload: scripts/cmp_acc.sh "<impl list>" "<segment counts>"
# (LARND_SORTED_SPLIT = 0 / 1 / 2 selects how the class-sorted kernels split the tile table over register-footprint variants)
for n in ${2:-10000000}; do
  for impl in ${1:-sorted chunk}; do
    LARND_ACC_IMPL=$impl python bench.py --no-cpu-baseline --segments $n > gpurun_out/bench_${impl}_$n.json 2> gpurun_out/bench_${impl}_$n.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${impl}_$n.json").read().strip().splitlines()[-1])
    print("$impl", $n, "ms/step %.3f" % d["ms_per_step"], "fwd+grad %.3f" % d["fwd_grad"]["ms_per_step"], {k: round(v, 3) for k, v in d["kernels_ms"].items()})
except Exception as e:
    print("$impl", $n, "FAILED", e)
    print(open("gpurun_out/bench_${impl}_$n.err").read()[-1500:])
PY
  done
done
