#!/bin/bash
# Compare the forward accumulate variants: LARND_ACC_KP4 = 0 (one kernel, 6 positions in registers), 1 / 2 (small-span tiles on
# the 3-CTAs-per-SM kernel with two / one window in flight).  scripts/cmp_kp4.sh "<modes>" "<segment counts>"
for n in ${2:-10000000}; do
  for m in ${1:-0 1 2}; do
    LARND_ACC_KP4=$m python bench.py --no-cpu-baseline --segments $n > gpurun_out/bench_kp4_${m}_$n.json 2> gpurun_out/bench_kp4_${m}_$n.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_kp4_${m}_$n.json").read().strip().splitlines()[-1])
    print("kp4=$m", $n, "ms/step %.3f" % d["ms_per_step"], "fwd+grad %.3f" % d["fwd_grad"]["ms_per_step"], {k: round(v, 3) for k, v in d["kernels_ms"].items()}, "hits", d["config"]["hits"])
except Exception as e:
    print("kp4=$m", $n, "FAILED", e)
    print(open("gpurun_out/bench_kp4_${m}_$n.err").read()[-1500:])
PY
  done
done
