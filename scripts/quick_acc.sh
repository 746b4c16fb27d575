#!/bin/bash
# parity of the accumulate kernels + timing at the bench size (on the GPU box): bash scripts/quick_acc.sh <tag>
TAG=${1:-q}
python -m pytest tests/test_gpu_parity.py -x -q -k "lut_forward or unsorted or skip_garbage or stride or spill or variants or full_fixture" 2>&1 | tail -4
python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || tail -5 gpurun_out/${TAG}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['ms_per_step'], d['kernels_ms'], d['fwd_grad']['ms_per_step'])"
