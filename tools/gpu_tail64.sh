#!/bin/bash
# folded BatchNorm finalize at the small-batch configuration (C2, B = 64) and at B = 1024, on / off, two repeats each
mkdir -p gpurun_out
for rep in 1 2; do
for b in 64 1024; do
for off in 0 1; do
  WF_BN_TAIL=$((1-off)) timeout 300 python bench.py --steps 30 --warmup 5 --no-extras --batch $b 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch $b tail_off=$off ms/step', round(d['ms_per_step'],4), 'launches', d['gpu_launches_per_step'])"
done; done; done
