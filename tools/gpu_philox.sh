#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "philox" 2>&1 | tail -5
for r in philox torch philox torch; do
timeout 300 python bench.py --steps 30 --warmup 5 --no-extras --dropout-rng $r 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$r ms/step', round(d['ms_per_step'],4), 'samples/s', round(d['value']), 'launches', d['gpu_launches_per_step'], 'loss', d['final_loss'])"
done
