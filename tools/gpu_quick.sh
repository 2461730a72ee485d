#!/bin/bash
# quick A/B: bench line without the extra legs (value, ms/step, C2-less)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; tail -2 gpurun_out/quick_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/quick_bench.json').read().strip().splitlines()[-1])
print('samples/s', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'frac', round(d['step_roofline']['frac'],4), 'launches/step', d['gpu_launches_per_step'])
print(d['kernel_breakdown_ms'])
PY
