#!/bin/bash
mkdir -p gpurun_out
timeout 300 tests/native/slab_selftest 4 > gpurun_out/r2d_selftest.txt 2>&1; tail -2 gpurun_out/r2d_selftest.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.txt 2>&1; tail -5 gpurun_out/r2d_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['step_roofline']['frac'])
print(d['kernel_breakdown_ms'])
print(d['also']['C2_train_b64'], d['also']['C3_infer_b8192'])
PY
