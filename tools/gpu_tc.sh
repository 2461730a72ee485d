#!/bin/bash
mkdir -p gpurun_out
timeout 300 tests/native/tc_selftest 8 | tail -1
for i in 1 2; do echo "--- old"; tests/native/tc_selftest_old p 2>&1 | grep "perf conv"; echo "--- new"; tests/native/tc_selftest p 2>&1 | grep "perf conv"; done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do bash tools/gpu_quick.sh 2>&1 | head -1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/quick_bench.json').read().strip().splitlines()[-1]); k=d['kernel_breakdown_ms']; print({x:k[x] for x in ['tc_fwd','tc_dgrad','tc_wgrad']})
PY
done
