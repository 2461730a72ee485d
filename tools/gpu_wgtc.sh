#!/bin/bash
mkdir -p gpurun_out
timeout 300 tests/native/tc_selftest 8 | grep -E "wgrad|SELFTEST"
for i in 1 2; do echo "--- old"; tests/native/tc_selftest_old p 2>&1 | grep "perf wgrad"; echo "--- new"; tests/native/tc_selftest p 2>&1 | grep "perf wgrad"; done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/gpu_quick.sh
