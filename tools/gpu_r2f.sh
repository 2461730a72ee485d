#!/bin/bash
mkdir -p gpurun_out
timeout 600 tests/native/slab_selftest 4 > gpurun_out/r2f_selftest.txt 2>&1; tail -2 gpurun_out/r2f_selftest.txt
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.txt 2>&1; tail -5 gpurun_out/r2f_pytest.txt
cat gpurun_out/grad_l2_ratio_b64.txt 2>/dev/null
bash tools/gpu_quick.sh
