#!/bin/bash
mkdir -p gpurun_out
timeout 300 tests/native/slab_selftest wgrad 2>&1 | tail -12
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/gpu_quick.sh
