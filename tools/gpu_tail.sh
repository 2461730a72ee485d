#!/bin/bash
# folded BatchNorm finalize (BnTail): parity tests, then step time with and without
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "--- tail on"; WF_BN_TAIL=1 bash tools/gpu_quick.sh
echo "--- tail off"; bash tools/gpu_quick.sh
