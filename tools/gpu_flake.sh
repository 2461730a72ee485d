#!/bin/bash
# repeat the B = 64 step-parity test and print the gradient error ratio of every run (flake hunting)
for mode in default notc; do
  for i in $(seq 1 14); do
    if [ $mode = notc ]; then export WF_DISABLE_TC=1; else unset WF_DISABLE_TC; fi
    r=$(timeout 300 python -m pytest tests/test_gpu_parity_sizes.py -m gpu -x -q -k "b64_vs_oracle" 2>&1 | grep -E "passed|failed|AssertionError" | tr '\n' ' ')
    echo "$mode $i ratio $(cat gpurun_out/grad_l2_ratio_b64.txt) :: $r"
  done
done
