#!/bin/bash
mkdir -p gpurun_out
timeout 300 tests/native/tc_selftest 8 > gpurun_out/r2_tc_selftest.txt 2>&1; tail -12 gpurun_out/r2_tc_selftest.txt
bash tools/gpu_quick.sh
