#!/bin/bash
mkdir -p gpurun_out
timeout 300 tests/native/tc_selftest 8 > gpurun_out/r2_tc_selftest.txt 2>&1; tail -17 gpurun_out/r2_tc_selftest.txt
timeout 120 tests/native/tc_selftest p 2>&1 | grep "perf"
bash tools/gpu_quick.sh
