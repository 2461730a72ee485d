#!/bin/bash
mkdir -p gpurun_out
timeout 300 tests/native/slab_selftest 4 > gpurun_out/r2b_selftest.txt 2>&1; tail -4 gpurun_out/r2b_selftest.txt
timeout 600 tests/native/slab_selftest bench > gpurun_out/r2b_slab_bench.txt 2>&1; cat gpurun_out/r2b_slab_bench.txt
