#!/bin/bash
mkdir -p gpurun_out
timeout 600 tests/native/slab_selftest 4 > gpurun_out/r2_selftest.txt 2>&1; tail -2 gpurun_out/r2_selftest.txt
timeout 600 tests/native/slab_selftest bench 2>&1 | grep -v timeline | tail -12
