#!/bin/bash
mkdir -p gpurun_out
timeout 600 tests/native/slab_selftest wgrad > gpurun_out/r2_wgrad.txt 2>&1; cat gpurun_out/r2_wgrad.txt | tail -40
