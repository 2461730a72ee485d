#!/bin/bash
mkdir -p gpurun_out
timeout 300 tests/native/slab_selftest bench > gpurun_out/r2c_plain.txt 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slab_tc_kernel -s 1 -c 1 -o gpurun_out/r2c_dbg0 tests/native/slab_selftest bench > gpurun_out/r2c_ncu0.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slab_tc_kernel -s 38 -c 1 -o gpurun_out/r2c_dbg55 tests/native/slab_selftest bench > gpurun_out/r2c_ncu55.log 2>&1
tail -2 gpurun_out/r2c_ncu55.log
