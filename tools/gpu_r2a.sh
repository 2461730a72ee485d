#!/bin/bash
# round-2 first GPU call: slab kernel self-test (with the descriptor-stride fallback), then the GPU test-suite and a bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 300 tests/native/slab_selftest 8 > gpurun_out/r2a_slab_selftest.txt 2>&1; echo "exit $?" >> gpurun_out/r2a_slab_selftest.txt
if ! grep -q "SLAB SELFTEST PASSED" gpurun_out/r2a_slab_selftest.txt; then
  WF_SLABTC_SWAP_LBO=1 timeout 300 tests/native/slab_selftest 8 > gpurun_out/r2a_slab_selftest_swap.txt 2>&1; echo "exit $?" >> gpurun_out/r2a_slab_selftest_swap.txt
fi
tail -25 gpurun_out/r2a_slab_selftest.txt
if grep -q "SLAB SELFTEST PASSED" gpurun_out/r2a_slab_selftest.txt; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.txt 2>&1; tail -15 gpurun_out/r2a_pytest.txt
  timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json
else
  WF_DISABLE_SLABTC=1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_noslab.txt 2>&1; tail -5 gpurun_out/r2a_pytest_noslab.txt
fi
