#!/bin/bash
# ablations of pw_tc_kernel: WF_TC_DBG bits 1 no activation loads, 2 no MMAs, 4 no weight copies
for d in 15 47 79 32 64; do echo "--- dbg=$d"; WF_TC_DBG=$d timeout 120 tests/native/tc_selftest p 2>&1 | grep "perf conv" | head -3; done
