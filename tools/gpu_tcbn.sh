#!/bin/bash
for bn in 256 224 192 160 128 96; do echo "--- bn=$bn"; WF_TC_BN=$bn timeout 120 tests/native/tc_selftest p 2>&1 | grep "perf conv"; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
