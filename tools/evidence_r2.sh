# Round-2 evidence run on the GPU box (gpurun): GPU tests, the bench line, the reference arm, the ncu launch list of bench.py and one
# `ncu --set full` capture of the new TMA + tcgen05 kernels (conv forward / backward-data / backward-weights at the 64-channel layer
# shapes of residual_blocks.3, B = 1024, through tests/native/slab_selftest bench <case>), summarised on the box.
set -x
mkdir -p gpurun_out/ev
python -m pytest tests -m gpu -x -q > gpurun_out/ev/r2_pytest.log 2>&1; tail -3 gpurun_out/ev/r2_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/ev/r2_bench_n1_b1024.json 2> gpurun_out/ev/r2_bench.err; tail -c 300 gpurun_out/ev/r2_bench.err
cp gpurun_out/bench_profile_n1_b1024.json gpurun_out/ev/r2_per_launch_profile_n1_b1024.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/ev/r2_bench_reference_arm.json 2> gpurun_out/ev/r2_bench_ref.err
python bench.py --steps 2 --warmup 3 --no-extras > /tmp/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/ev/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ev/ncu_launch.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/ev/r2_launches.csv gpurun_out/ev/r2_launches_summary.txt
REPS=""
for spec in "slab_tc_kernel 7 slab_fwd" "slab_tc_kernel 8 slab_dgrad" "slab_wgrad_kernel 13 slab_wgrad"; do set -- $spec
  tests/native/slab_selftest bench $2 > /tmp/plain_$3.log 2>&1 &&
  ncu --set full --import-source on --clock-control none -k regex:$1 --launch-skip 0 --launch-count 1 -f -o /tmp/r2_$3 tests/native/slab_selftest bench $2 > /tmp/ncu_r2_$3.log 2>&1; tail -1 /tmp/ncu_r2_$3.log
  python tools/summarize_ncu.py full /tmp/r2_$3.ncu-rep gpurun_out/ev/r2_$3_ncu.txt
  python tools/sass_hist.py /tmp/r2_$3.ncu-rep > gpurun_out/ev/r2_$3_sass.txt 2>&1
  REPS="$REPS /tmp/r2_$3.ncu-rep"
done
# the pointwise tcgen05 kernels on the 540 -> 540 TCN layer at B = 1024 (tests/native/tc_selftest p: 3rd launch of each = warm)
for spec in "pw_tc_kernel pw_tc" "pw_wgrad_tc_kernel pw_wgrad_tc"; do set -- $spec
  ncu --set full --import-source on --clock-control none -k regex:$1 --launch-skip 2 --launch-count 1 -f -o /tmp/r2_$2 tests/native/tc_selftest p > /tmp/ncu_r2_$2.log 2>&1; tail -1 /tmp/ncu_r2_$2.log
  python tools/summarize_ncu.py full /tmp/r2_$2.ncu-rep gpurun_out/ev/r2_$2_ncu.txt
  python tools/sass_hist.py /tmp/r2_$2.ncu-rep > gpurun_out/ev/r2_$2_sass.txt 2>&1
  REPS="$REPS /tmp/r2_$2.ncu-rep"
done
python tools/summarize_ncu.py traffic gpurun_out/ev/r2_ncu_traffic.json $REPS
# ablations of pw_tc_kernel on the same layer (WF_TC_DBG bits: 1 no activation loads, 2 no MMAs, 4 no weight copies, 8 no operand
# stores, 16 main product only, 32 no epilogue memory traffic, 64 no epilogue)
( for d in 0 1 2 4 8 9 15 32 47 79; do echo "WF_TC_DBG=$d"; WF_TC_DBG=$d tests/native/tc_selftest p 2>&1 | grep "perf conv" | head -2; done ) > gpurun_out/ev/r2_pw_tc_ablation.txt 2>&1
tests/native/slab_selftest bench > gpurun_out/ev/r2_slab_microbench.txt 2>&1
cuobjdump -sass wiflow-*/libwiflow_b200.so | grep -oE "UTCHMMA|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|UTCBAR|HMMA\.[0-9A-Z.]+" | sort | uniq -c > gpurun_out/ev/r2_sass_mnemonics.txt; cat gpurun_out/ev/r2_sass_mnemonics.txt
du -sh gpurun_out
