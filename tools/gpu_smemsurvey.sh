#!/bin/bash
# per-kernel shared-memory conflict share, issue-slot use and tensor-pipe activity over one training step
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-extras > /tmp/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none --launch-skip 1300 -c 260 --csv --log-file gpurun_out/smemsurvey.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/smemsurvey.log 2>&1
wc -l gpurun_out/smemsurvey.csv
