#!/bin/bash
# per-kernel L1 request shape over one training step: requests, sectors, tag-set accesses of global loads / stores
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-extras > /tmp/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_red.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum \
  --clock-control none --launch-skip 1300 -c 260 --csv --log-file gpurun_out/l1survey.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/l1survey.log 2>&1
tail -3 gpurun_out/l1survey.log; wc -l gpurun_out/l1survey.csv
