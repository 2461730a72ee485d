# Evidence run on the GPU box (gpurun): GPU tests, the bench line, the reference arm, the ncu launch list of bench.py and one
# `ncu --set full` capture per kernel family, summarised on the box (reports are large; only text summaries travel back).
set -x
mkdir -p gpurun_out/ev
python -m pytest tests -m gpu -x -q > gpurun_out/ev/t_final.log 2>&1; tail -3 gpurun_out/ev/t_final.log
python bench.py --steps 20 --warmup 3 > gpurun_out/ev/bench_final.json 2> gpurun_out/ev/bench_final.err; tail -c 400 gpurun_out/ev/bench_final.err
cp gpurun_out/bench_profile_n1_b1024.json gpurun_out/ev/prof_final.json
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/ev/bench_ref.json 2> gpurun_out/ev/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/ev/launches_r1c.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ev/ncu_launch_b.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/ev/launches_r1c.csv gpurun_out/ev/r1c_launches_summary.txt
REPS=""
for spec in "slide_conv_kernel 10 slide_conv" "slide_thin_kernel 0 slide_thin" "slide_wgrad_kernel 2 slide_wgrad" "pw_tc_kernel 0 pw_tc" "pw_wgrad_tc_kernel 0 pw_wgrad_tc" "group_conv_kernel 1 group_conv" "group_wgrad_kernel 0 group_wgrad" "attn_kernel 5 attn_bwd" "thin_conv_kernel 0 thin_conv" "thin_wgrad_kernel 0 thin_wgrad" "join_bwd_kernel 0 join_bwd"; do set -- $spec
  ncu --set full --import-source on --clock-control none -k regex:$1 --launch-skip $2 --launch-count 1 -f -o /tmp/r1c_$3 python tools/profile_step.py --batch 1024 --steps 1 > /tmp/ncu_r1c_$3.log 2>&1; tail -1 /tmp/ncu_r1c_$3.log
  python tools/summarize_ncu.py full /tmp/r1c_$3.ncu-rep gpurun_out/ev/r1c_$3_ncu.txt
  python tools/sass_hist.py /tmp/r1c_$3.ncu-rep > gpurun_out/ev/r1c_$3_sass.txt 2>&1
  REPS="$REPS /tmp/r1c_$3.ncu-rep"
done
# input side: the gather + masking + statistics kernel and the noise/scale kernel at B = 1024 (tools/profile_input.py)
for spec in "window_load_kernel 2 window_load" "noise_scale_kernel 1 noise_scale"; do set -- $spec
  ncu --set full --import-source on --clock-control none -k regex:$1 --launch-skip $2 --launch-count 1 -f -o /tmp/r1c_$3 python tools/profile_input.py > /tmp/ncu_r1c_$3.log 2>&1; tail -1 /tmp/ncu_r1c_$3.log
  python tools/summarize_ncu.py full /tmp/r1c_$3.ncu-rep gpurun_out/ev/r1c_$3_ncu.txt
  REPS="$REPS /tmp/r1c_$3.ncu-rep"
done
python tools/summarize_ncu.py traffic gpurun_out/ev/ncu_traffic.json $REPS
du -sh gpurun_out
