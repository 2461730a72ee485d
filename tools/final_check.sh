# Final check of a build on the GPU box: the whole GPU suite, smoke(), the N=1 bench line and a one-line summary of it.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_final.log 2>&1; tail -1 gpurun_out/t_final.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 > gpurun_out/r1c_bench_n1_b1024.json 2> gpurun_out/bench_n1.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r1c_bench_n1_b1024.json'))
a = d['also']
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'clocks', d['clocks'])
print('C2', round(a['C2_train_b64']['samples_per_s']), round(a['C2_train_b64']['ms_per_step'], 3), 'C3', round(a['C3_infer_b8192']['samples_per_s']),
      'cpu', round(d['cpu_baseline']['value']), 'roofline', d['roofline']['kernel'][:28], round(d['roofline']['frac'], 3), 'step frac', round(d['step_roofline']['frac'], 4))
print({k: (round(v['ms'] * 1e3, 1), round(v['frac_of_hbm_peak'], 3)) for k, v in a['input_side'].items() if isinstance(v, dict)})
print(d['kernel_breakdown_ms'])
PY
