#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.txt 2>&1; tail -8 gpurun_out/r2e_pytest.txt
cat gpurun_out/grad_l2_ratio_b64.txt 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -3 gpurun_out/r2e_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['step_roofline']['frac'])
r=d['roofline']; print({k:r[k] for k in ('kernel','bound','achieved','peak','unit','frac','frac_of_tensor_peak','frac_of_hbm_peak') if k in r})
a=d['also']; print(a['C2_train_b64'], a['C3_infer_b8192'])
print(json.dumps(a.get('torch_eager_gpu'))[:1200])
print(a.get('C1_cpu_infer_b64'))
print(json.dumps(a.get('C5_microbench'))[:2500])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2e_ref.json 2>gpurun_out/r2e_ref.err; cut -c1-600 gpurun_out/r2e_ref.json
