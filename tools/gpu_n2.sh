#!/bin/bash
# two-GPU checks: the replica lock-step test and the bench line at N = 2
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_parity_sizes.py -m gpu -x -q -k "lock_step" 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-extras > gpurun_out/r2_bench_n2_b1024.json 2> gpurun_out/r2_bench_n2.err; tail -3 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2_b1024.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','n_gpus','replicas_identical','gpu_launches']}); print(d['e2e']['value'], d['config'].get('collective'))
PY
