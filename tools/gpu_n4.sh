#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 3 --no-extras > gpurun_out/r2_bench_n4_b1024.json 2> gpurun_out/r2_bench_n4.err; tail -2 gpurun_out/r2_bench_n4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n4_b1024.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','n_gpus','replicas_identical']}); print(d['e2e']['value'])
PY
