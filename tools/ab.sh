# A/B of one environment knob on the GPU box: tools/ab.sh NAME VALUE1 VALUE2 ... -> samples/s and the kernel families matching $AB_FILTER
name=$1; shift
for v in "$@"; do
  env $name=$v python bench.py --no-extras --steps 20 > gpurun_out/ab_$v.json 2>/dev/null
  python - "$name" "$v" <<'PY'
import json, os, sys
d = json.load(open(f'gpurun_out/ab_{sys.argv[2]}.json'))
k = d['kernel_breakdown_ms']
f = os.environ.get('AB_FILTER', '')
print(f'{sys.argv[1]}={sys.argv[2]}', round(d['value']), 'e2e', round(d['e2e']['value']), {n: k[n] for n in k if f and n.startswith(f)}, 'sum', round(sum(k.values()), 3))
PY
done
