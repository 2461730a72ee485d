"""Print the headline + per-launch timings of a bench.py JSON line and its per-kernel profile (written next to it by bench.py)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('samples/s', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'step frac', round(d['step_roofline']['frac'], 4), 'e2e', round(d['e2e']['value']))
print(d['kernel_breakdown_ms'])
if len(sys.argv) > 2:
    p = json.load(open('gpurun_out/bench_profile_n1_b1024.json'))
    for n, ms, fl in p['records']:
        if any(n.startswith(k) for k in sys.argv[2].split(',')):
            print(f"{n:55s} {ms:7.3f} ms  {fl/ms/1e9 if ms else 0:7.1f} TF/s")
