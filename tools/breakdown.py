"""Per-layer-class breakdown of gpurun_out/bench_profile_n1_b1024.json (written by bench.py)."""
import collections, json, sys
d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/bench_profile_n1_b1024.json'))
def kind(n):
    fam, _, layer = n.partition(' ')
    if 'tcn' in layer:
        if '_pw' in layer or 'downsample' in layer: return fam + ' tcn_pw'
        if '_group' in layer: return fam + ' tcn_group'
        return fam + ' tcn'
    if layer.startswith('up.') or 'residual_blocks.0' in layer or 'residual_blocks.1' in layer: return fam + ' cv_thin'
    if 'residual_blocks' in layer: return fam + ' cv_mid'
    if 'qkv' in layer: return fam + ' qkv'
    if 'decoder' in layer: return fam + ' decoder'
    return fam
cls = collections.OrderedDict()
for n, ms, fl in d['records']:
    a = cls.setdefault(kind(n), [0, 0, 0]); a[0] += ms; a[1] += fl; a[2] += 1
tot = sum(a[0] for a in cls.values())
for k, a in sorted(cls.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:26s} {a[0]:7.3f} ms {100*a[0]/tot:5.1f}%  n={a[2]:3d}  {a[1]/1e9:7.1f} GF  {a[1]/a[0]/1e9 if a[0] else 0:6.1f} TF/s")
print('total', round(tot, 3))
