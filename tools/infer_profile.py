"""Per-launch CUDA-event profile of one eval-mode forward (WF_FLAG_PROFILE), aggregated by kernel family."""
import collections, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wiflow_b200 as wf
from wiflow_b200 import _lib, ops
from oracle import wiflow_oracle as O
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = wf.WiFlowPoseModel(dropout=0.5).to(dev).eval()
inf = wf.InferStep(model, B, use_cuda_graph=False)
x, _ = O.synthetic_batch(B, 6); x = x.to(dev)
for _ in range(2): inf.step(x)
ops.block_forward(x, inf.params, inf.running, inf.nbt, [], [0, 0, 0, 0, 0], _lib.FLAG_PROFILE, inf.ws)
torch.cuda.synchronize()
fam = collections.OrderedDict()
for name, ms, fl in _lib.profile_records():
    a = fam.setdefault(name.split(' ')[0], [0.0, 0.0, 0]); a[0] += ms; a[1] += fl; a[2] += 1
tot = sum(a[0] for a in fam.values())
for k, a in sorted(fam.items(), key=lambda kv: -kv[1][0]):
    print(f'{k:18s} {a[0]:8.3f} ms {100*a[0]/tot:5.1f}%  n={a[2]:3d}  {a[1]/a[0]/1e9 if a[0] else 0:6.1f} TF/s')
print('total', round(tot, 3), 'ms for', B, 'windows ->', round(B / tot * 1e3), 'samples/s')
