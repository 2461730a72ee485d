"""Per-launch CUDA-event profile of one forward+backward at a small batch (where the step is latency bound): family sums and the
slowest launches.  python tools/profile_small.py [B]"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wiflow_b200 as wf
from wiflow_b200 import _lib, ops
from oracle import wiflow_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = wf.WiFlowPoseModel(dropout=0.5).to(dev)
ts = wf.TrainStep(model, B, use_cuda_graph=False)
x, y = O.synthetic_batch(B, 1)
x, y = x.to(dev), y.to(dev)
for _ in range(3):
    ts.step(x, y)
masks = model._wf_masks(B, dev)
pf = ts.flags | _lib.FLAG_PROFILE
for rep in range(2):
    pred = ops.block_forward(x, ts.params, ts.running, ts.nbt, masks, [0, 0, 0, 0, 0], pf, ts.ws)
    out3, dpred = ops.pose_loss(pred, y, 0, 1.0, 0.2, ts.loss_scratch, True)
    ops.block_backward(x, ts.params, masks, dpred, [0, 0, 0, 0, 0], pf, ts.ws, False)
    torch.cuda.synchronize(dev)
    recs = _lib.profile_records()
fam = collections.OrderedDict()
for n, ms, fl in recs:
    a = fam.setdefault(n.split(' ')[0], [0.0, 0])
    a[0] += ms; a[1] += 1
tot = sum(a[0] for a in fam.values())
print(f'B={B}: {len(recs)} launches, {tot:.3f} ms of kernel time')
for k, a in sorted(fam.items(), key=lambda kv: -kv[1][0]):
    print(f'  {k:18s} {a[0] * 1e3:8.1f} us  n={a[1]:3d}  {a[0] * 1e3 / a[1]:6.1f} us/launch')
for n, ms, fl in sorted(recs, key=lambda r: -r[1])[:12]:
    print(f'  {ms * 1e3:7.1f} us  {n}')
