import sys, os, torch
sys.path.insert(0, '/root/repo')
import wiflow_b200 as wf
from oracle import wiflow_oracle as O
dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = wf.WiFlowPoseModel(dropout=0.5).to(dev)
inf = wf.InferStep(model, 8192)
x, _ = O.synthetic_batch(8192, 6); x = x.to(dev)
for i in range(3): inf.step(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(5): inf.step(x)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(os.environ.get('WF_EVAL_CHUNK', '1024'), 'ms', round(ms, 3), 'samples/s', round(8192 / ms * 1e3), 'frac', round(8192 / ms * 1e3 * 154.81e6 / 74.45e12, 4))
