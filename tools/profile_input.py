"""A few launches of the input-side kernels at B = 1024 for ncu (gather out of a resident array + time masking + statistics, then
noise + scaling): python tools/profile_input.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wiflow_b200 import ops
from wiflow_b200.utils import augmentation as A

dev = torch.device('cuda', 0)
B, NWIN = 1024, 8192
torch.manual_seed(0)
resident = torch.randn(NWIN, 540, 20, device=dev)
noise = torch.randn(B, 540, 20, device=dev)
out = torch.empty(B, 540, 20, device=dev)
stats = torch.zeros(2, device=dev, dtype=torch.float64)
spans = A._spans_to_device(A.draw_time_masks(B, 540, 0.3), dev)
for i in range(4):
    idx = torch.randint(0, NWIN, (B,), device=dev)
    stats.zero_()
    ops.window_load(resident, idx, out, spans, stats)
    ops.noise_scale(out, noise, 0.02, 1.05, stats, out)
torch.cuda.synchronize()
print('ok', out.float().mean().item())
