"""GPU diagnostic (not a test): gradient error of the CUDA path against the fp64 fixture, next to the fp32 reference's own error,
for the dropout and no-dropout fixtures.  python tools/grad_accuracy_report.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wiflow_b200 as wf                                               # noqa: E402
from tests.util import golden_masks, is_dead, load_golden, sample     # noqa: E402
from tests.test_gpu_parity import _run_lib_train, make_model          # noqa: E402

g = load_golden()
for tag, use_masks in (('nodrop_f32', False), ('f32', True)):
    model = make_model(wf).train()
    x, y = torch.from_numpy(g['x']).cuda(), torch.from_numpy(g['y']).cuda()
    masks = golden_masks(g) if use_masks else None
    pred, out3, dpred, grads, ws, flags = _run_lib_train(wf, model, x, y, masks)
    tag64 = tag.replace('f32', 'f64')
    stride = int(g['meta'][3])
    g64, g32 = g[f'{tag64}.grad_samples'], g[f'{tag}.grad_samples']
    off = goff = 0
    rows = []
    for i, (n, p) in enumerate(model.named_parameters()):
        gg = grads[goff:goff + p.numel()]
        goff += p.numel()
        s = sample(gg, stride).double().cpu().numpy()
        t64, t32 = g64[off:off + s.size], g32[off:off + s.size]
        off += s.size
        if is_dead(n):
            continue
        rows.append((float(((s - t64) ** 2).sum()), float(((t32 - t64) ** 2).sum()), float((t64 ** 2).sum()), n))
    so, sr = sum(r[0] for r in rows) ** 0.5, sum(r[1] for r in rows) ** 0.5
    print(f'== {tag}: global L2 error ours {so:.3e}  fp32 reference {sr:.3e}  ratio {so / sr:.2f}')
    fam = {}
    for a, b, c, n in rows:
        k = n.split('.')[0] + ('.' + n.split('.')[1] if n.startswith(('tcn', 'residual', 'attention')) else '')
        f = fam.setdefault(k, [0.0, 0.0])
        f[0] += a; f[1] += b
    for k, (a, b) in sorted(fam.items(), key=lambda kv: -kv[1][0])[:8]:
        print(f'   {k:28s} ours {a ** 0.5:.2e} ref {b ** 0.5:.2e} ratio {(a / max(b, 1e-300)) ** 0.5:.2f}')
    for a, b, c, n in sorted(rows, reverse=True)[:10]:
        print(f'      {n:44s} ours {a ** 0.5:.2e} ref {b ** 0.5:.2e} ratio {(a / max(b, 1e-300)) ** 0.5:5.2f}  |g| {c ** 0.5:.2e}')
