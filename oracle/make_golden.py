"""Generate tests/golden/wiflow_golden_b4.npz by running the UNMODIFIED reference here.

TEST INFRASTRUCTURE.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py

The reference has no golden vectors of its own; these fixtures are outputs
of the reference's own modules (fp32, and fp64 for gradient truth) on seeded inputs, so the
oracle and the CUDA path can be pinned on the GPU box where the reference is absent.
Weights are not stored (8.9 MB of noise); they are re-drawn from the seed through
`wiflow_oracle.make_state` and verified by the per-tensor checksums stored here.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wiflow_oracle as O          # noqa: E402
from oracle import load_reference as L         # noqa: E402

SEED, MASK_SEED, B, P_TCN = 0, 123, 4, 0.5
SAMPLE_STRIDE = 997


def sample(t):
    t = t.detach().reshape(-1)
    return t.clone() if t.numel() <= 1024 else t[::SAMPLE_STRIDE].clone()


def run(R, dtype, masks_on):
    torch.manual_seed(SEED)
    model = R.WiFlowPoseModel(dropout=P_TCN).to(dtype)
    x, y = O.synthetic_batch(B, SEED, dtype)
    out = {}
    model.eval()
    with torch.no_grad():
        out['eval_pred'] = model(x)
    model.train()
    if masks_on:
        torch.manual_seed(MASK_SEED)
    else:
        for m in model.modules():
            if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout2d)):
                m.p = 0.0
    pred = model(x)
    crit = R.PoseLoss()
    loss, ld = crit(pred, y)
    loss.backward()
    out['train_pred'] = pred.detach()
    out['loss'] = torch.tensor([loss.item(), ld['position'], ld['bone']], dtype=torch.float64)
    out['pck'] = torch.tensor(list(R.calculate_pck(pred.detach(), y, [0.1, 0.2, 0.3, 0.4, 0.5]).values()), dtype=torch.float64)
    out['pck_shoulder'] = torch.tensor(list(R.calculate_pck(pred.detach(), y, [0.2, 0.5], use_torso_norm=False).values()), dtype=torch.float64)
    out['mpjpe'] = torch.tensor([R.calculate_mpjpe(pred.detach(), y)], dtype=torch.float64)
    names = [n for n, _ in model.named_parameters()]
    out['grad_norm'] = torch.stack([p.grad.double().norm() for p in model.parameters()])
    out['grad_absmax'] = torch.stack([p.grad.double().abs().max() for p in model.parameters()])
    out['grad_samples'] = torch.cat([sample(p.grad) for p in model.parameters()])
    bufs = [b for n, b in model.named_buffers() if 'running' in n]
    out['running'] = torch.cat([b.detach().reshape(-1) for b in bufs])
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=5e-5, betas=(0.9, 0.999))
    out['total_norm'] = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0).double().reshape(1)
    opt.step()
    out['post_step_samples'] = torch.cat([sample(p) for p in model.parameters()])
    return names, out


def main():
    R = L.load()
    res = {}
    st = O.make_state(SEED)
    pn = O.param_names(st)
    res['param_sum'] = np.array([st[n].double().sum().item() for n in pn])
    res['param_abssum'] = np.array([st[n].double().abs().sum().item() for n in pn])
    x, y = O.synthetic_batch(B, SEED)
    res['x'] = x.numpy()
    res['y'] = y.numpy()
    torch.manual_seed(MASK_SEED)
    masks = O.make_dropout_masks(B, P_TCN)
    res['mask_bits'] = np.packbits(np.concatenate([(m.reshape(-1) > 0).numpy() for m in masks]))
    res['mask_scales'] = np.array([m.max().item() for m in masks])
    for tag, dtype, masks_on in (('f32', torch.float32, True), ('f64', torch.float64, True),
                                 ('nodrop_f32', torch.float32, False), ('nodrop_f64', torch.float64, False)):
        names, out = run(R, dtype, masks_on)
        assert names == pn
        for k, v in out.items():
            res[f'{tag}.{k}'] = v.numpy()
    res['meta'] = np.array([SEED, MASK_SEED, B, SAMPLE_STRIDE], dtype=np.int64)
    res['p_tcn'] = np.array([P_TCN])
    path = os.path.join(ROOT, 'tests', 'golden', 'wiflow_golden_b4.npz')
    np.savez_compressed(path, **res)
    print('wrote', path, os.path.getsize(path), 'bytes')
    # SURVEY Appendix D anchors (B=64, default ctor) as a second, independent fixture
    torch.manual_seed(0)
    m = R.WiFlowPoseModel()
    xa, ya = torch.randn(64, 540, 20), torch.rand(64, 15, 2)
    m.eval()
    with torch.no_grad():
        o = m(xa)
    loss, ld = R.PoseLoss()(o, ya)
    anchors = dict(out_sum=o.sum().item(), out00=o[0, 0].numpy(), loss=np.array([loss.item(), ld['position'], ld['bone']]),
                   pck=np.array(list(R.calculate_pck(o, ya, [0.1, 0.2, 0.3, 0.4, 0.5]).values())),
                   mpjpe=R.calculate_mpjpe(o, ya), pred=o.numpy())
    path = os.path.join(ROOT, 'tests', 'golden', 'wiflow_anchor_b64.npz')
    np.savez_compressed(path, **anchors)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
