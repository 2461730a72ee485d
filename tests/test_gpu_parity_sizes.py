"""Parity at the batch sizes the benchmark reports (pytest -m gpu).

The tile geometry of every kernel family depends on N = 20*B (column-tile width, split-K, positions per CTA, accumulator
ring), so the B=4 fixtures alone do not pin the configurations that are timed:
  * C2 (BASELINE.json configs[1]): one full training step at B = 64 -- prediction, loss, PCK / MPJPE, every saved activation
    and activation gradient (localises a wrong tile geometry to its layer), all parameter gradients against the fp64 oracle
    with the fp32 oracle as the yardstick, and the weights after clip + AdamW;
  * C4's per-GPU batch: forward at B = 1024 in eval mode and in train mode with the 18 dropout masks on, plus the running
    statistics -- against the fp32 oracle on the host cores.
Reference semantics: train.py:196-237 (step), models/pose_model.py:71-97 (forward)."""
import numpy as np
import pytest
import torch

from oracle import wiflow_oracle as O
from tests.util import is_dead, oracle_key, rel_err, to_internal
from tests.test_gpu_parity import TOL, _run_lib_train, make_model, oracle_state_from

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def wf():
    import wiflow_b200
    return wiflow_b200


def _masks(B, seed):
    torch.manual_seed(seed)
    return O.make_dropout_masks(B, 0.5)


def test_train_step_b64_vs_oracle(wf):
    """BASELINE.json configs[1]: single training step, batch 64, numerics checked against the reference (oracle)"""
    from wiflow_b200 import _lib, ops
    B = 64
    model = make_model(wf, seed=3).train()
    st32, st64 = oracle_state_from(model, torch.float32), oracle_state_from(model, torch.float64)
    x, y = O.synthetic_batch(B, seed=11)
    masks = _masks(B, 5)
    rec = {}
    pred64, loss64, g64 = O.grads(st64, x.double(), y.double(), masks=[m.double() for m in masks], record=rec)
    pred32, loss32, g32 = O.grads(st32, x, y, masks=masks)
    pred, out3, dpred, grads, ws, flags = _run_lib_train(wf, model, x.cuda(), y.cuda(), masks)

    # outputs, loss, metrics
    assert rel_err(pred.cpu(), pred64) < TOL
    np.testing.assert_allclose(out3.cpu().numpy()[:3], np.array(loss64), rtol=1e-4)
    thr = [0.1, 0.2, 0.3, 0.4, 0.5]
    ref_pck = O.pck(pred32, y, thr)
    got_pck = wf.calculate_pck(pred, y.cuda(), thr)
    for k in thr:
        assert round(got_pck[k], 4) == round(ref_pck[k], 4), (k, got_pck[k], ref_pck[k])
    assert round(wf.calculate_mpjpe(pred, y.cuda()), 4) == round(O.mpjpe(pred32, y), 4)

    # every saved activation / activation gradient: a wrong tile geometry shows up in its own layer
    dbg = _lib.debug_tensors(_lib.BlockDesc(0, 0, 0, 0, 0), B, flags)
    report = []
    for name, (off, C, P) in dbg.items():
        if name.endswith('.coef') or off < 0 or (name.endswith('downsample.0.dy') and not name.startswith('tcn.')):
            continue
        key, want_grad = oracle_key(name)
        if key not in rec:
            continue
        t = rec[key].grad if want_grad else rec[key]
        ref = to_internal(key, t.detach(), B).contiguous()
        got = ws[off:off + C * P * B * 20 * 4].view(torch.float32).view(C, P, B, 20).cpu()
        report.append((rel_err(got, ref), name))
    import os
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/intermediates_b64.txt', 'w') as f:
        f.write('\n'.join(f'{e:9.2e}  {n}' for e, n in report) + '\n')
    bad = [(e, n) for e, n in report if e > 2e-3]
    assert len(report) > 100 and not bad, bad

    # gradients: per tensor max(5 x the fp32 oracle's own error, 1e-3 |g|inf); all live parameters together: L2 error <= 3 x the
    # fp32 oracle's (measured 1.4x - 2.5x, tools/grad_accuracy_report.py)
    goff, fails, sq_ours, sq_ref = 0, [], 0.0, 0.0
    for n, p in model.named_parameters():
        g = grads[goff:goff + p.numel()].double().cpu().reshape(-1)
        goff += p.numel()
        t64, t32 = g64[n].reshape(-1), g32[n].double().reshape(-1)
        if is_dead(n):
            assert g.abs().max() <= 1e-5 * max(1.0, max(v.abs().max().item() for v in g64.values())), n
            continue
        scale = t64.abs().max().item()
        err_ref, err = (t32 - t64).abs().max().item(), (g - t64).abs().max().item()
        sq_ours += float(((g - t64) ** 2).sum())
        sq_ref += float(((t32 - t64) ** 2).sum())
        if err > max(5 * err_ref, 10 * TOL * scale) + 1e-12:
            fails.append((n, err, err_ref, scale))
    assert not fails, fails
    ratio = sq_ours ** 0.5 / max(sq_ref ** 0.5, 1e-30)
    with open('gpurun_out/grad_l2_ratio_b64.txt', 'w') as f:
        f.write(f'{ratio:.3f}\n')
    assert ratio <= 3.0, ratio

    # clip_grad_norm_(1.0) + AdamW on the flat buffers vs the oracle's restatement driven by the fp64 gradients
    flat, _, _ = model._wf_state()
    m, v = torch.zeros_like(flat), torch.zeros_like(flat)
    state = ops.adam_state('cuda')
    before = {n: p.detach().cpu().clone() for n, p in model.named_parameters()}
    ops.clip_adamw(flat, grads, m, v, state, 1e-4, 0.9, 0.999, 1e-8, 5e-5, 1.0, 1.0)
    torch.cuda.synchronize()
    params = {n: st64[n].clone() for n in g64}
    m64 = {n: torch.zeros_like(t) for n, t in params.items()}
    v64 = {n: torch.zeros_like(t) for n, t in params.items()}
    total = O.clip_adamw_step(params, g64, m64, v64, 1)
    assert abs(state.view(torch.float32)[4].item() - total) / total < 1e-3
    for n, p in model.named_parameters():
        if is_dead(n):
            continue                      # Adam normalises noise-level gradients to +-lr: direction is not defined
        step_ref = params[n] - before[n].double()
        step_got = p.detach().cpu().double() - before[n].double()
        # Adam's update is lr * m / (sqrt(v) + eps): elements whose gradient is far above the rounding noise must move alike.
        # "Far above the noise" is measured against the fp32 oracle's own error on that tensor, not only against the tensor's
        # largest entry: up.downsample.0.weight (Conv2d(1, 8, 1) + BatchNorm: one weight per channel, and BatchNorm cancels its
        # scale) has a gradient of the order of the BatchNorm eps, i.e. ALL of it is rounding noise, and the first Adam step
        # turns noise into +-lr.  (Seen as a 2-in-14 flake of this test before the second condition was added.)
        noise = (g32[n].double() - g64[n]).abs().max()
        big = (g64[n].abs() > 1e-3 * g64[n].abs().max()) & (g64[n].abs() > 50 * noise)
        if big.any():
            assert (step_got[big] - step_ref[big]).abs().max().item() <= 2e-2 * 1e-4 + 1e-9, n


@pytest.mark.parametrize('train', [False, True])
def test_forward_b1024_vs_oracle(wf, train):
    """C4's per-GPU batch (the configuration bench.py times): forward against the fp32 oracle"""
    B = 1024
    model = make_model(wf, seed=1)
    st = oracle_state_from(model)
    g = torch.Generator().manual_seed(7)
    for k in st:                                                  # non-trivial running statistics for the eval pass
        if k.endswith('running_mean'):
            st[k] = torch.randn(st[k].shape, generator=g) * 0.1
        elif k.endswith('running_var'):
            st[k] = torch.rand(st[k].shape, generator=g) + 0.5
    model.load_state_dict(st)
    x, _ = O.synthetic_batch(B, seed=21)
    if not train:
        model.eval()
        with torch.no_grad():
            out = model(x.cuda())
            ref = O.forward(st, x)
        assert rel_err(out.cpu(), ref) < TOL
        return
    from wiflow_b200 import _lib, ops
    model.train()
    masks = _masks(B, 9)
    desc = [0, 0, 0, 0, 0]
    flags = _lib.FLAG_TRAIN
    flat, running, nbt = model._wf_state()
    ws = torch.zeros(ops.workspace_bytes(desc, B, flags), device='cuda', dtype=torch.uint8)
    lm = [m.cuda().reshape(m.shape[0], m.shape[1], -1).squeeze(-1).contiguous() if m.dim() == 4 else m.cuda().contiguous() for m in masks]
    pred = ops.block_forward(x.cuda(), flat, running, nbt, lm, desc, flags, ws)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = O.forward(st, x, train=True, update_buffers=True, masks=masks)
    assert rel_err(pred.cpu(), ref) < TOL
    ref_running = torch.cat([torch.cat([st[k[:-len('weight')] + 'running_mean'], st[k[:-len('weight')] + 'running_var']])
                             for k in st if k.endswith('.weight') and k[:-len('weight')] + 'running_mean' in st])
    assert rel_err(running.cpu(), ref_running) < TOL


def test_replicas_stay_in_lock_step():
    """two ranks (NCCL) that start from the same seed hold bit-identical weights after three data-parallel steps on different
    shards: the all-reduced gradient is the same on every rank and the gradient-norm reduction is deterministic"""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import os, sys, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {root!r})\n"
        "import wiflow_b200 as wf\n"
        "r = int(os.environ['LOCAL_RANK']); torch.cuda.set_device(r)\n"
        "dist.init_process_group('nccl', device_id=torch.device('cuda', r))\n"
        "torch.manual_seed(0); m = wf.WiFlowPoseModel(dropout=0.5).cuda()\n"
        "ts = wf.TrainStep(m, 16, process_group=dist.group.WORLD)\n"
        "for i in range(3):\n"
        "    g = torch.Generator().manual_seed(100 * r + i)\n"
        "    ts.step(torch.randn(16, 540, 20, generator=g).cuda(), torch.rand(16, 15, 2, generator=g).cuda())\n"
        "torch.cuda.synchronize()\n"
        "c = ts.params.view(torch.int32).to(torch.int64).sum().reshape(1)\n"
        "lo, hi = c.clone(), c.clone()\n"
        "dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)\n"
        "assert lo.item() == hi.item(), (lo.item(), hi.item())\n"
        "if r == 0: print('LOCKSTEP OK')\n"
        "dist.destroy_process_group()\n")
    res = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
                          '--master-port', '29577', '--no-python', sys.executable, '-c', code], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and 'LOCKSTEP OK' in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


def test_eval_mode_backward_vs_oracle(wf):
    """model.eval() with grad enabled (frozen-BatchNorm fine-tuning, input saliency): BatchNorm back-propagates as the fixed affine
    of its running statistics, exactly as the reference nn.Module does; conv biases in front of a BatchNorm have real gradients here"""
    B = 4
    model = make_model(wf, seed=2).eval()
    st = oracle_state_from(model, torch.float64)
    g = torch.Generator().manual_seed(3)
    for k in st:
        if k.endswith('running_mean'):
            st[k] = (torch.randn(st[k].shape, generator=g) * 0.1).double()
        elif k.endswith('running_var'):
            st[k] = (torch.rand(st[k].shape, generator=g) + 0.5).double()
    model.load_state_dict({k: (v.float() if v.is_floating_point() else v) for k, v in st.items()})
    x, y = O.synthetic_batch(B, seed=31)
    names = O.param_names(st)
    for n in names:
        st[n].requires_grad_(True)
    x64 = x.double().requires_grad_(True)
    pred64 = O.forward(st, x64, train=False)
    O.pose_loss(pred64, y.double())[0].backward()
    xc = x.cuda().requires_grad_(True)
    out = model(xc)
    loss, _ = wf.PoseLoss()(out, y.cuda())
    loss.backward()
    assert rel_err(out.detach().cpu(), pred64.detach()) < TOL
    assert rel_err(xc.grad.cpu(), x64.grad) < 1e-3
    worst = []
    gmax = max(st[n].grad.abs().max().item() for n in names)
    for n, p in model.named_parameters():
        if n.endswith('bn_similarity.bias'):      # a shift of the logits is a softmax no-op: zero gradient in eval mode too
            assert p.grad.abs().max().item() <= 1e-5 * gmax, n
            continue
        e = rel_err(p.grad.cpu(), st[n].grad)
        if e > 2e-3:
            worst.append((n, e))
    assert not worst, worst
    # running statistics untouched by an eval-mode pass
    for k, v in model.state_dict().items():
        if k.endswith('running_mean') or k.endswith('running_var'):
            assert torch.equal(v.cpu().double(), st[k]) or rel_err(v.cpu(), st[k]) < 1e-7


def test_train_step_b64_through_slab_kernels():
    """the model switches the conv stack to the TMA + tcgen05 slab kernels from 4096 columns (B >= 205) up; this re-runs the B = 64
    step test (activations, activation gradients, parameter gradients, AdamW) and the B = 4 fixture tests with the switch forced on, so
    the slab forward AND backward-data kernels are compared with the oracle layer by layer inside the full model"""
    import os
    import subprocess
    import sys
    if os.environ.get('WF_SLABTC_MIN_N') == '0':
        pytest.skip('already running with the slab kernels forced on')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, WF_SLABTC_MIN_N='0')
    r = subprocess.run([sys.executable, '-m', 'pytest', '-x', '-q', '-m', 'gpu',
                        'tests/test_gpu_parity_sizes.py::test_train_step_b64_vs_oracle',
                        'tests/test_gpu_parity.py::test_train_step_matches_reference_fixture',
                        'tests/test_gpu_parity.py::test_train_intermediates_vs_oracle',
                        'tests/test_gpu_blocks.py'], cwd=root, env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]


def test_train_step_b64_with_folded_bn_finalize():
    """WF_BN_TAIL=1: the BatchNorm finalizes ride on the last CTA of the kernel that completes their sums (forward: every conv kernel
    family; backward: the backward-data kernels and the residual joins).  Off by default (measured slower) but kept
    working: the B = 64 step, the B = 4 fixture and the layer-by-layer comparison must hold with it on"""
    import os
    import subprocess
    import sys
    if os.environ.get('WF_BN_TAIL') == '1':
        pytest.skip('already running with the folded finalize')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, WF_BN_TAIL='1')
    r = subprocess.run([sys.executable, '-m', 'pytest', '-x', '-q', '-m', 'gpu',
                        'tests/test_gpu_parity_sizes.py::test_train_step_b64_vs_oracle',
                        'tests/test_gpu_parity_sizes.py::test_eval_mode_backward_vs_oracle',
                        'tests/test_gpu_parity.py::test_train_step_matches_reference_fixture',
                        'tests/test_gpu_parity.py::test_train_intermediates_vs_oracle'], cwd=root, env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
