"""GPU tests of the training-loop glue (wiflow_b200.train_loop / engine.TrainStep extensions) against the CPU oracle:
gradient accumulation with the reference's loss/k rule, device-side epoch statistics, the ragged last batch, learning-rate changes
between optimizer steps and a short fit() with validation, scheduler and best-state bookkeeping."""
import copy

import pytest
import torch

from oracle import wiflow_oracle as O
from tests.util import is_dead, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def wf():
    import wiflow_b200
    return wiflow_b200


def _model(wf, seed=0):
    torch.manual_seed(seed)
    m = wf.WiFlowPoseModel(dropout=0.5).cuda()
    for mod in m.modules():                      # dropout off: the oracle then needs no mask stream
        if isinstance(mod, (torch.nn.Dropout, torch.nn.Dropout2d)):
            mod.p = 0.0
    return m


def _solid(acc, total_norm):
    """entries whose clipped gradient is far enough above AdamW's eps that the direction of the step is not rounding noise
    (same rule as tests/test_oracle.py::check_post_step)"""
    coef = min(1.0, 1.0 / (total_norm + 1e-6))
    return {n: (g.abs() * coef) > 1e-5 for n, g in acc.items()}


def _check_weights(model, params, solid, nsteps, lr=1e-4):
    """after one step: solid entries to 2e-6.  After several steps the update lr*m/(sqrt(v)+eps) of an entry whose gradient changes
    sign between steps amplifies the fp32 rounding noise of the 3-4-window gradients (tests/tools/calibrate_grad_noise.py), so single
    entries only have to stay within the 2.5*lr-per-step bound, while the typical entry must still agree to 5e-7 and fewer than one
    entry in a thousand may be off by more than 5e-6."""
    tot, cnt, far = 0.0, 0, 0
    for n, p in model.named_parameters():
        if is_dead(n):
            continue
        d = (p.detach().cpu() - params[n]).abs()
        assert d.max().item() < 2.5 * lr * nsteps, n
        if solid[n].any():
            ds = d[solid[n]]
            if nsteps == 1:
                assert ds.max().item() < 2e-6, (n, ds.max().item())
            tot += ds.sum().item(); cnt += ds.numel(); far += int((ds > 5e-6).sum())
    assert cnt > 1_000_000 and tot / cnt < 5e-7 and far < 1e-3 * cnt, (tot / cnt, far, cnt)


def _oracle_state(model):
    return {k: (v.detach().cpu().clone()) for k, v in model.state_dict().items()}


@pytest.mark.parametrize('use_graph', [False, True])
def test_accumulation_matches_oracle(wf, use_graph):
    """two micro-batches of 4 windows, k=2: grads of loss/2 summed, one clip+AdamW step; running stats updated per micro-batch"""
    model = _model(wf)
    st = _oracle_state(model)
    B, k = 4, 2
    ts = wf.TrainStep(model, B, accumulation_steps=k, metric_thresholds=(0.2, 0.5), dropout=False, use_cuda_graph=use_graph)
    batches = [O.synthetic_batch(B, seed=10 + i) for i in range(k)]
    # oracle
    names = O.param_names(st)
    acc = {n: torch.zeros_like(st[n]) for n in names}
    losses, preds = [], []
    for x, y in batches:
        pred, l3, g = O.grads(st, x, y, masks=None, update_buffers=True)
        losses.append(l3); preds.append((pred, y))
        for n in names:
            acc[n] += g[n] / k
    params = {n: st[n] for n in names}
    m = {n: torch.zeros_like(p) for n, p in params.items()}
    v = {n: torch.zeros_like(p) for n, p in params.items()}
    total_norm = O.clip_adamw_step(params, acc, m, v, 1)
    solid = _solid(acc, total_norm)
    # CUDA
    before = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()
    for i, (x, y) in enumerate(batches):
        ts.step(x.cuda(), y.cuda())
        if i == 0:      # no optimizer step yet
            torch.cuda.synchronize()
            assert torch.equal(before, torch.cat([p.detach().reshape(-1) for p in model.parameters()]))
    torch.cuda.synchronize()
    _check_weights(model, params, solid, 1)
    assert abs(ts.out[3].item() - total_norm) < 1e-3 * total_norm        # pre-clip norm of the (loss / k) gradient
    sd = model.state_dict()
    for key in ('tcn.network.0.bn1_group.running_mean', 'decoder.1.running_var', 'attention.width_axis.bn_output.running_var'):
        assert rel_err(sd[key].cpu(), st[key]) < 1e-4, key
    assert int(sd['decoder.1.num_batches_tracked']) == k
    # epoch statistics: window-weighted means of the per-micro-batch values
    s = ts.read_sums()
    assert s['windows'] == B * k
    assert abs(s['loss'] - sum(l[0] for l in losses) / k) < 1e-4 * max(1.0, abs(s['loss']))
    assert abs(s['position'] - sum(l[1] for l in losses) / k) < 1e-4
    assert abs(s['bone'] - sum(l[2] for l in losses) / k) < 1e-4
    mp = sum(O.mpjpe(p, y) for p, y in preds) / k
    pk = sum(O.pck(p, y, (0.2, 0.5))[0.5] for p, y in preds) / k
    assert abs(s['mpjpe'] - mp) < 5e-5 and abs(s['pck@0.5'] - pk) < 5e-5


def test_ragged_batch_flush_and_lr_change(wf):
    """an epoch of 4 + 4 + 3 windows with k=2: the third (ragged) micro-batch is flushed alone with the loss/k rule; then the
    learning rate is halved and the next step must use it"""
    model = _model(wf, seed=1)
    st = _oracle_state(model)
    B, k = 4, 2
    ts = wf.TrainStep(model, B, accumulation_steps=k, dropout=False)
    names = O.param_names(st)
    params = {n: st[n] for n in names}
    m = {n: torch.zeros_like(p) for n, p in params.items()}
    v = {n: torch.zeros_like(p) for n, p in params.items()}
    sizes, lrs = [4, 4, 3, 4, 4], [1e-4, 1e-4, 1e-4, 5e-5, 5e-5]
    batches = [O.synthetic_batch(b, seed=30 + i) for i, b in enumerate(sizes)]
    groups = [[0, 1], [2], [3, 4]]
    solid = None
    for step, grp in enumerate(groups, 1):
        acc = {n: torch.zeros_like(st[n]) for n in names}
        for i in grp:
            _, _, g = O.grads(st, batches[i][0], batches[i][1], masks=None, update_buffers=True)
            for n in names:
                acc[n] += g[n] / k
        s_now = _solid(acc, O.clip_adamw_step(params, acc, m, v, step, lr=lrs[grp[0]]))
        solid = s_now if solid is None else {n: solid[n] & s_now[n] for n in names}
    # CUDA: first epoch (3 micro-batches + flush), lr change, two more
    for i in range(3):
        ts.step(batches[i][0].cuda(), batches[i][1].cuda())
    ts.flush()
    ts.set_lr(5e-5)
    for i in (3, 4):
        ts.step(batches[i][0].cuda(), batches[i][1].cuda())
    torch.cuda.synchronize()
    _check_weights(model, params, solid, 3)


def test_fit_bookkeeping(wf):
    """fit(): history keys of the reference, scheduler driven by the validation MPJPE, best state restored, checkpoint in the
    reference's state_dict format (295 keys)"""
    import os
    import tempfile
    model = _model(wf, seed=2)
    tr = wf.Trainer(model, 8, lr=1e-3, accumulation_steps=1, patience=2)
    train = [O.synthetic_batch(8, seed=50 + i) for i in range(3)] + [O.synthetic_batch(5, seed=59)]
    val = [O.synthetic_batch(8, seed=70), O.synthetic_batch(6, seed=71)]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, 'best.pth')
        hist = tr.fit(lambda: train, lambda: val, n_epochs=3, checkpoint_path=path)
        assert set(hist) == {'train_loss', 'val_loss', 'train_position_loss', 'train_bone_loss', 'train_mpe', 'val_mpe', 'train_pck',
                             'val_pck', 'train_pck50', 'val_pck50', 'lr'}
        n = len(hist['val_mpe'])
        assert 1 <= n <= 3 and all(len(v) == n for v in hist.values())
        assert all(map(lambda x: x == x and x < 1e3, hist['train_loss'] + hist['val_loss']))
        assert min(hist['val_mpe']) == pytest.approx(tr.best_val_mpe)
        saved = torch.load(path)
        assert len(saved) == 295 and set(saved) == set(model.state_dict())
        # the model holds the best state again
        for key, val_t in saved.items():
            assert torch.equal(val_t.cpu(), model.state_dict()[key].cpu()), key
    # validation of the restored model reproduces the best MPJPE (eval mode, running statistics)
    again = tr.validate(val)
    assert again['mpjpe'] == pytest.approx(tr.best_val_mpe, rel=1e-5)
    # against the oracle's eval forward
    st = _oracle_state(model)
    tot = sum(O.mpjpe(O.forward(st, x, train=False), y) * x.shape[0] for x, y in val) / sum(x.shape[0] for x, _ in val)
    assert abs(again['mpjpe'] - tot) < 5e-5


def test_host_batches_take_the_staged_upload_path_and_match_device_batches():
    """TrainStep.step with pinned HOST tensors (upload through two staging slots on a copy stream, engine.py::_upload) gives the same
    steps as with device tensors; the staging slots are reused safely over more calls than there are slots"""
    import wiflow_b200 as wf
    from oracle import wiflow_oracle as O
    dev = torch.device('cuda', 0)
    B = 16
    batches = [O.synthetic_batch(B, 40 + i) for i in range(5)]
    outs = []
    for host in (False, True):
        torch.manual_seed(3)
        model = wf.WiFlowPoseModel(dropout=0.0).to(dev)
        for mod in model.modules():
            if isinstance(mod, torch.nn.Dropout2d):
                mod.p = 0.0
        ts = wf.TrainStep(model, B, dropout=False)
        res = []
        for x, y in batches:
            if host:
                out = ts.step(x.pin_memory(), y.pin_memory())
            else:
                out = ts.step(x.to(dev), y.to(dev))
            res.append(out.clone())
        torch.cuda.synchronize(dev)
        assert (ts._copy_stream is not None) == host
        outs.append((torch.stack(res).cpu(), ts.params.detach().cpu().clone()))
    assert torch.allclose(outs[0][0], outs[1][0], rtol=2e-4, atol=1e-6), (outs[0][0], outs[1][0])
    # weights after 5 AdamW steps: the two runs differ by the summation order of the atomics in the weight-gradient kernels, and Adam
    # turns a rounding-level difference of a near-zero gradient into up to lr = 1e-4 per step for that one weight
    diff = (outs[0][1] - outs[1][1]).abs()
    assert diff.mean().item() <= 1e-6 and diff.max().item() <= 5e-4, (diff.mean().item(), diff.max().item())
