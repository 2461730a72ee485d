"""Multi-process path on CPU (gloo, world size 2): the batch sharding and the gradient exchange of the data-parallel step
(engine.shard_bounds / engine.allreduce_gradients are what TrainStep uses with NCCL).  The arithmetic on each rank
is the CPU oracle -- the kernels need a B200 -- so this checks the host logic: every window is owned by exactly one rank, the
all-reduced flat gradient times the returned scale equals the mean of the per-shard gradients (BatchNorm statistics local per
rank, exactly as under the reference's nn.DataParallel, train.py:91-93), and every rank ends the step with identical weights."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import wiflow_oracle as O

WORLD, B = 2, 6


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _shard_grads(rank, world):
    import wiflow_b200 as wf
    st = O.make_state(0)
    x, y = O.synthetic_batch(B, 3)
    b0, b1 = wf.shard_bounds(B, rank, world)
    _, _, g = O.grads(st, x[b0:b1], y[b0:b1])
    names = O.param_names(st)
    return st, names, torch.cat([g[n].reshape(-1) for n in names])


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import wiflow_b200 as wf
        torch.set_num_threads(2)
        st, names, flat = _shard_grads(rank, world)
        local = flat.clone()
        scale = wf.allreduce_gradients(flat, None, world)
        assert scale == 1.0 / world
        # the optimizer step every rank runs on the summed gradient (grad_scale = 1/world inside the fused kernel)
        params = {n: st[n] for n in names}
        g, off = {}, 0
        for n in names:
            k = params[n].numel()
            g[n] = (flat[off:off + k] * scale).view_as(params[n])
            off += k
        m = {n: torch.zeros_like(p) for n, p in params.items()}
        v = {n: torch.zeros_like(p) for n, p in params.items()}
        O.clip_adamw_step(params, g, m, v, 1)
        torch.save({'local': local, 'reduced': flat, 'post': torch.cat([params[n].reshape(-1) for n in names])},
                   os.path.join(out_dir, f'rank{rank}.pt'))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_every_window_once():
    import wiflow_b200 as wf
    for n in (1, 5, 64, 1024, 1027):
        for world in (1, 2, 3, 8):
            spans = [wf.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_gloo_world2_gradient_exchange(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(WORLD, port, str(tmp_path)), nprocs=WORLD, join=True)
    r = [torch.load(os.path.join(tmp_path, f'rank{i}.pt')) for i in range(WORLD)]
    # all ranks hold the same reduced gradient = sum of the per-shard gradients
    assert torch.equal(r[0]['reduced'], r[1]['reduced'])
    want = r[0]['local'] + r[1]['local']
    assert torch.allclose(r[0]['reduced'], want, rtol=0, atol=0)
    # single-process emulation: oracle on each shard, average
    emu = sum(_shard_grads(i, WORLD)[2] for i in range(WORLD)) / WORLD
    assert (r[0]['reduced'] / WORLD - emu).abs().max() <= 1e-5 * emu.abs().max()      # other thread count, other summation order
    # identical post-step weights on every rank (replicas stay in lock step)
    assert torch.equal(r[0]['post'], r[1]['post'])
