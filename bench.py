"""WiFlow hot-path benchmark (driver contract: `python bench.py --gpus N --steps K --warmup W [--impl reference]`).

One "step" = one training step of the pose model on one batch of synthetic 540x20 CSI windows:
forward + PoseLoss + backward + clip_grad_norm_(1.0) + AdamW, dropout on, fp32 -- BASELINE.json config C4's per-GPU
batch (1024 windows per GPU, weak scaling; at N=1 the same per-GPU batch so the driver's own scaling efficiency is
meaningful).  The JSON line also carries C2 (train, B=64) and C3 (inference, B=8192) under "also" at N=1.

  value    : whole-job samples/s, inputs already resident in HBM, CUDA-graph replayed step, CUDA-event timed, max over ranks
  e2e      : same metric through the public API (`TrainStep.step`) with HOST pinned inputs: H2D of x,y and D2H of the loss
             inside every timed step
  roofline : dominant kernel family (profiled live with CUDA events around every launch, WF_FLAG_PROFILE) against the
             FP32-FMA peak the path is bound by (148 SM x 128 lanes x 2 x sm_max_mhz); the measured
             bf16 tensor peak of MEASURED_PEAKS.json is reported beside it
  cpu_baseline : the oracle port of the reference's train step on this box's host cores (bounded sample)
  also     : C2 (train B=64), C3 (inference B=8192), the input side, `torch_eager_gpu` (the oracle port of the reference on the same GPU,
             eager PyTorch: SURVEY 8d's "kernel to beat"), C1 (CPU inference B=64) and C5 (attention / ResBlock microbench, tensor-core
             vs CUDA-core kernels vs eager PyTorch)
  replicas_identical (N > 1): MIN == MAX over ranks of an integer checksum of the parameter bits after the timed steps
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FWD_FLOPS = 153.56e6            # per sample, zero-pad taps excluded
TRAIN_FLOPS = 3 * 154.81e6      # fwd + dX + dW, dense count
METRIC = 'WiFlow train-step throughput (fwd+PoseLoss+bwd+clip+AdamW), synthetic 540x20 CSI, fp32'


def peaks():
    p = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
    except Exception:
        pass
    src = 'measured' if p else 'fallback'
    sm_mhz = p.get('sm_max_mhz', 1965.0)
    return dict(hbm_gbs=p.get('hbm_gbs', 6650.0), bf16_tflops=p.get('bf16_tflops', 1590.0),
                bf16_tflops_sustained=p.get('bf16_tflops_sustained', 1400.0), sm_max_mhz=sm_mhz,
                fp32_tflops=148 * 128 * 2 * sm_mhz * 1e6 / 1e12, source=src)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs"""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '50'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm)}


def cpu_train_step_rate(B, max_seconds, warmup, steps=None):
    """oracle port of the reference's training step (train.py:196-237) on the host cores; returns samples/s, n steps"""
    import torch
    from oracle import wiflow_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    st = O.make_state(0)
    pn = O.param_names(st)
    params = {n: st[n] for n in pn}
    m = {n: torch.zeros_like(p) for n, p in params.items()}
    v = {n: torch.zeros_like(p) for n, p in params.items()}
    x, y = O.synthetic_batch(B, 0)
    it = 0

    def one():
        nonlocal it
        it += 1
        masks = O.make_dropout_masks(B, 0.5)
        _, _, g = O.grads(st, x, y, masks=masks, update_buffers=True)
        O.clip_adamw_step(params, g, m, v, it)
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    n = 0
    while True:
        one()
        n += 1
        if steps is not None and n >= steps:
            break
        if steps is None and (time.perf_counter() - t0 > max_seconds or n >= 200):
            break
    dt = time.perf_counter() - t0
    return B * n / dt, n, dt, torch.get_num_threads()


def cpu_model_name():
    try:
        with open('/proc/cpuinfo') as f:
            for line in f:
                if line.startswith('model name'):
                    return line.split(':', 1)[1].strip()
    except Exception:
        pass
    return 'unknown'


def reference_batch(args):
    """windows per reference step: the arm's own per-GPU batch when the host has the memory for its autograd graph
    (~12 MB of fp32 activations per window) and the run stays within minutes, else a bounded sample"""
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    want = args.batch
    if args.ref_batch:
        return args.ref_batch
    if avail > 3 * want * 12e6 and (args.steps + args.warmup) * want / 350.0 < 240.0:      # ~350 samples/s on 16 host cores
        return want
    return min(want, 256 if avail > 3 * 256 * 12e6 else 64)


def run_reference(args, rank, world):
    """reference arm: the reference's CPU implementation of the path (oracle port -- the reference itself is Python and
    /root/reference does not exist on the GPU box) on all host cores, same metric, bounded sample per step."""
    if rank != 0:
        return
    import torch
    Bs = reference_batch(args)
    rate, n, dt, cores = cpu_train_step_rate(Bs, 0, args.warmup, steps=args.steps)
    same = Bs == args.batch
    line = {'metric': METRIC, 'value': rate, 'unit': 'samples/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * dt / n, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': workload_name(args), 'per_gpu_batch': args.batch, 'parallelism': f'dp{args.gpus}',
                       'reference_windows_per_step': Bs, 'same_config': same,
                       'note': ('one host process regardless of --gpus: the CPU arm does not scale with N' if args.gpus > 1 else '')},
            'cpu_baseline': {'value': rate, 'unit': 'samples/s', 'cores': cores, 'kind': 'port', 'cpu_model': cpu_model_name(),
                             'torch': torch.__version__,
                             'sample': (f'{Bs} windows per step' + ('' if same else f' of the {args.batch}') +
                                        f', {n} steps (torch CPU ops through the oracle port of train.py:196-237, all host threads)')},
            'e2e': {'value': rate, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def synthetic_batch(B, seed=0):
    """x ~ N(0,1) [B,540,20] (amplitude-normalised CSI stand-in), y ~ U(0,1) [B,15,2] (key points / 1000)"""
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 540, 20, generator=g), torch.rand(B, 15, 2, generator=g)


def workload_name(args):
    return (f'C4 WiFlow data-parallel training step, {args.batch} windows of 540x20 per GPU (fwd+PoseLoss+bwd+clip(1.0)+AdamW, '
            'dropout on: TCN p=0.5 / conv p=0.3, fp32, random-init weights)')


def timed_loop(fn, steps, dev):
    import torch
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    start.record()
    for i in range(steps):
        fn(i)
    end.record()
    torch.cuda.synchronize(dev)
    return start.elapsed_time(end)


def torch_eager_gpu(dev, pk, steps=6):
    """PyTorch-eager comparator on the same GPU (SURVEY 8d "the kernel to beat"): the oracle port of the reference model -- the same
    torch.nn.functional calls the reference modules make, torch.optim.AdamW + clip_grad_norm_ as train.py:105-110,235 -- on cuda:0.
    The reference itself cannot travel to the GPU box; the oracle is checked against it op for op (tests/test_oracle.py)."""
    import torch
    from oracle import wiflow_oracle as O
    out = {'how': 'oracle port of the reference forward (torch.nn.functional on CUDA, cuDNN/cuBLAS/ATen kernels) + torch.optim.AdamW + '
                  'clip_grad_norm_(1.0), eager mode, CUDA-event timed after 3 warm-up steps; masks drawn per step as nn.Dropout does',
           'torch': torch.__version__}

    def train_rate(B, strict):
        torch.backends.cudnn.allow_tf32 = not strict
        torch.backends.cuda.matmul.allow_tf32 = False
        st = {k: v.to(dev) for k, v in O.make_state(0).items()}
        names = O.param_names(st)
        for n in names:
            st[n].requires_grad_(True)
        params = [st[n] for n in names]
        opt = torch.optim.AdamW(params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=5e-5)
        x, y = synthetic_batch(B, 3)
        x, y = x.to(dev), y.to(dev)

        def one(_):
            masks = O.make_dropout_masks(B, 0.5, device=dev)
            pred = O.forward(st, x, train=True, update_buffers=True, masks=masks)
            total, _, _ = O.pose_loss(pred, y)
            total.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            opt.zero_grad(set_to_none=True)
        for i in range(3):
            one(i)
        ms = timed_loop(one, steps, dev) / steps
        return {'samples_per_s': B / (ms / 1e3), 'ms_per_step': ms,
                'frac_of_fp32_roofline': B / (ms / 1e3) * TRAIN_FLOPS / 1e12 / pk['fp32_tflops']}

    def infer_rate(B, strict):
        torch.backends.cudnn.allow_tf32 = not strict
        st = {k: v.to(dev) for k, v in O.make_state(0).items()}
        x, _ = synthetic_batch(B, 4)
        x = x.to(dev)

        def one(_):
            with torch.no_grad():
                O.forward(st, x)
        for i in range(2):
            one(i)
        ms = timed_loop(one, 3, dev) / 3
        return {'samples_per_s': B / (ms / 1e3), 'ms_per_step': ms,
                'frac_of_fp32_roofline': B / (ms / 1e3) * FWD_FLOPS / 1e12 / pk['fp32_tflops']}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for strict, tag in ((False, 'default_flags_cudnn_tf32_convs'), (True, 'strict_fp32')):
            out[tag] = {'train_b1024': train_rate(1024, strict), 'train_b64': train_rate(64, strict), 'infer_b8192': infer_rate(8192, strict)}
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    return out


def c1_cpu_inference(seconds=4.0):
    """BASELINE.json configs[0]: WiFlow inference, batch 64, fp32 on the host cores (the reference's run.py model path)"""
    import torch
    from oracle import wiflow_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    st = O.make_state(0)
    x, _ = O.synthetic_batch(64, 0)
    with torch.no_grad():
        for _ in range(3):
            O.forward(st, x)
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < seconds or n < 5:
            O.forward(st, x)
            n += 1
        dt = time.perf_counter() - t0
    return {'samples_per_s': 64 * n / dt, 'ms_per_batch': 1e3 * dt / n, 'iterations': n, 'cores': torch.get_num_threads(),
            'cpu_model': cpu_model_name(), 'kind': 'port (oracle forward, eval mode, no_grad)'}


def c5_worker(tag):
    """BASELINE.json configs[4]: DualAxialAttention on [B,64,15,20] and the four AsymmetricConvBlocks on [B,8,20,240], forward +
    backward through the drop-in modules; `tag` = 'tensor' (tcgen05 / mma.sync kernels), 'cuda_core' (FP32 SIMT kernels: this
    process was started with WF_DISABLE_TC=1 WF_DISABLE_SLABTC=1 WF_DISABLE_SLIDE=1) or 'torch_eager' (oracle port on CUDA)."""
    import torch
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    res = {}
    if tag == 'torch_eager':
        from oracle import wiflow_oracle as O
        st = {k: v.to(dev) for k, v in O.make_state(0).items()}
        for n in O.param_names(st):
            st[n].requires_grad_(True)
        ctx = O._Ctx(st, True, True, None, None)

        def attn(x):
            return O._axial(ctx, O._axial(ctx, x, 'attention.width_axis', True), 'attention.height_axis', False)

        def res_stack(x):
            for i in range(4):
                x = O._conv_block(ctx, x, f'residual_blocks.{i}', 2)
            return x
    else:
        import wiflow_b200 as wf
        torch.manual_seed(0)
        att = wf.DualAxialAttention(64, 64, groups=8).to(dev).train()
        chans = (8, 8, 16, 32, 64)
        blocks = torch.nn.ModuleList([wf.AsymmetricConvBlock(chans[i], chans[i + 1]) for i in range(4)]).to(dev).train()
        for m in blocks.modules():
            if isinstance(m, torch.nn.Dropout2d):
                m.p = 0.0

        def attn(x):
            return att(x)

        def res_stack(x):
            for b in blocks:
                x = b(x)
            return x
    for name, fn, shape in (('dual_axial_attention_64x15x20', attn, (64, 15, 20)), ('resblock_stack_8x20x240', res_stack, (8, 20, 240))):
        for B in (64, 1024, 8192):
            x = torch.randn(B, *shape, device=dev, requires_grad=True)
            y = fn(x)
            gy = torch.randn_like(y)

            def one(_):
                x.grad = None
                fn(x).backward(gy)
            for i in range(2):
                one(i)
            reps = 20 if B <= 64 else 5 if B <= 1024 else 2
            ms = timed_loop(one, reps, dev) / reps
            res[f'{name}_b{B}'] = {'us_per_sample_fwd_bwd': 1e3 * ms / B, 'ms': ms}
            del x, y, gy
            torch.cuda.empty_cache()
    print(json.dumps(res), flush=True)


def c5_microbench():
    res = {}
    variants = (('tensor', {}), ('cuda_core', {'WF_DISABLE_TC': '1', 'WF_DISABLE_SLABTC': '1', 'WF_DISABLE_SLIDE': '1'}), ('torch_eager', {}))
    for tag, env in variants:
        e = dict(os.environ)
        e.update(env)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--c5-worker', tag], env=e, capture_output=True, text=True, timeout=600)
            res[tag] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {'error': (r.stderr or r.stdout)[-300:]}
        except Exception as ex:
            res[tag] = {'error': repr(ex)[:300]}
    res['note'] = ('forward + backward per call through the nn.Module drop-ins, train mode, dropout off; tensor = tcgen05 (qkv projection, conv '
                   'blocks >= 16 channels) + mma.sync kernels, cuda_core = the FP32 SIMT kernels of wf_conv.cu / wf_thin.cu (the attention '
                   'softmax/QK/AV kernel is CUDA-core in both), torch_eager = oracle port on the same GPU')
    return res


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import wiflow_b200 as wf
    from wiflow_b200 import _lib, ops

    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    pg = None
    if world > 1:
        # NCCL prints its version banner on fd 1 when the first communicator comes up: keep stdout to the one JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            pg = dist.group.WORLD
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    B = args.batch
    torch.manual_seed(0)
    model = wf.WiFlowPoseModel(dropout=0.5).to(dev)
    ts = wf.TrainStep(model, B, process_group=pg, dropout_rng=args.dropout_rng)
    nbatch = 4
    xs, ys = [], []
    for i in range(nbatch):
        x, y = synthetic_batch(B, seed=1000 * rank + i)
        xs.append(x.to(dev)); ys.append(y.to(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_dev(i):
        ts.step(xs[i % nbatch], ys[i % nbatch])

    for i in range(max(args.warmup, 3)):
        step_dev(i)
    barrier()
    with ClockSampler(local_rank) as clk:
        ms = timed_loop(step_dev, args.steps, dev)
    barrier()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    value = world * B * args.steps / (ms / 1e3)
    loss_val = ts.out.tolist()

    # ---- e2e: host pinned inputs -> H2D -> step -> D2H of the loss, every step ----
    hx = [x.cpu().pin_memory() for x in xs]
    hy = [y.cpu().pin_memory() for y in ys]
    # one step in flight, as a training loop that logs asynchronously runs: step i's inputs are uploaded (copy stream) and its
    # kernels enqueued, then the host waits for the device->host copy of step i-1's result before it goes on to step i+1
    res = [torch.zeros(4).pin_memory() for _ in range(2)]
    pending = [None]

    def step_e2e(i):
        out = ts.step(hx[i % nbatch], hy[i % nbatch])
        res[i & 1].copy_(out, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        if pending[0] is not None:
            pending[0].synchronize()
        pending[0] = ev
    for i in range(3):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(i)
    pending[0].synchronize()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / t.item()
    h2d = hx[0].numel() * 4 + hy[0].numel() * 4
    d2h = 16

    line = {'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': workload_name(args), 'per_gpu_batch': B, 'global_batch': B * world, 'parallelism': f'dp{world}',
                       'l2': f'inputs rotate over {nbatch} batches ({nbatch * B * 43200 / 1e6:.0f} MB) and every step streams '
                             f'{ts.ws.numel() / 1e9:.1f} GB of saved activations, far above the 126 MB L2',
                       'collective': 'NCCL all-reduce of the flat 8.9 MB fp32 gradient per step' if world > 1 else 'none',
                       'dropout_rng': "philox: all 18 masks of a step drawn inside the timed step by one wf_dropout_masks launch" if args.dropout_rng == 'philox'
                                      else "torch: one F.dropout(ones) draw per site inside the timed step (the parity mode)"},
            'e2e': {'value': e2e_value, 'unit': 'samples/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'how': 'TrainStep.step(pinned host x, y) + device->host read of its [loss, position, bone, grad_norm] every step; one step '
                           'in flight (the read of step i completes while step i+1 runs), wall clock over the timed steps'},
            'final_loss': loss_val[0], 'grad_norm': loss_val[3]}

    # ---- launches per step: one eager (un-graphed) step on EVERY rank (it contains the all-reduce) ----
    c0 = _lib.lib().wf_launch_count()
    ts.use_graph = False
    ts.step(xs[0], ys[0])
    torch.cuda.synchronize(dev)
    per_step = _lib.lib().wf_launch_count() - c0
    replicas_identical = None
    if world > 1:          # every rank must hold bit-identical weights: MIN == MAX of an integer checksum of the parameter bits
        csum = ts.params.view(torch.int32).to(torch.int64).sum().reshape(1)
        lo, hi = csum.clone(), csum.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        replicas_identical = bool(lo.item() == hi.item())
    if rank == 0:
        if replicas_identical is not None:
            line['replicas_identical'] = replicas_identical
        line['clocks'] = clk.summary()
        pk = peaks()
        # ---- per-kernel profile of one forward+backward (CUDA events around every launch; no collective inside) ----
        line['gpu_launches'] = int(per_step * args.steps)
        line['gpu_launches_per_step'] = int(per_step)
        masks = model._wf_masks(B, dev)
        pf = ts.flags | _lib.FLAG_PROFILE
        pred = ops.block_forward(xs[0], ts.params, ts.running, ts.nbt, masks, [0, 0, 0, 0, 0], pf, ts.ws)
        out3, dpred = ops.pose_loss(pred, ys[0], 0, 1.0, 0.2, ts.loss_scratch, True)
        ops.block_backward(xs[0], ts.params, masks, dpred, [0, 0, 0, 0, 0], pf, ts.ws, False)
        torch.cuda.synchronize(dev)
        recs = _lib.profile_records(with_bytes=True)
        fam = {}
        for name, t_ms, fl, by in recs:
            k = name.split(' ')[0]
            a = fam.setdefault(k, [0.0, 0.0, 0, 0.0])
            a[0] += t_ms; a[1] += fl; a[2] += 1; a[3] += by
        total_ms = sum(a[0] for a in fam.values())
        # kernels behind the family tags (csrc/wf_model.cu Scope names); fwd and dgrad launches of a conv share one kernel
        kernels = {'pw_tc_kernel (tcgen05 3xTF32 pointwise conv, fwd + dgrad)': ('tc_fwd', 'tc_dgrad'),
                   'pw_wgrad_tc_kernel (tcgen05 3xTF32 pointwise wgrad)': ('tc_wgrad',),
                   'slab_tc_kernel (TMA + tcgen05 3xTF32 position-tap conv, fwd + dgrad)': ('slab_fwd', 'slab_dgrad'),
                   'slab_wgrad_kernel (TMA + tcgen05 3xTF32 position-tap wgrad)': ('slab_wgrad',),
                   'slide_conv_kernel (sliding-window mma.sync 3xTF32 position-tap conv, fwd + dgrad)': ('slide_fwd', 'slide_dgrad'),
                   'slide_thin_kernel (sliding-window mma.sync 3xTF32 conv, <= 8 output channels, fwd + dgrad)': ('slidethin_fwd', 'slidethin_dgrad'),
                   'slide_wgrad_kernel (sliding-window mma.sync 3xTF32 wgrad)': ('slide_wgrad',),
                   'conv_gemm_kernel (FP32 implicit-GEMM conv, fwd + dgrad)': ('conv_fwd', 'conv_dgrad'),
                   'conv_wgrad_kernel (FP32 conv wgrad)': ('conv_wgrad',),
                   'thin_conv_kernel (direct conv <=16 channels, fwd + dgrad)': ('thin_fwd', 'thin_dgrad'),
                   'thin_wgrad_kernel': ('thin_wgrad',),
                   'group_conv_kernel (grouped causal conv, mma.sync 3xTF32, fwd + dgrad)': ('group_fwd', 'group_dgrad'),
                   'group_wgrad_kernel (grouped causal conv wgrad, mma.sync 3xTF32)': ('group_wgrad',)}
        tf32_peak = pk['bf16_tflops'] / 2.0            # dense tf32 tensor peak = half the measured bf16 peak (MEASURED_PEAKS.json)
        traffic = {}
        try:
            with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
                traffic = json.load(f)
        except Exception:
            pass
        roofs = []
        for kname, tags in kernels.items():
            k_ms = sum(fam.get(t, [0, 0, 0, 0])[0] for t in tags)
            k_fl = sum(fam.get(t, [0, 0, 0, 0])[1] for t in tags)
            k_n = sum(fam.get(t, [0, 0, 0, 0])[2] for t in tags)
            k_by = sum(fam.get(t, [0, 0, 0, 0])[3] for t in tags)
            if not k_n:
                continue
            ach = k_fl / (k_ms * 1e-3) / 1e12                  # ALGORITHMIC flops (2 per multiply-add of the conv) per second
            gbs = k_by / (k_ms * 1e-3) / 1e9
            tensor = 'tcgen05' in kname or 'mma.sync' in kname
            # the primary denominator is the FP32-FMA peak (the straightforward arithmetic that meets 1e-4); a kernel
            # whose algorithmic bytes / HBM peak is the longer time is HBM-bound and reported against the measured copy rate
            t_flop, t_byte = k_fl / (pk['fp32_tflops'] * 1e12), k_by / (pk['hbm_gbs'] * 1e9)
            hbm_bound = t_byte > t_flop
            r = {'kernel': kname, 'bound': 'hbm' if hbm_bound else 'tensor' if tensor else 'fp32_fma',
                 'achieved': gbs if hbm_bound else ach, 'peak': pk['hbm_gbs'] if hbm_bound else pk['fp32_tflops'],
                 'unit': 'GB/s' if hbm_bound else 'TFLOP/s',
                 'frac': (gbs / pk['hbm_gbs']) if hbm_bound else ach / pk['fp32_tflops'],
                 'peak_source': (f"measured copy bandwidth {pk['hbm_gbs']} GB/s ({pk['source']})" if hbm_bound else
                                 f"FP32-FMA peak 148 SM x 128 lanes x 2 x {pk['sm_max_mhz']:.0f} MHz ({pk['source']} sm_max_mhz): SURVEY 8(d) primary denominator"),
                 'launches_per_step': k_n, 'avg_launch_ms': k_ms / k_n, 'share_of_step': k_ms / total_ms,
                 'algorithmic_flops_per_launch_avg': k_fl / k_n, 'algorithmic_bytes_per_launch_avg': k_by / k_n,
                 'fp32_equivalent_tflops': ach, 'frac_of_fp32_fma_peak': ach / pk['fp32_tflops'],
                 'hbm_gbs': gbs, 'frac_of_hbm_peak': gbs / pk['hbm_gbs'],
                 'traffic': traffic.get(kname.split(' ')[0]),
                 'traffic_source': 'static: one ncu --set full capture per kernel, committed as profiles/ncu_traffic.json (not measured in this run)'}
            if tensor:      # secondary: against the measured tensor peak, algorithmic (no credit for the 3 tf32 MMAs per fp32 product)
                r.update(tensor_peak_tflops=tf32_peak, frac_of_tensor_peak=ach / tf32_peak, issued_tensor_frac=3 * ach / tf32_peak,
                         tensor_peak_source=f"measured bf16 {pk['bf16_tflops']} TFLOP/s / 2 ({pk['source']}); issued_tensor_frac counts the 3 tf32 MMAs of the 3xTF32 split")
            roofs.append(r)
        roofs.sort(key=lambda r: -r['share_of_step'])
        line['roofline'] = roofs[0]
        line['roofline_all'] = roofs
        step_tflops = value / world * TRAIN_FLOPS / 1e12
        line['step_roofline'] = {'flops_per_sample': TRAIN_FLOPS, 'achieved_tflops_per_gpu': step_tflops, 'peak_tflops': pk['fp32_tflops'],
                                 'frac': step_tflops / pk['fp32_tflops'],
                                 'roofline_samples_per_s_per_gpu': pk['fp32_tflops'] * 1e12 / TRAIN_FLOPS}
        line['kernel_breakdown_ms'] = {k: round(a[0], 4) for k, a in sorted(fam.items(), key=lambda kv: -kv[1][0])}
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', f'bench_profile_n{world}_b{B}.json'), 'w') as f:
            json.dump({'records': recs, 'families': fam, 'B': B}, f)

        if world == 1 and not args.no_extras:
            also = {}
            # C2: train step B=64
            torch.manual_seed(0)
            m2 = wf.WiFlowPoseModel(dropout=0.5).to(dev)
            t2 = wf.TrainStep(m2, 64)
            x2, y2 = synthetic_batch(64, 5)
            x2, y2 = x2.to(dev), y2.to(dev)
            for i in range(5):
                t2.step(x2, y2)
            ms2 = timed_loop(lambda i: t2.step(x2, y2), 50, dev)
            also['C2_train_b64'] = {'samples_per_s': 64 * 50 / (ms2 / 1e3), 'ms_per_step': ms2 / 50,
                                    'frac_of_fp32_roofline': 64 * 50 / (ms2 / 1e3) * TRAIN_FLOPS / 1e12 / pk['fp32_tflops']}
            # C3: inference B=8192
            inf = wf.InferStep(model, 8192)
            x3, _ = synthetic_batch(8192, 6)
            x3 = x3.to(dev)
            for i in range(3):
                inf.step(x3)
            ms3 = timed_loop(lambda i: inf.step(x3), 5, dev)
            also['C3_infer_b8192'] = {'samples_per_s': 8192 * 5 / (ms3 / 1e3), 'ms_per_step': ms3 / 5,
                                      'frac_of_fp32_roofline': 8192 * 5 / (ms3 / 1e3) * FWD_FLOPS / 1e12 / pk['fp32_tflops']}
            # input side: resident-window gather and the train.py:187-193 augmentation, HBM-bound
            from wiflow_b200.utils import augmentation as A
            from oracle import data_oracle as DO          # cpu_baseline leg of the input side: the only use of oracle/ in this arm
            nwin = 8192
            resident = torch.randn(nwin, 540, 20, device=dev)
            gidx = [torch.randint(0, nwin, (B,), device=dev) for _ in range(8)]
            # every buffer rotates over 4 copies (177 MB each of noise and output on top of the 354 MB source), so that no launch
            # finds its operands in the 126 MB L2
            noises = [torch.randn(B, 540, 20, device=dev) for _ in range(4)]
            gouts = [torch.empty(B, 540, 20, device=dev) for _ in range(4)]
            for gbuf in gouts:
                gbuf.normal_()
            stats = torch.zeros(2, device=dev, dtype=torch.float64)
            spans_h = A.draw_time_masks(B, 540, 0.3)
            spans_d = A._spans_to_device(spans_h, dev)
            wbytes = B * 43200
            inp = {}
            for name, fn, nbytes in (
                    ('gather', lambda i: ops.window_load(resident, gidx[i % 8], gouts[i % 4]), 2 * wbytes),
                    ('gather_mask_stats', lambda i: ops.window_load(resident, gidx[i % 8], gouts[i % 4], spans_d, stats), 2 * wbytes),
                    ('noise_scale', lambda i: ops.noise_scale(gouts[i % 4], noises[(i + 1) % 4], 0.02, 1.0001, stats, gouts[i % 4]), 3 * wbytes)):
                for i in range(3):
                    fn(i)
                msk = timed_loop(fn, 48, dev) / 48
                inp[name] = {'ms': msk, 'windows_per_s': B / (msk / 1e3), 'algorithmic_bytes': nbytes,
                             'achieved_gbs': nbytes / (msk / 1e3) / 1e9, 'frac_of_hbm_peak': nbytes / (msk / 1e3) / 1e9 / pk['hbm_gbs']}
            noise = noises[0]
            xc = resident[:64].cpu()
            nc = noise[:64].cpu()
            torch.manual_seed(0)
            t0 = time.perf_counter()
            reps = 0
            while time.perf_counter() - t0 < 3.0:
                DO.augment_step(xc, nc)
                reps += 1
            inp['cpu_port_augment_windows_per_s'] = 64 * reps / (time.perf_counter() - t0)
            inp['note'] = (f'B={B} windows gathered out of a resident array of {nwin} (354 MB > L2); gather_mask_stats = gather + time masking '
                           '(30 % of the windows) + sum/sumsq for add_noise; noise_scale = add_noise + random_scaling in place; '
                           'cpu port = oracle augment_step (the reference\'s Python loops) on 64 windows')
            also['input_side'] = inp
            del resident, noise, noises, gouts
            del ts, t2, inf, m2
            torch.cuda.empty_cache()
            try:
                also['torch_eager_gpu'] = torch_eager_gpu(dev, pk)
            except Exception as ex:                      # a comparator must never take the product's line down
                also['torch_eager_gpu'] = {'error': repr(ex)[:300]}
            torch.cuda.empty_cache()
            also['C1_cpu_infer_b64'] = c1_cpu_inference()
            also['C5_microbench'] = c5_microbench()
            line['also'] = also
            rate, n, dt, cores = cpu_train_step_rate(64, args.cpu_seconds, 2)
            line['cpu_baseline'] = {'value': rate, 'unit': 'samples/s', 'cores': cores, 'kind': 'port', 'cpu_model': cpu_model_name(),
                                    'sample': f'oracle port of the reference train step, B=64, {n} steps in {dt:.1f} s on the host cores'}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=1024, help='windows per GPU per step')
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--no-extras', action='store_true')
    ap.add_argument('--dropout-rng', default='philox', choices=['philox', 'torch'],
                    help="how TrainStep draws the dropout masks: the library's one-launch Philox generator (default) or torch's generator")
    ap.add_argument('--ref-batch', type=int, default=0, help='windows per step of the reference arm (0: choose)')
    ap.add_argument('--c5-worker', default='', help='internal: run the C5 microbench in this process and print JSON')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.c5_worker:
        c5_worker(args.c5_worker)
        return
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        os.execvp(sys.executable, [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}',
                                   '--master-addr', '127.0.0.1', '--master-port', '29531', os.path.abspath(__file__)] + sys.argv[1:])
    run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
