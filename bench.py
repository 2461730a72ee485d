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


def run_reference(args, rank, world):
    """reference arm: the reference's CPU implementation of the path (oracle port -- the reference itself is Python and
    /root/reference does not exist on the GPU box) on all host cores, same metric, bounded sample per step."""
    if rank != 0:
        return
    Bs = 64
    rate, n, dt, cores = cpu_train_step_rate(Bs, 0, args.warmup, steps=args.steps)
    line = {'metric': METRIC, 'value': rate, 'unit': 'samples/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * dt / n, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': workload_name(args), 'per_gpu_batch': args.batch, 'parallelism': f'dp{args.gpus}'},
            'cpu_baseline': {'value': rate, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
                             'sample': f'{Bs} of the {args.batch} windows per step, {n} steps (torch CPU ops, all host threads)'},
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
    ts = wf.TrainStep(model, B, process_group=pg)
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
                       'collective': 'NCCL all-reduce of the flat 8.9 MB fp32 gradient per step' if world > 1 else 'none'},
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
    if rank == 0:
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
        recs = _lib.profile_records()
        fam = {}
        for name, t_ms, fl in recs:
            k = name.split(' ')[0]
            a = fam.setdefault(k, [0.0, 0.0, 0])
            a[0] += t_ms; a[1] += fl; a[2] += 1
        total_ms = sum(a[0] for a in fam.values())
        # kernels behind the family tags (csrc/wf_model.cu Scope names); fwd and dgrad launches of a conv share one kernel
        kernels = {'pw_tc_kernel (tcgen05 3xTF32 pointwise conv, fwd + dgrad)': ('tc_fwd', 'tc_dgrad'),
                   'pw_wgrad_tc_kernel (tcgen05 3xTF32 pointwise wgrad)': ('tc_wgrad',),
                   'slide_conv_kernel (sliding-window mma.sync 3xTF32 position-tap conv, fwd + dgrad)': ('slide_fwd', 'slide_dgrad'),
                   'slide_thin_kernel (sliding-window mma.sync 3xTF32 conv, <= 8 output channels, fwd + dgrad)': ('slidethin_fwd', 'slidethin_dgrad'),
                   'slide_wgrad_kernel (sliding-window mma.sync 3xTF32 wgrad)': ('slide_wgrad',),
                   'conv_gemm_kernel (FP32 implicit-GEMM conv, fwd + dgrad)': ('conv_fwd', 'conv_dgrad'),
                   'conv_wgrad_kernel (FP32 conv wgrad)': ('conv_wgrad',),
                   'thin_conv_kernel (direct conv <=16 channels, fwd + dgrad)': ('thin_fwd', 'thin_dgrad'),
                   'thin_wgrad_kernel': ('thin_wgrad',),
                   'group_conv_kernel (grouped causal conv, mma.sync 3xTF32, fwd + dgrad)': ('group_fwd', 'group_dgrad'),
                   'group_wgrad_kernel (grouped causal conv wgrad, mma.sync 3xTF32)': ('group_wgrad',)}
        tf32_peak = pk['bf16_tflops'] / 2.0            # dense tf32 (tcgen05) = half the measured bf16 tensor peak
        # warp-level mma.sync tf32: 1024 flop/clk/SM (ncu sm__ops_path_tensor_op_hmma_src_tf32 peak_sustained, profiles/README.md)
        mma_peak = 148 * 1024 * pk['sm_max_mhz'] * 1e6 / 1e12
        traffic = {}
        try:
            with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
                traffic = json.load(f)
        except Exception:
            pass
        roofs = []
        for kname, tags in kernels.items():
            k_ms = sum(fam.get(t, [0, 0, 0])[0] for t in tags)
            k_fl = sum(fam.get(t, [0, 0, 0])[1] for t in tags)
            k_n = sum(fam.get(t, [0, 0, 0])[2] for t in tags)
            if not k_n:
                continue
            ach = k_fl / (k_ms * 1e-3) / 1e12
            tcgen = 'tcgen05' in kname
            mma = 'mma.sync' in kname
            r = {'kernel': kname, 'bound': 'tensor' if (tcgen or mma) else 'fp32_fma', 'unit': 'TFLOP/s', 'launches_per_step': k_n,
                 'avg_launch_ms': k_ms / k_n, 'share_of_step': k_ms / total_ms, 'algorithmic_flops_per_launch_avg': k_fl / k_n,
                 'traffic': traffic.get(kname.split(' ')[0])}
            if tcgen or mma:      # the tensor pipe executes 3 tf32 MMAs per algorithmic fp32 multiply-add (3xTF32 split)
                peak = tf32_peak if tcgen else mma_peak
                src = (f"tcgen05 tf32 peak = measured bf16 {pk['bf16_tflops']} TFLOP/s / 2 ({pk['source']})" if tcgen else
                       f"mma.sync tf32 peak = 148 SM x 1024 flop/clk (ncu hmma tf32 peak_sustained) x {pk['sm_max_mhz']:.0f} MHz")
                r.update(achieved=3 * ach, peak=peak, frac=3 * ach / peak, fp32_equivalent_tflops=ach,
                         frac_of_fp32_fma_peak=ach / pk['fp32_tflops'],
                         peak_source=src + '; achieved counts the 3 tf32 MMAs issued per fp32 multiply-add')
            else:
                r.update(achieved=ach, peak=pk['fp32_tflops'], frac=ach / pk['fp32_tflops'],
                         peak_source=f"148 SM x 128 FP32 lanes x 2 x {pk['sm_max_mhz']:.0f} MHz ({pk['source']} sm_max_mhz)")
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
            line['also'] = also
            rate, n, dt, cores = cpu_train_step_rate(64, args.cpu_seconds, 2)
            line['cpu_baseline'] = {'value': rate, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
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
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
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
