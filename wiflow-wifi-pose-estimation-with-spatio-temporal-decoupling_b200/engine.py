"""Fused training / inference steps on top of the C ABI: the restated hot loop of the reference's train.py
(:183-239: forward, PoseLoss, backward, clip_grad_norm_(1.0), AdamW(lr 1e-4, wd 5e-5)) without autograd, with
all device work of one step captured in a CUDA graph and, under torch.distributed, the flat 8.9 MB gradient
all-reduced over NCCL (one process per GPU, batch-sharded; BatchNorm statistics stay local per rank exactly as under
the reference's nn.DataParallel, train.py:91-93)."""
import ctypes

import torch
import torch.distributed as dist

from . import _lib, ops
from .models.pose_model import WiFlowPoseModel

MODEL_DESC = [_lib.BLOCK_MODEL, 0, 0, 0, 0]


def shard_bounds(n_items: int, rank: int, world: int):
    """[begin, end) of the contiguous shard of `n_items` windows that `rank` of `world` processes (the batch is
    sharded, nothing else).  Remainders go to the lowest ranks, so shards differ by at most one window."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_gradients(flat_grads: torch.Tensor, group=None, world: int = None):
    """The one exchange step of data-parallel training: SUM all-reduce of the flat gradient buffer over the process group
    (NCCL over NVLink on GPUs, gloo in the CPU tests).  Returns the factor the optimizer must scale the sum by (1/world):
    the fused clip+AdamW kernel applies it, so no extra pass over the 8.9 MB buffer is needed."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


class TrainStep:
    """One data-parallel training step of WiFlowPoseModel at a fixed per-rank batch size.

    step(x, y) consumes device tensors x [B,540,20], y [B,15,2] and returns a device tensor
    [total, position, bone, grad_norm] without synchronising."""

    def __init__(self, model: WiFlowPoseModel, batch_size: int, lr=1e-4, weight_decay=5e-5, betas=(0.9, 0.999), eps=1e-8,
                 max_norm=1.0, position_weight=1.0, bone_weight=0.2, loss_type='smooth_l1', process_group=None,
                 use_cuda_graph=True, dropout=True, accumulation_steps=1, metric_thresholds=None, dropout_rng='torch', seed=None):
        self.model = model
        self.B = int(batch_size)
        self.k = max(1, int(accumulation_steps))            # micro-batches per optimizer step (train.py:80, :243-249)
        self.thresholds = tuple(float(t) for t in metric_thresholds) if metric_thresholds else ()
        self.hp = dict(lr=lr, wd=weight_decay, b1=betas[0], b2=betas[1], eps=eps, max_norm=max_norm)
        self.loss = (_lib.LOSS_TYPES[loss_type], float(position_weight), float(bone_weight))
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (process_group is not None or (dist.is_available() and dist.is_initialized())) else 1
        self.dropout = dropout
        if dropout_rng not in ('torch', 'philox'):
            raise ValueError("dropout_rng must be 'torch' (masks drawn by torch's generator, as nn.Dropout would: the parity mode) or "
                             "'philox' (the library's one-launch generator: the perf mode)")
        self.dropout_rng = dropout_rng
        model.train()
        self.params, self.running, self.nbt = model._wf_state()
        dev = self.params.device
        if dev.type != 'cuda':
            raise RuntimeError('TrainStep needs the model on a CUDA (sm_100a) device')
        self.dev = dev
        n = self.params.numel()
        self.grads = torch.zeros(n, device=dev)
        self.exp_avg = torch.zeros(n, device=dev)
        self.exp_avg_sq = torch.zeros(n, device=dev)
        self.adam_state = ops.adam_state(dev)
        self.rng_state = None
        if dropout_rng == 'philox':
            # device-side draw counter (every graph replay draws fresh masks); the key mixes in the rank: each replica sees other data
            self.rng_state = torch.zeros(2, device=dev, dtype=torch.int64)
            base = torch.initial_seed() if seed is None else int(seed)
            rank = dist.get_rank(process_group) if self.world > 1 else 0
            self.rng_seed = (base * 0x9E3779B97F4A7C15 + (rank + 1) * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF
        self.flags = _lib.FLAG_TRAIN | _lib.FLAG_SAVE
        self.ws = torch.empty(ops.workspace_bytes(MODEL_DESC, self.B, self.flags), device=dev, dtype=torch.uint8)
        self.x = torch.zeros(self.B, 540, 20, device=dev)
        self.y = torch.zeros(self.B, 15, 2, device=dev)
        self.loss_scratch = torch.zeros(2, device=dev, dtype=torch.float64)
        self.out = torch.zeros(4, device=dev)
        self.pred = None
        self.use_graph = use_cuda_graph
        self._graph_a = self._graph_b = None
        self._copy_stream = None                             # host-input path (_upload): created on first use
        self.kernel_launches = 0
        self.micro = 0                                       # micro-batches accumulated since the last optimizer step
        self.acc = torch.zeros(n, device=dev) if self.k > 1 else None
        # device-side epoch accumulators (train.py:221-230 keeps them on the host with 7 .item() syncs per step):
        # [total, position, bone, mpjpe, pck(thr_0), ..., windows], each weighted by the batch size
        self.sums = torch.zeros(5 + len(self.thresholds), device=dev, dtype=torch.float64)
        self._metric_scratch = torch.zeros(16, device=dev, dtype=torch.float64)
        if self.world > 1:
            self.sync_replicas()

    def sync_replicas(self):
        """Rank 0's parameters, BatchNorm running statistics and optimizer state become every rank's (what nn.DataParallel's
        replicate() does each forward, train.py:91-93: replica 0 wins).  After this the ranks stay bit-identical: they apply the
        same all-reduced gradient through a deterministic clip + AdamW.  Called at construction; call it again after loading a
        checkpoint on rank 0 only, or at epoch boundaries to give every rank rank 0's running statistics before validation."""
        if self.world <= 1:
            return
        g = self.pg
        src = dist.get_global_rank(g, 0) if g is not None else 0
        for t in (self.params, self.running, self.nbt, self.exp_avg, self.exp_avg_sq, self.adam_state):
            dist.broadcast(t, src=src, group=g)

    # -- pieces -----------------------------------------------------------------------------------------
    def _fwd_bwd_on(self, x, y):
        """forward + PoseLoss + backward of one micro-batch (any B <= self.B: the workspace was sized for self.B)"""
        m = self.model
        B = x.shape[0]
        if not self.dropout:
            masks = []
        elif self.rng_state is not None:
            masks = m._wf_masks(B, self.dev, self.rng_state, self.rng_seed)
        else:
            masks = m._wf_masks(B, self.dev)
        self.pred = ops.block_forward(x, self.params, self.running, self.nbt, masks, MODEL_DESC, self.flags, self.ws)
        out3, dpred = ops.pose_loss(self.pred, y, self.loss[0], self.loss[1], self.loss[2], self.loss_scratch, True)
        self.out[:3].copy_(out3)
        grads, _ = ops.block_backward(x, self.params, masks, dpred, MODEL_DESC, self.flags, self.ws, False)
        if self.acc is not None:
            self.acc.add_(grads)
        else:
            self.grads.copy_(grads)
        # epoch statistics stay on the device
        self.sums[:3].add_(out3.double(), alpha=float(B))
        if self.thresholds:
            mt = ops.pose_metrics(self.pred.detach(), y, list(self.thresholds), True, self._metric_scratch)
            self.sums[3:4].add_(mt[-1:].double(), alpha=float(B))
            self.sums[4:4 + len(self.thresholds)].add_(mt[:-1].double(), alpha=float(B))
        self.sums[-1:].add_(float(B))

    def _fwd_bwd(self):
        self._fwd_bwd_on(self.x, self.y)

    def _optim(self):
        h = self.hp
        g = self.acc if self.acc is not None else self.grads
        # the reference divides every micro-batch loss by k (train.py:200), also in an incomplete last group
        ops.clip_adamw(self.params, g, self.exp_avg, self.exp_avg_sq, self.adam_state, h['lr'], h['b1'], h['b2'],
                       h['eps'], h['wd'], h['max_norm'], 1.0 / (self.world * self.k))
        self.out[3:4].copy_(self.adam_state.view(torch.float32)[4:5])
        if self.acc is not None:
            self.acc.zero_()

    def _allreduce(self):
        allreduce_gradients(self.acc if self.acc is not None else self.grads, self.pg, self.world)

    def _capture(self):
        saved_sums = self.sums.clone()
        saved_acc = self.acc.clone() if self.acc is not None else None      # micro-batches accumulated before the first full-size step
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(2):                    # warm-up outside capture (lazy init, allocator)
                self._fwd_bwd()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self._graph_a = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph_a):
            self._fwd_bwd()
        if self.acc is not None:
            self.acc.copy_(saved_acc)
        self.sums.copy_(saved_sums)

    def _capture_optim(self):
        self._graph_b = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph_b):
            self._optim()

    def set_lr(self, lr: float):
        """new learning rate for the following optimizer steps (ReduceLROnPlateau lives on the host); the optimizer graph is
        re-captured lazily because lr is a kernel argument"""
        if float(lr) != self.hp['lr']:
            self.hp['lr'] = float(lr)
            self._graph_b = None

    def reset_sums(self):
        self.sums.zero_()

    def read_sums(self):
        """one device->host copy: dict of epoch means (total, position, bone, mpjpe, pck@thr..., windows)"""
        v = self.sums.tolist()
        n = v[-1]
        names = ['loss', 'position', 'bone', 'mpjpe'] + [f'pck@{t:g}' for t in self.thresholds]
        out = {k: (x / n if n > 0 else float('inf')) for k, x in zip(names, v[:-1])}
        out['windows'] = int(n)
        return out

    def _apply(self):
        """all-reduce + clip + AdamW on what has been accumulated"""
        self._allreduce()
        if self.use_graph:
            if self._graph_b is None:
                self._capture_optim()          # capture runs nothing: replay below does the step
            self._graph_b.replay()
        else:
            self._optim()
        self.micro = 0

    def flush(self):
        """optimizer step on an incomplete accumulation group (train.py:251-257)"""
        if self.micro > 0:
            self._apply()

    def _upload(self, x, y):
        """Host batch -> the step's input buffers through one of two device staging slots on a copy stream: the upload of batch i+1
        is enqueued while step i still runs (the copy engine works next to the SMs), and the compute stream only pays a device-to-
        device copy of 44 MB.  Pinned host tensors make the upload asynchronous; pageable ones work but serialise on the host."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.dev)
            self._xs = [torch.empty_like(self.x) for _ in range(2)]
            self._ys = [torch.empty_like(self.y) for _ in range(2)]
            self._slot_free = [None, None]
            self._slot = 0
        k = self._slot
        self._slot ^= 1
        cur = torch.cuda.current_stream(self.dev)
        cs = self._copy_stream
        if self._slot_free[k] is not None:
            cs.wait_event(self._slot_free[k])            # the step that consumed this slot two calls ago has read it
        with torch.cuda.stream(cs):
            self._xs[k].copy_(x, non_blocking=True)
            self._ys[k].copy_(y, non_blocking=True)
            up = torch.cuda.Event()
            up.record(cs)
        cur.wait_event(up)
        self.x.copy_(self._xs[k], non_blocking=True)
        self.y.copy_(self._ys[k], non_blocking=True)
        free = torch.cuda.Event()
        free.record(cur)
        self._slot_free[k] = free

    # -- public -----------------------------------------------------------------------------------------
    def step(self, x, y):
        """one micro-batch; every `accumulation_steps`-th call also runs the exchange + optimizer step"""
        B = x.shape[0]
        if B != self.B:                       # ragged last batch of an epoch: same kernels, eager launches
            if B > self.B:
                raise RuntimeError(f'batch of {B} windows exceeds the workspace sized for {self.B}')
            self._fwd_bwd_on(x.to(self.dev, non_blocking=True).contiguous(), y.to(self.dev, non_blocking=True).contiguous())
        else:
            if x.is_cuda:
                self.x.copy_(x, non_blocking=True)
                self.y.copy_(y, non_blocking=True)
            else:
                self._upload(x, y)
            if self.use_graph:
                if self._graph_a is None:
                    # the warm-up iterations inside _capture must not disturb the weights: they only run fwd/bwd
                    saved = (self.running.clone(), self.nbt.clone())
                    self._capture()
                    self.running.copy_(saved[0]); self.nbt.copy_(saved[1])
                self._graph_a.replay()
            else:
                self._fwd_bwd()
        self.micro += 1
        if self.micro >= self.k:
            self._apply()
        return self.out


class InferStep:
    """Eval-mode forward at a fixed batch size with a persistent workspace (CUDA-graph replayed)."""

    def __init__(self, model: WiFlowPoseModel, batch_size: int, use_cuda_graph=True):
        model.eval()
        self.model, self.B = model, int(batch_size)
        self.params, self.running, self.nbt = model._wf_state()
        dev = self.params.device
        self.dev = dev
        self.ws = torch.empty(ops.workspace_bytes(MODEL_DESC, self.B, 0), device=dev, dtype=torch.uint8)
        self.x = torch.zeros(self.B, 540, 20, device=dev)
        self.pred = None
        self.use_graph = use_cuda_graph
        self._graph = None

    def _run(self):
        self.pred = ops.block_forward(self.x, self.params, self.running, self.nbt, [], MODEL_DESC, 0, self.ws)

    def step(self, x):
        self.x.copy_(x, non_blocking=True)
        if self.use_graph:
            if self._graph is None:
                self._run()
                torch.cuda.synchronize(self.dev)
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._run()
            self._graph.replay()
        else:
            self._run()
        return self.pred
