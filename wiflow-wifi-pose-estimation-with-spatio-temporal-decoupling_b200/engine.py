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
                 use_cuda_graph=True, dropout=True):
        self.model = model
        self.B = int(batch_size)
        self.hp = dict(lr=lr, wd=weight_decay, b1=betas[0], b2=betas[1], eps=eps, max_norm=max_norm)
        self.loss = (_lib.LOSS_TYPES[loss_type], float(position_weight), float(bone_weight))
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (process_group is not None or (dist.is_available() and dist.is_initialized())) else 1
        self.dropout = dropout
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
        self.adam_state = torch.zeros(8, device=dev, dtype=torch.float64)
        self.flags = _lib.FLAG_TRAIN | _lib.FLAG_SAVE
        self.ws = torch.empty(ops.workspace_bytes(MODEL_DESC, self.B, self.flags), device=dev, dtype=torch.uint8)
        self.x = torch.zeros(self.B, 540, 20, device=dev)
        self.y = torch.zeros(self.B, 15, 2, device=dev)
        self.loss_scratch = torch.zeros(2, device=dev, dtype=torch.float64)
        self.out = torch.zeros(4, device=dev)
        self.pred = None
        self.use_graph = use_cuda_graph
        self._graph_a = self._graph_b = None
        self.kernel_launches = 0

    # -- pieces -----------------------------------------------------------------------------------------
    def _fwd_bwd(self):
        m = self.model
        masks = m._wf_masks(self.B, self.dev) if self.dropout else []
        self.pred = ops.block_forward(self.x, self.params, self.running, self.nbt, masks, MODEL_DESC, self.flags, self.ws)
        out3, dpred = ops.pose_loss(self.pred, self.y, self.loss[0], self.loss[1], self.loss[2], self.loss_scratch, True)
        self.out[:3].copy_(out3)
        grads, _ = ops.block_backward(self.x, self.params, masks, dpred, MODEL_DESC, self.flags, self.ws, False)
        self.grads.copy_(grads) if grads.data_ptr() != self.grads.data_ptr() else None

    def _optim(self):
        h = self.hp
        ops.clip_adamw(self.params, self.grads, self.exp_avg, self.exp_avg_sq, self.adam_state, h['lr'], h['b1'], h['b2'],
                       h['eps'], h['wd'], h['max_norm'], 1.0 / self.world)
        self.out[3:4].copy_(self.adam_state.view(torch.float32)[4:5])

    def _allreduce(self):
        allreduce_gradients(self.grads, self.pg, self.world)

    def _capture(self):
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
        self._graph_b = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph_b):
            self._optim()

    # -- public -----------------------------------------------------------------------------------------
    def step(self, x, y):
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        if self.use_graph:
            if self._graph_a is None:
                # the warm-up iterations inside _capture must not disturb the weights: they only run fwd/bwd
                saved = (self.running.clone(), self.nbt.clone())
                self._capture()
                self.running.copy_(saved[0]); self.nbt.copy_(saved[1])
            self._graph_a.replay()
            self._allreduce()
            self._graph_b.replay()
        else:
            self._fwd_bwd()
            self._allreduce()
            self._optim()
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
