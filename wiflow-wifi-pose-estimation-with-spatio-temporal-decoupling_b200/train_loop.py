"""Training-loop glue of the reference (`train.py:48-400`) on top of the fused steps of `engine.py`.

What is kept from the reference loop: gradient accumulation over `accumulation_steps` micro-batches with the loss divided by
that count (`train.py:80,200,243-257`), `clip_grad_norm_(1.0)` + AdamW(lr, wd 5e-5) per optimizer step, the epoch statistics
(loss, position/bone parts, MPJPE, PCK@0.2, PCK@0.5), a validation pass in eval mode each epoch (`:288-330`),
`ReduceLROnPlateau(mode='min', factor=0.5, patience=3, min_lr=lr/1000, cooldown=1, threshold=1e-4)` stepped on the validation
MPJPE (`:112-121,358`), the best-by-validation-MPJPE `state_dict` (reference format, 295 keys) kept / saved (`:361-376`) and
early stopping after `patience` epochs without improvement (`:379-385`).

What changes: all per-step statistics stay on the device (the reference pays seven `.item()` synchronisations per step); one
device->host copy per epoch reads them.  The AMP / GradScaler path of the reference (`:96,196-202,234-237`) is NOT reproduced:
this implementation computes in fp32 (3xTF32 on the tensor cores), which is what the reference validates and tests in."""
import copy

import torch
import torch.distributed as dist
from torch.optim.lr_scheduler import ReduceLROnPlateau

from . import _lib, ops
from .engine import MODEL_DESC, InferStep, TrainStep


class _Eval:
    """eval-mode forward + PoseLoss + metrics with device-side accumulators"""

    def __init__(self, model, batch_size, loss, thresholds):
        self.inf = InferStep(model, batch_size)
        self.loss, self.thresholds = loss, tuple(thresholds)
        dev = self.inf.dev
        self.sums = torch.zeros(5 + len(self.thresholds), device=dev, dtype=torch.float64)
        self.loss_scratch = torch.zeros(2, device=dev, dtype=torch.float64)
        self.metric_scratch = torch.zeros(16, device=dev, dtype=torch.float64)

    def step(self, x, y):
        inf = self.inf
        B = x.shape[0]
        x = x.to(inf.dev, non_blocking=True).contiguous()
        y = y.to(inf.dev, non_blocking=True).contiguous()
        if B == inf.B:
            pred = inf.step(x)
        else:
            if B > inf.B:
                raise RuntimeError(f'batch of {B} windows exceeds the workspace sized for {inf.B}')
            pred = ops.block_forward(x, inf.params, inf.running, inf.nbt, [], MODEL_DESC, 0, inf.ws)
        out3, _ = ops.pose_loss(pred, y, self.loss[0], self.loss[1], self.loss[2], self.loss_scratch, False)
        mt = ops.pose_metrics(pred, y, list(self.thresholds), True, self.metric_scratch)
        self.sums[:3].add_(out3.double(), alpha=float(B))
        self.sums[3:4].add_(mt[-1:].double(), alpha=float(B))
        self.sums[4:4 + len(self.thresholds)].add_(mt[:-1].double(), alpha=float(B))
        self.sums[-1:].add_(float(B))

    def read(self):
        v = self.sums.tolist()
        n = v[-1]
        names = ['loss', 'position', 'bone', 'mpjpe'] + [f'pck@{t:g}' for t in self.thresholds]
        out = {k: (a / n if n > 0 else float('inf')) for k, a in zip(names, v[:-1])}
        out['windows'] = int(n)
        return out


class Trainer:
    """fit(train_batches, val_batches, n_epochs): `*_batches` are callables returning an iterable of (x [B,540,20], y [B,15,2])
    tensors (host pinned or device) for one epoch; the last batch of an epoch may be smaller."""

    def __init__(self, model, batch_size, lr=1e-4, weight_decay=5e-5, accumulation_steps=1, patience=5, thresholds=(0.2, 0.5),
                 position_weight=1.0, bone_weight=0.2, loss_type='smooth_l1', process_group=None, use_cuda_graph=True):
        self.model = model
        self.thresholds = tuple(thresholds)
        self.ts = TrainStep(model, batch_size, lr=lr, weight_decay=weight_decay, position_weight=position_weight,
                            bone_weight=bone_weight, loss_type=loss_type, process_group=process_group,
                            use_cuda_graph=use_cuda_graph, accumulation_steps=accumulation_steps, metric_thresholds=thresholds)
        self._loss = (_lib.LOSS_TYPES[loss_type], float(position_weight), float(bone_weight))
        self._eval = None
        self.batch_size = int(batch_size)
        # the reference's scheduler, driven through a one-parameter stand-in optimizer so its semantics are torch's own
        self._lr_holder = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr)
        self.scheduler = ReduceLROnPlateau(self._lr_holder, mode='min', factor=0.5, patience=3, min_lr=lr / 1000, cooldown=1,
                                           threshold=1e-4)
        self.patience = patience
        self.history = {k: [] for k in ('train_loss', 'val_loss', 'train_position_loss', 'train_bone_loss', 'train_mpe', 'val_mpe',
                                        'train_pck', 'val_pck', 'train_pck50', 'val_pck50', 'lr')}
        self.best_val_mpe = float('inf')
        self.best_state = None

    @property
    def lr(self):
        return self._lr_holder.param_groups[0]['lr']

    def train_epoch(self, batches):
        self.model.train()
        ts = self.ts
        ts.reset_sums()
        for x, y in batches:
            ts.step(x, y)
        ts.flush()
        return ts.read_sums()

    def validate(self, batches):
        if self._eval is None:
            self._eval = _Eval(self.model, self.batch_size, self._loss, self.thresholds)
        self.model.eval()
        self._eval.sums.zero_()
        for x, y in batches:
            self._eval.step(x, y)
        out = self._eval.read()
        self.model.train()
        return out

    def _all_ranks(self, stats):
        """window-weighted mean of per-rank epoch statistics over the process group: every rank then takes the same scheduler,
        best-checkpoint and early-stop decisions (a rank that stopped alone would leave the others hanging in the all-reduce)"""
        ts = self.ts
        if ts.world <= 1:
            return stats
        keys = [k for k in stats if k != 'windows']
        n = float(stats['windows'])
        v = torch.tensor([stats[k] * n if n > 0 else 0.0 for k in keys] + [n], device=ts.dev, dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=ts.pg)
        v = v.tolist()
        tot = v[-1]
        out = {k: (x / tot if tot > 0 else float('inf')) for k, x in zip(keys, v[:-1])}
        out['windows'] = int(tot)
        return out

    def fit(self, train_batches, val_batches, n_epochs, checkpoint_path=None, log=None):
        bad_epochs = 0
        t0, t1 = (f'pck@{t:g}' for t in self.thresholds[:2])
        for epoch in range(n_epochs):
            tr = self._all_ranks(self.train_epoch(train_batches()))
            if self.ts.world > 1:
                # BatchNorm statistics are local per rank (nn.DataParallel semantics, train.py:91-93: replica 0's persist):
                # validate, checkpoint and continue with rank 0's running buffers on every rank
                src = dist.get_global_rank(self.ts.pg, 0) if self.ts.pg is not None else 0
                dist.broadcast(self.ts.running, src=src, group=self.ts.pg)
                dist.broadcast(self.ts.nbt, src=src, group=self.ts.pg)
            va = self._all_ranks(self.validate(val_batches()))
            h = self.history
            h['train_loss'].append(tr['loss']); h['val_loss'].append(va['loss'])
            h['train_position_loss'].append(tr['position']); h['train_bone_loss'].append(tr['bone'])
            h['train_mpe'].append(tr['mpjpe']); h['val_mpe'].append(va['mpjpe'])
            h['train_pck'].append(tr[t0]); h['val_pck'].append(va[t0])
            h['train_pck50'].append(tr[t1]); h['val_pck50'].append(va[t1])
            h['lr'].append(self.lr)
            if log:
                log(f"epoch {epoch + 1}/{n_epochs}  train loss {tr['loss']:.4f} mpe {tr['mpjpe']:.4f}  "
                    f"val loss {va['loss']:.4f} mpe {va['mpjpe']:.4f} pck@0.2 {va[t0]:.4f}  lr {self.lr:.6f}")
            self.scheduler.step(va['mpjpe'])
            self.ts.set_lr(self.lr)
            if va['mpjpe'] < self.best_val_mpe:
                self.best_val_mpe = va['mpjpe']
                self.best_state = copy.deepcopy(self.model.state_dict())
                if checkpoint_path and (self.ts.world <= 1 or dist.get_rank(self.ts.pg) == 0):
                    torch.save(self.best_state, checkpoint_path)
                bad_epochs = 0
            else:
                bad_epochs += 1
                if bad_epochs >= self.patience:
                    break
        if self.best_state is not None:
            self.model.load_state_dict(self.best_state)
        return self.history
