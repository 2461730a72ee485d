"""Base class of the drop-in modules: keeps the reference's module tree (same child names, same state_dict keys,
same init RNG stream) purely as a parameter container, and routes forward/backward through the CUDA library.

Parameters of a block live in ONE flat fp32 buffer in named_parameters() order (what the C ABI consumes); the
individual nn.Parameters are views into it, so optimizers, clip_grad_norm_, state_dict()/load_state_dict() and
DataParallel-style `.module` unwrapping keep working.  `.to()/.cuda()` break the aliasing; it is restored lazily
on the next forward."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.modules.batchnorm import _BatchNorm

from . import _lib, ops


def _is_flat(ts):
    base = ts[0]
    sp = base.untyped_storage().data_ptr()
    off = 0
    for t in ts:
        if (t.device != base.device or t.dtype != base.dtype or not t.is_contiguous()
                or t.untyped_storage().data_ptr() != sp or t.data_ptr() != base.data_ptr() + off * base.element_size()):
            return False
        off += t.numel()
    return True


def _flatten(ts, dtype):
    total = sum(t.numel() for t in ts)
    flat = torch.empty(total, device=ts[0].device, dtype=dtype)
    off = 0
    for t in ts:
        n = t.numel()
        flat[off:off + n].copy_(t.detach().reshape(-1))
        t.data = flat[off:off + n].view(t.shape)
        off += n
    return flat


def _flat_view(ts):
    base = ts[0].detach()
    total = sum(t.numel() for t in ts)
    return base.as_strided((total,), (1,), base.storage_offset())


_ONES = {}

class _BlockFunction(torch.autograd.Function):
    """autograd bridge: forward saves the library workspace, backward returns per-parameter views of one flat gradient."""

    @staticmethod
    def forward(ctx, blk, x, *params):
        flat, running, nbt = blk._wf_state()
        train = blk.training
        need_grad = any(ctx.needs_input_grad)          # eval mode too: BatchNorm back-propagates as the fixed affine it is
        flags = (_lib.FLAG_TRAIN if train else 0) | (_lib.FLAG_SAVE if need_grad else 0)
        desc = list(blk._wf_desc_key())
        B = x.shape[0]
        masks = blk._wf_masks(B, x.device) if train else []
        ws = torch.empty(ops.workspace_bytes(desc, B, flags), device=x.device, dtype=torch.uint8)
        x = x.contiguous()
        y = ops.block_forward(x, flat, running, nbt, masks, desc, flags, ws)
        if need_grad:
            ctx.blk, ctx.desc, ctx.flags, ctx.ws, ctx.masks, ctx.x, ctx.flat = blk, desc, flags, ws, masks, x, flat
            ctx.shapes = [p.shape for p in params]
        return y

    @staticmethod
    def backward(ctx, dy):
        need_dx = ctx.needs_input_grad[1]
        grads, dx = ops.block_backward(ctx.x, ctx.flat, ctx.masks, dy.contiguous(), ctx.desc, ctx.flags, ctx.ws, need_dx)
        out, off = [], 0
        for shp in ctx.shapes:
            n = shp.numel()
            out.append(grads[off:off + n].view(shp))
            off += n
        ctx.ws = None
        return (None, dx if need_dx else None, *out)


class WFBlock(nn.Module):
    """A reference-compatible module whose forward/backward run in libwiflow_b200.so."""

    def _wf_desc_key(self):           # (block, cin, cout, width, dilation)
        raise NotImplementedError

    def _wf_dropout_sites(self):      # [(p, kind, C)] in the reference's forward order; kind 'elem' ([B,C,20]) or 'plane' ([B,C])
        return []

    def _wf_input_shape(self):        # per-sample input shape, for error messages
        raise NotImplementedError

    # -- flat state ---------------------------------------------------------------------------------
    def _wf_state(self):
        params = list(self.parameters())
        if not _is_flat(params) or params[0].dtype != torch.float32:
            _flatten(params, torch.float32)
        bns = [m for m in self.modules() if isinstance(m, _BatchNorm)]
        run = [t for m in bns for t in (m.running_mean, m.running_var)]
        if not _is_flat(run) or run[0].dtype != torch.float32:
            _flatten(run, torch.float32)
        nbts = [m.num_batches_tracked for m in bns]
        if not _is_flat(nbts):
            _flatten(nbts, torch.int64)
        return _flat_view(params), _flat_view(run), _flat_view(nbts)

    def _wf_masks(self, B, device, rng_state=None, seed=0):
        """the step's dropout masks in the reference's forward order.  Default: torch's generator (the draws nn.Dropout /
        nn.Dropout2d would make, so a seeded run matches the oracle); with rng_state (2 int64 on the device): the library's
        one-launch Philox generator (ops.dropout_masks) -- different stream of random numbers, same distribution"""
        sites = self._wf_dropout_sites()
        if all(p == 0 for p, _, _ in sites):
            return []
        if rng_state is not None:
            shapes = [(B, C, 20) if kind == 'elem' else (B, C) for _, kind, C in sites]
            return ops.dropout_masks(shapes, [p for p, _, _ in sites], seed, rng_state, device)
        masks = []
        cache = _ONES       # the all-ones inputs of the draws, filled once per shape (not per step); module-level: never pickled / deep-copied
        for p, kind, C in sites:
            shape = (B, C, 20) if kind == 'elem' else (B, C, 1, 1)
            key = (shape, str(device))
            ones = cache.get(key)
            if ones is None:
                if len(cache) > 64:
                    cache.clear()
                ones = cache[key] = torch.ones(*shape, device=device)
            if kind == 'elem':
                masks.append(F.dropout(ones, p, True))
            else:
                masks.append(F.dropout2d(ones, p, True).view(B, C))
        return masks

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError(f'{type(self).__name__}: the B200 kernels need CUDA tensors (sm_100a); there is no CPU fallback')
        if x.dtype != torch.float32:
            x = x.float()
        exp = tuple(self._wf_input_shape())
        if tuple(x.shape[1:]) != exp:
            raise RuntimeError(f'{type(self).__name__}: expected input [B, {", ".join(map(str, exp))}], got {list(x.shape)}')
        if x.shape[0] == 0:
            # the reference's modules on an empty batch: eval mode returns an empty tensor, train-mode BatchNorm refuses
            if self.training:
                raise ValueError('Expected more than 1 value per channel when training, got an empty batch')
            return x.new_empty(ops.out_shape(list(self._wf_desc_key()), 0))
        return _BlockFunction.apply(self, x, *self.parameters())
