"""Thin torch custom-op layer over the C ABI (include/wiflow_b200.h).

Each op only validates tensors, picks the current CUDA stream and forwards raw device pointers to
libwiflow_b200.so.  The ops are registered with torch.library (namespace `wiflow_b200`) so they are visible
to the dispatcher (fake-tensor shape inference, CUDA-graph capture); autograd is wired in `block.py` /
`losses/pose_loss.py`.  There is deliberately no CPU implementation: CPU tensors raise."""
import ctypes
from typing import List

import torch

from . import _lib
from ._lib import BlockDesc


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


def _need_cuda(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError('wiflow_b200 ops run on sm_100a CUDA devices only (no CPU fallback); got a CPU tensor')
        if not t.is_contiguous():
            raise RuntimeError('wiflow_b200 ops need contiguous tensors')


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _desc(d: List[int]) -> BlockDesc:
    return BlockDesc(*d)


def out_shape(d: List[int], B: int):
    blk, cin, cout, width, _ = d
    if blk == _lib.BLOCK_MODEL:
        return (B, 15, 2)
    if blk == _lib.BLOCK_TCN:
        return (B, 240, 20)
    if blk == _lib.BLOCK_INNER_TCN:
        return (B, cout, 20)
    if blk == _lib.BLOCK_CONVBLOCK1:
        return (B, cout, 20, width)
    if blk == _lib.BLOCK_ASYMCONV:
        return (B, cout, 20, (width - 1) // 2 + 1)
    return (B, 64, 15, 20)


def workspace_bytes(d: List[int], B: int, flags: int) -> int:
    desc = _desc(d)
    n = _lib.lib().wf_workspace_bytes(ctypes.byref(desc), B, flags)
    if n == 0:
        raise RuntimeError('wf_workspace_bytes: bad block descriptor ' + _lib.lib().wf_last_error_string().decode())
    return n


def _mask_array(masks):
    if not masks:
        return None, ctypes.c_void_p(0)
    arr = (ctypes.c_void_p * len(masks))(*[m.data_ptr() for m in masks])
    return arr, ctypes.cast(arr, ctypes.c_void_p)


@torch.library.custom_op('wiflow_b200::block_forward', mutates_args=('running', 'nbt', 'workspace'))
def block_forward(x: torch.Tensor, params: torch.Tensor, running: torch.Tensor, nbt: torch.Tensor, masks: List[torch.Tensor],
                  desc: List[int], flags: int, workspace: torch.Tensor) -> torch.Tensor:
    _need_cuda(x, params, running, nbt, workspace, *masks)
    B = x.shape[0]
    y = torch.empty(out_shape(desc, B), device=x.device, dtype=torch.float32)
    d = _desc(desc)
    keep, mp = _mask_array(masks)
    with torch.cuda.device(x.device):
        rc = _lib.lib().wf_block_forward(ctypes.byref(d), _ptr(x), _ptr(params), _ptr(running), _ptr(nbt), mp, _ptr(y),
                                         _ptr(workspace), workspace.numel() * workspace.element_size(), B, flags, _stream())
    _lib.check(rc, 'wf_block_forward')
    return y


@block_forward.register_fake
def _(x, params, running, nbt, masks, desc, flags, workspace):
    return x.new_empty(out_shape(desc, x.shape[0]))


@torch.library.custom_op('wiflow_b200::block_backward', mutates_args=('workspace',))
def block_backward(x: torch.Tensor, params: torch.Tensor, masks: List[torch.Tensor], dy: torch.Tensor, desc: List[int], flags: int,
                   workspace: torch.Tensor, need_dx: bool) -> List[torch.Tensor]:
    _need_cuda(x, params, dy, workspace, *masks)
    B = x.shape[0]
    grads = torch.empty_like(params)
    dx = torch.empty_like(x) if need_dx else x.new_empty(0)
    d = _desc(desc)
    keep, mp = _mask_array(masks)
    with torch.cuda.device(x.device):
        rc = _lib.lib().wf_block_backward(ctypes.byref(d), _ptr(x), _ptr(params), mp, _ptr(dy), _ptr(grads), _ptr(dx),
                                          _ptr(workspace), workspace.numel() * workspace.element_size(), B, flags, _stream())
    _lib.check(rc, 'wf_block_backward')
    return [grads, dx]


@block_backward.register_fake
def _(x, params, masks, dy, desc, flags, workspace, need_dx):
    return [torch.empty_like(params), torch.empty_like(x) if need_dx else x.new_empty(0)]


@torch.library.custom_op('wiflow_b200::pose_loss', mutates_args=('scratch',))
def pose_loss(pred: torch.Tensor, target: torch.Tensor, loss_type: int, position_weight: float, bone_weight: float,
              scratch: torch.Tensor, want_grad: bool) -> List[torch.Tensor]:
    """returns [out3 = (total, position, bone), dpred (d total / d pred, or empty)]"""
    _need_cuda(pred, target, scratch)
    B = pred.shape[0]
    out3 = torch.empty(3, device=pred.device, dtype=torch.float32)
    dpred = torch.empty_like(pred) if want_grad else pred.new_empty(0)
    with torch.cuda.device(pred.device):
        rc = _lib.lib().wf_pose_loss(_ptr(pred), _ptr(target), B, loss_type, position_weight, bone_weight, ctypes.c_void_p(0),
                                     _ptr(dpred), _ptr(out3), _ptr(scratch), _stream())
    _lib.check(rc, 'wf_pose_loss')
    return [out3, dpred]


@pose_loss.register_fake
def _(pred, target, loss_type, position_weight, bone_weight, scratch, want_grad):
    return [pred.new_empty(3), torch.empty_like(pred) if want_grad else pred.new_empty(0)]


@torch.library.custom_op('wiflow_b200::pose_metrics', mutates_args=('scratch',))
def pose_metrics(pred: torch.Tensor, target: torch.Tensor, thresholds: List[float], use_torso_norm: bool,
                 scratch: torch.Tensor) -> torch.Tensor:
    """returns [pck(thr_0), ..., pck(thr_k-1), mpjpe] on the device"""
    _need_cuda(pred, target, scratch)
    B = pred.shape[0]
    out = torch.empty(len(thresholds) + 1, device=pred.device, dtype=torch.float32)
    thr = (ctypes.c_float * max(1, len(thresholds)))(*thresholds)
    with torch.cuda.device(pred.device):
        rc = _lib.lib().wf_pose_metrics(_ptr(pred), _ptr(target), B, thr, len(thresholds), 1 if use_torso_norm else 0,
                                        _ptr(out), _ptr(scratch), _stream())
    _lib.check(rc, 'wf_pose_metrics')
    return out


@pose_metrics.register_fake
def _(pred, target, thresholds, use_torso_norm, scratch):
    return pred.new_empty(len(thresholds) + 1)


ADAM_STATE_BYTES = 64 + 8 * 256          # WF_ADAM_STATE_BYTES of include/wiflow_b200.h


def adam_state(device) -> torch.Tensor:
    """zeroed optimizer-state block of wf_clip_adamw: step counter, clip results, per-block partial sums of the gradient norm"""
    return torch.zeros(ADAM_STATE_BYTES // 8, device=device, dtype=torch.float64)


@torch.library.custom_op('wiflow_b200::clip_adamw', mutates_args=('params', 'exp_avg', 'exp_avg_sq', 'state'))
def clip_adamw(params: torch.Tensor, grads: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, state: torch.Tensor,
               lr: float, beta1: float, beta2: float, eps: float, weight_decay: float, max_norm: float, grad_scale: float) -> None:
    _need_cuda(params, grads, exp_avg, exp_avg_sq, state)
    if state.numel() * state.element_size() < ADAM_STATE_BYTES:
        raise RuntimeError(f'clip_adamw: state must hold {ADAM_STATE_BYTES} bytes (adam_state(device) allocates it)')
    with torch.cuda.device(params.device):
        rc = _lib.lib().wf_clip_adamw(_ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq), params.numel(), _ptr(state),
                                      lr, beta1, beta2, eps, weight_decay, max_norm, grad_scale, _stream())
    _lib.check(rc, 'wf_clip_adamw')


# ---- input side: plain functions, no autograd (data preparation has no gradient) ----
def window_load(windows: torch.Tensor, idx, out: torch.Tensor = None, spans: torch.Tensor = None, stats: torch.Tensor = None,
                t_major: bool = True) -> torch.Tensor:
    """out[b] = windows[idx[b]] (idx None: identity, `out` may be `windows` itself) with the optional time-masking spans applied and
    sum / sum-of-squares of the result accumulated into `stats` (2 doubles).  windows: [N, T, C] (t_major) or [N, C, T]."""
    _need_cuda(windows, idx, out, spans, stats)
    if windows.dim() != 3 or windows.dtype != torch.float32:
        raise RuntimeError(f'window_load needs a float32 [N, {"T, C" if t_major else "C, T"}] tensor, got {list(windows.shape)} {windows.dtype}')
    N = windows.shape[0]
    T, C = (windows.shape[1], windows.shape[2]) if t_major else (windows.shape[2], windows.shape[1])
    B = N if idx is None else idx.numel()
    if idx is not None and idx.dtype != torch.int64:
        raise RuntimeError('window indices must be int64')
    if out is None:
        out = torch.empty((B,) + tuple(windows.shape[1:]), device=windows.device, dtype=torch.float32)
    elif out.shape[0] != B or out.shape[1:] != windows.shape[1:] or out.dtype != torch.float32:
        raise RuntimeError(f'window_load: output of shape {list(out.shape)} does not hold {B} windows of {list(windows.shape[1:])}')
    if spans is not None and (spans.dtype != torch.int32 or spans.numel() != 4 * B):
        raise RuntimeError('spans must be an int32 [B, 2, 2] tensor of (start, length)')
    if stats is not None and (stats.dtype != torch.float64 or stats.numel() < 2):
        raise RuntimeError('stats must hold 2 doubles')
    if B == 0:                                   # an empty batch is an empty tensor, not an error (nothing to launch)
        return out
    with torch.cuda.device(windows.device):
        rc = _lib.lib().wf_window_load(_ptr(windows), N, _ptr(idx), _ptr(out), B, C, T, 1 if t_major else 0, _ptr(spans), _ptr(stats), _stream())
    _lib.check(rc, 'wf_window_load')
    return out


def noise_scale(x: torch.Tensor, noise, level: float, scale: float, stats, out: torch.Tensor = None) -> torch.Tensor:
    """(x + (noise*level)*std(x)) * scale; noise None: scaling only.  stats: the 2 doubles window_load filled for x."""
    _need_cuda(x, noise, stats, out)
    if x.dtype != torch.float32 or (noise is not None and (noise.dtype != torch.float32 or noise.numel() != x.numel())):
        raise RuntimeError('noise_scale needs float32 tensors of equal size')
    if out is None:
        out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = _lib.lib().wf_noise_scale(_ptr(x), _ptr(noise), _ptr(out), x.numel(), float(level), float(scale), _ptr(stats), x.numel(), _stream())
    _lib.check(rc, 'wf_noise_scale')
    return out


def dropout_masks(shapes, ps, seed: int, state: torch.Tensor, device) -> List[torch.Tensor]:
    """Dropout masks of one step in one launch (wf_dropout_masks): one float32 tensor per (shape, p), entries 0 or 1/(1-p).
    state: 2 int64 on the device, zeroed once (the draw counter advances on the device, so graph replays draw fresh masks)."""
    _need_cuda(state)
    if state.dtype != torch.int64 or state.numel() < 2:
        raise RuntimeError('dropout_masks needs an int64 state tensor of 2 elements')
    outs = [torch.empty(*sh, device=device, dtype=torch.float32) for sh in shapes]
    n = len(outs)
    if n == 0:
        return outs
    ptrs = (ctypes.c_void_p * n)(*[o.data_ptr() for o in outs])
    numel = (ctypes.c_longlong * n)(*[o.numel() for o in outs])
    pp = (ctypes.c_float * n)(*[float(p) for p in ps])
    with torch.cuda.device(device):
        rc = _lib.lib().wf_dropout_masks(ptrs, numel, pp, n, int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(state), _stream())
    _lib.check(rc, 'wf_dropout_masks')
    return outs


def keypoint_batch(frames: torch.Tensor, idx, clean: bool = True, out: torch.Tensor = None) -> torch.Tensor:
    """y[b] = frames[idx[b]] ([K,2]; zeros for indices outside the array) with all-zero joints replaced by the mean of the others."""
    _need_cuda(frames, idx, out)
    if frames.dim() != 3 or frames.shape[2] != 2 or frames.dtype != torch.float32:
        raise RuntimeError(f'keypoint_batch needs a float32 [F, K, 2] tensor, got {list(frames.shape)} {frames.dtype}')
    if idx is not None and idx.dtype != torch.int64:
        raise RuntimeError('frame indices must be int64')
    B = frames.shape[0] if idx is None else idx.numel()
    K = frames.shape[1]
    if out is None:
        out = torch.empty(B, K, 2, device=frames.device, dtype=torch.float32)
    with torch.cuda.device(frames.device):
        rc = _lib.lib().wf_keypoint_batch(_ptr(frames), frames.shape[0], _ptr(idx), _ptr(out), B, K, 1 if clean else 0, _stream())
    _lib.check(rc, 'wf_keypoint_batch')
    return out


def keypoint_sequences_(frames: torch.Tensor, seq_off: torch.Tensor) -> torch.Tensor:
    """in place: zero joints interpolated along each sequence frames[seq_off[s]:seq_off[s+1]] (dataset.py:159-206)"""
    _need_cuda(frames, seq_off)
    if frames.dim() != 3 or frames.shape[2] != 2 or frames.dtype != torch.float32 or seq_off.dtype != torch.int64:
        raise RuntimeError('keypoint_sequences_ needs float32 [F, K, 2] frames and int64 offsets')
    with torch.cuda.device(frames.device):
        rc = _lib.lib().wf_keypoint_sequences(_ptr(frames), _ptr(seq_off), seq_off.numel() - 1, frames.shape[1], _stream())
    _lib.check(rc, 'wf_keypoint_sequences')
    return frames
