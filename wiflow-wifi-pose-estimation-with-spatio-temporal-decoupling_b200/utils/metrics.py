"""Drop-in for the reference `utils/metrics.py:3-47`: PCK (torso- or shoulder-normalised) and MPJPE, computed by one
CUDA kernel (csrc/wf_elem.cu metrics_kernel) and read back with a single device->host copy."""
import torch

from .. import ops

_scratch = {}


def _prep(pred, target):
    B = pred.shape[0]
    if pred.dim() == 2 and pred.shape[1] == 30:
        pred = pred.reshape(B, 15, 2)
        target = target.reshape(B, 15, 2)
    if tuple(pred.shape[1:]) != (15, 2) or pred.shape != target.shape:
        raise RuntimeError(f'expected [B,15,2] (or [B,30]) tensors, got {list(pred.shape)} and {list(target.shape)}')
    key = (pred.device.type, pred.device.index)
    if key not in _scratch:
        _scratch[key] = torch.zeros(16, device=pred.device, dtype=torch.float64)
    return pred.detach().float().contiguous(), target.detach().float().contiguous(), _scratch[key]


def pose_metrics_device(pred, target, thresholds=(0.2,), use_torso_norm=True):
    """Device tensor [pck(thr) for thr in thresholds] + [mpjpe]; no host synchronisation."""
    pred, target, scratch = _prep(pred, target)
    out = []
    thr = [float(t) for t in thresholds]
    for i in range(0, max(len(thr), 1), 8):        # the kernel takes up to 8 thresholds per launch
        out.append(ops.pose_metrics(pred, target, thr[i:i + 8], bool(use_torso_norm), scratch))
    if len(out) == 1:
        return out[0]
    return torch.cat([o[:-1] for o in out] + [out[-1][-1:]])


def calculate_pck(pred, target, thresholds=[0.2], use_torso_norm=True):
    vals = pose_metrics_device(pred, target, thresholds, use_torso_norm).tolist()
    return {t: v for t, v in zip(thresholds, vals[:-1])}


def calculate_mpjpe(pred, target):
    return pose_metrics_device(pred, target, (), True).tolist()[-1]
