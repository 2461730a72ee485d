"""Drop-in for the reference `losses/pose_loss.py:5-88` PoseLoss: position term + 0.2 x bone-length term, one CUDA
kernel for value and gradient (csrc/wf_elem.cu pose_loss_kernel) instead of ~35 tiny ops."""
import torch
import torch.nn as nn

from .. import _lib, ops

BONE_CONNECTIONS = [(0, 1), (1, 8), (1, 2), (2, 3), (3, 4), (1, 5), (5, 6), (6, 7), (8, 9), (8, 12),
                    (9, 10), (10, 11), (12, 13), (13, 14)]      # pose_loss.py:20-24


class _PoseLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, loss_type, pw, bw, scratch):
        want = ctx.needs_input_grad[0]
        out3, dpred = ops.pose_loss(pred, target, loss_type, pw, bw, scratch, want)
        if want:
            ctx.save_for_backward(dpred)
        ctx.mark_non_differentiable(out3)
        return out3[0].clone(), out3

    @staticmethod
    def backward(ctx, g_total, _g3):
        (dpred,) = ctx.saved_tensors
        return dpred * g_total, None, None, None, None, None


class PoseLoss(nn.Module):
    def __init__(self, position_weight: float = 1.0, bone_weight: float = 0.2, loss_type: str = 'smooth_l1'):
        super().__init__()
        self.position_weight = position_weight
        self.bone_weight = bone_weight
        self.loss_type = loss_type
        self.bone_connections = list(BONE_CONNECTIONS)
        self._scratch = {}

    def compute_bone_lengths(self, keypoints):
        """[.., 15, 2] -> [.., 14] (pose_loss.py:26-33); plain torch, only for callers that use it directly."""
        s = torch.tensor([a for a, _ in self.bone_connections], device=keypoints.device)
        e = torch.tensor([b for _, b in self.bone_connections], device=keypoints.device)
        v = keypoints.index_select(-2, e) - keypoints.index_select(-2, s)
        return torch.sqrt((v ** 2).sum(-1) + 1e-8)

    def forward_device(self, pred, target):
        """(total with grad, out3 device tensor [total, position, bone]) without any host synchronisation."""
        if self.loss_type not in _lib.LOSS_TYPES:
            raise ValueError(f"Unknown loss type: {self.loss_type}")
        B = pred.shape[0]
        if pred.dim() == 2 and pred.shape[1] == 30:
            pred = pred.reshape(B, 15, 2)
        if target.dim() == 2 and target.shape[1] == 30:
            target = target.reshape(B, 15, 2)
        if tuple(pred.shape[1:]) != (15, 2) or pred.shape != target.shape:
            raise RuntimeError(f'PoseLoss: expected [B,15,2] (or [B,30]) tensors, got {list(pred.shape)} and {list(target.shape)}')
        key = (pred.device.type, pred.device.index)
        if key not in self._scratch:
            self._scratch[key] = torch.zeros(2, device=pred.device, dtype=torch.float64)
        total, out3 = _PoseLossFunction.apply(pred.float().contiguous(), target.float().contiguous(),
                                              _lib.LOSS_TYPES[self.loss_type], float(self.position_weight),
                                              float(self.bone_weight), self._scratch[key])
        return total, out3

    def forward(self, pred, target):
        total, out3 = self.forward_device(pred, target)
        vals = out3.tolist()              # the reference contract returns Python floats (pose_loss.py:83-86): one sync
        return total, {'position': vals[1], 'bone': vals[2]}
