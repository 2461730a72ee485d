"""Drop-in for the reference `models/attention.py` (AxialAttention :7-80, DualAxialAttention :83-98).
Children `qkv_transform, bn_qkv, bn_similarity, bn_output`; arithmetic in csrc/wf_attn.cu + csrc/wf_conv.cu.
The kernels implement the configuration the model uses: 64 planes, 8 groups, stride 1, a 15x20 grid."""
import math

import torch.nn as nn

from .. import _lib
from ..block import WFBlock


def _check(in_planes, out_planes, groups, stride):
    assert (in_planes % groups == 0) and (out_planes % groups == 0)
    if (in_planes, out_planes, groups, stride) != (64, 64, 8, 1):
        raise ValueError('the B200 axial-attention kernels implement in_planes=out_planes=64, groups=8, stride=1 '
                         '(models/pose_model.py:39-41)')


class AxialAttention(WFBlock):
    def __init__(self, in_planes, out_planes, groups=8, stride=1, bias=False, width=False):
        _check(in_planes, out_planes, groups, stride)
        super().__init__()
        self.in_planes, self.out_planes, self.groups = in_planes, out_planes, groups
        self.group_planes = out_planes // groups
        self.stride, self.bias, self.width = stride, bias, width
        self.qkv_transform = nn.Conv1d(in_planes, out_planes * 3, kernel_size=1, stride=1, padding=0, bias=False)
        self.bn_qkv = nn.BatchNorm1d(out_planes * 3)
        self.bn_similarity = nn.BatchNorm2d(groups)
        self.bn_output = nn.BatchNorm1d(out_planes)
        nn.init.normal_(self.qkv_transform.weight.data, 0, math.sqrt(1. / in_planes))      # attention.py:34-35

    def _wf_desc_key(self):
        return (_lib.BLOCK_AXIAL_W if self.width else _lib.BLOCK_AXIAL_H, 0, 0, 0, 0)

    def _wf_input_shape(self):
        return (64, 15, 20)


class DualAxialAttention(WFBlock):
    def __init__(self, in_planes, out_planes, groups=8, stride=1, bias=False):
        super().__init__()
        self.width_axis = AxialAttention(in_planes, out_planes, groups, stride, bias, width=True)
        self.height_axis = AxialAttention(out_planes, out_planes, groups, stride, bias, width=False)

    def _wf_desc_key(self):
        return (_lib.BLOCK_DUAL_AXIAL, 0, 0, 0, 0)

    def _wf_input_shape(self):
        return (64, 15, 20)
