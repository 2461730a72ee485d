"""Drop-in for the reference `models/pose_model.py:9-97` WiFlowPoseModel: [B,540,20] CSI -> [B,15,2] keypoints.

The module tree (tcn, up, residual_blocks, attention, decoder, avg_pool) only holds state; forward and backward of
the whole network run as one chained schedule of sm_100a kernels (csrc/wf_model.cu), activations in the internal
[channel][position][b*20+t] layout so the three layout changes of the reference forward (:79,87,95) cost nothing."""
import torch.nn as nn

from .. import _lib
from ..block import WFBlock
from .attention import DualAxialAttention
from .convnet import AsymmetricConvBlock, ConvBlock1
from .tcn import TemporalBlock


class WiFlowPoseModel(WFBlock):
    def __init__(self, dropout=0.3):
        super().__init__()
        self.tcn = TemporalBlock(num_inputs=540, num_channels=[540, 440, 340, 240], kernel_size=3, dropout=dropout,
                                 attention_type='none')
        self.up = ConvBlock1(1, 8)
        widths = [8, 8, 16, 32, 64]
        self.residual_blocks = nn.ModuleList(AsymmetricConvBlock(a, b) for a, b in zip(widths[:-1], widths[1:]))
        self.attention = DualAxialAttention(in_planes=64, out_planes=64, groups=8)
        self.decoder = nn.Sequential(nn.Conv2d(64, 32, kernel_size=3, padding=1), nn.BatchNorm2d(32), nn.SiLU(inplace=True),
                                     nn.Conv2d(32, 2, kernel_size=1), nn.BatchNorm2d(2), nn.SiLU(inplace=True))
        self.avg_pool = nn.AdaptiveAvgPool2d((15, 1))
        self._initialize_weights()

    def _initialize_weights(self):
        """Same initial distribution and RNG order as pose_model.py:57-69 (Conv1d: kaiming-normal fan_out, which also
        overrides the attention's own qkv init; BatchNorm1d: 1/0; Conv2d keeps torch's default)."""
        for m in self.modules():
            if isinstance(m, nn.Conv1d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm1d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
            # (the reference also lists nn.Linear / nn.LayerNorm; the model has neither, and constant_ draws no random numbers)

    def _wf_desc_key(self):
        return (_lib.BLOCK_MODEL, 0, 0, 0, 0)

    def _wf_input_shape(self):
        return (540, 20)

    def _wf_dropout_sites(self):
        sites = self.tcn._wf_dropout_sites()
        for blk in [self.up] + list(self.residual_blocks):
            sites += blk._wf_dropout_sites()
        return sites


WiFlow = WiFlowPoseModel      # the name BASELINE.json's north_star uses
