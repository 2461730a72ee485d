"""Drop-in for the reference `models/convnet.py` (AsymmetricConvBlock :4-38, ConvBlock1 :41-74): three (1x3)
convolutions along the feature axis with BatchNorm/SiLU/Dropout2d, a 1x1 shortcut, add, SiLU.  Same child names
(`block.{0,1,4,5,8,9}`, `downsample.{0,1}`, `activation`) and state_dict keys; arithmetic in csrc/wf_conv.cu."""
import torch.nn as nn

from .. import _lib
from ..block import WFBlock


def _main_path(cin, cout, first_stride, p):
    layers, c = [], cin
    for i in range(3):
        layers.append(nn.Conv2d(c, cout, kernel_size=(1, 3), stride=(1, first_stride if i == 0 else 1), padding=(0, 1)))
        layers.append(nn.BatchNorm2d(cout))
        if i < 2:
            layers += [nn.SiLU(inplace=True), nn.Dropout2d(p)]
        c = cout
    return nn.Sequential(*layers)


class _ResidualRowConv(WFBlock):
    _stride = 1
    _block_id = _lib.BLOCK_CONVBLOCK1

    def __init__(self, in_channels, out_channels, dropout=0.3):
        super().__init__()
        self._cin, self._cout = in_channels, out_channels
        self.block = _main_path(in_channels, out_channels, self._stride, dropout)
        self.downsample = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=(1, self._stride), bias=False),
            nn.BatchNorm2d(out_channels))
        self.activation = nn.SiLU(inplace=True)

    def _wf_desc_key(self):
        return (self._block_id, self._cin, self._cout, self._width, 0)

    def _wf_input_shape(self):
        return (self._cin, 20, self._width)

    def _wf_dropout_sites(self):
        return [(self.block[3].p, 'plane', self._cout), (self.block[7].p, 'plane', self._cout)]

    def forward(self, x):
        if x.dim() != 4 or x.shape[2] != 20:
            raise RuntimeError(f'{type(self).__name__}: expected [B, {self._cin}, 20, W] (20 time rows), got {list(x.shape)}')
        self._width = x.shape[3]
        return super().forward(x)


class AsymmetricConvBlock(_ResidualRowConv):
    _stride = 2
    _block_id = _lib.BLOCK_ASYMCONV


class ConvBlock1(_ResidualRowConv):
    _stride = 1
    _block_id = _lib.BLOCK_CONVBLOCK1
