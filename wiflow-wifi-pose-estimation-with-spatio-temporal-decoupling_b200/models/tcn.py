"""Drop-in for the reference `models/tcn.py` (Chomp1d :6-12, InnerGroupedTemporalBlock :14-74, TemporalBlock :76-97).

Same constructor signatures, child names and state_dict keys; the arithmetic is the fused implicit-GEMM kernels
of csrc/wf_conv.cu (causal taps are column shifts with t - k*d < 0 masked, so no padded columns and no Chomp copy)."""
import torch.nn as nn

from .. import _lib
from ..block import WFBlock

GROUPS = 20      # hard-coded in the reference (tcn.py:18)


class Chomp1d(nn.Module):
    """Kept for API compatibility (tcn.py:6-12): drops the last `chomp_size` steps.  The fused TCN kernels never
    materialise the padded columns, so this module is only used when someone calls it directly."""

    def __init__(self, chomp_size):
        super().__init__()
        self.chomp_size = chomp_size

    def forward(self, x):
        return x[:, :, :x.shape[2] - self.chomp_size].contiguous()


class InnerGroupedTemporalBlock(WFBlock):
    def __init__(self, n_inputs, n_outputs, kernel_size, stride, dilation, padding, dropout=0.2, attention_type='none'):
        super().__init__()
        if kernel_size != 3 or stride != 1 or padding != (kernel_size - 1) * dilation:
            raise ValueError('the B200 TCN kernel implements the reference configuration: kernel_size=3, stride=1, '
                             'padding=(kernel_size-1)*dilation (tcn.py:88-91)')
        self.groups = GROUPS
        self._cin, self._cout, self._dil, self._p = n_inputs, n_outputs, dilation, dropout
        c = n_inputs
        for i, (stage, cout) in enumerate((('1', n_outputs), ('2', n_outputs)), 1):
            self.add_module(f'conv{stage}_group', nn.Conv1d(c, c, kernel_size, stride=1, padding=padding, dilation=dilation,
                                                          groups=GROUPS, bias=False))
            self.add_module(f'chomp{stage}', Chomp1d(padding) if padding > 0 else nn.Identity())
            self.add_module(f'bn{stage}_group', nn.BatchNorm1d(c))
            self.add_module(f'relu{stage}_group', nn.SiLU(inplace=True))
            self.add_module(f'conv{stage}_pw', nn.Conv1d(c, cout, 1, bias=False))
            self.add_module(f'bn{stage}_pw', nn.BatchNorm1d(cout))
            self.add_module(f'relu{stage}_pw', nn.SiLU(inplace=True))
            self.add_module(f'dropout{stage}', nn.Dropout(dropout))
            c = cout
        if n_inputs != n_outputs:
            self.downsample = nn.Sequential(nn.Conv1d(n_inputs, n_outputs, 1, bias=False), nn.BatchNorm1d(n_outputs))
        else:
            self.downsample = nn.Identity()

    def _wf_desc_key(self):
        return (_lib.BLOCK_INNER_TCN, self._cin, self._cout, 0, self._dil)

    def _wf_input_shape(self):
        return (self._cin, 20)

    def _wf_dropout_sites(self):
        return [(self.dropout1.p, 'elem', self._cout), (self.dropout2.p, 'elem', self._cout)]


class TemporalBlock(WFBlock):
    def __init__(self, num_inputs, num_channels, kernel_size=3, dropout=0.2, attention_type='none'):
        super().__init__()
        chans = [num_inputs] + list(num_channels)
        self._chans = chans
        self.network = nn.Sequential(*[
            InnerGroupedTemporalBlock(chans[i], chans[i + 1], kernel_size, stride=1, dilation=2 ** i,
                                      padding=(kernel_size - 1) * 2 ** i, dropout=dropout, attention_type=attention_type)
            for i in range(len(num_channels))])

    def _wf_desc_key(self):
        if self._chans != [540, 540, 440, 340, 240]:
            return None
        return (_lib.BLOCK_TCN, 0, 0, 0, 0)

    def _wf_input_shape(self):
        return (self._chans[0], 20)

    def _wf_dropout_sites(self):
        return [s for blk in self.network for s in blk._wf_dropout_sites()]

    def forward(self, x):
        if self._wf_desc_key() is None:        # non-reference channel plan: chain the per-block kernels
            return self.network(x)
        return super().forward(x)


TemporalConvNet = TemporalBlock     # the name the reference's models/__init__.py:7 still exports
