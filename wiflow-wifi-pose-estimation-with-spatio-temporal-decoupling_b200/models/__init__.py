"""Model components (mirrors the reference's models/__init__.py:5-8 exports, plus the aliases it lost)."""
from .attention import AxialAttention, DualAxialAttention
from .convnet import AsymmetricConvBlock, ConvBlock1
from .pose_model import WiFlow, WiFlowPoseModel
from .tcn import Chomp1d, InnerGroupedTemporalBlock, TemporalBlock, TemporalConvNet

__all__ = ['WiFlowPoseModel', 'WiFlow', 'AxialAttention', 'DualAxialAttention', 'TemporalConvNet', 'TemporalBlock',
           'InnerGroupedTemporalBlock', 'Chomp1d', 'AsymmetricConvBlock', 'ConvBlock1']
