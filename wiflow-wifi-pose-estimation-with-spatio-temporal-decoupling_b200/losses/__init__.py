"""Loss functions (mirrors the reference's losses/__init__.py:5)."""
from .pose_loss import PoseLoss

__all__ = ['PoseLoss']
