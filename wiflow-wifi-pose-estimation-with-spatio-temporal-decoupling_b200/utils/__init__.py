"""Utility functions (mirrors the reference's utils/__init__.py:5; the augmentation helpers are host-side data
preparation outside the hot path and are not re-implemented here -- SURVEY.md section 8f-4)."""
from .metrics import calculate_mpjpe, calculate_pck, pose_metrics_device

__all__ = ['calculate_pck', 'calculate_mpjpe', 'pose_metrics_device']
