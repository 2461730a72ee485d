"""Utility functions (mirrors the reference's utils/__init__.py:5-6): metrics and the augmentation helpers, both on the GPU."""
from .augmentation import add_noise, augment_batch, draw_augmentation, random_scaling, time_masking
from .metrics import calculate_mpjpe, calculate_pck, pose_metrics_device

__all__ = ['calculate_pck', 'calculate_mpjpe', 'pose_metrics_device', 'time_masking', 'add_noise', 'random_scaling',
           'augment_batch', 'draw_augmentation']
