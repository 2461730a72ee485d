"""wiflow_b200: a from-scratch, B200-native (sm_100a) implementation of the WiFlow hot path -- forward and backward of the
540x20-CSI pose model, its pose loss, PCK/MPJPE and the clip+AdamW step -- behind the reference's own Python interfaces.

    from wiflow_b200.models import WiFlowPoseModel, TemporalBlock, ConvBlock1, AsymmetricConvBlock, AxialAttention, DualAxialAttention
    from wiflow_b200.losses import PoseLoss
    from wiflow_b200.utils import calculate_pck, calculate_mpjpe
    from wiflow_b200.engine import TrainStep, InferStep
    from wiflow_b200.train_loop import Trainer
    from wiflow_b200.data import PreprocessedCSIKeypointsDataset, create_preprocessed_train_val_test_loaders, DeviceBatchLoader
    from wiflow_b200.utils import time_masking, add_noise, random_scaling, augment_batch

All arithmetic runs in libwiflow_b200.so (hand-written CUDA, C ABI in include/wiflow_b200.h).  There is no CPU, Triton or
eager-PyTorch fallback: importing works anywhere, calling needs the built library and a B200."""
from . import _lib, ops                      # noqa: F401
from . import losses, models, utils           # noqa: F401
from .engine import InferStep, TrainStep, allreduce_gradients, shard_bounds      # noqa: F401
from .train_loop import Trainer              # noqa: F401
from .data import DeviceBatchLoader, PreprocessedCSIKeypointsDataset, create_preprocessed_train_val_test_loaders      # noqa: F401
from .losses import PoseLoss                  # noqa: F401
from .models import (AsymmetricConvBlock, AxialAttention, ConvBlock1, DualAxialAttention, InnerGroupedTemporalBlock,  # noqa: F401
                     TemporalBlock, TemporalConvNet, WiFlow, WiFlowPoseModel)
from .utils import calculate_mpjpe, calculate_pck     # noqa: F401

__version__ = '0.1.0'
