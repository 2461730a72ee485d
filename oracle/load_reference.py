"""Import the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE.  The reference's `models/__init__.py:7` imports a class that no longer
exists (`TemporalConvNet`), so `import models` raises; we register an empty `models`
package whose __path__ points at the reference directory and import the sub-modules
directly.  Nothing here is reachable on the GPU box, where /root/reference
does not exist -- callers must check `available()` first.
"""
import os
import sys
import types

REF_ROOT = os.environ.get('WIFLOW_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'models', 'pose_model.py'))


def load():
    """Returns a namespace with WiFlowPoseModel, TemporalBlock, ConvBlock1, AsymmetricConvBlock,
    AxialAttention, DualAxialAttention, PoseLoss, calculate_pck, calculate_mpjpe."""
    if not available():
        raise RuntimeError(f'reference not found under {REF_ROOT}')
    saved = {k: sys.modules.get(k) for k in ('models', 'losses', 'utils')}
    saved_path = list(sys.path)
    try:
        for k in ('models', 'losses', 'utils'):
            for mk in [m for m in sys.modules if m == k or m.startswith(k + '.')]:
                del sys.modules[mk]
        pkg = types.ModuleType('models')
        pkg.__path__ = [os.path.join(REF_ROOT, 'models')]
        sys.modules['models'] = pkg
        upkg = types.ModuleType('utils')           # utils/__init__ is importable, but keep it hermetic
        upkg.__path__ = [os.path.join(REF_ROOT, 'utils')]
        sys.modules['utils'] = upkg
        sys.path.insert(0, REF_ROOT)
        from models.pose_model import WiFlowPoseModel
        from models.tcn import TemporalBlock, InnerGroupedTemporalBlock, Chomp1d
        from models.convnet import ConvBlock1, AsymmetricConvBlock
        from models.attention import AxialAttention, DualAxialAttention
        from losses.pose_loss import PoseLoss
        from utils.metrics import calculate_pck, calculate_mpjpe
        ns = types.SimpleNamespace(**{k: v for k, v in locals().items() if k[0].isupper() or k.startswith('calculate')})
        return ns
    finally:
        sys.path[:] = saved_path
        for k in ('models', 'losses', 'utils'):
            for mk in [m for m in sys.modules if m == k or m.startswith(k + '.')]:
                del sys.modules[mk]
            if saved[k] is not None:
                sys.modules[k] = saved[k]


def load_data():
    """The reference's input side, loaded by file path (no package imports, nothing left in sys.modules): a namespace with
    time_masking, add_noise, random_scaling (utils/augmentation.py:3-35), PreprocessedCSIKeypointsDataset and
    create_preprocessed_train_val_test_loaders (dataset.py:16,256)."""
    if not available():
        raise RuntimeError(f'reference not found under {REF_ROOT}')
    import importlib.util
    out = {}
    for name, rel in (('ref_augmentation', 'utils/augmentation.py'), ('ref_dataset', 'dataset.py')):
        spec = importlib.util.spec_from_file_location('_wiflow_' + name, os.path.join(REF_ROOT, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        out[name] = mod
    a, d = out['ref_augmentation'], out['ref_dataset']
    return types.SimpleNamespace(time_masking=a.time_masking, add_noise=a.add_noise, random_scaling=a.random_scaling,
                                 PreprocessedCSIKeypointsDataset=d.PreprocessedCSIKeypointsDataset,
                                 create_preprocessed_train_val_test_loaders=d.create_preprocessed_train_val_test_loaders)
