"""Shared helpers for the parity tests (test infrastructure; may import the oracle)."""
import os
import re

import numpy as np
import torch

from oracle import wiflow_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# parameters whose gradient is analytically zero in train mode because a BatchNorm / softmax cancels them
# (SURVEY.md section 7-H3): the reference's own fp32 values are rounding noise.
DEAD_PARAM_RE = re.compile(r'(^up\.block\.[048]\.bias$|^residual_blocks\.\d\.block\.[048]\.bias$|^decoder\.[03]\.bias$|'
                           r'bn_similarity\.bias$|^attention\.width_axis\.bn_output\.bias$)')


def is_dead(name):
    return DEAD_PARAM_RE.search(name) is not None


def load_golden():
    g = np.load(os.path.join(GOLDEN, 'wiflow_golden_b4.npz'))
    return {k: g[k] for k in g.files}


def golden_masks(g, B=4, device='cpu'):
    """Rebuild the 18 multiplicative dropout masks stored (bit-packed) in the fixture."""
    shapes = O.dropout_mask_shapes(B)
    bits = np.unpackbits(g['mask_bits'])
    masks, off = [], 0
    for shp, scale in zip(shapes, g['mask_scales']):
        n = int(np.prod(shp))
        m = torch.from_numpy(bits[off:off + n].astype(np.float32)).reshape(shp) * float(scale)
        masks.append(m.to(device))
        off += n
    return masks


def sample(t, stride):
    t = t.detach().reshape(-1)
    return t if t.numel() <= 1024 else t[::stride]


def rel_err(a, b):
    """max-norm relative error of a against the truth b"""
    a, b = a.double(), b.double()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


# ---- mapping between the library's named workspace tensors and the oracle's recorded intermediates ----
_CONV_ALIASES = (('conv1_group', 'g1'), ('conv1_pw', 'pw1'), ('conv2_group', 'g2'), ('conv2_pw', 'pw2'), ('downsample.0', 'ds'),
                 ('block.0', 'c1'), ('block.4', 'c2'), ('block.8', 'c3'), ('qkv_transform', 'qkv'))


def oracle_key(dbg_name):
    """debug tensor name -> (oracle record name, wants_grad)"""
    kind = dbg_name.rsplit('.', 1)[1]
    stem = dbg_name.rsplit('.', 1)[0]
    if stem == 'decoder.0':
        stem = 'decoder.d1'
    elif stem == 'decoder.3':
        stem = 'decoder.d2'
    for a, b in _CONV_ALIASES:
        if stem.endswith(a):
            stem = stem[:-len(a)] + b
            break
    if kind == 'raw':
        return stem + '.raw', False
    if kind == 'dy':
        return stem + '.y', True
    if kind == 'out':
        return stem + '.out', False
    if kind == 'dout':
        return stem + '.out', True
    raise KeyError(dbg_name)


def to_internal(name, t, B):
    """oracle tensor (reference layout) -> the library's [C][P][B][20] layout"""
    if name.startswith('tcn.'):                       # [B, C, 20]
        return t.permute(1, 0, 2).unsqueeze(1)
    if name.startswith('up.') or name.startswith('residual_blocks.'):   # [B, C, 20, W]
        return t.permute(1, 3, 0, 2)
    if name.startswith('attention.width_axis.'):      # [B*15, C, 20]
        return t.reshape(B, 15, t.shape[1], 20).permute(2, 1, 0, 3)
    if name.startswith('attention.height_axis.'):     # [B*20, C, 15]
        return t.reshape(B, 20, t.shape[1], 15).permute(2, 3, 0, 1)
    if name.startswith('decoder.'):                   # [B, C, 15, 20]
        return t.permute(1, 2, 0, 3)
    raise KeyError(name)
