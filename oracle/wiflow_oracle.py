"""CPU oracle for the WiFlow hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional restatement (torch, CPU, fp32 or fp64) of the reference's pose-model forward,
pose loss, PCK/MPJPE metrics and one clip+AdamW training step.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import it; the product package never does (it fails loudly without its CUDA library).

Parity pinning: the reference ships no tests and no golden vectors, so
this restatement is pinned against the reference's own modules executed in the build
container (`oracle/make_golden.py`, `tests/test_oracle_vs_reference.py`) and against the
fixtures those runs produced (`tests/golden/*.npz`).

Every function cites the reference file:line it follows (paths relative to the
reference root).  The arithmetic itself lives in the third-party dependency `torch`
(pinned torch==2.3.1 in the reference's requirements.txt:10; 2.11.0 in this image).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------
# Architecture constants (models/pose_model.py:16-53, models/tcn.py:18, train.py:88,105-110)
# ----------------------------------------------------------------------------------------
T_STEPS = 20
TCN_CHANNELS = (540, 540, 440, 340, 240)          # input + num_channels, pose_model.py:17-18
TCN_GROUPS = 20                                   # tcn.py:18
RES_CHANNELS = (8, 8, 16, 32, 64)                 # up out + residual blocks, pose_model.py:25-36
ATT_PLANES, ATT_GROUPS = 64, 8                    # pose_model.py:39-41
CONV_DROPOUT = 0.3                                # convnet.py:7,44 (never forwarded by pose_model.py:25,34)
BN_EPS, BN_MOMENTUM = 1e-5, 0.1
BONES = ((0, 1), (1, 8), (1, 2), (2, 3), (3, 4), (1, 5), (5, 6), (6, 7), (8, 9), (8, 12),
         (9, 10), (10, 11), (12, 13), (13, 14))   # losses/pose_loss.py:20-24


# ----------------------------------------------------------------------------------------
# State factory: same construction order (hence same RNG stream) as the reference ctor.
# ----------------------------------------------------------------------------------------
def _tcn_block_layers(cin, cout, d):
    """Registration order of InnerGroupedTemporalBlock (models/tcn.py:20-49)."""
    pad = 2 * d
    layers = OrderedDict()
    layers['conv1_group'] = nn.Conv1d(cin, cin, 3, stride=1, padding=pad, dilation=d, groups=TCN_GROUPS, bias=False)
    layers['bn1_group'] = nn.BatchNorm1d(cin)
    layers['conv1_pw'] = nn.Conv1d(cin, cout, 1, bias=False)
    layers['bn1_pw'] = nn.BatchNorm1d(cout)
    layers['conv2_group'] = nn.Conv1d(cout, cout, 3, stride=1, padding=pad, dilation=d, groups=TCN_GROUPS, bias=False)
    layers['bn2_group'] = nn.BatchNorm1d(cout)
    layers['conv2_pw'] = nn.Conv1d(cout, cout, 1, bias=False)
    layers['bn2_pw'] = nn.BatchNorm1d(cout)
    if cin != cout:
        layers['downsample.0'] = nn.Conv1d(cin, cout, 1, bias=False)
        layers['downsample.1'] = nn.BatchNorm1d(cout)
    return layers


def _conv_block_layers(cin, cout, stride):
    """Registration order of ConvBlock1 / AsymmetricConvBlock (models/convnet.py:10-29,47-65)."""
    layers = OrderedDict()
    layers['block.0'] = nn.Conv2d(cin, cout, (1, 3), stride=(1, stride), padding=(0, 1))
    layers['block.1'] = nn.BatchNorm2d(cout)
    layers['block.4'] = nn.Conv2d(cout, cout, (1, 3), padding=(0, 1))
    layers['block.5'] = nn.BatchNorm2d(cout)
    layers['block.8'] = nn.Conv2d(cout, cout, (1, 3), padding=(0, 1))
    layers['block.9'] = nn.BatchNorm2d(cout)
    layers['downsample.0'] = nn.Conv2d(cin, cout, 1, stride=(1, stride), bias=False)
    layers['downsample.1'] = nn.BatchNorm2d(cout)
    return layers


def _axial_layers():
    """Registration order of AxialAttention (models/attention.py:22-35) incl. its own init."""
    layers = OrderedDict()
    layers['qkv_transform'] = nn.Conv1d(ATT_PLANES, 3 * ATT_PLANES, 1, bias=False)
    layers['bn_qkv'] = nn.BatchNorm1d(3 * ATT_PLANES)
    layers['bn_similarity'] = nn.BatchNorm2d(ATT_GROUPS)
    layers['bn_output'] = nn.BatchNorm1d(ATT_PLANES)
    nn.init.normal_(layers['qkv_transform'].weight.data, 0, math.sqrt(1.0 / ATT_PLANES))  # attention.py:34-35
    return layers


def make_state(seed=None, dtype=torch.float32):
    """Random-init state_dict (295 entries) drawn exactly as
    WiFlowPoseModel.__init__ does (models/pose_model.py:12-69)."""
    if seed is not None:
        torch.manual_seed(seed)
    mods = OrderedDict()
    for i in range(4):
        for k, m in _tcn_block_layers(TCN_CHANNELS[i], TCN_CHANNELS[i + 1], 2 ** i).items():
            mods[f'tcn.network.{i}.{k}'] = m
    for k, m in _conv_block_layers(1, RES_CHANNELS[0], 1).items():
        mods[f'up.{k}'] = m
    for i in range(4):
        for k, m in _conv_block_layers(RES_CHANNELS[i], RES_CHANNELS[i + 1], 2).items():
            mods[f'residual_blocks.{i}.{k}'] = m
    for axis in ('width_axis', 'height_axis'):
        for k, m in _axial_layers().items():
            mods[f'attention.{axis}.{k}'] = m
    mods['decoder.0'] = nn.Conv2d(ATT_PLANES, 32, 3, padding=1)
    mods['decoder.1'] = nn.BatchNorm2d(32)
    mods['decoder.3'] = nn.Conv2d(32, 2, 1)
    mods['decoder.4'] = nn.BatchNorm2d(2)
    # _initialize_weights (pose_model.py:57-69): module-tree order == registration order here.
    for m in mods.values():
        if isinstance(m, nn.Conv1d):
            nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
        elif isinstance(m, nn.BatchNorm1d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)
    state = OrderedDict()
    for name, m in mods.items():
        for k, v in m.state_dict().items():
            state[f'{name}.{k}'] = v.detach().clone().to(dtype) if v.is_floating_point() else v.detach().clone()
    return state


def param_names(state):
    return [k for k in state if not (k.endswith('running_mean') or k.endswith('running_var')
                                     or k.endswith('num_batches_tracked'))]


# ----------------------------------------------------------------------------------------
# Forward
# ----------------------------------------------------------------------------------------
class _Ctx:
    def __init__(self, state, train, update_buffers, masks, record):
        self.s, self.train, self.upd, self.masks, self.rec = state, train, update_buffers, masks, record
        self.mask_i = 0

    def bn(self, x, name):
        s = self.s
        rm, rv = s[name + '.running_mean'], s[name + '.running_var']
        if self.train and not self.upd:
            rm, rv = rm.clone(), rv.clone()
        y = F.batch_norm(x, rm, rv, s[name + '.weight'], s[name + '.bias'], self.train, BN_MOMENTUM, BN_EPS)
        if self.train and self.upd:
            s[name + '.num_batches_tracked'] += 1
        return y

    def drop(self, x):
        """Dropout / Dropout2d: multiply by the next mask (values 0 or 1/(1-p)), reference call order."""
        if not self.train or self.masks is None:
            return x
        m = self.masks[self.mask_i]
        self.mask_i += 1
        return x * m

    def note(self, name, t):
        if self.rec is not None:
            if t.requires_grad:
                t.retain_grad()
            self.rec[name] = t
        return t


def _tcn_block(c, x, i):
    """InnerGroupedTemporalBlock.forward (models/tcn.py:51-74); Chomp1d (tcn.py:11-12)."""
    p = f'tcn.network.{i}'
    d = 2 ** i
    w = c.s
    cin, cout = TCN_CHANNELS[i], TCN_CHANNELS[i + 1]
    if cin != cout:
        r = c.note(f'{p}.ds.raw', F.conv1d(x, w[p + '.downsample.0.weight']))
        res = c.note(f'{p}.ds.y', c.bn(r, p + '.downsample.1'))
    else:
        res = x

    def gconv(h, name):
        o = F.conv1d(h, w[f'{p}.{name}.weight'], padding=2 * d, dilation=d, groups=TCN_GROUPS)
        return o[:, :, :-2 * d].contiguous()

    o = c.note(f'{p}.g1.raw', gconv(x, 'conv1_group'))
    o = F.silu(c.note(f'{p}.g1.y', c.bn(o, p + '.bn1_group')))
    o = c.note(f'{p}.pw1.raw', F.conv1d(o, w[p + '.conv1_pw.weight']))
    o = c.drop(F.silu(c.note(f'{p}.pw1.y', c.bn(o, p + '.bn1_pw'))))
    o = c.note(f'{p}.g2.raw', gconv(o, 'conv2_group'))
    o = F.silu(c.note(f'{p}.g2.y', c.bn(o, p + '.bn2_group')))
    o = c.note(f'{p}.pw2.raw', F.conv1d(o, w[p + '.conv2_pw.weight']))
    o = c.drop(F.silu(c.note(f'{p}.pw2.y', c.bn(o, p + '.bn2_pw'))))
    return c.note(f'{p}.out', F.silu(o + res))


def _conv_block(c, x, p, stride):
    """ConvBlock1.forward / AsymmetricConvBlock.forward (models/convnet.py:33-38,69-74)."""
    w = c.s
    r = c.note(f'{p}.ds.raw', F.conv2d(x, w[p + '.downsample.0.weight'], stride=(1, stride)))
    identity = c.note(f'{p}.ds.y', c.bn(r, p + '.downsample.1'))
    o = c.note(f'{p}.c1.raw', F.conv2d(x, w[p + '.block.0.weight'], w[p + '.block.0.bias'], stride=(1, stride), padding=(0, 1)))
    o = c.drop(F.silu(c.note(f'{p}.c1.y', c.bn(o, p + '.block.1'))))
    o = c.note(f'{p}.c2.raw', F.conv2d(o, w[p + '.block.4.weight'], w[p + '.block.4.bias'], padding=(0, 1)))
    o = c.drop(F.silu(c.note(f'{p}.c2.y', c.bn(o, p + '.block.5'))))
    o = c.note(f'{p}.c3.raw', F.conv2d(o, w[p + '.block.8.weight'], w[p + '.block.8.bias'], padding=(0, 1)))
    o = c.note(f'{p}.c3.y', c.bn(o, p + '.block.9'))
    return c.note(f'{p}.out', F.silu(o + identity))


def _axial(c, x, p, width):
    """AxialAttention.forward (models/attention.py:37-80), stride 1."""
    w = c.s
    B, C, H, W = x.shape
    x = x.permute(0, 2, 1, 3) if width else x.permute(0, 3, 1, 2)
    N, R, C, L = x.shape
    x = x.contiguous().view(N * R, C, L)
    raw = c.note(f'{p}.qkv.raw', F.conv1d(x, w[p + '.qkv_transform.weight']))
    qkv = c.note(f'{p}.qkv.y', c.bn(raw, p + '.bn_qkv'))
    qkv = qkv.reshape(N * R, 3, ATT_PLANES, L).permute(1, 0, 2, 3)
    gp = ATT_PLANES // ATT_GROUPS
    q = qkv[0].reshape(N * R, ATT_GROUPS, gp, L)
    k = qkv[1].reshape(N * R, ATT_GROUPS, gp, L)
    v = qkv[2].reshape(N * R, ATT_GROUPS, gp, L)
    qk = c.note(f'{p}.sim.raw', torch.einsum('bgci,bgcj->bgij', q, k))     # no 1/sqrt(d), attention.py:61
    qk = c.note(f'{p}.sim.y', c.bn(qk, p + '.bn_similarity'))
    sim = F.softmax(qk, dim=-1)
    sv = torch.einsum('bgij,bgcj->bgci', sim, v).reshape(N * R, ATT_PLANES, L)
    sv = c.note(f'{p}.sv.raw', sv)
    out = c.note(f'{p}.sv.y', c.bn(sv, p + '.bn_output')).view(N, R, ATT_PLANES, L)
    return out.permute(0, 2, 1, 3) if width else out.permute(0, 2, 3, 1)


def forward(state, x, train=False, update_buffers=False, masks=None, record=None):
    """WiFlowPoseModel.forward (models/pose_model.py:71-97).  x [B,540,20] -> [B,15,2].

    masks: None (all dropout p=0) or the 18 multiplicative masks from `make_dropout_masks`.
    record: optional dict that receives named intermediates (with retain_grad)."""
    c = _Ctx(state, train, update_buffers, masks, record)
    h = x
    for i in range(4):
        h = _tcn_block(c, h, i)                                      # pose_model.py:76
    h = h.transpose(1, 2).unsqueeze(1)                              # :79  [B,1,20,240]
    h = _conv_block(c, h, 'up', 1)                                  # :82
    for i in range(4):
        h = _conv_block(c, h, f'residual_blocks.{i}', 2)            # :83-84
    h = h.permute(0, 1, 3, 2)                                       # :87  [B,64,15,20]
    h = _axial(c, h, 'attention.width_axis', True)                  # attention.py:95-98
    h = _axial(c, h, 'attention.height_axis', False)
    w = state
    h = c.note('decoder.d1.raw', F.conv2d(h, w['decoder.0.weight'], w['decoder.0.bias'], padding=1))
    h = F.silu(c.note('decoder.d1.y', c.bn(h, 'decoder.1')))
    h = c.note('decoder.d2.raw', F.conv2d(h, w['decoder.3.weight'], w['decoder.3.bias']))
    h = F.silu(c.note('decoder.d2.y', c.bn(h, 'decoder.4')))
    h = h.mean(dim=3)                                               # AdaptiveAvgPool2d((15,1)), :53,94
    return h.transpose(1, 2)                                        # :95  [B,15,2]


def dropout_mask_shapes(B):
    """Shapes of the 18 dropout sites in reference call order:
    8x nn.Dropout in the TCN (tcn.py:30,43), 10x nn.Dropout2d in up + residual blocks (convnet.py:15,20,51,56)."""
    shapes = []
    for i in range(4):
        shapes += [(B, TCN_CHANNELS[i + 1], T_STEPS)] * 2
    for cout in RES_CHANNELS:
        shapes += [(B, cout, 1, 1)] * 2
    return shapes


def make_dropout_masks(B, p_tcn, p_conv=CONV_DROPOUT, device='cpu', dtype=torch.float32):
    """Draw the 18 masks with torch's own RNG in the order the reference's forward consumes them.
    Dropout2d draws one Bernoulli per (b, c) plane (aten feature_dropout)."""
    masks = []
    for i, shp in enumerate(dropout_mask_shapes(B)):
        p = p_tcn if i < 8 else p_conv
        ones = torch.ones(shp, device=device, dtype=dtype)
        masks.append(F.dropout(ones, p, True) if i < 8 else F.dropout2d(ones, p, True))
    return masks


# ----------------------------------------------------------------------------------------
# Loss and metrics
# ----------------------------------------------------------------------------------------
def bone_lengths(kp):
    """PoseLoss.compute_bone_lengths (losses/pose_loss.py:26-33)."""
    s = torch.tensor([b[0] for b in BONES])
    e = torch.tensor([b[1] for b in BONES])
    v = kp[..., e, :] - kp[..., s, :]
    return torch.sqrt((v ** 2).sum(-1) + 1e-8)


def pose_loss(pred, target, position_weight=1.0, bone_weight=0.2, loss_type='smooth_l1'):
    """PoseLoss.forward (losses/pose_loss.py:35-88) -> (total, position, bone) tensors."""
    B = pred.shape[0]
    if pred.dim() == 2 and pred.shape[1] == 30:
        pred = pred.reshape(B, 15, 2)
    if target.dim() == 2 and target.shape[1] == 30:
        target = target.reshape(B, 15, 2)
    lp, lt = bone_lengths(pred), bone_lengths(target)
    if loss_type == 'mse':
        pos, bone = F.mse_loss(pred, target), F.mse_loss(lp, lt)
    elif loss_type == 'l1':
        pos, bone = F.l1_loss(pred, target), F.l1_loss(lp, lt)
    elif loss_type == 'smooth_l1':
        pos, bone = F.smooth_l1_loss(pred, target, beta=0.1), F.smooth_l1_loss(lp, lt, beta=0.05)
    else:
        raise ValueError(f"Unknown loss type: {loss_type}")
    return position_weight * pos + bone_weight * bone, pos, bone


def pck(pred, target, thresholds=(0.2,), use_torso_norm=True):
    """calculate_pck (utils/metrics.py:3-33)."""
    B = pred.shape[0]
    if pred.dim() == 2 and pred.shape[1] == 30:
        pred, target = pred.reshape(B, 15, 2), target.reshape(B, 15, 2)
    a, b = (2, 12) if use_torso_norm else (2, 5)
    norm = torch.sqrt(((target[:, a] - target[:, b]) ** 2).sum(1)).clamp(min=0.01)
    d = torch.sqrt(((pred - target) ** 2).sum(2)) / norm.unsqueeze(1)
    return {t: (d <= t).float().mean().item() for t in thresholds}


def mpjpe(pred, target):
    """calculate_mpjpe (utils/metrics.py:36-47)."""
    B = pred.shape[0]
    if pred.dim() == 2 and pred.shape[1] == 30:
        pred, target = pred.reshape(B, 15, 2), target.reshape(B, 15, 2)
    return torch.sqrt(((pred - target) ** 2).sum(2)).mean().item()


# ----------------------------------------------------------------------------------------
# One training step: fwd + loss + bwd + clip_grad_norm_(1.0) + AdamW (train.py:105-110,196-237)
# ----------------------------------------------------------------------------------------
def grads(state, x, y, masks=None, update_buffers=False, record=None):
    """Loss and d loss / d param for every parameter (train mode, batch-stat BN)."""
    names = param_names(state)
    for n in names:
        state[n].requires_grad_(True)
        state[n].grad = None
    pred = forward(state, x, train=True, update_buffers=update_buffers, masks=masks, record=record)
    if record is not None:
        pred.retain_grad()
        record['pred'] = pred
    total, pos, bone = pose_loss(pred, y)
    total.backward()
    g = OrderedDict((n, state[n].grad.detach().clone()) for n in names)
    for n in names:
        state[n].requires_grad_(False)
        state[n].grad = None
    return pred.detach(), (total.item(), pos.item(), bone.item()), g


def clip_adamw_step(params, grads_, m, v, step, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, wd=5e-5, max_norm=1.0):
    """clip_grad_norm_(max_norm) then AdamW, restated (train.py:105-110,235-236).
    params/grads_/m/v are dicts name->tensor, updated in place; returns the pre-clip total norm."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads_.values())).item()
    coef = min(1.0, max_norm / (total + 1e-6))
    b1, b2 = betas
    for n, p in params.items():
        g = grads_[n] * coef
        p.mul_(1 - lr * wd)
        m[n].mul_(b1).add_(g, alpha=1 - b1)
        v[n].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v[n] / (1 - b2 ** step)).sqrt_().add_(eps)
        p.addcdiv_(m[n] / (1 - b1 ** step), denom, value=-lr)
    return total


def synthetic_batch(B, seed=0, dtype=torch.float32):
    """x ~ N(0,1) [B,540,20], y ~ U(0,1) [B,15,2]."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 540, T_STEPS, generator=g, dtype=torch.float32).to(dtype)
    y = torch.rand(B, 15, 2, generator=g, dtype=torch.float32).to(dtype)
    return x, y
