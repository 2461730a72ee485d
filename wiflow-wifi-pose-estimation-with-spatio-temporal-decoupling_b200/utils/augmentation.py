"""Drop-in for the reference `utils/augmentation.py:3-35` (time_masking, add_noise, random_scaling) plus `augment_batch`,
the whole train.py:187-193 sequence in at most two kernel launches (csrc/wf_data.cu).

The random DECISIONS are drawn on the host from the same torch generators, with the same calls in the same order as the
reference (so `torch.manual_seed(s)` reproduces the reference's augmented batch); the arithmetic -- per-row means written
over the masked spans, the batch-wide unbiased std, noise and scale -- runs on the GPU.  The reference loops over
B x masks x 20 rows in Python with a `.mean()` and a slice assignment each (up to 40 launches + syncs per masked window)."""
import numpy as np
import torch

from .. import ops


def draw_time_masks(B, T, mask_ratio=0.3, mask_len_range=(5, 10)):
    """Host draws of augmentation.py:9-14 in the reference's order: int32 [B, 2, 2] of (start, length), length 0 = unused."""
    spans = np.zeros((B, 2, 2), dtype=np.int32)
    for i in range(B):
        if torch.rand(1).item() < mask_ratio:
            num_masks = torch.randint(1, 3, (1,)).item()
            for k in range(num_masks):
                mask_len = torch.randint(mask_len_range[0], mask_len_range[1], (1,)).item()
                start = torch.randint(0, T - mask_len, (1,)).item()
                spans[i, k, 0], spans[i, k, 1] = start, mask_len
    return spans


def _spans_to_device(spans, device):
    return torch.from_numpy(np.ascontiguousarray(spans, dtype=np.int32)).to(device)


def _as_windows(x):
    """[B, C, T] tensor -> (dense [B, ., .] storage view, t_major, restore-view function)"""
    if x.dim() != 3:
        raise RuntimeError(f'expected a [B, C, T] tensor, got {list(x.shape)}')
    if x.transpose(1, 2).is_contiguous() and not x.is_contiguous():     # train.py:189: x = batch_x.permute(0, 2, 1)
        return x.transpose(1, 2), True, (lambda o: o.transpose(1, 2))
    return x.contiguous(), False, (lambda o: o)


def time_masking(x, mask_ratio=0.3, mask_len_range=(5, 10), spans=None):
    """x: [B, C, T] float32 CUDA tensor (a permuted view of the [B, T, C] batch, as train.py:189 passes it, is handled without
    a copy).  Returns a new tensor of the same shape/strides.  `spans` overrides the random draws (tests)."""
    B, C, T = x.shape
    if spans is None:
        spans = draw_time_masks(B, T, mask_ratio, mask_len_range)
    win, t_major, restore = _as_windows(x.float())
    out = ops.window_load(win, None, None, _spans_to_device(spans, x.device), None, t_major)
    return restore(out)


def add_noise(x, noise_level=0.05, noise=None):
    """x + randn_like(x) * noise_level * std(x) (unbiased std over the whole batch).  `noise` overrides randn_like (tests)."""
    xc = x.float().contiguous()
    if noise is None:
        noise = torch.randn_like(xc)
    stats = torch.zeros(2, device=x.device, dtype=torch.float64)
    _accumulate_stats(xc, stats)
    return ops.noise_scale(xc, noise.float().contiguous(), noise_level, 1.0, stats).view_as(x)


def _accumulate_stats(xc, stats):
    """sum / sum of squares of a contiguous tensor through window_load's statistics path (one CTA per row; nothing is written):
    one row per leading index when that is 16-byte aligned, else the largest aligned divisor below 48 000 floats"""
    n = xc.numel()
    if xc.dim() >= 2 and (n // xc.shape[0]) % 4 == 0:
        rows = xc.shape[0]
    else:
        rows = 0
        for w in range(min(n, 48000) // 4 * 4, 3, -4):
            if n % w == 0:
                rows = n // w
                break
        rows = rows or 1                   # no aligned tiling (odd sizes): one CTA walks the whole tensor, scalar tail
    v = xc.reshape(rows, 1, n // rows)
    ops.window_load(v, None, v, None, stats, False)


def random_scaling(x, scale_range=(0.9, 1.1)):
    if torch.rand(1).item() < 0.5:
        scale_factor = torch.FloatTensor(1).uniform_(scale_range[0], scale_range[1]).item()
        xc = x.float().contiguous()
        return ops.noise_scale(xc, None, 0.0, scale_factor, None).view_as(x)
    return x


def draw_augmentation(B, T=540, p_mask=0.6, p_noise=0.6, p_scale=0.5, mask_ratio=0.3, mask_len_range=(5, 10), scale_range=(0.9, 1.1)):
    """All host-side draws of train.py:188-193 for one batch, in the reference's order (the device-side randn_like of add_noise is
    not a host draw and happens in augment_batch).  Returns (spans or None, use_noise, scale or None)."""
    spans = draw_time_masks(B, T, mask_ratio, mask_len_range) if torch.rand(1).item() < p_mask else None
    use_noise = torch.rand(1).item() < p_noise
    scale = None
    if torch.rand(1).item() < p_scale:
        if torch.rand(1).item() < 0.5:
            scale = torch.FloatTensor(1).uniform_(scale_range[0], scale_range[1]).item()
    return spans, use_noise, scale


def augment_batch(x, noise_level=0.02, plan=None, noise=None, out=None, windows=None, idx=None):
    """train.py:187-193 on a [B, 540, 20] CUDA batch: time masking (p 0.6), noise (p 0.6, level 0.02), scaling (p 0.5 x 0.5), as
    two launches: window_load (masks + statistics, optionally fused with the gather `windows[idx]` of the batch itself) and
    noise_scale.  `plan` = draw_augmentation(...) (drawn here if None)."""
    if windows is not None:
        B, T = idx.numel(), windows.shape[1]
    else:
        B, T = x.shape[0], x.shape[1]
    spans, use_noise, scale = plan if plan is not None else draw_augmentation(B, T)
    dev = windows.device if windows is not None else x.device
    stats = torch.zeros(2, device=dev, dtype=torch.float64) if use_noise else None
    d_spans = _spans_to_device(spans, dev) if spans is not None and spans[:, :, 1].any() else None
    if windows is not None:
        y = ops.window_load(windows, idx, out, d_spans, stats, True)
    elif d_spans is not None or stats is not None:
        xc = x.contiguous()
        y = ops.window_load(xc, None, out if out is not None else (torch.empty_like(xc) if d_spans is not None else xc), d_spans, stats, True)
    else:
        y = x
    if use_noise or scale is not None:
        if use_noise and noise is None:
            noise = torch.randn_like(y)
        dst = out if out is not None else (y if y is not x else torch.empty_like(y))
        y = ops.noise_scale(y, noise if use_noise else None, noise_level, 1.0 if scale is None else scale, stats, dst)
    return y
