"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's input side: utils/augmentation.py and the
key-point handling of dataset.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file;
the product (wiflow_b200.data, wiflow_b200.utils.augmentation) never does.

Pinned two ways: tests/test_oracle_data.py compares every function with the reference's own code imported from /root/reference
when that tree is mounted (the build container), and with tests/golden/wiflow_data_golden.npz, which oracle/make_golden_data.py
produced by running the reference's functions, dataset class and torch DataLoader.  The reference ships no tests for this path."""
import numpy as np
import torch


# ---- utils/augmentation.py -------------------------------------------------------------------------------------------------
def draw_time_masks(B, T, mask_ratio=0.3, mask_len_range=(5, 10)):
    """the random draws of augmentation.py:9-14 (default CPU generator, reference call order) -> [(start, len), ...] per window"""
    plan = []
    for _ in range(B):
        spans = []
        if torch.rand(1).item() < mask_ratio:                               # augmentation.py:10
            for _ in range(torch.randint(1, 3, (1,)).item()):               # :11
                mask_len = torch.randint(mask_len_range[0], mask_len_range[1], (1,)).item()   # :13
                start = torch.randint(0, T - mask_len, (1,)).item()         # :14
                spans.append((start, mask_len))
        plan.append(spans)
    return plan


def apply_time_masks(x, plan):
    """augmentation.py:15-17 on x [B, C, T]: every row's current mean over T overwrites the span; spans apply in order"""
    out = x.clone()
    for i, spans in enumerate(plan):
        for start, mask_len in spans:
            for c in range(out.shape[1]):
                out[i, c, start:start + mask_len] = out[i, c, :].mean()
    return out


def time_masking(x, mask_ratio=0.3, mask_len_range=(5, 10)):
    return apply_time_masks(x, draw_time_masks(x.shape[0], x.shape[2], mask_ratio, mask_len_range))


def add_noise(x, noise_level=0.05, noise=None):
    """augmentation.py:22-26; `noise` stands for torch.randn_like(x)"""
    if noise is None:
        noise = torch.randn_like(x)
    return x + noise * noise_level * torch.std(x)


def draw_scale(scale_range=(0.9, 1.1)):
    """augmentation.py:32-33 -> python float or None"""
    if torch.rand(1).item() < 0.5:
        return torch.FloatTensor(1).uniform_(scale_range[0], scale_range[1])
    return None


def random_scaling(x, scale_range=(0.9, 1.1)):
    s = draw_scale(scale_range)
    return x if s is None else x * s


def augment_step(x, noise=None):
    """train.py:187-193 on a [B, 540, 20] batch.  Returns (augmented batch, dict of the decisions taken)."""
    info = {'plan': None, 'noise': False, 'scale': None}
    if torch.rand(1).item() < 0.6:                                          # train.py:188
        info['plan'] = draw_time_masks(x.shape[0], x.shape[1], 0.3)
        x = apply_time_masks(x.permute(0, 2, 1), info['plan']).permute(0, 2, 1)
    if torch.rand(1).item() < 0.6:                                          # :190
        info['noise'] = True
        x = add_noise(x, 0.02, noise)
    if torch.rand(1).item() < 0.5:                                          # :192
        s = draw_scale((0.9, 1.1))
        if s is not None:
            info['scale'] = float(s.item())
            x = x * s
    return x, info


# ---- dataset.py key points -------------------------------------------------------------------------------------------------
def clean_single_frame_zeros(frame):
    """dataset.py:105-120 on one [K, 2] float32 frame: joints that are (0, 0) take the mean of the others (fp32, joint order)"""
    frame = np.asarray(frame, dtype=np.float32)
    out = frame.copy()
    sx = np.float32(0.0)
    sy = np.float32(0.0)
    cnt = 0
    for j in range(frame.shape[0]):
        if frame[j, 0] != 0 or frame[j, 1] != 0:
            sx = np.float32(sx + frame[j, 0])
            sy = np.float32(sy + frame[j, 1])
            cnt += 1
    if cnt:
        mx, my = np.float32(sx / np.float32(cnt)), np.float32(sy / np.float32(cnt))
        for j in range(frame.shape[0]):
            if frame[j, 0] == 0 and frame[j, 1] == 0:
                out[j, 0], out[j, 1] = mx, my
    return out


def keypoint_batch(frames, frame_index, clean=True):
    """dataset.py:80-103 for a batch: frames[frame_index[b]] or zeros when the index is outside the table"""
    K = frames.shape[1]
    out = np.zeros((len(frame_index), K, 2), dtype=np.float32)
    for b, f in enumerate(frame_index):
        if 0 <= f < len(frames):
            out[b] = clean_single_frame_zeros(frames[f]) if clean else frames[f]
    return out


def clean_zero_keypoints(seq):
    """dataset.py:159-206 on one [frames, K, 2] float32 sequence (in-place semantics of the reference loop preserved: a repaired
    frame is a valid predecessor for the next zero frame; alpha is a python float that numpy rounds to float32 at the products)"""
    coords = np.array(seq, dtype=np.float32, copy=True)
    n, K, _ = coords.shape
    for k in range(K):
        zeros = [t for t in range(n) if coords[t, k, 0] == 0 and coords[t, k, 1] == 0]
        for t in zeros:
            prev = nxt = None
            for p in range(t - 1, -1, -1):
                if not (coords[p, k, 0] == 0 and coords[p, k, 1] == 0):
                    prev = p
                    break
            for q in range(t + 1, n):
                if not (coords[q, k, 0] == 0 and coords[q, k, 1] == 0):
                    nxt = q
                    break
            if prev is not None and nxt is not None:
                a = (t - prev) / (nxt - prev)
                wa, wb = np.float32(a), np.float32(1 - a)
                coords[t, k] = (wb * coords[prev, k]).astype(np.float32) + (wa * coords[nxt, k]).astype(np.float32)
            elif prev is not None:
                coords[t, k] = coords[prev, k]
            elif nxt is not None:
                coords[t, k] = coords[nxt, k]
    return coords


# ---- DataLoader order / split ------------------------------------------------------------------------------------------------
def loader_order(n, shuffle):
    """order in which torch's DataLoader visits a dataset of n items (RandomSampler when shuffle), consuming the default generator
    like iter(loader) + next() do"""
    torch.empty((), dtype=torch.int64).random_()
    if not shuffle:
        return list(range(n))
    g = torch.Generator()
    g.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))
    return torch.randperm(n, generator=g).tolist()
