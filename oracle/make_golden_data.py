"""Generate tests/golden/wiflow_data_golden.npz by running the UNMODIFIED reference input side here.

TEST INFRASTRUCTURE.  Run in the build container (needs /root/reference):

    python oracle/make_golden_data.py

Contents: outputs of the reference's utils/augmentation.py functions, of the train.py:187-193 augmentation sequence (restated
around the reference's functions, with torch.randn_like replaced by a stored noise tensor so that the host generator sees the
same draws as in a CUDA run), of dataset.py's two key-point repair routines, and of the reference dataset class + torch
DataLoader + file-level split on a tiny synthetic dataset directory whose files are stored in the fixture as well."""
import contextlib
import io
import os
import pickle
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import load_reference as L         # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'wiflow_data_golden.npz')
T_SMALL = 64                                    # masked axis of the small fixtures (the kernels are generic in T; 540 is tested against the oracle)


def synth_keypoints(rng, n, p_zero=0.25):
    kp = rng.uniform(0.05, 0.95, size=(n, 15, 2)).astype(np.float32)
    kp[rng.uniform(size=(n, 15)) < p_zero] = 0.0
    return kp


def write_dataset(d, g):
    """tiny dataset directory in the reference's on-disk format (dataset.py:23-60)"""
    np.save(os.path.join(d, 'csi_windows.npy'), g['ds_csi'])
    np.savez(os.path.join(d, 'window_info.npz'), window_to_file=g['ds_w2file'], window_to_frame=g['ds_w2frame'])
    files = np.array([f'file_{i}.csv' for i in range(len(g['ds_ranges']))], dtype=object)
    np.savez(os.path.join(d, 'file_info.npz'), keypoints_files=files, file_ids=np.arange(len(files)), window_ranges=g['ds_ranges'])
    np.savez(os.path.join(d, 'config.npz'), window_size=np.int64(g['ds_csi'].shape[1]), stride=np.int64(1))
    np.save(os.path.join(d, 'all_keypoints.npy'), g['ds_all_kp'])
    maps = {f'file_{i}.csv': {'start_idx': int(s)} for i, s in enumerate(g['ds_starts']) if s >= 0}
    with open(os.path.join(d, 'file_mappings.pkl'), 'wb') as f:
        pickle.dump(maps, f)


def main():
    R = L.load_data()
    g = {}
    # -- time_masking (augmentation.py:3-19) called like train.py:189
    torch.manual_seed(11)
    x = torch.randn(8, T_SMALL, 20)
    torch.manual_seed(12)
    g['tm_x'] = x.numpy()
    g['tm_out'] = R.time_masking(x.permute(0, 2, 1), mask_ratio=0.7).permute(0, 2, 1).contiguous().numpy()
    # ... and on a contiguous [B, C, T] tensor
    torch.manual_seed(13)
    g['tm_out_ct'] = R.time_masking(x.permute(0, 2, 1).contiguous(), mask_ratio=0.7).numpy()
    # -- add_noise / random_scaling
    torch.manual_seed(14)
    noise = torch.randn(8, T_SMALL, 20)
    real = torch.randn_like
    torch.randn_like = lambda t: noise
    try:
        g['noise'] = noise.numpy()
        g['an_out'] = R.add_noise(x, noise_level=0.05).numpy()
        outs, flags = [], []
        torch.manual_seed(15)
        for _ in range(6):
            y = R.random_scaling(x[:2])
            flags.append(y is not x[:2] and not torch.equal(y, x[:2]))
            outs.append(y.numpy().copy())
        g['rs_out'] = np.stack(outs)
        # -- the train.py:187-193 sequence, 6 consecutive batches from one seed
        seq = []
        torch.manual_seed(16)
        for _ in range(6):
            b = x
            if torch.rand(1).item() < 0.6:
                b = R.time_masking(b.permute(0, 2, 1), mask_ratio=0.3).permute(0, 2, 1)
            if torch.rand(1).item() < 0.6:
                b = R.add_noise(b, noise_level=0.02)
            if torch.rand(1).item() < 0.5:
                b = R.random_scaling(b, scale_range=(0.9, 1.1))
            seq.append(b.contiguous().numpy().copy())
        g['aug_seq'] = np.stack(seq)
    finally:
        torch.randn_like = real
    # -- key points
    rng = np.random.default_rng(21)
    frames = synth_keypoints(rng, 48)
    frames[5] = 0.0                                 # a frame with no valid joint stays zero
    frames[6] = np.abs(frames[6]) + 0.1             # a frame with nothing to repair
    frames[7, :, 0] = 0.0                           # x == 0 alone does not make a joint "zero"
    frames[7, :, 1] += 0.1
    dummy = type('D', (), {})()
    g['kp_frames'] = frames
    g['kp_single'] = np.stack([R.PreprocessedCSIKeypointsDataset._clean_single_frame_zeros(dummy, f) for f in frames])
    lens = [30, 17, 1, 9]
    seqs = [synth_keypoints(rng, n, 0.4) for n in lens]
    seqs[0][:4, 0] = 0.0                            # leading zeros: copy of the first valid frame
    seqs[0][-5:, 1] = 0.0                           # trailing zeros: copy of the last valid frame
    seqs[0][:, 2] = 0.0                             # a joint that is never seen stays zero
    seqs[1][3:12, 4] = 0.0                          # a long interior run: interpolation through repaired predecessors
    g['kp_seq_in'] = np.concatenate(seqs, 0)
    g['kp_seq_off'] = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    g['kp_seq_out'] = np.concatenate([R.PreprocessedCSIKeypointsDataset._clean_zero_keypoints(dummy, s) for s in seqs], 0)
    # -- dataset class + DataLoader + split on a tiny directory
    n_files, per = 9, [7, 5, 9, 4, 6, 8, 3, 5, 6]
    ranges = np.zeros((n_files, 2), dtype=np.int64)
    ranges[:, 1] = np.cumsum(per)
    ranges[1:, 0] = ranges[:-1, 1]
    N = int(ranges[-1, 1])
    g['ds_csi'] = rng.standard_normal((N, 12, 20)).astype(np.float32)
    g['ds_w2file'] = np.concatenate([np.full(p, i) for i, p in enumerate(per)]).astype(np.int64)
    g['ds_w2frame'] = np.concatenate([np.arange(p) for p in per]).astype(np.int64)
    g['ds_ranges'] = ranges
    starts = ranges[:, 0].copy()
    starts[4] = -1                                  # a file missing from file_mappings.pkl -> zero key points (dataset.py:103)
    g['ds_starts'] = starts
    g['ds_all_kp'] = synth_keypoints(rng, N - 2)    # table two frames short: the last windows fall off its end (dataset.py:92)
    with tempfile.TemporaryDirectory() as d, contextlib.redirect_stdout(io.StringIO()):
        write_dataset(d, g)
        ds = R.PreprocessedCSIKeypointsDataset(d)
        g['ds_items_y'] = np.stack([ds[i][1].numpy() for i in range(N)])
        tr, va, te = R.create_preprocessed_train_val_test_loaders(ds, batch_size=4, random_seed=42)
        g['split_train'] = np.array(tr.dataset.indices, dtype=np.int64)
        g['split_val'] = np.array(va.dataset.indices, dtype=np.int64)
        g['split_test'] = np.array(te.dataset.indices, dtype=np.int64)
        torch.manual_seed(31)
        for e in range(2):                          # two epochs from one seed: the generator state carries over
            xs, ys = zip(*[(bx.numpy(), by.numpy()) for bx, by in tr])
            g[f'ep{e}_train_x'] = np.concatenate(xs)
            g[f'ep{e}_train_y'] = np.concatenate(ys)
            xs, ys = zip(*[(bx.numpy(), by.numpy()) for bx, by in va])
            g[f'ep{e}_val_x'] = np.concatenate(xs)
            g[f'ep{e}_val_y'] = np.concatenate(ys)
    np.savez_compressed(OUT, **g)
    print(OUT, os.path.getsize(OUT), 'bytes;', {k: v.shape for k, v in g.items()})


if __name__ == '__main__':
    main()
