"""Data path of the reference (`dataset.py:16-325`) rebuilt for a 180 GB HBM device.

The reference keeps `csi_windows.npy` in host memory and builds every batch on the host: `__getitem__` per window
(`dataset.py:206-241`: one numpy slice, one keypoint lookup + zero-joint repair in numpy, two `torch.from_numpy`), the default
collate, `pin_memory`, then `.to(device)` (`train.py:183-184`).  At the measured 50 k windows/s of the training step that is
2.2 GB/s of CSI through a single Python thread.  Here:

* `PreprocessedCSIKeypointsDataset` reads the same files, keeps the key-point table `[F,15,2]` on the GPU and, when the
  window array fits (`resident=True`; a 360 k-window dataset is 15.6 GB of the 180 GB), the CSI windows as well;
* `DeviceBatchLoader` yields `(x [B,540,20], y [B,15,2])` CUDA tensors: resident mode is one gather kernel per batch
  (`wf_window_load`, optionally fused with the train.py:187-193 augmentation), streaming mode gathers from the memory-mapped
  file into double-buffered pinned staging on a worker thread and copies on a side stream while the previous batch trains;
  key points are gathered and repaired by `wf_keypoint_batch` (NPY mode, `dataset.py:80-120`) or were repaired once over
  whole sequences by `wf_keypoint_sequences` (CSV mode, `dataset.py:159-206`);
* the file-level 70/15/15 split (`dataset.py:254-325`) and the shuffled epoch order of `DataLoader(shuffle=True)` are
  reproduced draw for draw, so a seeded run sees the reference's batches.
No arithmetic happens on the host: without the CUDA library the loader raises."""
import os
import pickle
import random
import threading

import numpy as np
import torch

from . import ops
from .utils import augmentation as aug

KEEP_KEYPOINTS = list(range(15))                     # dataset.py:13


def sampler_order(n, shuffle):
    """Index order of one epoch of torch's DataLoader over n items: range(n), or -- shuffle=True -- what RandomSampler yields.
    Consumes the default CPU generator exactly like `iter(DataLoader(...))` + the first `next()` do (one draw for the iterator's
    base seed, one for the sampler's private generator), in both cases."""
    torch.empty((), dtype=torch.int64).random_()                       # _BaseDataLoaderIter._base_seed
    if not shuffle:
        return np.arange(n, dtype=np.int64)
    seed = int(torch.empty((), dtype=torch.int64).random_().item())    # RandomSampler.__iter__
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g).numpy().astype(np.int64)


def split_files(n_files, window_ranges, random_seed=42):
    """dataset.py:256-293: file-level 70/15/15 split -> three int64 arrays of window indices (python `random`, as the reference)."""
    random.seed(random_seed)                                           # dataset.py:259
    files = list(range(n_files))
    random.shuffle(files)
    a = int(np.floor(0.7 * n_files))
    b = int(np.floor(0.15 * n_files))
    parts = (files[:a], files[a:a + b], files[a + b:])
    out = []
    for part in parts:
        idx = [np.arange(window_ranges[f][0], window_ranges[f][1], dtype=np.int64) for f in part]
        out.append(np.concatenate(idx) if idx else np.zeros(0, dtype=np.int64))
    return out, parts


def shard_epoch(indices, batch_size, shuffle, drop_last=False, rank=0, world=1, order_seed=None):
    """This rank's batches of one epoch: the DataLoader order over `indices` cut into global batches of batch_size * world windows,
    of which rank keeps its contiguous shard (at most one window more or less than the other ranks in a ragged last batch).

    order_seed: None replays torch's DataLoader draws on the default CPU generator (one process: the reference's batches bit for
    bit).  With several ranks that generator does not stay in lock step (a ragged shard makes draw_augmentation consume a
    data-dependent number of draws per rank), so the order then comes from a private generator seeded by order_seed -- the same
    on every rank whatever else the ranks drew -- and a ragged global batch with fewer windows than ranks is dropped (a rank
    with an empty shard would leave its peers waiting in the gradient all-reduce)."""
    from .engine import shard_bounds
    indices = np.asarray(indices, dtype=np.int64)
    if order_seed is None:
        order = indices[sampler_order(len(indices), shuffle)]
    elif shuffle:
        g_ = torch.Generator()
        g_.manual_seed(int(order_seed))
        order = indices[torch.randperm(len(indices), generator=g_).numpy().astype(np.int64)]
    else:
        order = indices
    g = batch_size * max(1, world)
    nb = len(order) // g if drop_last else (len(order) + g - 1) // g
    out = []
    for i in range(nb):
        glob = order[i * g:(i + 1) * g]
        if world > 1 and len(glob) < world:
            break
        lo, hi = shard_bounds(len(glob), rank, world)
        out.append(glob[lo:hi])
    return out


class PreprocessedCSIKeypointsDataset:
    """Same constructor arguments, files and indexing behaviour as dataset.py:16; tensors come back on `device`."""

    def __init__(self, data_dir, keypoint_scale=1000.0, transform=None, enable_temporal_clean=True, device='cuda', resident=None,
                 resident_budget_bytes=120 << 30):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('the wiflow_b200 data path runs on a CUDA (sm_100a) device; there is no host implementation')
        self.csi_windows = np.load(os.path.join(data_dir, 'csi_windows.npy'), mmap_mode='r')
        window_info = np.load(os.path.join(data_dir, 'window_info.npz'))
        self.window_to_file = np.asarray(window_info['window_to_file']).astype(np.int64)
        self.window_to_frame = np.asarray(window_info['window_to_frame']).astype(np.int64)
        file_info = np.load(os.path.join(data_dir, 'file_info.npz'), allow_pickle=True)
        self.keypoints_files = file_info['keypoints_files']
        self.file_ids = file_info['file_ids']
        self.window_ranges = file_info['window_ranges']
        config = np.load(os.path.join(data_dir, 'config.npz'))
        self.window_size = config['window_size']
        self.stride = config['stride']
        self.keypoint_scale = keypoint_scale
        self.transform = transform
        self.enable_temporal_clean = enable_temporal_clean

        kp_path = os.path.join(data_dir, 'all_keypoints.npy')
        map_path = os.path.join(data_dir, 'file_mappings.pkl')
        self.use_npy_mode = os.path.exists(kp_path) and os.path.exists(map_path)
        n_files = len(self.keypoints_files)
        if self.use_npy_mode:
            all_kp = np.load(kp_path)
            with open(map_path, 'rb') as f:
                self.file_mappings = pickle.load(f)
            starts = np.full(n_files, -1, dtype=np.int64)
            for i, name in enumerate(self.keypoints_files):
                if name in self.file_mappings:
                    starts[i] = int(self.file_mappings[name]['start_idx'])
            st = starts[self.window_to_file]
            # dataset.py:91-101: unknown file or frame past the table -> zeros (the kernel zero-fills indices outside [0, F))
            self.frame_index = np.where(st >= 0, st + self.window_to_frame, -1).astype(np.int64)
            self.frames = torch.from_numpy(np.ascontiguousarray(all_kp, dtype=np.float32)).to(self.device)
            self._clean_per_batch = bool(enable_temporal_clean)
        else:
            seqs = [self._load_raw_keypoints(i) for i in range(n_files)]
            off = np.zeros(n_files + 1, dtype=np.int64)
            off[1:] = np.cumsum([len(s) for s in seqs])
            self.frames = torch.from_numpy(np.ascontiguousarray(np.concatenate(seqs, 0), dtype=np.float32)).to(self.device)
            if enable_temporal_clean:                       # dataset.py:208-218, once for all files instead of a 10-file cache
                ops.keypoint_sequences_(self.frames, torch.from_numpy(off).to(self.device))
            self.frame_index = (off[:-1][self.window_to_file] + self.window_to_frame).astype(np.int64)
            self._clean_per_batch = False
        self.all_keypoints = self.frames

        W = int(np.prod(self.csi_windows.shape[1:]))
        nbytes = len(self.csi_windows) * W * 4
        self.resident = (nbytes <= resident_budget_bytes) if resident is None else bool(resident)
        self.windows = None
        if self.resident:
            self.windows = torch.empty((len(self.csi_windows),) + tuple(self.csi_windows.shape[1:]), device=self.device, dtype=torch.float32)
            step = max(1, (256 << 20) // (W * 4))
            stage = torch.empty((step,) + tuple(self.csi_windows.shape[1:]), dtype=torch.float32).pin_memory()
            for i in range(0, len(self.csi_windows), step):
                n = min(step, len(self.csi_windows) - i)
                stage.numpy()[:n] = self.csi_windows[i:i + n]
                self.windows[i:i + n].copy_(stage[:n], non_blocking=True)
                torch.cuda.current_stream(self.device).synchronize()

    # -- reference surface ----------------------------------------------------------------------------------------
    def __len__(self):
        return len(self.csi_windows)

    def _load_raw_keypoints(self, file_idx):
        """dataset.py:122-157 (host: CSV parsing is I/O, not arithmetic on the path): last 50 columns / scale -> [frames, 15, 2]"""
        import pandas as pd
        data = pd.read_csv(self.keypoints_files[file_idx], header=0).values
        if data.shape[1] > 50:
            data = data[:, -50:]
        data = data.astype(np.float32) / self.keypoint_scale
        return data.reshape(len(data), 25, 2)[:, KEEP_KEYPOINTS, :]

    def batch(self, indices, out_x=None, plan=None, noise=None):
        """(x [B,540,20], y [B,15,2]) CUDA tensors of the given window indices (resident mode), `plan`: optional augmentation draws"""
        if not self.resident:
            raise RuntimeError('batch() needs the resident window array; use DeviceBatchLoader for streaming datasets')
        idx = np.asarray(indices, dtype=np.int64).reshape(-1)
        both = torch.from_numpy(np.stack([idx, self.frame_index[idx]])).to(self.device)
        if plan is not None:
            x = aug.augment_batch(None, plan=plan, noise=noise, out=out_x, windows=self.windows, idx=both[0])
        else:
            x = ops.window_load(self.windows, both[0], out_x)
        y = ops.keypoint_batch(self.frames, both[1], self._clean_per_batch)
        return x, y

    def __getitem__(self, idx):
        if idx < 0:
            idx += len(self)
        if not 0 <= idx < len(self):
            raise IndexError(idx)
        if self.resident:
            x, y = self.batch([idx])
            x = x[0]
        else:
            x = torch.from_numpy(np.array(self.csi_windows[idx], dtype=np.float32)).to(self.device)
            y = ops.keypoint_batch(self.frames, torch.tensor([self.frame_index[idx]], device=self.device), self._clean_per_batch)
        if self.transform:
            x = self.transform(x)
        return x, y[0]

    def get_file_indices(self):
        return list(range(len(self.keypoints_files)))

    def get_samples_from_file(self, file_idx):
        start_idx, end_idx = self.window_ranges[file_idx]
        return list(range(start_idx, end_idx))


class DeviceBatchLoader:
    """Iterates (x, y) CUDA batches over `indices` of a dataset in DataLoader order (batch_size, shuffle, drop_last as torch's).
    Yielded tensors stay valid until the second following batch is requested (two device slots)."""

    def __init__(self, dataset, indices=None, batch_size=64, shuffle=False, drop_last=False, augment=False, rank=0, world=1, seed=0):
        """rank / world: data-parallel training with one process per GPU.  Every rank walks the SAME epoch order (seed
        the default torch generator identically on all ranks) in global batches of batch_size * world windows and keeps its
        contiguous shard of each (engine.shard_bounds), so world processes together see exactly the batches one process would."""
        self.rank, self.world = int(rank), max(1, int(world))
        self.seed, self.epoch = int(seed), 0         # world > 1: the epoch order comes from a generator seeded by (seed, epoch)
        self.dataset = dataset
        self.indices = np.arange(len(dataset), dtype=np.int64) if indices is None else np.asarray(indices, dtype=np.int64)
        self.batch_size = int(batch_size)
        self.shuffle = bool(shuffle)
        self.drop_last = bool(drop_last)
        self.augment = bool(augment)                 # train.py:187 `use_augmentation and epoch > 0`: the caller flips it per epoch
        dev = dataset.device
        shape = (self.batch_size,) + tuple(dataset.csi_windows.shape[1:])
        self._x = [torch.empty(shape, device=dev, dtype=torch.float32) for _ in range(2)]
        self._pinned = self._copy_stream = None
        if not dataset.resident:
            self._pinned = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._copied = [None, None]

    def __len__(self):
        n, g = len(self.indices), self.batch_size * self.world
        nb = n // g if self.drop_last else (n + g - 1) // g
        if self.world > 1 and not self.drop_last and 0 < n % g < self.world:
            nb -= 1                                  # a last global batch smaller than the number of ranks is dropped
        return nb

    def epoch_batches(self):
        order_seed = None if self.world == 1 else self.seed * 1000003 + self.epoch
        self.epoch += 1
        return shard_epoch(self.indices, self.batch_size, self.shuffle, self.drop_last, self.rank, self.world, order_seed)

    # streaming mode: worker thread fills pinned slot k (gather out of the memory map), then the copy stream moves it
    def _fill(self, slot, idx):
        if self._copied[slot] is not None:
            self._copied[slot].synchronize()         # the previous copy out of this pinned slot must have finished
        dst = self._pinned[slot].numpy()[:len(idx)]
        src = self.dataset.csi_windows
        if src.dtype == np.float32:
            np.take(src, idx, axis=0, out=dst, mode='clip')        # one native gather out of the memory map
        else:
            dst[...] = src[idx]                                      # other dtypes on disk: torch.from_numpy(...).float(), dataset.py:232

    def __iter__(self):
        ds, dev = self.dataset, self.dataset.device
        batches = self.epoch_batches()
        worker = None
        if not ds.resident and batches:
            worker = threading.Thread(target=self._fill, args=(0, batches[0]))
            worker.start()
        for k, idx in enumerate(batches):
            slot = k & 1
            B = len(idx)
            plan = aug.draw_augmentation(B, ds.csi_windows.shape[1]) if self.augment else None
            if ds.resident:
                x, y = ds.batch(idx, out_x=self._x[slot][:B], plan=plan)
            else:
                worker.join()
                if k + 1 < len(batches):
                    worker = threading.Thread(target=self._fill, args=(slot ^ 1, batches[k + 1]))
                    worker.start()
                cur = torch.cuda.current_stream(dev)
                self._copy_stream.wait_stream(cur)   # slot's previous consumer kernels are ordered before the overwrite
                with torch.cuda.stream(self._copy_stream):
                    self._x[slot][:B].copy_(self._pinned[slot][:B], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self._copy_stream)
                self._copied[slot] = ev
                cur.wait_event(ev)
                x = self._x[slot][:B]
                if plan is not None:
                    x = aug.augment_batch(x, plan=plan, out=x)
                fi = torch.from_numpy(ds.frame_index[idx]).to(dev)
                y = ops.keypoint_batch(ds.frames, fi, ds._clean_per_batch)
            if ds.transform:
                x = torch.stack([ds.transform(w) for w in x])
            yield x, y


def create_preprocessed_train_val_test_loaders(dataset, batch_size=64, num_workers=0, random_seed=42, augment=False):
    """dataset.py:254-325: file-level split, shuffled train loader, ordered val/test loaders (num_workers is accepted and unused:
    there are no host workers to feed)."""
    (tr, va, te), parts = split_files(len(dataset.get_file_indices()), dataset.window_ranges, random_seed)
    print(f'train: {len(tr)} windows ({len(parts[0])} files)  val: {len(va)} ({len(parts[1])})  test: {len(te)} ({len(parts[2])})')
    return (DeviceBatchLoader(dataset, tr, batch_size, shuffle=True, augment=augment),
            DeviceBatchLoader(dataset, va, batch_size, shuffle=False),
            DeviceBatchLoader(dataset, te, batch_size, shuffle=False))
