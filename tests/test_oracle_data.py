"""CPU tests of the input side: the oracle restatement against the fixtures produced by the reference
(tests/golden/wiflow_data_golden.npz, oracle/make_golden_data.py) and against the live reference when it is mounted; the host-side
logic of the product (random draws in the reference's order, DataLoader epoch order, file-level split) against the same fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import data_oracle as D
from oracle import load_reference as L

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'wiflow_data_golden.npz')


@pytest.fixture(scope='module')
def g():
    z = np.load(GOLDEN)
    return {k: z[k] for k in z.files}


def test_oracle_time_masking_matches_golden(g):
    x = torch.from_numpy(g['tm_x'])
    torch.manual_seed(12)
    assert np.array_equal(D.time_masking(x.permute(0, 2, 1), mask_ratio=0.7).permute(0, 2, 1).numpy(), g['tm_out'])
    torch.manual_seed(13)
    assert np.array_equal(D.time_masking(x.permute(0, 2, 1).contiguous(), mask_ratio=0.7).numpy(), g['tm_out_ct'])
    assert (g['tm_out'] != g['tm_x']).any()


def test_oracle_noise_scale_sequence_match_golden(g):
    x, noise = torch.from_numpy(g['tm_x']), torch.from_numpy(g['noise'])
    assert np.array_equal(D.add_noise(x, 0.05, noise).numpy(), g['an_out'])
    torch.manual_seed(15)
    got = np.stack([D.random_scaling(x[:2]).numpy() for _ in range(6)])
    assert np.array_equal(got, g['rs_out'])
    assert 0 < sum(np.array_equal(o, g['tm_x'][:2]) for o in got) < 6      # both branches were taken
    torch.manual_seed(16)
    kinds = set()
    for want in g['aug_seq']:
        got, info = D.augment_step(x, noise)
        assert np.array_equal(got.contiguous().numpy(), want)
        kinds.add((info['plan'] is not None, info['noise'], info['scale'] is not None))
    assert len(kinds) >= 3


def test_oracle_keypoints_match_golden(g):
    got = np.stack([D.clean_single_frame_zeros(f) for f in g['kp_frames']])
    assert np.array_equal(got, g['kp_single'])
    off = g['kp_seq_off']
    got = np.concatenate([D.clean_zero_keypoints(g['kp_seq_in'][a:b]) for a, b in zip(off[:-1], off[1:])], 0)
    assert np.array_equal(got, g['kp_seq_out'])
    assert (g['kp_seq_out'] != g['kp_seq_in']).any()


def _frame_index(g):
    st = g['ds_starts'][g['ds_w2file']]
    return np.where(st >= 0, st + g['ds_w2frame'], -1)


def test_oracle_dataset_items_and_epochs_match_golden(g):
    fi = _frame_index(g)
    assert np.array_equal(D.keypoint_batch(g['ds_all_kp'], fi, True), g['ds_items_y'])
    assert (g['ds_items_y'][g['ds_w2file'] == 4] == 0).all() and (g['ds_items_y'][-2:] == 0).all()
    torch.manual_seed(31)
    for e in range(2):
        order = g['split_train'][D.loader_order(len(g['split_train']), True)]
        assert np.array_equal(g['ds_csi'][order], g[f'ep{e}_train_x'])
        assert np.array_equal(g['ds_items_y'][order], g[f'ep{e}_train_y'])
        order = g['split_val'][D.loader_order(len(g['split_val']), False)]
        assert np.array_equal(g['ds_csi'][order], g[f'ep{e}_val_x'])
    assert not np.array_equal(g['ep0_train_x'], g['ep1_train_x'])


# ---- host logic of the product (no GPU needed) ----
def test_product_draws_follow_the_reference_order(g):
    from wiflow_b200.utils import augmentation as A
    for seed in (12, 99):
        torch.manual_seed(seed)
        plan = D.draw_time_masks(8, 64, 0.7)
        torch.manual_seed(seed)
        spans = A.draw_time_masks(8, 64, 0.7)
        for i, sp in enumerate(plan):
            want = np.zeros((2, 2), dtype=np.int32)
            for k, (s, n) in enumerate(sp):
                want[k] = (s, n)
            assert np.array_equal(spans[i], want)
    x = torch.from_numpy(g['tm_x'])
    torch.manual_seed(16)
    infos = [D.augment_step(x, torch.from_numpy(g['noise']))[1] for _ in range(6)]
    torch.manual_seed(16)
    for info in infos:
        spans, use_noise, scale = A.draw_augmentation(8, 64)
        assert (spans is not None) == (info['plan'] is not None) and use_noise == info['noise']
        assert (scale is None) == (info['scale'] is None) and (scale is None or scale == info['scale'])


def test_product_epoch_order_and_split(g):
    from wiflow_b200 import data as P
    from torch.utils.data import DataLoader, TensorDataset
    ds = TensorDataset(torch.arange(37))
    torch.manual_seed(5)
    want = [torch.cat([b[0] for b in DataLoader(ds, batch_size=8, shuffle=sh)]).numpy() for sh in (True, False, True)]
    torch.manual_seed(5)
    got = [P.sampler_order(37, sh) for sh in (True, False, True)]
    for w, o in zip(want, got):
        assert np.array_equal(w, o)
    (tr, va, te), _ = P.split_files(len(g['ds_ranges']), g['ds_ranges'], 42)
    assert np.array_equal(tr, g['split_train']) and np.array_equal(va, g['split_val']) and np.array_equal(te, g['split_test'])


def test_sharded_epoch_covers_the_single_process_batches():
    """SURVEY 8e x 8f-3: world processes that seed identically see, together, exactly the batches of one process"""
    from wiflow_b200 import data as P
    idx = np.arange(100, 100 + 53)
    for world in (2, 4):
        torch.manual_seed(9)
        single = P.shard_epoch(idx, 4 * world, True)
        shards = []
        for rank in range(world):
            torch.manual_seed(9)
            shards.append(P.shard_epoch(idx, 4, True, False, rank, world))
        assert all(len(s) == len(single) for s in shards)
        for b, glob in enumerate(single):
            parts = [s[b] for s in shards]
            assert np.array_equal(np.concatenate(parts), glob)
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    torch.manual_seed(9)
    assert sum(len(b) for b in P.shard_epoch(idx, 4, True, True, 1, 2)) == (53 // 8) * 4       # drop_last drops the ragged global batch


def test_product_data_path_has_no_host_fallback(tmp_path):
    from wiflow_b200 import data as P
    from wiflow_b200.utils import augmentation as A
    with pytest.raises(RuntimeError):
        P.PreprocessedCSIKeypointsDataset(str(tmp_path), device='cpu')
    for fn in (lambda: A.time_masking(torch.zeros(2, 20, 64)), lambda: A.add_noise(torch.zeros(2, 64, 20))):
        with pytest.raises(RuntimeError):
            fn()


@pytest.mark.skipif(not L.available(), reason='reference tree not mounted (GPU box)')
def test_oracle_vs_live_reference():
    R = L.load_data()
    x = torch.randn(5, 540, 20, generator=torch.Generator().manual_seed(3))
    for seed in (1, 2):
        torch.manual_seed(seed)
        want = R.time_masking(x.permute(0, 2, 1), mask_ratio=0.8)
        torch.manual_seed(seed)
        got = D.time_masking(x.permute(0, 2, 1), mask_ratio=0.8)
        assert torch.equal(want, got)
        torch.manual_seed(seed)
        want = R.random_scaling(x)
        torch.manual_seed(seed)
        assert torch.equal(want, D.random_scaling(x))
    rng = np.random.default_rng(0)
    dummy = type('D', (), {})()
    for _ in range(20):
        seq = rng.uniform(0.1, 0.9, size=(int(rng.integers(1, 40)), 15, 2)).astype(np.float32)
        seq[rng.uniform(size=seq.shape[:2]) < 0.5] = 0
        assert np.array_equal(R.PreprocessedCSIKeypointsDataset._clean_zero_keypoints(dummy, seq), D.clean_zero_keypoints(seq))
        for f in seq[:5]:
            assert np.array_equal(R.PreprocessedCSIKeypointsDataset._clean_single_frame_zeros(dummy, f), D.clean_single_frame_zeros(f))
