"""GPU parity tests of the input side through the C ABI: wf_window_load, wf_noise_scale, wf_keypoint_batch,
wf_keypoint_sequences and the Python drop-ins above them (wiflow_b200.utils.augmentation, wiflow_b200.data) against the oracle and
the fixtures produced by the reference (oracle/make_golden_data.py).

Tolerances: gathers, key-point repair, scaling and the DataLoader order are bit exact.  A time-masked span holds the fp32 mean of 64 /
540 values: the kernel sums in a different order than torch's .mean(), so masked values are compared to 2e-6 absolute (|mean| < 1);
add_noise inherits the last-bit difference of the batch-wide std: 1e-6 relative."""
import os

import numpy as np
import pytest
import torch

from oracle import data_oracle as D
from oracle.make_golden_data import write_dataset

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'wiflow_data_golden.npz')
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def g():
    z = np.load(GOLDEN)
    return {k: z[k] for k in z.files}


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def check_masked(got, want, src):
    got, want, src = got.cpu().numpy(), np.asarray(want), np.asarray(src)
    untouched = want == src
    assert np.array_equal(got[untouched], want[untouched])           # everything outside the spans is a bit-exact copy
    assert np.abs(got - want).max() <= 2e-6


def test_time_masking_golden_both_layouts(g):
    from wiflow_b200.utils import augmentation as A
    x = cu(g['tm_x'])
    torch.manual_seed(12)
    y = A.time_masking(x.permute(0, 2, 1), mask_ratio=0.7)
    assert y.shape == (8, 20, 64) and y.permute(0, 2, 1).is_contiguous()   # same strides as the reference's clone of the view
    check_masked(y.permute(0, 2, 1), g['tm_out'], g['tm_x'])
    assert torch.equal(x.cpu(), torch.from_numpy(g['tm_x']))              # input untouched (the reference clones)
    xc = x.permute(0, 2, 1).contiguous()
    torch.manual_seed(13)
    check_masked(A.time_masking(xc, mask_ratio=0.7), g['tm_out_ct'], g['tm_x'].transpose(0, 2, 1))


def test_time_masking_full_size_vs_oracle():
    from wiflow_b200.utils import augmentation as A
    x = torch.randn(96, 540, 20, generator=torch.Generator().manual_seed(3))
    torch.manual_seed(7)
    want = D.time_masking(x.permute(0, 2, 1), 0.5).permute(0, 2, 1)
    torch.manual_seed(7)
    got = A.time_masking(x.to(DEV).permute(0, 2, 1), 0.5).permute(0, 2, 1)
    assert (want != x).any()
    check_masked(got, want.numpy(), x.numpy())
    # span at the very end of the axis and two overlapping spans (the second sees the first one's result)
    spans = np.zeros((96, 2, 2), dtype=np.int32)
    spans[0] = [[531, 9], [0, 5]]
    spans[1] = [[100, 9], [104, 9]]
    plan = [[tuple(s) for s in sp if s[1] > 0] for sp in spans.tolist()]
    want = D.apply_time_masks(x.permute(0, 2, 1), plan).permute(0, 2, 1)
    got = A.time_masking(x.to(DEV).permute(0, 2, 1), spans=spans).permute(0, 2, 1)
    check_masked(got, want.numpy(), x.numpy())


def test_add_noise_and_scaling_golden(g):
    from wiflow_b200.utils import augmentation as A
    x, noise = cu(g['tm_x']), cu(g['noise'])
    got = A.add_noise(x, 0.05, noise=noise).cpu().numpy()
    assert np.abs(got - g['an_out']).max() <= 1e-6 * np.abs(g['an_out']).max()
    torch.manual_seed(15)
    for want in g['rs_out']:
        assert np.array_equal(A.random_scaling(x[:2]).cpu().numpy(), want)        # fp32 product with the same drawn factor: bit exact
    # odd sizes: statistics rows + scalar tail of the elementwise kernel
    xo = torch.randn(3, 7, 11, generator=torch.Generator().manual_seed(1))
    no = torch.randn(3, 7, 11, generator=torch.Generator().manual_seed(2))
    got = A.add_noise(xo.to(DEV), 0.1, noise=no.to(DEV)).cpu()
    assert torch.allclose(got, D.add_noise(xo, 0.1, no), rtol=1e-6, atol=1e-7)


def test_augmentation_sequence_golden(g):
    """train.py:187-193, six consecutive batches from one seed: same decisions, same arithmetic"""
    from wiflow_b200.utils import augmentation as A
    x, noise = cu(g['tm_x']), cu(g['noise'])
    torch.manual_seed(16)
    for want in g['aug_seq']:
        got = A.augment_batch(x, plan=A.draw_augmentation(8, 64), noise=noise).cpu().numpy()
        assert np.abs(got - want).max() <= 3e-6
    assert torch.equal(x.cpu(), torch.from_numpy(g['tm_x']))


def test_augmentation_full_batch_vs_oracle_and_launch_count():
    from wiflow_b200 import _lib
    from wiflow_b200.utils import augmentation as A
    B = 1024
    x = torch.randn(B, 540, 20, generator=torch.Generator().manual_seed(5))
    noise = torch.randn(B, 540, 20, generator=torch.Generator().manual_seed(6))
    xd, nd = x.to(DEV), noise.to(DEV)
    torch.manual_seed(40)
    seen = set()
    for _ in range(5):
        st = torch.get_rng_state()
        want, info = D.augment_step(x, noise)
        torch.set_rng_state(st)
        plan = A.draw_augmentation(B, 540)
        n0 = _lib.lib().wf_launch_count()
        got = A.augment_batch(xd, plan=plan, noise=nd)
        assert _lib.lib().wf_launch_count() - n0 <= 2
        seen.add((info['plan'] is not None, info['noise'], info['scale'] is not None))
        assert (got.cpu() - want).abs().max().item() <= 3e-6
    assert len(seen) >= 2


def test_keypoint_repair_bit_exact(g):
    from wiflow_b200 import ops
    frames = cu(g['kp_frames'])
    assert np.array_equal(ops.keypoint_batch(frames, None, True).cpu().numpy(), g['kp_single'])
    assert np.array_equal(ops.keypoint_batch(frames, None, False).cpu().numpy(), g['kp_frames'])
    idx = torch.tensor([3, -1, 47, 48, 0, 3, 10 ** 12], device=DEV)
    want = D.keypoint_batch(g['kp_frames'], idx.cpu().numpy(), True)
    assert np.array_equal(ops.keypoint_batch(frames, idx, True).cpu().numpy(), want)
    seq = cu(g['kp_seq_in'])
    ops.keypoint_sequences_(seq, cu(g['kp_seq_off']))
    assert np.array_equal(seq.cpu().numpy(), g['kp_seq_out'])
    rng = np.random.default_rng(9)
    lens = rng.integers(1, 300, size=40)
    big = rng.uniform(0.1, 0.9, size=(int(lens.sum()), 15, 2)).astype(np.float32)
    big[rng.uniform(size=big.shape[:2]) < 0.6] = 0
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    want = np.concatenate([D.clean_zero_keypoints(big[a:b]) for a, b in zip(off[:-1], off[1:])], 0)
    got = ops.keypoint_sequences_(cu(big), cu(off)).cpu().numpy()
    assert np.array_equal(got, want)


def test_window_gather_bit_exact_and_errors():
    from wiflow_b200 import ops
    w = torch.randn(300, 540, 20, generator=torch.Generator().manual_seed(8))
    wd = w.to(DEV)
    idx = torch.randint(0, 300, (1024,), generator=torch.Generator().manual_seed(9))
    got = ops.window_load(wd, idx.to(DEV))
    assert torch.equal(got.cpu(), w[idx])
    stats = torch.zeros(2, device=DEV, dtype=torch.float64)
    ops.window_load(wd, None, wd, None, stats)                                    # statistics only, in place
    assert torch.equal(wd.cpu(), w)
    s = stats.cpu()
    assert abs(s[0].item() - w.double().sum().item()) <= 1e-6 * w.numel() ** 0.5
    assert abs(s[1].item() / w.double().pow(2).sum().item() - 1) <= 1e-6
    assert ops.window_load(wd, torch.zeros(0, dtype=torch.int64, device=DEV)).shape == (0, 540, 20)
    with pytest.raises(RuntimeError):
        ops.window_load(w, idx)                                                   # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        ops.window_load(wd.double(), idx.to(DEV))
    with pytest.raises(RuntimeError):
        ops.window_load(wd, idx.to(DEV).int())


@pytest.mark.parametrize('resident', [True, False])
def test_dataset_and_loader_reproduce_the_reference_batches(g, tmp_path, resident):
    from wiflow_b200 import data as P
    write_dataset(str(tmp_path), g)
    ds = P.PreprocessedCSIKeypointsDataset(str(tmp_path), device=DEV, resident=resident)
    assert len(ds) == len(g['ds_csi']) and ds.use_npy_mode and ds.resident == resident
    for i in (0, 17, len(ds) - 1, -1):
        x, y = ds[i]
        assert np.array_equal(x.cpu().numpy(), g['ds_csi'][i]) and np.array_equal(y.cpu().numpy(), g['ds_items_y'][i])
    assert ds.get_samples_from_file(2) == list(range(int(g['ds_ranges'][2, 0]), int(g['ds_ranges'][2, 1])))
    tr, va, te = P.create_preprocessed_train_val_test_loaders(ds, batch_size=4, random_seed=42)
    assert np.array_equal(tr.indices, g['split_train']) and np.array_equal(te.indices, g['split_test'])
    assert len(tr) == 9 and len(va) == 2
    torch.manual_seed(31)
    for e in range(2):
        for name, loader in (('train', tr), ('val', va)):
            xs, ys = [], []
            for x, y in loader:
                xs.append(x.cpu().numpy()); ys.append(y.cpu().numpy())
            assert [len(a) for a in xs][:-1] == [4] * (len(xs) - 1)
            assert np.array_equal(np.concatenate(xs), g[f'ep{e}_{name}_x'])
            assert np.array_equal(np.concatenate(ys), g[f'ep{e}_{name}_y'])


def test_loader_feeds_the_train_step(g, tmp_path):
    """the loader's batches go straight into TrainStep.step (device tensors, ragged last batch included)"""
    import wiflow_b200 as wf
    from wiflow_b200 import data as P
    rng = np.random.default_rng(2)
    gg = dict(g)
    gg['ds_csi'] = rng.standard_normal((len(g['ds_csi']), 540, 20)).astype(np.float32)
    write_dataset(str(tmp_path), gg)
    ds = P.PreprocessedCSIKeypointsDataset(str(tmp_path), device=DEV)
    loader = P.DeviceBatchLoader(ds, None, batch_size=16, shuffle=True, augment=True)
    torch.manual_seed(0)
    model = wf.WiFlowPoseModel().to(DEV)
    step = wf.TrainStep(model, 16)
    n = 0
    for x, y in loader:
        out = step.step(x, y)
        n += x.shape[0]
    assert n == len(ds) and torch.isfinite(out).all()
    assert step.read_sums()['windows'] == len(ds)
