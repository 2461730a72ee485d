"""CPU tests of host-side logic that needs no GPU: the C-ABI library loads and exports every symbol include/wiflow_b200.h declares,
the product path refuses to run without a CUDA device (no CPU fallback), and the training-loop glue drives the reference's
ReduceLROnPlateau settings (train.py:112-121)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import wiflow_b200
    from wiflow_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'wiflow_b200.h')).read()
    names = sorted(set(re.findall(r'\b(wf_[a-z0-9_]+)\s*\(', hdr)))
    assert len(names) >= 10
    lib = ctypes.CDLL(_lib.LIB_PATH) if hasattr(_lib, 'LIB_PATH') else _lib.lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback():
    import wiflow_b200 as wf
    m = wf.WiFlowPoseModel()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 540, 20))
    with pytest.raises(RuntimeError):
        wf.TrainStep(m, 2)
    with pytest.raises(RuntimeError):
        wf.calculate_mpjpe(torch.zeros(2, 15, 2), torch.zeros(2, 15, 2))


def test_scheduler_settings_match_reference():
    """the Trainer's scheduler is torch's ReduceLROnPlateau with the reference's arguments: after patience=3 bad epochs (plus the
    cooldown rule) the rate halves, never below lr/1000"""
    from torch.optim.lr_scheduler import ReduceLROnPlateau
    lr = 1e-4
    ref_opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr)
    ref = ReduceLROnPlateau(ref_opt, mode='min', factor=0.5, patience=3, min_lr=lr / 1000, cooldown=1, threshold=1e-4)
    from wiflow_b200.train_loop import Trainer
    tr = Trainer.__new__(Trainer)                     # host-side parts only: no CUDA objects
    tr._lr_holder = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr)
    tr.scheduler = ReduceLROnPlateau(tr._lr_holder, mode='min', factor=0.5, patience=3, min_lr=lr / 1000, cooldown=1, threshold=1e-4)
    seq = [1.0, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.5] + [0.5] * 60
    got = []
    for v in seq:
        ref.step(v); tr.scheduler.step(v)
        assert tr.lr == ref_opt.param_groups[0]['lr']
        got.append(tr.lr)
    assert got[4] == lr and got[5] == lr / 2          # 4th bad epoch after the best one triggers the first halving
    assert min(got) == pytest.approx(lr / 1000)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the reference's CPU path on the host cores) needs no GPU: one JSON line with the contract's keys"""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype',
                'data', 'config', 'impl', 'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['value'] > 0 and d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['h2d_bytes_per_step'] == 0 and 'workload' in d['config']
