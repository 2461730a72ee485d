"""Native self-test of the tcgen05 / TMEM kernels (tests/native/tc_selftest.cu, built by build.py): the pointwise-conv
forward/backward-data and backward-weights kernels against a CPU loop, first with structured operands (identity weights,
position-coded activations) that expose any operand-layout or descriptor mistake, then random data at the real layer shapes."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, 'tests', 'native', 'tc_selftest')


def test_tc_selftest():
    assert os.path.exists(EXE), 'tests/native/tc_selftest missing: run `python __graft_entry__.py` (build)'
    r = subprocess.run([EXE, '8'], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'SELFTEST PASSED' in r.stdout, r.stdout[-3000:] + r.stderr[-1000:]


SLAB_EXE = os.path.join(ROOT, 'tests', 'native', 'slab_selftest')


def test_slab_selftest():
    """TMA + tcgen05 conv-stack kernels (wf_slabtc.cu) against a CPU restatement of the ConvP contract."""
    assert os.path.exists(SLAB_EXE), 'tests/native/slab_selftest missing: run `python __graft_entry__.py` (build)'
    r = subprocess.run([SLAB_EXE, '8'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'SLAB SELFTEST PASSED' in r.stdout, r.stdout[-3000:] + r.stderr[-1000:]
