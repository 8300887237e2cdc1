"""pytest configuration: registers the `gpu` marker and makes the repo root importable.

`-m "not gpu"` : oracle vs golden vectors, host logic, C-ABI symbol export (no CUDA calls).
`-m gpu`       : parity tests proper -- CUDA path (through the C-ABI) vs oracle / golden vectors.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False))


def rel_err(a, b, scale=None):
    """max |a-b| / max |b|  -- the 'relative' of the north_star's 1e-5 bound, taken over the tensor.

    `scale` overrides the denominator where the reference's own fp32 rounding is relative to a larger
    operand: the straight-through output is fl(x + fl(q - x)) (vector_quantizer_ema.py:169), whose
    rounding error scales with max(|x|, |q|), not with |q|."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if not b.size:
        return 0.0
    denom = max(float(np.max(np.abs(b))) if scale is None else float(scale), 1e-30)
    return float(np.max(np.abs(a - b))) / denom


@pytest.fixture(scope='session')
def golden():
    return load_golden
