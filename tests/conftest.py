"""pytest configuration: registers the `gpu` marker and shares seeded datasets.

`-m "not gpu"` tests run on CPU: the oracle against analytic known answers, the host logic (including
the world_size-2 gloo run of the pipeline) and the C-ABI export check.  `-m gpu` tests are the parity
tests proper and call the CUDA path through the C ABI.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    from fetalreconstruction_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def small_ds():
    from fetalreconstruction_b200.phantom import make_dataset, small_config
    return make_dataset(small_config())


@pytest.fixture(scope="session")
def tiny_ds():
    from fetalreconstruction_b200.phantom import make_dataset, small_config
    return make_dataset(small_config(seed=11, vol=24, n_stacks=2, slices=5, size=20, inplane=1.2, spacing=2.5))


def rel_stats(a, b, scale=None):
    """max-abs and RMS difference relative to `scale` (default: RMS of b over its non-zeros)."""
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    if scale is None:
        nz = b[b != 0]
        scale = float(np.sqrt(np.mean(nz ** 2))) if nz.size else 1.0
    d = np.abs(a - b)
    return float(d.max() / scale), float(np.sqrt(np.mean(d ** 2)) / scale)
