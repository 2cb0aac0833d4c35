"""CPU: the C-ABI library builds, loads, exports every symbol include/svr_abi.h declares, refuses to run
without a GPU (no fallback), and its pure-host helpers agree with the oracle."""
import ctypes as C

import numpy as np
import pytest

from fetalreconstruction_b200 import _lib
from fetalreconstruction_b200 import reconstruction as R
from oracle import oracle as orc


def test_library_exports_every_declared_symbol(built_lib):
    lib = C.CDLL(built_lib)
    declared = _lib.declared_symbols()
    assert len(declared) >= 40
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in include/svr_abi.h but not exported: {missing}"
    # and the Python binding table covers the header
    assert sorted(_lib._SIGNATURES) == declared


def test_abi_version(built_lib):
    assert _lib.load().svr_abi_version() == 2


def test_create_fails_loudly_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(R.SVRError, match="no CUDA device"):
        R.Reconstruction(0)


def test_host_slice_em_matches_oracle(built_lib):
    rng = np.random.default_rng(3)
    for trial in range(20):
        S = int(rng.integers(1, 60))
        pot = rng.uniform(0, 0.6, S).astype(np.float32)
        pot[rng.uniform(size=S) < 0.1] = -1
        scale = rng.uniform(0.1, 6.0, S).astype(np.float32) if trial % 2 else np.ones(S, np.float32)
        sw = rng.uniform(0, 1, S).astype(np.float32)
        fe = rng.integers(0, S, 2).astype(np.int32)
        sm = rng.integers(0, S, 3).astype(np.int32)
        st = np.array([0.025, 0.9, 0.1, 0.3, 0.01], np.float32)
        a = (pot.copy(), sw.copy(), st.copy())
        b = (pot.copy(), sw.copy(), st.copy())
        R.host_slice_em(a[0], scale, a[1], fe, sm, 1e-4, a[2])
        orc.host_slice_em(b[0], scale, b[1], fe, sm, 1e-4, b[2])
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)


def test_host_slice_em_all_equal_potentials_keep_all_slices(built_lib):
    S = 12
    pot = np.full(S, 0.2, np.float32)
    sw = np.ones(S, np.float32)
    st = np.array([0.025, 0.9, 0, 0, 0], np.float32)
    R.host_slice_em(pot, np.ones(S, np.float32), sw, [], [], 1e-4, st)
    np.testing.assert_array_equal(sw, 1.0)          # den2 == 0 -> mean_s2 = (max + mean)/2 == mean_s -> weight 1
    assert st[1] == pytest.approx(1.0)


def test_small_slices_rule(built_lib):
    vn = np.array([100, 120, 5, 90, 0, 110, 95], np.int32)
    # median element = sorted[round(7*0.5)] = sorted[4] = 100 -> threshold 10
    np.testing.assert_array_equal(R.host_small_slices(vn), [2, 4])
    assert R.host_small_slices(np.zeros(0, np.int32)).size == 0


@pytest.mark.parametrize("stacks,nranks", [([128] * 8, 1), ([128] * 8, 2), ([128] * 8, 4), ([128] * 8, 8),
                                           ([98, 84, 84, 84], 2), ([98, 84, 84, 84], 4), ([10, 10, 10], 8), ([7], 3)])
def test_partition_covers_all_slices_once(built_lib, stacks, nranks):
    total = sum(stacks)
    cuts = [R.host_partition(stacks, nranks, r) for r in range(nranks)]
    assert cuts[0][0] == 0 and cuts[-1][1] == total
    for (b0, e0), (b1, e1) in zip(cuts[:-1], cuts[1:]):
        assert e0 == b1 and b0 <= e0
    if len(stacks) >= nranks:                         # whole stacks only
        bounds = set(np.cumsum([0] + stacks).tolist())
        assert all(b in bounds and e in bounds for b, e in cuts)
        assert all(e > b for b, e in cuts)
    if stacks == [128] * 8:
        assert all(e - b == total // nranks for b, e in cuts)


@pytest.mark.parametrize("stacks,nranks", [([128] * 8, 1), ([128] * 8, 8), ([98, 84, 84, 84], 4), ([10, 10, 10], 8), ([7], 3), ([], 2)])
def test_strided_partition_is_a_balanced_cover(built_lib, stacks, nranks):
    """svr_host_partition_strided: every slice exactly once, every rank the same share (+-1) of EVERY stack."""
    total = sum(stacks)
    parts = [R.host_partition_strided(stacks, nranks, r) for r in range(nranks)]
    allidx = np.concatenate(parts) if parts else np.zeros(0, np.int32)
    assert sorted(allidx.tolist()) == list(range(total))
    bounds = np.cumsum([0] + stacks)
    for st in range(len(stacks)):
        per_rank = [int(np.count_nonzero((p >= bounds[st]) & (p < bounds[st + 1]))) for p in parts]
        assert max(per_rank) - min(per_rank) <= 1
    for p in parts:
        assert np.all(np.diff(p) > 0)                 # ascending global order within a rank
