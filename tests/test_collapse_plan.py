"""Host half of the coalition collapse (iq_collapse_plan, models_common.cu::collapse_plan): which compacted size every
masked cloud is evaluated at and in which order.  Pure host code, so it is held to its contract here on the CPU:
sizes are multiples of 128 that keep every kept point and min(M, copies) copies of the masking location (a kNN list can
see at most `copies` of them; with fewer than `copies` masked points the cloud is not compacted at all), the order is a
stable sort by size (largest first), and `extra` is the multiplicity the average pool still owes the last copy."""
import ctypes

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from interpret_quality_b200 import _lib


def plan(kept, N, copies):
    lib = _lib.load()
    kept = np.ascontiguousarray(kept, np.int32)
    B = kept.size
    src, size = np.empty(B, np.int32), np.empty(B, np.int32)
    extra = np.empty(B, np.float32)
    count = np.zeros(N // 128 + 1, np.int64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.iq_collapse_plan(p(kept), B, N, copies, p(src), p(size), p(extra), p(count))
    if rc != 0:
        raise _lib.IQError(_lib.last_error())
    return src, size, extra, count


@settings(max_examples=200, deadline=None)
@given(st.sampled_from([128, 1024, 2048]), st.sampled_from([1, 5, 20]), st.data())
def test_plan_contract(N, copies, data):
    kept = np.array(data.draw(st.lists(st.integers(0, N), min_size=0, max_size=60)), np.int32)
    src, size, extra, count = plan(kept, N, copies)
    assert sorted(src.tolist()) == list(range(kept.size))                     # a permutation of the clouds
    U = kept[src].astype(np.int64)
    M = N - U
    n = size.astype(np.int64)
    assert np.all(n % 128 == 0) and np.all(n >= 128) and np.all(n <= N)
    assert np.all(n >= U + np.minimum(M, copies))                             # every kept point + enough copies
    assert np.all(n - 128 < U + np.minimum(M, copies))                        # and not a tile more than that
    assert np.all(n[M == 0] == N)
    small = (M > 0) & (M < copies)                                            # fewer masked points than a kNN list is long:
    assert np.all(n[small] == N) and np.all(extra[small] == 0)                # exactly the M copies the cloud has
    assert np.all(extra == (M - (n - U)).astype(np.float32)) and np.all(extra >= 0)
    assert np.all(np.diff(n) <= 0)                                            # largest first ...
    for t in np.unique(n):
        assert np.all(np.diff(src[n == t]) > 0)                               # ... original order inside a size
        assert count[t // 128] == int((n == t).sum())
    assert count.sum() == kept.size and count[0] == 0


def test_shapley_batch_sizes():
    """The 33 rows of one permutation with 32 regions of 32 points: row r keeps r regions."""
    kept = np.arange(33, dtype=np.int32) * 32
    src, size, extra, count = plan(kept, 1024, 20)
    n = np.empty(33, np.int64)
    n[src] = size
    assert n[0] == 128 and n[3] == 128 and n[4] == 256 and n[31] == 1024 and n[32] == 1024
    assert size.sum() / (33 * 1024) < 0.62
    e = np.empty(33, np.float32)
    e[src] = extra
    assert e[0] == 1024 - 128 and e[32] == 0 and e[31] == 0                   # row 31: 32 masked points, 1024 - 992 = 32 copies


def test_plan_rejects_bad_arguments():
    with pytest.raises(_lib.IQError):
        plan(np.array([5], np.int32), 1000, 20)                               # N not a multiple of 128
    with pytest.raises(_lib.IQError):
        plan(np.array([2000], np.int32), 1024, 20)                            # more kept points than points
