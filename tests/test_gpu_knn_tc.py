"""Fused tcgen05 feature-space kNN (csrc/knn_tc.cu) against a float64 exhaustive kNN of the same features.

knn() of the reference (models/dgcnn.py:12-18) ranks by -|xi|^2 + 2 xi.xj - |xj|^2 in fp32; ours nominates candidates
with the 3xTF32 Gram keys and decides on sum_c (xi[c]-xj[c])^2 in float64, ordered by (distance, index).  The
yardstick below is that exact rule evaluated exhaustively."""
import numpy as np
import pytest
import torch

from interpret_quality_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def exhaustive(x, k):
    """x (B,N,C) float32 -> (B,N,k) indices ordered by (float64 squared distance of the fp32 differences, index)."""
    xd = torch.from_numpy(x).to(DEV)
    out = []
    for b in range(x.shape[0]):
        diff = (xd[b][:, None, :] - xd[b][None, :, :]).double()       # fp32 subtraction first, like the kernel
        d = (diff * diff).sum(-1)
        out.append(torch.sort(d, dim=1, stable=True)[1][:, :k])        # stable: ties by lower index
        del diff
    return torch.stack(out).cpu().numpy(), None


def dist_of(x, idx):
    xd = torch.from_numpy(x).to(DEV)
    j = torch.from_numpy(idx.astype(np.int64)).to(DEV)
    out = []
    for b in range(x.shape[0]):
        nb = xd[b][j[b]]                                               # (N,k,C)
        diff = (xd[b][:, None, :] - nb).double()
        out.append((diff * diff).sum(-1))
    return torch.stack(out).cpu().numpy()


def check(x, k, expect_exhaustive_rows=None):
    got, cnt = ops.knn_features(torch.from_numpy(x).to(DEV), k, return_counts=True)
    got, cnt = got.cpu().numpy(), cnt.cpu().numpy()
    want, _ = exhaustive(x, k)
    # The SET of neighbours is the float64 decision (the fp32 fast path defers to float64 whenever the k-th and
    # (k+1)-th distances are within its error bound).  The order inside the set is unspecified: the consumer takes a max
    # over the neighbourhood, and the fast path removes the n - k farthest candidates instead of sorting them all.
    same_set = np.sort(got, -1) == np.sort(want, -1)
    if not same_set.all():
        # float64 sums in two different orders may still swap two points at the boundary when they agree to ~1e-15
        dg, dw = np.sort(dist_of(x, got), -1), np.sort(dist_of(x, want), -1)
        assert np.allclose(dg, dw, rtol=1e-12, atol=0), "neighbour set differs from the exhaustive float64 kNN"
        assert same_set.mean() > 0.9999
    assert (np.sort(got, -1)[..., 1:] != np.sort(got, -1)[..., :-1]).all(), "a neighbour is listed twice"
    redo = (cnt > 64) | (cnt < k)
    if expect_exhaustive_rows is not None:
        assert bool(redo.any()) == expect_exhaustive_rows
    return cnt


@pytest.mark.parametrize("B,N,C", [(2, 1024, 64), (3, 1024, 128), (1, 2048, 64), (2, 2048, 128), (5, 128, 64), (2, 256, 128)])
def test_random_features(B, N, C):
    rs = np.random.RandomState(B * N + C)
    x = (rs.normal(size=(B, N, C)) * rs.uniform(0.2, 2.0, size=(1, 1, C)) + rs.normal(size=(1, 1, C))).astype(np.float32)
    cnt = check(x, 20, expect_exhaustive_rows=False)
    assert cnt.min() >= 24 and cnt.mean() < 48               # the block-maxima threshold keeps the lists short


def test_masked_cloud_coincident_points():
    """Masked regions collapse hundreds of points onto one coordinate: identical feature rows, exact key ties."""
    rs = np.random.RandomState(5)
    x = rs.normal(size=(4, 1024, 64)).astype(np.float32)
    for b, m in enumerate((1, 40, 500, 1000)):
        sel = rs.permutation(1024)[:m]
        x[b, sel] = x[b, sel[0]]
    check(x, 20, expect_exhaustive_rows=False)
    x[3] = x[3, 0]                                               # a fully masked cloud: every key ties
    check(x, 20, expect_exhaustive_rows=False)


def test_smaller_k_and_surface_like_features():
    rs = np.random.RandomState(9)
    t = rs.uniform(size=(2, 1024, 2)).astype(np.float32)
    w = rs.normal(size=(2, 64)).astype(np.float32)
    x = np.tanh(t @ w).astype(np.float32)                        # a smooth 2-d sheet embedded in 64 dims
    for k in (1, 5, 20):
        check(x, k, expect_exhaustive_rows=False)


def test_pathological_column_order_takes_exhaustive_path():
    """96 points of one tight cluster sitting in 6 of the 64 strided column blocks defeat the block-maxima threshold
    (lists overflow); those rows must be redone exhaustively and still be exact."""
    rs = np.random.RandomState(3)
    x = (rs.normal(size=(1, 1024, 64)) * 3.0).astype(np.float32)
    cols = np.array([j for j in range(1024) if j % 64 < 6])      # 96 points
    x[0, cols] = (0.01 * rs.normal(size=(len(cols), 64))).astype(np.float32)
    check(x, 20, expect_exhaustive_rows=True)


def test_dgcnn_forward_uses_the_fused_knn_and_matches_fp32_engine():
    import types
    from interpret_quality_b200 import _lib, synthetic
    from interpret_quality_b200.tools import final_util
    a = types.SimpleNamespace(model="dgcnn", k=20, dataset="shapenet", device=DEV)
    model = final_util.build_model(a, synthetic.make_state_dict("dgcnn"))
    x = torch.from_numpy(synthetic.make_cloud(1024)).to(DEV).expand(3, -1, -1).contiguous()
    _lib.profile_enable(True)
    tc = model.forward_point_major(x).cpu().numpy()
    rep = _lib.profile_report()
    _lib.profile_enable(False)
    assert rep["tc_gram_knn_c64"][1] == 2 and rep["tc_gram_knn_c128"][1] == 1
    model.set_engine("fp32")
    fp = model.forward_point_major(x).cpu().numpy()
    assert np.abs(tc - fp).max() / np.abs(fp).max() <= 1e-3
