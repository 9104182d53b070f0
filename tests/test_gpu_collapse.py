"""Coalition collapse (csrc/collapse.cu, iq_model_forward_coalitions): a masked cloud is evaluated on its kept points
plus a few copies of the location the masking rule moved the absent regions to (tools/final_common.py:56-60 of the
reference).  The result must be the plain forward's: identical neighbour sets, identical max pools, the average pool
up to its summation order -- held here to 2e-6 of the logit scale against the uncollapsed path and to the north
star's 1e-3 against the reference's golden vectors."""
import types

import numpy as np
import pytest
import torch

from interpret_quality_b200 import ops, synthetic
from interpret_quality_b200.tools import final_util
from oracle import coalition, geom

pytestmark = pytest.mark.gpu
R, LBL = 32, 3
DEV = "cuda:0"
SAME = 2e-6            # collapsed vs uncollapsed, of the logit scale


def relmax(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def make(name, N=1024, k=20):
    a = types.SimpleNamespace(model=name, k=k, dataset="shapenet", feature_transform=True, device=DEV,
                              num_points=N, num_regions=R, softmax_type="modified")
    return final_util.build_model(a, synthetic.make_state_dict(name)), a


def shapley_batch(n_perm, N=1024, seed=1):
    data = synthetic.make_cloud(N)
    rid = geom.region_id(data[0], geom.fps(data, R)[0])
    center = coalition.center_of(data)
    orders = synthetic.make_orders(max(n_perm, 1), R, seed=seed)[:n_perm]
    masked = ops.mask_shapley(torch.from_numpy(data[0]).to(DEV), torch.from_numpy(center).to(DEV),
                              torch.from_numpy(orders).to(DEV), torch.from_numpy(rid).to(DEV))
    return data, rid, torch.from_numpy(center).to(DEV), orders, masked


@pytest.mark.parametrize("name", ["dgcnn", "gcnn", "pointnet"])
def test_collapsed_forward_equals_plain_forward(name):
    model, _ = make(name)
    _, _, center, _, masked = shapley_batch(6)
    plain = model.forward_point_major(masked).cpu().numpy()
    got = model.forward_point_major(masked, masked_to=center).cpu().numpy()
    frac = model.last_row_fraction()
    assert 0.4 < frac < 0.75, frac                       # a Shapley batch keeps half of its points on average
    err = relmax(got, plain)
    print("%s: collapsed vs plain %.2e of scale, rows evaluated %.3f of the batch" % (name, err, frac))
    if name == "pointnet":
        assert np.array_equal(got, plain)                # max pools only: bit-identical
    assert err <= SAME


@pytest.mark.parametrize("name", ["dgcnn", "gcnn", "pointnet"])
def test_collapsed_forward_vs_reference_golden(golden, name):
    g = golden(name)
    model, _ = make(name)
    _, rid, center, _, masked = shapley_batch(int(g["shapley_nperm"]))
    assert np.array_equal(rid, golden("geometry")["region_id_1024"])
    got = model.forward_point_major(masked, masked_to=center).cpu().numpy()
    assert model.last_row_fraction() < 0.75
    assert relmax(got, g["shapley_logits"]) <= 1e-3


@pytest.mark.parametrize("name", ["dgcnn", "gcnn"])
def test_fewer_coincident_points_than_neighbours(name):
    """A masked region smaller than k = 20 points: the cloud must keep exactly its M copies (a k-NN list would see an
    extra one), i.e. it is not compacted; a region of >= k points is."""
    model, _ = make(name)
    data = synthetic.make_cloud(1024)
    rid = geom.region_id(data[0], geom.fps(data, R)[0])
    sizes = np.bincount(rid, minlength=R)
    small, big = int(np.argmin(sizes)), int(np.argmax(sizes))
    assert sizes[small] < 20 <= sizes[big], sizes
    center = torch.from_numpy(coalition.center_of(data)).to(DEV)
    rest = [r for r in range(R) if r not in (small, big)]
    orders = np.array([rest + [big, small], rest + [small, big]], dtype=np.int64)      # row 31 masks the last region only
    masked = ops.mask_shapley(torch.from_numpy(data[0]).to(DEV), center, torch.from_numpy(orders).to(DEV),
                              torch.from_numpy(rid).to(DEV))
    x = masked[[31, 30, 64, 63]].contiguous()            # {small}, {big, small}, {big}, {small, big} masked
    plain = model.forward_point_major(x).cpu().numpy()
    got = model.forward_point_major(x, masked_to=center).cpu().numpy()
    assert relmax(got, plain) <= SAME
    assert model.last_row_fraction() == 1.0              # 1024 - 64 - 18 + 20 copies still needs 8 tiles of 128
    x1 = masked[31:32].contiguous()
    assert relmax(model.forward_point_major(x1, masked_to=center).cpu().numpy(), plain[:1]) <= SAME


@pytest.mark.parametrize("name", ["dgcnn", "gcnn", "pointnet"])
def test_sparse_and_empty_coalitions(name):
    """|S| in {0, 1, 3, R-1, R}: the all-centre cloud (128 copies), nearly empty clouds, the unmasked cloud."""
    model, _ = make(name)
    _, _, center, _, masked = shapley_batch(3, seed=5)
    rows = [0, 1, 3, 31, 32, 33, 34, 36, 65, 66, 69, 98]
    x = masked[rows].contiguous()
    plain = model.forward_point_major(x).cpu().numpy()
    got = model.forward_point_major(x, masked_to=center).cpu().numpy()
    assert relmax(got, plain) <= SAME
    # single clouds and odd batch sizes walk the one-cloud-per-bucket corner of the planner
    for r in (0, 1, 32):
        one = model.forward_point_major(masked[r:r + 1].contiguous(), masked_to=center).cpu().numpy()
        assert relmax(one, plain[rows.index(r)][None]) <= SAME


def test_no_coincident_points_and_foreign_location():
    """Clouds without any point on masked_to run exactly as in the plain forward."""
    model, _ = make("dgcnn")
    data = torch.from_numpy(synthetic.make_cloud(1024)).to(DEV)
    x = data.repeat(3, 1, 1).contiguous()
    far = torch.tensor([9.0, 9.0, 9.0], device=DEV)
    plain = model.forward_point_major(x).cpu().numpy()
    got = model.forward_point_major(x, masked_to=far).cpu().numpy()
    assert model.last_row_fraction() == 1.0
    assert np.array_equal(got, plain)


@pytest.mark.parametrize("name", ["pointnet2", "pointconv"])
def test_multiplicity_dependent_models_ignore_the_hint(name):
    """FPS, ball query and density see multiplicities: PointNet++ / PointConv must run the plain path."""
    model, _ = make(name)
    _, _, center, _, masked = shapley_batch(1)
    x = masked[::4].contiguous()
    plain = model.forward_point_major(x).cpu().numpy()
    got = model.forward_point_major(x, masked_to=center).cpu().numpy()
    assert model.last_row_fraction() == 1.0
    assert np.array_equal(got, plain)


@pytest.mark.parametrize("name", ["dgcnn", "pointnet"])
def test_channel_first_layout_and_lanes(name):
    model, _ = make(name)
    _, _, center, _, masked = shapley_batch(2)
    want = model.forward_point_major(masked, masked_to=center).cpu().numpy()
    cf = masked.permute(0, 2, 1).contiguous()
    got = model.forward_coalitions(cf, center).cpu().numpy()
    assert np.array_equal(got, want)
    for lanes in (1, 3):
        model.set_lanes(lanes)
        assert np.array_equal(model.forward_point_major(masked, masked_to=center).cpu().numpy(), want)
    model.set_chunk(5)
    assert np.array_equal(model.forward_point_major(masked, masked_to=center).cpu().numpy(), want)


def test_collapse_at_2048_points_and_small_k():
    model, _ = make("dgcnn", 2048, k=5)
    _, _, center, _, masked = shapley_batch(1, N=2048)
    x = masked[::3].contiguous()
    plain = model.forward_point_major(x).cpu().numpy()
    got = model.forward_point_major(x, masked_to=center).cpu().numpy()
    assert model.last_row_fraction() < 0.8
    assert relmax(got, plain) <= SAME


def test_masked_to_argument_checks():
    model, _ = make("pointnet")
    x = torch.zeros((2, 1024, 3), device=DEV)
    with pytest.raises(ValueError):
        model.forward_point_major(x, masked_to=torch.zeros(3))                       # host tensor
    with pytest.raises(ValueError):
        model.forward_point_major(x, masked_to=torch.zeros(4, device=DEV))
    with pytest.raises(ValueError):
        model.forward_point_major(x, out=torch.zeros((3, 10), device=DEV))           # wrong shape
    with pytest.raises(ValueError):
        model.forward_point_major(x, out=torch.zeros((2, 10), device=DEV, dtype=torch.float64))
