"""GPU parity of the integer / byte stages through the C ABI: FPS, region ids, squared distances,
coalition masks (bit-exact vs oracle and vs the reference's golden vectors), reward and reductions."""
import hashlib
import types

import numpy as np
import pytest
import torch

from interpret_quality_b200 import ops, synthetic
from interpret_quality_b200 import final_save_fps, final_shapley_value
from interpret_quality_b200.tools import final_common, final_util
from oracle import coalition, geom

pytestmark = pytest.mark.gpu
R, LBL = 32, 3
DEV = "cuda:0"


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.to(DEV) if dtype is None else t.to(DEV, dtype)


@pytest.fixture(scope="module")
def geo(golden):
    return golden("geometry")


@pytest.mark.parametrize("N", [1024, 2048])
def test_region_fps_and_ids_match_reference_golden(geo, N):
    data = synthetic.make_cloud(N)
    idx = final_save_fps.farthest_point_sample(cu(data), R)
    assert idx.dtype == torch.int64 and tuple(idx.shape) == (1, R)
    assert np.array_equal(idx.cpu().numpy()[0], geo["fps_idx_%d" % N])
    rid = final_shapley_value.cal_region_id(cu(data), geo["fps_idx_%d" % N], None, save=False)
    assert rid.dtype == np.int64 and np.array_equal(rid, geo["region_id_%d" % N])


def _masked_clouds(geo):
    data = synthetic.make_cloud(1024)
    center = coalition.center_of(data)
    md = geom.mask_shapley(data[0], center, synthetic.make_orders(3, R), geo["region_id_1024"])
    rows = [r for r in geo["geo_cloud_rows"].tolist() if r >= 0]
    return np.concatenate([md[rows], data], 0)


def test_in_model_fps_on_masked_clouds_bit_exact(geo):
    clouds = _masked_clouds(geo)
    f1 = ops.fps(cu(clouds), 512).cpu().numpy()
    assert np.array_equal(f1, geo["fps512"].astype(np.int64))
    new_xyz = np.take_along_axis(clouds, f1[:, :, None].repeat(3, 2), 1)
    assert np.array_equal(ops.fps(cu(new_xyz), 128).cpu().numpy(), geo["fps128"].astype(np.int64))


@pytest.mark.parametrize("B,N,npoint", [(1, 128, 128), (7, 1000, 33), (3, 4096, 64), (2, 2048, 512)])
def test_fps_vs_oracle_ragged_sizes(B, N, npoint):
    rs = np.random.RandomState(N + npoint)
    xyz = rs.uniform(-1, 1, (B, N, 3)).astype(np.float32)
    xyz[0, N // 2:] = xyz[0, 0]                      # heavy duplicates -> ties and exhaustion
    assert np.array_equal(ops.fps(cu(xyz), npoint).cpu().numpy(), geom.fps(xyz, npoint))


def test_square_distance_bit_exact(geo):
    clouds = _masked_clouds(geo)
    f1 = geo["fps512"].astype(np.int64)
    new_xyz = np.take_along_axis(clouds, f1[:, :, None].repeat(3, 2), 1)[:, :64]
    sq = final_util.square_distance(cu(new_xyz), cu(clouds)).cpu().numpy()
    assert sha(sq) == str(geo["sqdist_sha1"])
    assert np.array_equal(sq, geom.square_distance3(new_xyz, clouds))


def test_ball_query_bit_exact(geo):
    clouds = _masked_clouds(geo)
    f1 = geo["fps512"].astype(np.int64)
    new_xyz = np.take_along_axis(clouds, f1[:, :, None].repeat(3, 2), 1)
    for radius, K in ((0.1, 16), (0.2, 32), (0.4, 128)):
        got = ops.ball_query(radius, K, cu(clouds), cu(new_xyz)).cpu().numpy().astype(np.int64)
        assert np.array_equal(got, geo["ball_r%d" % int(radius * 10)].astype(np.int64)), radius
        assert np.array_equal(got, geom.ball_query(radius, K, clouds, new_xyz))


def test_shapley_mask_bit_exact_fused_and_in_place(geo):
    data = synthetic.make_cloud(1024)
    center = coalition.center_of(data)
    orders = synthetic.make_orders(3, R)
    rid = geo["region_id_1024"]
    fused = ops.mask_shapley(cu(data[0]), cu(center), cu(orders), cu(rid)).cpu().numpy()
    assert sha(fused) == str(geo["mask_shapley_sha1"])
    args = types.SimpleNamespace(num_regions=R)
    md = cu(data).expand((R + 1) * 3, 1024, 3).clone()
    out = final_common.mask_data_batch(md, cu(center), orders, rid, args)
    assert out.data_ptr() == md.data_ptr()
    assert sha(md.cpu().numpy()) == str(geo["mask_shapley_sha1"])
    one = cu(data).expand(R + 1, 1024, 3).clone()
    final_shapley_value.mask_data(one, cu(center), orders[0], rid)
    assert np.array_equal(one.cpu().numpy(), fused[:R + 1])


@pytest.mark.parametrize("bs,Rr,N", [(1, 32, 1024), (50, 32, 1024), (5, 7, 2048), (2, 255, 512), (3, 1, 64)])
def test_shapley_mask_vs_oracle_shapes(bs, Rr, N):
    rs = np.random.RandomState(bs * 1000 + Rr)
    data = rs.uniform(-1, 1, (N, 3)).astype(np.float32)
    rid = rs.randint(0, Rr, N).astype(np.int64)
    orders = np.stack([rs.permutation(Rr) for _ in range(bs)]).astype(np.int64)
    center = data.mean(0).astype(np.float32)
    got = ops.mask_shapley(cu(data), cu(center), cu(orders), cu(rid)).cpu().numpy()
    assert np.array_equal(got, geom.mask_shapley(data, center, orders, rid))


@pytest.mark.parametrize("m", [0, 3, 30])
def test_interaction_mask_bit_exact(geo, m):
    data = synthetic.make_cloud(1024)
    center = coalition.center_of(data)
    pairs, ctxs = geo["inter_pairs"], geo["inter_ctx_m%d" % m]
    blocks = []
    for p, (ri, rj) in enumerate(pairs):
        ctx = ctxs[p].reshape(ctxs[p].shape[0], m)
        for s in range(0, ctx.shape[0], 3):
            cb = ctx[s:s + 3]
            cf = ops.mask_interaction(cu(data[0]), cu(center), cu(cb).reshape(cb.shape[0], m), ri, rj,
                                      cu(geo["region_id_1024"]), R)
            pm = ops.mask_interaction(cu(data[0]), cu(center), cu(cb).reshape(cb.shape[0], m), ri, rj,
                                      cu(geo["region_id_1024"]), R, point_major=True)
            assert torch.equal(pm.permute(0, 2, 1), cf)
            assert np.array_equal(cf.cpu().numpy(), geom.mask_interaction(data[0], center, cb, ri, rj,
                                                                        geo["region_id_1024"], R))
            blocks.append(cf.cpu().numpy())
    assert sha(np.concatenate(blocks, 0)) == str(geo["mask_inter_m%d_sha1" % m])


def test_negative_zero_follows_multiply_form():
    data = np.array([[-0.0, 1.0, -2.0], [0.5, -0.0, 0.25]] * 2, np.float32)
    center = np.zeros(3, np.float32)
    rid = np.array([0, 1, 0, 1], np.int64)
    ctx = np.array([[1]], np.int64)
    got = ops.mask_interaction(cu(data), cu(center), cu(ctx), 0, 2, cu(rid), 3).cpu().numpy()
    want = geom.mask_interaction(data, center, ctx, 0, 2, rid, 3)
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("softmax", ["modified", "normal"])
def test_reward_vs_oracle_and_golden(golden, softmax):
    g = golden("pointnet")
    logits = g["shapley_logits"]
    a = types.SimpleNamespace(softmax_type=softmax)
    v = final_common.get_reward(cu(logits), torch.tensor([LBL]), a).cpu().numpy()
    ref = g["reward_%s" % softmax]
    assert np.abs(v - ref).max() <= 1e-5 * np.abs(ref).max()
    want = coalition.reward(torch.from_numpy(logits), LBL, softmax).numpy()
    assert np.abs(v - want).max() <= 1e-5 * np.abs(want).max()
    if softmax == "modified":          # identity of SURVEY.md section 4: v = log p/(1-p)
        p = torch.softmax(torch.from_numpy(logits).double(), 1)[:, LBL].numpy()
        assert np.abs(v - np.log(p / (1 - p))).max() <= 1e-4 * np.abs(ref).max()


@pytest.mark.parametrize("bs,Rr", [(1, 32), (4, 32), (100, 32), (300, 7), (5, 255)])
def test_shapley_accumulate_matches_float64_numpy(bs, Rr):
    rs = np.random.RandomState(bs + Rr)
    v = rs.normal(size=(bs * (Rr + 1))).astype(np.float32)
    orders = np.stack([rs.permutation(Rr) for _ in range(bs)]).astype(np.int64)
    want = np.zeros(Rr)
    for p in range(bs):
        vv = v[(Rr + 1) * p:(Rr + 1) * (p + 1)]
        want[orders[p]] += (vv[1:] - vv[:-1])
    phi = torch.zeros(Rr, dtype=torch.float64, device=DEV)
    ops.shapley_accumulate(cu(v), cu(orders), phi)
    assert np.array_equal(phi.cpu().numpy(), want)          # same fp32 differences, same f64 addition order


def test_interaction_reduce_vs_golden(golden):
    from interpret_quality_b200.final_cal_interactions import compute_order_interaction
    g = golden("dgcnn")
    a = types.SimpleNamespace(softmax_type="modified")
    for m in (0, 3, 30):
        got = compute_order_interaction(cu(g["inter_logits_m%d" % m]), torch.tensor([LBL]), a)
        ref = g["inter_m%d" % m]
        assert got.dtype == np.float64 and got.shape == ref.shape
        assert np.abs(got - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-3)


def test_topk_exact_with_heavy_ties():
    rs = np.random.RandomState(0)
    for N, k in ((1024, 20), (2048, 20), (512, 64), (1024, 32), (100, 5)):
        keys = rs.normal(size=(64, N)).astype(np.float32)
        keys[1, :700] = 3.0                      # giant tie group above everything
        keys[2, ::3] = keys[2, 0]                # tie group straddling the boundary
        keys[3] = 1.0                            # all equal
        keys[4, :15] = 9.0                       # fewer than k in the top group
        keys[5] = np.round(keys[5], 1)           # many small tie groups
        for largest in (True, False):
            idx = ops.topk_rows(cu(keys), k, largest).cpu().numpy().astype(np.int64)
            kk = keys if largest else -keys
            for r in range(keys.shape[0]):
                sel = idx[r]
                assert len(set(sel.tolist())) == k and sel.min() >= 0 and sel.max() < N
                kth = np.sort(kk[r])[-k]
                assert (kk[r][sel] >= kth).all()                       # nothing below the k-th value
                assert (kk[r] > kth).sum() == (kk[r][sel] > kth).sum() # everything strictly above is taken
                ties = np.where(kk[r] == kth)[0]
                took = np.sort(sel[kk[r][sel] == kth])
                assert np.array_equal(took, ties[:len(took)])          # lowest indices among the tied


def test_knn_xyz_sets_match_reference_golden(geo):
    data = synthetic.make_cloud(1024)
    idx = np.sort(ops.knn_xyz(cu(data), 20).cpu().numpy()[0].astype(np.int64), axis=1)
    assert np.array_equal(idx, geo["knn_xyz_unmasked"].astype(np.int64))


def test_linear_fp32_engine_vs_torch():
    rs = np.random.RandomState(1)
    for M, N, K in ((300, 70, 3), (1000, 130, 6), (257, 64, 64), (4096, 256, 128), (33, 10, 256), (128, 1024, 512)):
        x = rs.normal(size=(M, K)).astype(np.float32)
        w = rs.normal(size=(N, K)).astype(np.float32)
        b = rs.normal(size=(N,)).astype(np.float32)
        want = torch.nn.functional.leaky_relu(torch.from_numpy(x).double() @ torch.from_numpy(w).double().T
                                              + torch.from_numpy(b).double(), 0.2).numpy()
        got = ops.linear(cu(x), cu(w), cu(b), act=2).cpu().numpy()
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()


def test_cpu_tensors_are_rejected_loudly():
    from interpret_quality_b200._lib import IQError
    with pytest.raises(IQError):
        ops.fps(torch.zeros(1, 16, 3), 4)
    a = types.SimpleNamespace(model="dgcnn", k=20, dataset="shapenet", device=DEV)
    model = final_util.build_model(a, synthetic.make_state_dict("dgcnn"))
    with pytest.raises(IQError):
        model(torch.zeros(1, 3, 1024))
    with pytest.raises(RuntimeError):
        model.train()(torch.zeros(1, 3, 1024, device=DEV))
