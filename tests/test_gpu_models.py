"""GPU parity of the masked forward pass and the whole coalition path against the reference's golden
vectors (tests/golden/*.npz) and the oracle.  Tolerance of the north star: max|ours - ref| <= 1e-3 *
max|ref| per output array (logits batch, phi vector, interaction matrix), fp32 accumulation."""
import types

import numpy as np
import pytest
import torch

from interpret_quality_b200 import synthetic
from interpret_quality_b200.final_cal_interactions import compute_order_interaction
from interpret_quality_b200.final_point_binary_interaction_logits import compute_order_interaction_logits
from interpret_quality_b200.final_shapley_value import cal_norm_factor
from interpret_quality_b200.tools import final_common, final_util
from oracle import coalition, geom, nets
from _gates import interaction_gate

pytestmark = pytest.mark.gpu
R, LBL, TOL = 32, 3, 1e-3
DEV = "cuda:0"
# models whose forward pass is implemented in libiq_b200.so
MODELS = ["pointnet", "dgcnn", "gcnn", "pointnet2", "pointconv"]


def relmax(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def make(name, N=1024):
    a = types.SimpleNamespace(model=name, k=20, dataset="shapenet", feature_transform=True, device=DEV,
                              num_points=N, num_regions=R, softmax_type="modified")
    return final_util.build_model(a, synthetic.make_state_dict(name)), a


@pytest.mark.parametrize("name", MODELS)
def test_shapley_path_vs_reference_golden(golden, name):
    g = golden(name)
    model, a = make(name)
    data = torch.from_numpy(synthetic.make_cloud(1024))                # host tensor: e2e entry
    rid = golden("geometry")["region_id_1024"]
    orders = synthetic.make_orders(16, R)
    a.shapley_batch_size, a.num_samples = 2, int(g["shapley_nperm"])
    phi, logits = final_common.shap_sampling_all_regions_batch(model, data, torch.tensor([LBL]), rid, orders, a)
    ref_logits = g["shapley_logits"]
    spread = np.abs(ref_logits - ref_logits[0:1]).max()
    assert spread > 1e-3 * np.abs(ref_logits).max()                   # non-degenerate network
    assert phi.dtype == np.float64 and phi.shape == (R,)
    assert logits.is_cuda and tuple(logits.shape) == ref_logits.shape
    assert relmax(logits.cpu().numpy(), ref_logits) <= TOL
    assert relmax(phi, g["shapley_phi"]) <= TOL
    # efficiency: sum phi = v(N) - v(empty)
    nf = cal_norm_factor(model, data.to(DEV), torch.tensor([LBL]), torch.from_numpy(coalition.center_of(data)), None,
                         a, save=False)
    assert abs(nf - float(g["norm_factor"])) <= TOL * max(1.0, abs(float(g["norm_factor"])))
    assert abs(phi.sum() - nf) <= 1e-4 * max(1.0, abs(nf))


@pytest.mark.parametrize("name", MODELS)
def test_batch_knob_and_chunk_do_not_change_results(name):
    model, a = make(name)
    data = torch.from_numpy(synthetic.make_cloud(1024)).to(DEV)
    rid = geom.region_id(synthetic.make_cloud(1024)[0], geom.fps(synthetic.make_cloud(1024), R)[0])
    orders = synthetic.make_orders(6, R)
    a.num_samples = 6
    outs = []
    for bs, chunk in ((1, None), (3, 7), (6, 200)):
        a.shapley_batch_size = bs
        if chunk:
            model.set_chunk(chunk)
        phi, logits = final_common.shap_sampling_all_regions_batch(model, data, torch.tensor([LBL]), rid, orders, a)
        outs.append((phi, logits.cpu().numpy()))
    for phi, lg in outs[1:]:
        assert np.array_equal(lg, outs[0][1])
        assert np.array_equal(phi, outs[0][0])


@pytest.mark.parametrize("name", MODELS)
def test_interaction_path_vs_reference_golden(golden, name):
    g, geo = golden(name), golden("geometry")
    model, a = make(name)
    a.interaction_batch_size = 3
    data = torch.from_numpy(synthetic.make_cloud(1024))
    for m in (0, 3, 30):
        ctx = geo["inter_ctx_m%d" % m]
        if m == 0:
            ctx = ctx.astype(np.float64)                 # an empty context list loads as float64 in the reference
        il = compute_order_interaction_logits(model, data, geo["region_id_1024"], geo["inter_pairs"], ctx, a)
        assert tuple(il.shape) == g["inter_logits_m%d" % m].shape
        assert relmax(il.cpu().numpy(), g["inter_logits_m%d" % m]) <= TOL
        inter = compute_order_interaction(il, torch.tensor([LBL]), a)
        err, bound = interaction_gate(inter, g["inter_m%d" % m], golden("interactions_f64")["%s_m%d" % (name, m)],
                                      "%s m=%-2d" % (name, m))
        assert err <= bound, (name, m, err, bound)


@pytest.mark.parametrize("name", MODELS)
def test_sparse_coalitions_vs_oracle(name):
    """|S| in {0,1,3} and |S| = R-1: collapsed clouds with huge tie groups."""
    model, a = make(name)
    sd = synthetic.make_state_dict(name)
    data = synthetic.make_cloud(1024)
    rid = geom.region_id(data[0], geom.fps(data, R)[0])
    orders = synthetic.make_orders(40, R, seed=5)[37:40]
    masked = geom.mask_shapley(data[0], coalition.center_of(data), orders, rid)
    rows = [0, 1, 3, 31, 32, 33, 34, 36, 65, 66, 69]
    x = torch.from_numpy(masked[rows]).permute(0, 2, 1).contiguous()
    want = nets.forward(name, x, sd).numpy()
    got = model(x.to(DEV))
    got = got[0] if isinstance(got, tuple) else got
    assert relmax(got.cpu().numpy(), want) <= TOL
    got_pm = model.forward_point_major(torch.from_numpy(masked[rows]).to(DEV)).cpu().numpy()
    assert np.array_equal(got_pm, got.cpu().numpy())


def test_pointnet_aux_outputs(golden):
    g = golden("pointnet")
    model, a = make("pointnet")
    x = torch.from_numpy(synthetic.make_cloud(1024)).permute(0, 2, 1).contiguous().to(DEV)
    logits, trans_feat, crt = model(x)
    assert tuple(trans_feat.shape) == (1, 64, 64) and crt.dtype == torch.int64 and tuple(crt.shape) == (1, 1024)
    assert relmax(trans_feat.cpu().numpy(), g["pointnet_trans_feat"]) <= TOL
    # the unmasked cloud has no duplicated points, so the arg-max points are unique
    assert (crt.cpu().numpy() == g["pointnet_crt"].astype(np.int64)).mean() > 0.99


def test_dgcnn_2048_points_vs_reference_golden(golden):
    g = golden("dgcnn_2048")
    model, a = make("dgcnn", 2048)
    data = torch.from_numpy(synthetic.make_cloud(2048))
    a.shapley_batch_size, a.num_samples = 1, 1
    phi, logits = final_common.shap_sampling_all_regions_batch(model, data, torch.tensor([LBL]),
                                                               golden("geometry")["region_id_2048"],
                                                               synthetic.make_orders(16, R), a)
    assert relmax(logits.cpu().numpy(), g["shapley_logits"]) <= TOL
    assert relmax(phi, g["shapley_phi"]) <= TOL


@pytest.mark.parametrize("name", ["dgcnn", "pointnet"])
def test_full_size_properties(name):
    """BASELINE size (100 permutations x 33 clouds): size-independent properties instead of an oracle run."""
    model, a = make(name)
    data = torch.from_numpy(synthetic.make_cloud(1024))
    rid = geom.region_id(data[0].numpy(), geom.fps(data.numpy(), R)[0])
    orders = synthetic.make_orders(100, R)
    a.shapley_batch_size, a.num_samples = 5, 100
    phi, logits = final_common.shap_sampling_all_regions_batch(model, data, torch.tensor([LBL]), rid, orders, a)
    lg = logits.cpu().numpy().reshape(100, R + 1, -1)
    # row 0 of every permutation is the all-centre cloud, row R the unmasked cloud: identical logits across perms
    assert np.array_equal(lg[:, 0], np.broadcast_to(lg[0, 0], lg[:, 0].shape))
    assert np.array_equal(lg[:, R], np.broadcast_to(lg[0, R], lg[:, R].shape))
    v = coalition.reward(torch.from_numpy(lg.reshape(-1, lg.shape[-1])), LBL).numpy().reshape(100, R + 1)
    assert abs(phi.sum() - (v[0, R] - v[0, 0])) <= 1e-4 * max(1.0, abs(v[0, R] - v[0, 0]))
    want = coalition.shapley_from_logits(logits.cpu(), LBL, orders, R, 100)
    assert np.abs(phi - want).max() <= 1e-5 * max(np.abs(want).max(), 1e-6)


@pytest.mark.parametrize("name", MODELS)
def test_error_vs_float64_oracle_is_fp32_noise(name):
    """The reference's own fp32 CPU run sits up to ~1e-3 from a float64 evaluation of the same network on some
    masked clouds (a near-tie kNN neighbour flips in DGCNN); the CUDA path must sit at fp32 noise from float64."""
    model, a = make(name)
    sd = synthetic.make_state_dict(name)
    data = synthetic.make_cloud(1024)
    rid = geom.region_id(data[0], geom.fps(data, R)[0])
    masked = geom.mask_shapley(data[0], coalition.center_of(data), synthetic.make_orders(1, R), rid)[::2]
    x = torch.from_numpy(masked).permute(0, 2, 1).contiguous()
    sd64 = {k: torch.from_numpy(v).double() if v.dtype == np.float32 else torch.from_numpy(v) for k, v in sd.items()}
    torch.set_default_dtype(torch.float64)
    try:
        want = nets.forward(name, x.double(), sd64).numpy()
    finally:
        torch.set_default_dtype(torch.float32)
    got = model(x.to(DEV))
    got = (got[0] if isinstance(got, tuple) else got).cpu().numpy()
    per_cloud = np.abs(got - want).max(1) / np.abs(want).max()
    if name == "dgcnn":
        # the dynamic kNN is discontinuous: fp32 vs float64 *features* can pick a different neighbour at a near tie
        # (the reference's own fp32 run does); typical clouds must still sit at fp32 noise
        assert np.median(per_cloud) <= 1e-5 and per_cloud.max() <= 1e-3
    else:
        assert per_cloud.max() <= 5e-5


@pytest.mark.parametrize("name", ["gcnn", "pointnet2", "pointconv"])
def test_fp32_engine_agrees_with_tcgen05_engine(name):
    model, a = make(name)
    data = synthetic.make_cloud(1024)
    rid = geom.region_id(data[0], geom.fps(data, R)[0])
    masked = geom.mask_shapley(data[0], coalition.center_of(data), synthetic.make_orders(1, R), rid)[::3]
    x = torch.from_numpy(masked).to(DEV)
    model.set_engine("3xtf32")
    tc = model.forward_point_major(x).cpu().numpy()
    model.set_engine("fp32")
    fp = model.forward_point_major(x).cpu().numpy()
    assert relmax(tc, fp) <= 5e-5


def test_dgcnn_wide_sample_vs_reference_golden(golden):
    """528 masked clouds (16 permutations) against the reference's fp32 CPU logits.  The dynamic kNN makes single
    clouds jump by a few 1e-4 when a near-tie neighbour is chosen differently (coincident masked points flip together);
    the bar stays 1e-3 of scale for every cloud and almost all clouds must sit at fp32 noise."""
    from interpret_quality_b200 import ops
    model, a = make("dgcnn")
    geo = golden("geometry")
    ref = np.concatenate([golden("dgcnn")["shapley_logits"], golden("dgcnn_more")["shapley_logits"]], 0)
    data = synthetic.make_cloud(1024)
    masked = ops.mask_shapley(torch.from_numpy(data[0]).to(DEV), torch.from_numpy(np.asarray(geo["center"])).to(DEV),
                              torch.from_numpy(synthetic.make_orders(16, R)).to(DEV),
                              torch.from_numpy(geo["region_id_1024"]).to(DEV))
    out = model.forward_point_major(masked).cpu().numpy()
    err = np.abs(out - ref).max(1) / np.abs(ref).max()
    assert err.max() <= TOL
    assert np.median(err) <= 1e-5
    assert (err > 1e-4).mean() <= 0.03
