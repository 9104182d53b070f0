"""Parity at the sizes BASELINE.json quotes, against outputs of the UNMODIFIED reference (tests/golden/make_golden.py
wide_shapley / c4_interactions, generated in the build container): the full headline call (DGCNN, 100 permutations x 33
clouds), DGCNN at 2048 points, PointNet++ at 100 permutations, and config C4 (8 pairs x all 13 orders x <= 12 contexts,
DGCNN and GCNN).  Tolerance of the north star: 1e-3 of the array's scale; interactions per order, see _gates.py."""
import types

import numpy as np
import pytest
import torch

from _gates import flip_audit, interaction_gate, relmax
from interpret_quality_b200 import synthetic
from interpret_quality_b200.final_cal_interactions import compute_order_interaction
from interpret_quality_b200.final_point_binary_interaction_logits import compute_order_interaction_logits
from interpret_quality_b200.tools import final_common, final_util
from oracle import coalition, geom, nets

pytestmark = pytest.mark.gpu
R, LBL, TOL = 32, 3, 1e-3
DEV = "cuda:0"


def make(name, N=1024):
    a = types.SimpleNamespace(model=name, k=20, dataset="shapenet", feature_transform=True, device=DEV,
                              num_points=N, num_regions=R, softmax_type="modified", interaction_batch_size=25)
    return final_util.build_model(a, synthetic.make_state_dict(name)), a


@pytest.mark.parametrize("tag,name,N,bs", [("dgcnn_headline", "dgcnn", 1024, 5), ("dgcnn_2048_4", "dgcnn", 2048, 1),
                                           ("pointnet2_100", "pointnet2", 1024, 5)])
def test_shapley_call_vs_reference(golden, tag, name, N, bs):
    g = golden(tag)
    n_perm = int(g["shapley_nperm"])
    model, a = make(name, N)
    a.shapley_batch_size, a.num_samples = bs, n_perm
    data = torch.from_numpy(synthetic.make_cloud(N))
    rid = golden("geometry")["region_id_%d" % N]
    phi, logits = final_common.shap_sampling_all_regions_batch(model, data, torch.tensor([LBL]), rid,
                                                               synthetic.make_orders(1000, R), a)
    ref = g["shapley_logits"]
    assert tuple(logits.shape) == ref.shape == (n_perm * (R + 1), 10)
    ours = logits.cpu().numpy()
    e_phi = relmax(phi, g["shapley_phi"])
    print("%s: phi err %.2e of scale; rows evaluated %.3f" % (tag, e_phi, model.last_row_fraction()))
    assert e_phi <= TOL
    if name != "dgcnn":
        per_cloud = np.abs(ours - ref).max(1) / np.abs(ref).max()
        print("%s: %d clouds, logits err vs reference: max %.2e median %.2e" % (tag, len(per_cloud), per_cloud.max(),
                                                                              np.median(per_cloud)))
        assert per_cloud.max() <= TOL
        return
    data_np = synthetic.make_cloud(N)
    orders = synthetic.make_orders(1000, R)[:n_perm]

    def f64_of(idx):
        masked = geom.mask_shapley(data_np[0], coalition.center_of(data_np), orders, rid)[idx]
        return forward_f64(name, torch.from_numpy(masked).permute(0, 2, 1).contiguous())

    flip_audit(ours, ref, f64_of, tag)


def forward_f64(name, x):
    """Float64 evaluation of the network (oracle) on fp32 masked clouds x (B,3,N)."""
    sd = synthetic.make_state_dict(name)
    sd64 = {k: torch.from_numpy(v).double() if v.dtype == np.float32 else torch.from_numpy(v) for k, v in sd.items()}
    torch.set_default_dtype(torch.float64)
    try:
        return nets.forward(name, x.double(), sd64).numpy()
    finally:
        torch.set_default_dtype(torch.float32)


@pytest.mark.parametrize("name", ["dgcnn", "gcnn"])
def test_c4_interactions_all_orders_vs_reference(golden, name):
    g, f64 = golden("c4_" + name), golden("interactions_f64")
    model, a = make(name)
    data = torch.from_numpy(synthetic.make_cloud(1024))
    rid = golden("geometry")["region_id_1024"]
    worst = 0.0
    data_np = synthetic.make_cloud(1024)
    center = coalition.center_of(data_np)
    for m in g["orders_m"]:
        ctx = g["ctx_m%d" % m]
        il = compute_order_interaction_logits(model, data, rid, g["pairs"], ctx.astype(np.float64) if m == 0 else ctx, a)
        ref_l = g["logits_m%d" % m]
        assert tuple(il.shape) == ref_l.shape
        P, rows, C = ref_l.shape
        ours_l = il.cpu().numpy()
        touched = np.zeros((P, rows // 4), bool)
        if name == "dgcnn":
            # dynamic graph: audit the clouds away from the reference in float64, gate the interactions on the others
            def f64_of(idx):
                x = np.stack([geom.mask_interaction(data_np[0], center, ctx[i // rows][(i % rows) // 4][None].astype(np.int64),
                                                    g["pairs"][i // rows][0], g["pairs"][i // rows][1], rid, R)[i % 4]
                              for i in idx])
                return forward_f64(name, torch.from_numpy(x))
            away = flip_audit(ours_l.reshape(P * rows, C), ref_l.reshape(P * rows, C), f64_of, "%s m=%d" % (name, m))
            touched[away // rows, (away % rows) // 4] = True
            e_l = relmax(np.delete(ours_l.reshape(P * rows, C), away, 0), np.delete(ref_l.reshape(P * rows, C), away, 0))
        else:
            e_l = relmax(ours_l, ref_l)
            assert e_l <= TOL, (m, e_l)
        worst = max(worst, e_l)
        inter = compute_order_interaction(il, torch.tensor([LBL]), a)
        keep = ~touched
        assert keep.mean() >= 0.75
        err, bound = interaction_gate(inter[keep], g["inter_m%d" % m][keep], f64["c4_%s_m%d" % (name, m)][keep],
                                      "%s m=%-2d logits err %.1e, %d of %d contexts audited separately |"
                                      % (name, m, e_l, int(touched.sum()), touched.size))
        assert err <= bound, (m, err, bound)
    print("%s: worst logits error over the 13 orders %.2e of scale (clouds within 1e-4 of the reference)" % (name, worst))
