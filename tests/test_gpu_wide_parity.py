"""Parity at the sizes BASELINE.json quotes, against outputs of the UNMODIFIED reference (tests/golden/make_golden.py
wide_shapley / c4_interactions, generated in the build container): the full headline call (DGCNN, 100 permutations x 33
clouds), DGCNN at 2048 points, PointNet++ at 100 permutations, and config C4 (8 pairs x all 13 orders x <= 12 contexts,
DGCNN and GCNN).  Tolerance of the north star: 1e-3 of the array's scale; interactions per order, see _gates.py."""
import types

import numpy as np
import pytest
import torch

from _gates import interaction_gate, relmax
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
    scale = np.abs(ref).max()
    per_cloud = np.abs(ours - ref).max(1) / scale
    e_phi = relmax(phi, g["shapley_phi"])
    print("%s: %d clouds, logits err vs reference: max %.2e median %.2e, clouds above 1e-4: %d, above 1e-3: %d; phi err "
          "%.2e of scale; rows evaluated %.3f" % (tag, ref.shape[0], per_cloud.max(), np.median(per_cloud),
                                                 int((per_cloud > 1e-4).sum()), int((per_cloud > TOL).sum()), e_phi,
                                                 model.last_row_fraction()))
    assert e_phi <= TOL
    assert np.median(per_cloud) <= 1e-5
    if name != "dgcnn":
        assert per_cloud.max() <= TOL
        return
    # DGCNN recomputes its kNN graph in feature space: a near-tie between the k-th and (k+1)-th neighbour is decided by
    # fp32 rounding, and in a masked cloud the coincident points flip together, so a single cloud can move by > 1e-3
    # between two correct fp32 evaluations.  Every cloud further than 1e-4 from the reference is
    # therefore re-evaluated in float64 (oracle, same masked input): it passes if we are within 1e-3 of the reference OR
    # within 1e-3 of float64 while the reference's own fp32 run is the one that left it.
    out = np.nonzero(per_cloud > 1e-4)[0]
    assert len(out) <= 0.1 * len(per_cloud)
    data_np = synthetic.make_cloud(N)
    orders = synthetic.make_orders(1000, R)[:n_perm]
    sd = synthetic.make_state_dict(name)
    sd64 = {k: torch.from_numpy(v).double() if v.dtype == np.float32 else torch.from_numpy(v) for k, v in sd.items()}
    masked = geom.mask_shapley(data_np[0], coalition.center_of(data_np), orders, rid)[out]
    torch.set_default_dtype(torch.float64)
    try:
        f64 = nets.forward(name, torch.from_numpy(masked).permute(0, 2, 1).contiguous().double(), sd64).numpy()
    finally:
        torch.set_default_dtype(torch.float32)
    ours_f64 = np.abs(ours[out] - f64).max(1) / scale
    ref_f64 = np.abs(ref[out] - f64).max(1) / scale
    bad = 0
    for i, c in enumerate(out):
        ok = per_cloud[c] <= TOL or (ours_f64[i] <= TOL and ref_f64[i] >= 0.5 * per_cloud[c])
        bad += not ok
        if per_cloud[c] > TOL or not ok:
            print("  cloud %4d: ours-ref %.2e | ours-f64 %.2e | ref-f64 %.2e %s" % (c, per_cloud[c], ours_f64[i], ref_f64[i],
                                                                                "" if ok else "<-- FAIL"))
    print("%s: %d clouds re-evaluated in float64: ours-f64 max %.2e median %.2e | reference-f64 max %.2e median %.2e | "
          "clouds where the reference is the outlier (> 1e-3 from us, we within 1e-3 of float64): %d"
          % (tag, len(out), ours_f64.max(initial=0), np.median(ours_f64) if len(out) else 0, ref_f64.max(initial=0),
             np.median(ref_f64) if len(out) else 0, int((per_cloud[out] > TOL).sum()) - bad))
    assert bad == 0


@pytest.mark.parametrize("name", ["dgcnn", "gcnn"])
def test_c4_interactions_all_orders_vs_reference(golden, name):
    g, f64 = golden("c4_" + name), golden("interactions_f64")
    model, a = make(name)
    data = torch.from_numpy(synthetic.make_cloud(1024))
    rid = golden("geometry")["region_id_1024"]
    worst = 0.0
    for m in g["orders_m"]:
        ctx = g["ctx_m%d" % m]
        il = compute_order_interaction_logits(model, data, rid, g["pairs"], ctx.astype(np.float64) if m == 0 else ctx, a)
        ref_l = g["logits_m%d" % m]
        assert tuple(il.shape) == ref_l.shape
        e_l = relmax(il.cpu().numpy(), ref_l)
        worst = max(worst, e_l)
        assert e_l <= TOL, (m, e_l)
        inter = compute_order_interaction(il, torch.tensor([LBL]), a)
        err, bound = interaction_gate(inter, g["inter_m%d" % m], f64["c4_%s_m%d" % (name, m)],
                                      "%s m=%-2d logits err %.1e |" % (name, m, e_l))
        assert err <= bound, (m, err, bound)
    print("%s: worst logits error over the 13 orders %.2e of scale" % (name, worst))
