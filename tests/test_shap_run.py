"""final_shapley_value.shap_sampling / save_shapley: the 100-permutation run of the
synthetic cloud through PointNet against the reference's own run (tests/golden/shap_run.npz), file for file."""
import os
import types

import numpy as np
import pytest
import torch

from interpret_quality_b200 import synthetic

R, LBL = 32, 3


def test_save_shapley_layout(tmp_path):
    from interpret_quality_b200.final_shapley_value import save_shapley
    a = types.SimpleNamespace(num_points=8, num_regions=2)
    rid = np.array([0, 1, 1, 0, 0, 1, 0, 1])
    save_shapley(np.array([4.0, -2.0]), 3, 2, str(tmp_path) + "/", rid, a)
    pts = np.load(str(tmp_path) + "/shapley/3_2.npy")
    reg = np.load(str(tmp_path) + "/region_shapley/3_2.npy")
    assert pts.dtype == np.float64 and np.array_equal(pts, np.where(rid == 0, 2.0, -1.0))
    assert np.array_equal(reg, np.array([2.0, -1.0]))


@pytest.mark.gpu
def test_shap_sampling_run_matches_the_reference(golden, tmp_path):
    from interpret_quality_b200 import final_shapley_value as fsv
    from interpret_quality_b200.tools import final_util
    g, geo = golden("shap_run"), golden("geometry")
    dev = torch.device("cuda:0")
    data = torch.from_numpy(synthetic.make_cloud(1024))
    exp = str(tmp_path) + "/exp/"
    a = types.SimpleNamespace(model="pointnet", k=20, dataset="shapenet", feature_transform=True, device=dev, num_points=1024,
                              num_regions=R, num_samples_save=100, num_samples=100, shapley_batch_size=50,
                              softmax_type="modified", exp_folder=exp)
    model = final_util.build_model(a, synthetic.make_state_dict("pointnet"))
    final_util.set_random(1)
    fsv.shap_sampling(model, [(data, torch.tensor([LBL]))], a, ["cloud0"], fps_indices=geo["fps_idx_1024"].reshape(1, -1))
    out = exp + "cloud0/"
    assert np.array_equal(np.load(out + "region_id.npy"), g["region_id"])           # integers: bit exact
    assert np.array_equal(np.load(out + "all_orders.npy"), g["all_orders"])         # seed replay
    rel = lambda got, want: np.abs(got - want).max() / np.abs(want).max()
    assert abs(float(np.load(out + "norm_factor.npy")) - float(g["norm_factor"])) <= 1e-3 * abs(float(g["norm_factor"]))
    sv_all = np.load(out + "region_sv_all.npy")
    assert sv_all.shape == (100, R) and sv_all.dtype == np.float64 and rel(sv_all, g["region_sv_all"]) <= 1e-3
    assert rel(np.load(out + "region_shapley/0_100.npy"), g["region_shapley_100"]) <= 1e-3
    pts = np.load(out + "shapley/0_100.npy")
    assert pts.shape == (1024,) and rel(pts, g["shapley_100"]) <= 1e-3
    # efficiency: the marginal contributions of a permutation telescope to v(N) - v(empty)
    assert np.allclose(sv_all.sum(1), float(np.load(out + "norm_factor.npy")), rtol=1e-4, atol=1e-4)
    assert sorted(os.listdir(out + "shapley")) == ["0_100.npy"]
