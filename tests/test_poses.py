"""Pose-enumeration runner: pose grids and disturb functions against the reference's own
outputs (tests/golden/poses.npz, bitwise on CPU), and the runner's files / values on the GPU."""
import os
import types

import numpy as np
import pytest
import torch

from interpret_quality_b200 import final_rotate_center_enum_all as rot
from interpret_quality_b200 import final_scale_center_enum_all as sca
from interpret_quality_b200 import final_trans_center_enum_all as tra
from interpret_quality_b200 import synthetic

R, LBL = 32, 3


def pose_args():
    return types.SimpleNamespace(trans_dist_threshold=tra.TRANS_DIST_THRESHOLD, num_grid_enum_trans=tra.NUM_GRID_ENUM_TRANS,
                                 angle_threshold=rot.ANGLE_THRESHOLD, num_grid_enum_rotate=rot.NUM_GRID_ENUM_ROTATE,
                                 scale_lower=sca.SCALE_LOWER, scale_upper=sca.SCALE_UPPER,
                                 num_grid_enum_scale=sca.NUM_GRID_ENUM_SCALE)


def test_pose_grids_match_the_reference_bitwise(golden):
    g, a, cpu = golden("poses"), pose_args(), torch.device("cpu")
    tv, ra, sc = tra.generate_trans_vector(a, cpu), rot.generate_rotate_angle(a, cpu), sca.generate_scale(a, cpu)
    assert tv.dtype == ra.dtype == sc.dtype == torch.float32
    assert np.array_equal(tv.numpy(), g["trans_vector"]) and tv.shape == (216, 3)
    assert np.array_equal(ra.numpy(), g["rotate_angle"]) and ra.shape == (216, 3)
    assert np.array_equal(sc.numpy(), g["scale"]) and sc.shape == (30,)
    assert float(torch.norm(tv, dim=1).max()) <= 0.5 + 1e-6                 # clipped to the ball


def test_disturb_functions_match_the_reference_bitwise(golden):
    g = golden("poses")
    data = torch.from_numpy(synthetic.make_cloud(1024))
    assert np.array_equal(tra.translate_pc(data, torch.from_numpy(g["trans_vector"][7])).numpy(), g["translated_7"])
    assert np.array_equal(rot.rotate_xyz(data, torch.from_numpy(g["rotate_angle"][11])).numpy(), g["rotated_11"])
    assert np.array_equal(sca.scale_pc(data, torch.from_numpy(g["scale"])[3]).numpy(), g["scaled_3"])


def test_rotation_is_rigid():
    data = torch.from_numpy(synthetic.make_cloud(256, seed=3))
    out = rot.rotate_xyz(data, torch.tensor([0.3, -0.2, 0.7]))
    pair = lambda p: ((p[:, None, :].double() - p[None, :, :].double()) ** 2).sum(-1)
    assert torch.allclose(pair(data[0]), pair(out[0]), atol=1e-5)


def test_runner_needs_samples():
    from interpret_quality_b200.tools import final_common
    with pytest.raises(ValueError):
        final_common.test(types.SimpleNamespace(), None, None, None, None)


@pytest.mark.gpu
def test_runner_writes_the_reference_files_and_values(golden, tmp_path):
    """Three scale poses of the synthetic cloud through PointNet: values against the reference's own sampler
    (poses.npz), files and dtypes as tools/final_common.py:150-172 writes them."""
    from interpret_quality_b200 import ops
    from interpret_quality_b200.tools import final_common, final_util
    dev = torch.device("cuda:0")
    data = torch.from_numpy(synthetic.make_cloud(1024))
    geo = golden("geometry")
    exp = str(tmp_path) + "/exp/"
    folder = exp + "cloud0/"
    os.makedirs(folder)
    np.save(folder + "norm_factor.npy", 1.0)
    np.save(folder + "region_id.npy", geo["region_id_1024"])
    np.save(folder + "all_orders.npy", synthetic.make_orders(8, R))
    args = types.SimpleNamespace(model="pointnet", k=20, dataset="shapenet", feature_transform=True, device=dev,
                                 num_points=1024, num_regions=R, shapley_batch_size=2, num_samples=4,
                                 softmax_type="modified", mode=sca.MODE, exp_folder=exp, scale_lower=0.5, scale_upper=2.0,
                                 num_grid_enum_scale=3)
    model = final_util.build_model(args, synthetic.make_state_dict("pointnet"))
    final_common.test(args, sca.generate_scale, sca.scale_pc, sca.print_scale_info, sca.save_scale_info,
                      samples=[(data, torch.tensor([LBL]), "cloud0")], model=model)
    out = folder + "scale_all/"
    phi = np.load(out + "region_shapley_value.npy")
    orig = np.load(out + "orig_shapley_value.npy")
    logits = torch.load(out + "all_logits.pt")
    scale = np.load(out + "scale.npy")
    assert phi.shape == (3, R) and phi.dtype == np.float64 and orig.shape == (R,) and orig.dtype == np.float64
    assert tuple(logits.shape) == (3, 4 * (R + 1), 10) and logits.dtype == torch.float32
    assert np.array_equal(scale, np.array([0.5, 1.25, 2.0], dtype=np.float32))
    want = golden("poses")["scale3_pointnet_phi"]
    for i in range(3):
        assert np.abs(phi[i] - want[i]).max() <= 1e-3 * np.abs(want[i]).max()
    assert os.path.exists(out + "log.txt")
    # the reference's table code (final_result.py:83-102) consumes these files unchanged
    from interpret_quality_b200 import final_result
    sens = final_result.cal_sensitivity(folder, "scale")
    want_sens = final_result.sensitivity_of(want)
    assert sens.shape == (R,) and np.abs(sens - want_sens).max() <= 2e-3 * want_sens.max()
    _ = ops
