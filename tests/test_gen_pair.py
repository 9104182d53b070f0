"""final_gen_pair.py mirror: seed-replayed pairs / contexts against the reference's own
output (tests/golden/gen_pair.npz, bitwise), and the all-poses forward of check_adv_success on the GPU."""
import os
import types

import numpy as np
import pytest
import torch

from interpret_quality_b200 import final_gen_pair as gp
from interpret_quality_b200 import synthetic
from interpret_quality_b200.tools import final_util

R, LBL = 32, 3


def test_pairs_and_contexts_replay_the_reference_stream(golden, tmp_path):
    g = golden("gen_pair")
    a = types.SimpleNamespace(num_regions=R, num_pairs_random=6, ratio=[0.0, 0.04, 0.1, 0.5, 0.94, 1.0], num_save_context_max=100)
    final_util.set_random(1)
    pairs = gp.gen_pair_random(a)
    assert pairs.dtype == np.int64 and np.array_equal(pairs, g["pairs"]) and (pairs[:, 1] > pairs[:, 0]).all()
    d = str(tmp_path) + "/"
    ctx = gp.gen_context(pairs, d, a)
    for r in a.ratio:
        pct = int(r * 100)
        saved = np.load(d + "ratio%d_context_list.npy" % pct)
        assert saved.shape == g["ctx%d" % pct].shape and np.array_equal(saved, g["ctx%d" % pct])
        assert np.array_equal(ctx[pct], saved)
        for p in range(6):                                       # contexts never contain the pair itself
            assert not np.isin(saved[p], pairs[p]).any()


@pytest.mark.gpu
def test_check_adv_success_and_pred_label(tmp_path):
    """216 rotation poses in one forward (GCNN): the pose with the lowest ground-truth reward must also be (within
    tolerance) the oracle's, and the files carry the reference's names and shapes."""
    from interpret_quality_b200 import final_rotate_center_enum_all as rot
    from oracle import nets
    dev = torch.device("cuda:0")
    data = torch.from_numpy(synthetic.make_cloud(1024))
    exp = str(tmp_path) + "/exp/"
    os.makedirs(exp + "cloud0/rotate_all/")
    a = types.SimpleNamespace(model="gcnn", k=20, dataset="shapenet", device=dev, num_points=1024, num_regions=R,
                              softmax_type="modified", mode="rotate", exp_folder=exp, seed=1,
                              angle_threshold=rot.ANGLE_THRESHOLD, num_grid_enum_rotate=6)
    angles = rot.generate_rotate_angle(a, torch.device("cpu"))
    rot.save_rotate_info(angles, exp + "cloud0/rotate_all/")
    sd = synthetic.make_state_dict("gcnn")
    model = final_util.build_model(a, sd)
    res = gp.check_adv_success(a, rot.rotate_xyz, samples=[(data, torch.tensor([LBL]), "cloud0")], model=model)
    folder = exp + "cloud0/interaction_seed1/rotate_adv/"
    pose_idx = int(np.load(folder + "pose_idx.npy"))
    params = np.load(folder + "transform_params.npy")
    assert res[0][1] == pose_idx and params.shape == (3,) and np.array_equal(params, angles[pose_idx].numpy())
    # oracle rewards of all poses on the CPU
    clouds = torch.cat([rot.rotate_xyz(data, t) for t in angles], 0).permute(0, 2, 1).contiguous()
    with torch.no_grad():
        lg = nets.forward("gcnn", clouds, {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    z = lg.numpy().astype(np.float64)
    others = np.delete(z, LBL, axis=1)
    v = z[:, LBL] - (np.log(np.exp(others - others.max(1, keepdims=True)).sum(1)) + others.max(1))
    assert v[pose_idx] <= v.min() + 1e-3 * np.abs(v).max()
    assert res[0][0] == int((z.argmax(1) != LBL).sum()) or abs(res[0][0] - int((z.argmax(1) != LBL).sum())) <= 2
    np.save(folder + "transform_params.npy", params)
    pred = gp.gen_pred_label(model, data, torch.tensor([LBL]), rot.rotate_xyz, folder, a)
    saved = np.load(folder + "pred_labels.npy")
    assert saved.tolist() == [LBL, pred] and os.path.exists(folder + "pred_labels.txt")


@pytest.mark.gpu
def test_interaction_files_of_one_pose(tmp_path):
    """gen_pair_random -> gen_context -> save_logits_all_orders -> cal_interaction_all_orders: the reference's
    on-disk chain (interaction_seed<s>/region_pair_list.npy, ratio%d_context_list.npy, <pose>/ratio%d_all_logits.pt,
    <pose>/ratio%d_gt_interaction.npy) with the values of the direct calls."""
    from interpret_quality_b200 import final_cal_interactions as fci
    from interpret_quality_b200 import final_point_binary_interaction_logits as fpb
    dev = torch.device("cuda:0")
    geo = np.load(os.path.join(os.path.dirname(__file__), "golden", "geometry.npz"))
    data = torch.from_numpy(synthetic.make_cloud(1024))
    rid = geo["region_id_1024"]
    inter = str(tmp_path) + "/interaction_seed1/"
    os.makedirs(inter + "normal/")
    a = types.SimpleNamespace(model="gcnn", k=20, dataset="shapenet", device=dev, num_points=1024, num_regions=R,
                              softmax_type="modified", num_pairs_random=3, ratio=[0.0, 0.1, 1.0], num_save_context_max=10,
                              output_type="gt")
    final_util.set_random(1)
    pairs = gp.gen_pair_random(a)
    np.save(inter + "region_pair_list.npy", pairs)
    gp.gen_context(pairs, inter, a)
    model = final_util.build_model(a, synthetic.make_state_dict("gcnn"))
    fpb.save_logits_all_orders(model, data, rid, inter + "normal/", a)
    fci.cal_interaction_all_orders(torch.tensor([LBL]), inter + "normal/", a)
    for pct, ctx in ((0, 1), (10, 10), (100, 1)):
        lg = torch.load(inter + "normal/ratio%d_all_logits.pt" % pct)
        it = np.load(inter + "normal/ratio%d_gt_interaction.npy" % pct)
        assert tuple(lg.shape) == (3, 4 * ctx, 10) and lg.dtype == torch.float32
        assert it.shape == (3, ctx) and it.dtype == np.float64
        direct = fpb.compute_order_interaction_logits(model, data, rid, pairs, np.load(inter + "ratio%d_context_list.npy" % pct), a)
        assert torch.equal(direct.cpu(), lg.cpu())
        assert np.array_equal(fci.compute_order_interaction(direct, torch.tensor([LBL]), a), it)
    # m = R-2: the context is everything but the pair, so S+{i,j} is the full cloud for every pair
    full = torch.load(inter + "normal/ratio100_all_logits.pt")[:, 0]
    assert torch.allclose(full[0], full[1]) and torch.allclose(full[0], full[2])
