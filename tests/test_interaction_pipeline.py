"""The interaction chain end to end: final_gen_pair's folder walkers ->
final_point_binary_interaction_logits.save_logits -> final_cal_interactions.cal_interaction, against every file the
unmodified reference wrote for the same cloud, seed and poses (tests/golden/interaction_pipeline.npz, keyed by path).

CPU: everything that is sampling / bookkeeping (pairs, contexts, pose choices, folder names) must be bit-identical.
GPU: the adversarial pose, the predicted labels, the logits (1e-3 of their scale) and the interactions."""
import os
import types

import numpy as np
import pytest
import torch

from interpret_quality_b200 import final_gen_pair as gp
from interpret_quality_b200 import final_rotate_center_enum_all as rot
from interpret_quality_b200 import synthetic
from interpret_quality_b200.tools import final_util
from _gates import interaction_gate
from oracle import coalition

R, LBL = 32, 3
SEED_DIR = "cloud0/interaction_seed1/"


def pipeline_args(exp, **kw):
    a = types.SimpleNamespace(num_points=1024, num_regions=R, model="pointnet", dataset="shapenet", mode="rotate", seed=1,
                              gen_pair_seed=1, exp_folder=exp, ratio=[0.0, 0.1, 1.0], num_pairs_random=5, k=20,
                              num_save_context_max=3, softmax_type="modified", interaction_batch_size=2, output_type="gt",
                              feature_transform=True)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def lay_out_inputs(g, exp):
    os.makedirs(exp + "cloud0/rotate_all/")
    np.save(exp + "cloud0/region_id.npy", g["cloud0/region_id.npy"])
    np.save(exp + "cloud0/rotate_all/angle_tuple.npy", g["cloud0/rotate_all/angle_tuple.npy"])
    np.save(exp + "cloud0/rotate_all/region_shapley_value.npy", g["cloud0/rotate_all/region_shapley_value.npy"])


def written(exp, suffixes):
    out = {}
    for root, _, files in os.walk(exp):
        for f in files:
            if f.endswith(suffixes):
                out[os.path.relpath(os.path.join(root, f), exp)] = os.path.join(root, f)
    return out


def test_pairs_contexts_and_pose_choices_are_bit_identical(golden, tmp_path):
    g, exp = golden("interaction_pipeline"), str(tmp_path) + "/"
    lay_out_inputs(g, exp)
    a = pipeline_args(exp)
    data = torch.from_numpy(synthetic.make_cloud(1024))
    samples = [(data, torch.tensor([LBL]), "cloud0")]
    final_util.set_random(a.seed)
    gp.save_pair_random(a, ["cloud0"])                 # check_adv_success sits here in the reference; it draws nothing
    last = gp.save_pair_single_region(a, samples)
    gp.save_context(a, ["cloud0"])
    ours = written(exp, (".npy",))
    bookkeeping = [k for k in g.files if k.endswith(("region_pair_list.npy", "context_list.npy", "pose_idx.npy",
                                                     "transform_params.npy")) and "rotate_adv/" not in k]
    assert len(bookkeeping) == 1 + 3 + 32 * (1 + 3 + 4)
    for k in bookkeeping:
        got = np.load(ours[k])
        assert got.shape == g[k].shape and got.dtype == g[k].dtype and np.array_equal(got, g[k]), k
    ref_regions = {k.split("/")[3] for k in g.files if "rotate_adv_single_region/" in k}
    assert set(os.listdir(exp + SEED_DIR + "rotate_adv_single_region/")) == ref_regions
    assert np.array_equal(last, g[[k for k in g.files if "region31/region_pair_list" in k][0]])


def test_cal_rank_and_neighbour_pairs():
    v = np.array([0.3, -1.0, 2.0, 0.3])
    assert np.array_equal(final_util.cal_rank(v), np.argsort(np.argsort(v)))
    nb = np.array([[True, True, False], [True, True, True], [False, True, True]])
    a = types.SimpleNamespace(num_regions=3)
    assert np.array_equal(gp.gen_pair_single_region(1, nb, a), [[1, 0], [1, 2]])
    assert gp.gen_pair_single_region(0, np.eye(3, dtype=bool), a).shape == (0,)


@pytest.mark.gpu
def test_whole_chain_matches_the_reference(golden, tmp_path):
    from interpret_quality_b200 import final_cal_interactions as ci
    from interpret_quality_b200 import final_point_binary_interaction_logits as il
    dev = torch.device("cuda:0")
    g, exp = golden("interaction_pipeline"), str(tmp_path) + "/"
    lay_out_inputs(g, exp)
    a = pipeline_args(exp, device=dev)
    model = final_util.build_model(a, synthetic.make_state_dict("pointnet"))
    samples = [(torch.from_numpy(synthetic.make_cloud(1024)), torch.tensor([LBL]), "cloud0")]
    final_util.set_random(a.seed)
    gp.save_pair_random(a, ["cloud0"])
    (num_miscls, pose_idx), = gp.check_adv_success(a, rot.rotate_xyz, samples=samples, model=model)
    gp.save_pair_single_region(a, samples)
    gp.save_context(a, ["cloud0"])
    gp.save_pred_label(a, rot.rotate_xyz, samples, model=model)
    il.save_logits(a, rot.rotate_xyz, samples, model=model, selected_sample_idx=[0])
    ci.cal_interaction(a, samples, selected_sample_idx=[0])
    a.output_type = "pred"
    ci.cal_interaction(a, samples)
    ours = written(exp, (".npy", ".pt"))
    assert set(ours) == set(g.files)                                  # the same tree, file for file
    assert pose_idx == int(g[SEED_DIR + "rotate_adv/pose_idx.npy"]) and num_miscls >= 1
    sd = synthetic.make_state_dict("pointnet")
    worst_ratio = 0.0
    for k in g.files:
        want = g[k]
        if k.endswith(".pt"):
            got = torch.load(ours[k])
            assert got.dtype == torch.float32 and tuple(got.shape) == want.shape, k
            assert np.abs(got.cpu().numpy() - want).max() <= 1e-3 * np.abs(want).max(), k
        elif k.endswith("interaction.npy"):
            got = np.load(ours[k])
            assert got.dtype == np.float64 and got.shape == want.shape, k
            # per file: 1e-3 of max|I|, or twice the reference's own fp32-vs-float64 noise on these clouds (_gates.py)
            folder, fname = k.rsplit("/", 1)
            ratio, out_type = fname.split("_")[0], fname.split("_")[1]
            pose = samples[0][0].numpy()
            lbl_k = LBL
            if folder.endswith("rotate_adv"):
                params = torch.from_numpy(g[folder + "/transform_params.npy"].astype(np.float32))
                pose = rot.rotate_xyz(samples[0][0], params).numpy()
                if out_type == "pred":
                    lbl_k = int(g[folder + "/pred_labels.npy"][1])
            up = folder.rsplit("/", 1)[0]
            yard = coalition.interaction_float64_yardstick("pointnet", sd, pose, g["cloud0/region_id.npy"],
                                                           g[up + "/region_pair_list.npy"],
                                                           g[up + "/%s_context_list.npy" % ratio], R, lbl_k)
            err, bound = interaction_gate(got, want, yard, k.split("interaction_seed1/")[1])
            worst_ratio = max(worst_ratio, err / bound)
            assert err <= bound, (k, err, bound)
        else:
            got = np.load(ours[k])
            assert got.shape == want.shape and np.array_equal(got, want), k       # labels, poses, pairs, contexts
    adv = SEED_DIR + "rotate_adv/"
    assert not np.array_equal(np.load(ours[adv + "ratio10_gt_interaction.npy"]), np.load(ours[adv + "ratio10_pred_interaction.npy"]))
