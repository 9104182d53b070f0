"""Generates tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py [names...]

The reference ships no tests or golden vectors, so these
fixtures are outputs of the reference's own functions on the synthetic inputs of
interpret_quality_b200/synthetic.py:

  geometry.npz      farthest_point_sample, cal_region_id, generate_all_orders,
                    mask_data_batch, the interaction mask block, square_distance,
                    query_ball_point, in-model FPS on a masked cloud
  <model>.npz       logits of masked clouds, shap_sampling_all_regions_batch,
                    compute_order_interaction_logits, compute_order_interaction,
                    get_reward (both softmax types), cal_norm_factor

The fixtures pin oracle/ (tests/test_oracle_golden.py, CPU) and the CUDA path
(tests/test_gpu_*.py).
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("INTERPRET_QUALITY_REF", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
os.chdir(REF)

from interpret_quality_b200 import synthetic  # noqa: E402

import final_save_fps as ref_fps  # noqa: E402
import final_shapley_value as ref_sv  # noqa: E402
import final_point_binary_interaction_logits as ref_il  # noqa: E402
import final_cal_interactions as ref_ci  # noqa: E402
from tools import final_common as ref_common  # noqa: E402
from tools import final_util as ref_util  # noqa: E402
from models import pointnet2 as ref_pn2  # noqa: E402
from models import pointconv as ref_pc  # noqa: E402
from models import dgcnn as ref_dg  # noqa: E402

R = 32
LBL = 3


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def base_inputs(N):
    data = torch.from_numpy(synthetic.make_cloud(N))
    fps_idx = ref_fps.farthest_point_sample(data, R)[0].numpy()
    region_id = ref_sv.cal_region_id(data, fps_idx, None, save=False)
    return data, fps_idx, region_id


def geometry():
    out = {}
    for N in (1024, 2048):
        data, fps_idx, region_id = base_inputs(N)
        out["fps_idx_%d" % N] = fps_idx
        out["region_id_%d" % N] = region_id
    data, fps_idx, region_id = base_inputs(1024)
    # seed replay of the permutations (final_shapley_value.py:59-72 after set_random(1))
    ref_util.set_random(1)
    a = types.SimpleNamespace(num_samples_save=1000, num_regions=R)
    orders = ref_sv.generate_all_orders(None, a, save=False)
    out["orders_sha1"] = np.array(sha(orders.astype(np.int64)))
    out["orders_head"] = orders[:8]
    # pairs (final_gen_pair.py:288-300 after set_random(1))
    import final_gen_pair as ref_gp
    ref_util.set_random(1)
    out["pairs_head"] = ref_gp.gen_pair_random(types.SimpleNamespace(num_regions=R, num_pairs_random=300))[:8]
    # Shapley masks
    center = torch.mean(data, dim=1).squeeze()
    out["center"] = center.numpy()
    bs = 3
    a = types.SimpleNamespace(num_regions=R)
    md = data.expand((R + 1) * bs, 1024, 3).clone()
    md = ref_common.mask_data_batch(md, center, orders[:bs], region_id, a)
    out["mask_shapley_sha1"] = np.array(sha(md.numpy()))
    out["mask_shapley_rows"] = md.numpy()[[0, 1, 17, 32, 33, 40, 98]]
    out["mask_shapley_row_ids"] = np.array([0, 1, 17, 32, 33, 40, 98])
    # single-permutation form (final_shapley_value.py:74-88)
    md1 = data.expand(R + 1, 1024, 3).clone()
    md1 = ref_sv.mask_data(md1, center, orders[0], region_id)
    assert torch.equal(md1, md[:R + 1])
    # interaction mask block (final_point_binary_interaction_logits.py:42-56), captured through a probe model
    pairs, contexts = synthetic.make_pairs_and_contexts(2, R, orders_m=(0, 3, 30), max_contexts=4)
    for m in (0, 3, 30):
        seen = []

        class Probe(torch.nn.Module):
            def forward(self, x):
                seen.append(x.clone())
                return torch.zeros(x.shape[0], 10)

        a = types.SimpleNamespace(interaction_batch_size=3, model="dgcnn")
        ref_il.compute_order_interaction_logits(Probe(), data, region_id, pairs, contexts[m], a)
        blk = torch.cat(seen, 0).numpy()
        out["mask_inter_m%d_sha1" % m] = np.array(sha(blk))
        out["mask_inter_m%d_shape" % m] = np.array(blk.shape)
        if m == 3:
            out["mask_inter_m3_rows"] = blk[:8]
    out["inter_pairs"] = pairs
    for m in (0, 3, 30):
        out["inter_ctx_m%d" % m] = contexts[m].astype(np.int64)
    # masked cloud -> in-model FPS, square_distance, ball query, knn
    sparse = md[34:36]                         # permutation 1 rows 1,2: one / two regions kept
    dense = md[[30, 66]]
    clouds = torch.cat([sparse, dense, data], 0).contiguous()
    out["geo_cloud_rows"] = np.array([34, 35, 30, 66, -1])
    f1 = ref_pn2.farthest_point_sample(clouds, 512)
    out["fps512"] = f1.numpy().astype(np.int16)
    new_xyz = ref_pn2.index_points(clouds, f1)
    f2 = ref_pc.farthest_point_sample(new_xyz, 128)
    out["fps128"] = f2.numpy().astype(np.int16)
    sq = ref_pn2.square_distance(new_xyz[:, :64], clouds)
    out["sqdist_sha1"] = np.array(sha(sq.numpy()))
    out["sqdist_head"] = sq.numpy()[:, :4, :16]
    for radius, K in ((0.1, 16), (0.2, 32), (0.4, 128)):
        gi = ref_pn2.query_ball_point(radius, K, clouds, new_xyz)
        out["ball_r%d" % int(radius * 10)] = gi.numpy().astype(np.int16)
    # DGCNN layer-1 kNN sets on xyz (models/dgcnn.py:12-18): sorted member sets only for tie-free rows
    x = clouds[4:5].permute(0, 2, 1).contiguous()
    out["knn_xyz_unmasked"] = np.sort(ref_dg.knn(x, 20)[0].numpy(), axis=1).astype(np.int16)
    np.savez_compressed(os.path.join(HERE, "geometry.npz"), **out)
    print("geometry.npz written")


def load_ref_model(name):
    args = types.SimpleNamespace(model=name, k=20, dataset="shapenet", feature_transform=True,
                                 device=torch.device("cpu"))
    cls = {"pointnet2": ref_util.PointNet2ClsMsg, "pointnet": ref_util.PointNetCls, "dgcnn": ref_util.DGCNN_cls,
           "gcnn": ref_util.GCNN_cls, "pointconv": ref_util.PointConvDensityClsSsg}[name]
    m = cls(args)
    sd = synthetic.make_state_dict(name)
    assert list(m.state_dict().keys()) == [k for k, _, _ in synthetic.state_dict_spec(name)]
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return m.eval(), args


def model_golden(name, N=1024, tag=None):
    torch.manual_seed(0)
    model, margs = load_ref_model(name)
    data, fps_idx, region_id = base_inputs(N)
    orders = synthetic.make_orders(16, R)
    lbl = torch.tensor([LBL])
    out = {"N": np.array(N)}
    n_perm, bs = (4, 2) if N == 1024 else (1, 1)
    args = types.SimpleNamespace(num_points=N, num_regions=R, shapley_batch_size=bs, num_samples=n_perm,
                                 softmax_type="modified", model=name, device=torch.device("cpu"))
    with torch.no_grad():
        phi, logits = ref_common.shap_sampling_all_regions_batch(model, data, lbl, region_id, orders, args)
    out["shapley_phi"] = phi
    out["shapley_logits"] = logits.numpy()
    out["shapley_nperm"] = np.array(n_perm)
    out["reward_modified"] = ref_common.get_reward(logits, lbl, args).numpy()
    args_n = types.SimpleNamespace(softmax_type="normal")
    out["reward_normal"] = ref_common.get_reward(logits, lbl, args_n).numpy()
    # efficiency identity needs v(N) - v(empty)
    center = torch.mean(data, dim=1).squeeze()
    with torch.no_grad():
        out["norm_factor"] = np.array(ref_sv.cal_norm_factor(model, data, lbl, center, None, args, save=False))
    spread = float(np.abs(logits.numpy() - logits.numpy()[0:1]).max())
    assert spread > 1e-3 * float(np.abs(logits.numpy()).max()), "degenerate network: logits do not depend on the mask"
    if N == 1024:
        pairs, contexts = synthetic.make_pairs_and_contexts(2, R, orders_m=(0, 3, 30), max_contexts=4)
        iargs = types.SimpleNamespace(interaction_batch_size=3, model=name, softmax_type="modified")
        for m in (0, 3, 30):
            with torch.no_grad():
                il = ref_il.compute_order_interaction_logits(model, data, region_id, pairs, contexts[m], iargs)
            out["inter_logits_m%d" % m] = il.numpy()
            out["inter_m%d" % m] = ref_ci.compute_order_interaction(il, lbl, iargs)
        if name == "pointnet":
            x = data.permute(0, 2, 1).contiguous()
            with torch.no_grad():
                lg, tf, crt = model(x)
            out["pointnet_trans_feat"] = tf.numpy()
            out["pointnet_crt"] = crt.numpy().astype(np.int16)
    fn = os.path.join(HERE, "%s.npz" % (tag or name))
    np.savez_compressed(fn, **out)
    print(fn, "written; logits spread", spread, "phi", phi[:4], "norm", out["norm_factor"])


if __name__ == "__main__":
    names = sys.argv[1:] or ["geometry", "pointnet", "dgcnn", "gcnn", "pointnet2", "pointconv", "dgcnn_2048"]
    for n in names:
        if n == "geometry":
            geometry()
        elif n == "dgcnn_2048":
            model_golden("dgcnn", 2048, "dgcnn_2048")
        elif n in ("dgcnn_more", "poses", "gen_pair", "shap_run", "result_tables", "smoothness", "interaction_pipeline",
                   "dgcnn_headline", "dgcnn_2048_4", "pointnet2_100", "c4_dgcnn", "c4_gcnn", "interactions_f64"):
            pass                                   # handled at the bottom of the file
        else:
            model_golden(n)


def dgcnn_more():
    """A wider DGCNN sample (12 more permutations = 396 masked clouds, logits only): the dynamic kNN makes
    DGCNN the model whose parity is decided by near-tie neighbour choices, so it gets more coverage."""
    model, margs = load_ref_model("dgcnn")
    data, fps_idx, region_id = base_inputs(1024)
    orders = synthetic.make_orders(16, R)[4:16]
    args = types.SimpleNamespace(num_points=1024, num_regions=R, shapley_batch_size=4, num_samples=12,
                                 softmax_type="modified", model="dgcnn", device=torch.device("cpu"))
    with torch.no_grad():
        phi, logits = ref_common.shap_sampling_all_regions_batch(model, data, torch.tensor([LBL]), region_id, orders, args)
    np.savez_compressed(os.path.join(HERE, "dgcnn_more.npz"), shapley_phi=phi, shapley_logits=logits.numpy())
    print("dgcnn_more.npz written", logits.shape)


if __name__ == "__main__" and "dgcnn_more" in sys.argv[1:]:
    dgcnn_more()


def poses():
    """poses.npz: the pose grids and one disturbed cloud per mode from the reference's final_{trans,rotate,scale}_
    center_enum_all.py (generate_trans_vector / generate_rotate_angle / generate_scale, translate_pc / rotate_xyz /
    scale_pc), and the region Shapley values of three scale poses through the reference's own sampler (PointNet)."""
    import final_trans_center_enum_all as ref_t
    import final_rotate_center_enum_all as ref_r
    import final_scale_center_enum_all as ref_s
    a = types.SimpleNamespace(trans_dist_threshold=0.5, num_grid_enum_trans=6, angle_threshold=np.pi / 4,
                              num_grid_enum_rotate=6, scale_lower=0.5, scale_upper=2.0, num_grid_enum_scale=30)
    cpu = torch.device("cpu")
    tv, ra, sc = ref_t.generate_trans_vector(a, cpu), ref_r.generate_rotate_angle(a, cpu), ref_s.generate_scale(a, cpu)
    data, fps_idx, region_id = base_inputs(1024)
    model, margs = load_ref_model("pointnet")
    orders = synthetic.make_orders(8, R)
    args = types.SimpleNamespace(num_points=1024, num_regions=R, shapley_batch_size=2, num_samples=4,
                                 softmax_type="modified", model="pointnet", device=cpu)
    a3 = types.SimpleNamespace(scale_lower=0.5, scale_upper=2.0, num_grid_enum_scale=3)
    phis = []
    with torch.no_grad():
        for s in ref_s.generate_scale(a3, cpu):
            phi, _ = ref_common.shap_sampling_all_regions_batch(model, ref_s.scale_pc(data, s), torch.tensor([LBL]),
                                                                region_id, orders, args)
            phis.append(phi)
    np.savez_compressed(os.path.join(HERE, "poses.npz"), trans_vector=tv.numpy(), rotate_angle=ra.numpy(), scale=sc.numpy(),
                        translated_7=ref_t.translate_pc(data, tv[7]).numpy(), rotated_11=ref_r.rotate_xyz(data, ra[11]).numpy(),
                        scaled_3=ref_s.scale_pc(data, sc[3]).numpy(), scale3_pointnet_phi=np.stack(phis))
    print("poses.npz written")


if __name__ == "__main__" and "poses" in sys.argv[1:]:
    poses()


def gen_pair():
    """gen_pair.npz: gen_pair_random + gen_context of the reference's final_gen_pair.py after set_random(1)."""
    import tempfile
    import final_gen_pair as ref_gp
    from tools.final_util import set_random as ref_set_random
    a = types.SimpleNamespace(num_regions=R, num_pairs_random=6, ratio=[0.0, 0.04, 0.1, 0.5, 0.94, 1.0], num_save_context_max=100)
    ref_set_random(1)
    pairs = ref_gp.gen_pair_random(a)
    d = tempfile.mkdtemp() + "/"
    ref_gp.gen_context(pairs, d, a)
    out = {"pairs": pairs}
    for r in a.ratio:
        out["ctx%d" % int(r * 100)] = np.load(d + "ratio%d_context_list.npy" % int(r * 100))
    np.savez_compressed(os.path.join(HERE, "gen_pair.npz"), **out)
    print("gen_pair.npz written", {k: v.shape for k, v in out.items()})


if __name__ == "__main__" and "gen_pair" in sys.argv[1:]:
    gen_pair()


def shap_run():
    """shap_run.npz: the reference's final_shapley_value.shap_sampling (PointNet, CPU, 100 saved permutations) on
    the synthetic cloud: region_sv_all, the count-100 checkpoints and norm_factor."""
    import tempfile
    from tools.final_util import set_random as ref_set_random
    data, fps_idx, region_id = base_inputs(1024)
    model, margs = load_ref_model("pointnet")
    d = tempfile.mkdtemp() + "/"
    os.chdir(d)
    np.save("fps_shapenet_1024_32_index_final30.npy", np.asarray(fps_idx).reshape(1, -1))
    a = types.SimpleNamespace(num_points=1024, num_regions=R, num_samples_save=100, softmax_type="modified", model="pointnet",
                              dataset="shapenet", device=torch.device("cpu"), exp_folder=d + "exp/")
    ref_set_random(1)
    ref_sv.shap_sampling(model, [(data, torch.tensor([LBL]))], a, ["cloud0"])
    out = d + "exp/cloud0/"
    np.savez_compressed(os.path.join(HERE, "shap_run.npz"), region_sv_all=np.load(out + "region_sv_all.npy"),
                        shapley_100=np.load(out + "shapley/0_100.npy"), region_shapley_100=np.load(out + "region_shapley/0_100.npy"),
                        norm_factor=np.load(out + "norm_factor.npy"), all_orders=np.load(out + "all_orders.npy"),
                        region_id=np.load(out + "region_id.npy"))
    os.chdir(REF)
    print("shap_run.npz written")


if __name__ == "__main__" and "shap_run" in sys.argv[1:]:
    shap_run()


def result_tables():
    """result_tables.npz: the reference's final_result.py table code (cal_sensitivity :83-102, cal_correlation_coef
    :124-140, cal_shapley_smoothness_metric_single_pc :144-176) on seeded (poses, R) value arrays laid out in the
    reference's folders.  matplotlib is not installed here and final_result.py only needs it for its plots, so the
    import is satisfied by empty stand-in modules; no plotting function is called."""
    import tempfile
    from unittest import mock
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d", "matplotlib.ticker",
                 "matplotlib.patches", "matplotlib.colors", "matplotlib.cm"):
        sys.modules.setdefault(name, mock.MagicMock())
    import final_result as ref_res
    ref_res.num_regions, ref_res.num_points = R, 1024
    rng = np.random.RandomState(7)
    names = ["cloud%d" % i for i in range(4)]
    d = tempfile.mkdtemp() + "/"
    out = {}
    data, fps_idx, region_id = base_inputs(1024)
    for ci, name in enumerate(names):
        base = d + name + "/"
        # values with a per-region scale so that sensitivity and mean intensity correlate like real runs
        scale = rng.gamma(2.0, 0.05, size=R)
        for mode, n_pose in (("scale", 30), ("rotate", 24)):
            os.makedirs(base + "%s_all/" % mode)
            v = rng.randn(n_pose, R) * scale + rng.randn(R) * 0.02
            np.save(base + "%s_all/region_shapley_value.npy" % mode, v)
            out["%s_%s_values" % (name, mode)] = v
        for direction, n_pose in (("inc", 5), ("dec", 7)):
            os.makedirs(base + "linearity_all/allregion_%s/" % direction)
            v = rng.randn(n_pose, R) * scale
            np.save(base + "linearity_all/allregion_%s/region_shapley_value.npy" % direction, v)
            out["%s_linearity_%s_values" % (name, direction)] = v
        np.save(base + "region_id.npy", region_id)
        for mode in ("scale", "rotate", "linearity"):
            out["%s_%s_sensitivity" % (name, mode)] = ref_res.cal_sensitivity(base, mode)
        m, m_poses, den = ref_res.cal_shapley_smoothness_metric_single_pc(
            data[0].numpy(), out["%s_rotate_values" % name], region_id)
        out["%s_smooth" % name] = np.array([m, den])
        out["%s_smooth_poses" % name] = m_poses
    # Table 3 through the reference's own function: it reads a global namespace and the sample-name list
    ref_res.args = types.SimpleNamespace(dataset="shapenet")
    ref_res.num_pc = len(names)
    ref_res.get_exp_folder_name = lambda model_name, dataset: d
    ref_res.get_folder_name_list = lambda a: names
    for mode in ("scale", "rotate"):
        out["pearson_mean_%s" % mode] = np.array(ref_res.cal_correlation_coef("pointnet", mode))
        out["sens_all_%s" % mode] = ref_res.cal_sensitivity_all_pc("pointnet", mode)
        out["intensity_all_%s" % mode] = ref_res.cal_mean_sv_intensity("pointnet", mode)
    out["region_id"] = region_id
    np.savez_compressed(os.path.join(HERE, "result_tables.npz"), **out)
    print("result_tables.npz written", len(out), "arrays")


if __name__ == "__main__" and "result_tables" in sys.argv[1:]:
    result_tables()


def smoothness():
    """smoothness.npz: the reference's final_smoothness_center_enum_all.test_all_region (:281-350) on the synthetic cloud
    (PointNet, CPU, 3 epochs, 4 permutations) for the three modes and both objectives: the cloud after every epoch,
    the per-region smoothness, the Shapley values, and the per-region original info of get_original_region_info
    (:245-268).  torch.symeig (:41) no longer exists in this torch; it is provided here as a thin forwarder to
    torch.linalg.eigh (same ascending eigenvalue order, eigenvectors in columns), the reference file is unmodified."""
    import tempfile
    torch.symeig = lambda A, eigenvectors=True: torch.linalg.eigh(A)
    import final_smoothness_center_enum_all as ref_sm
    import time
    data, fps_idx, region_id = base_inputs(1024)
    model, margs = load_ref_model("pointnet")
    orders = synthetic.make_orders(8, R)
    out = {}
    d = tempfile.mkdtemp() + "/"
    for mode in ("linearity", "planarity", "scattering"):
        a = types.SimpleNamespace(num_points=1024, num_regions=R, shapley_batch_size=2, num_samples=4,
                                  softmax_type="modified", model="pointnet", device=torch.device("cpu"), mode=mode,
                                  step=ref_sm.STEP, enum_step=ref_sm.ENUM_STEP, epoch=3, var_threshold=ref_sm.VAR_THRESHOLD,
                                  dist_threshold=ref_sm.DIST_THRESHOLD, stop_ratio=ref_sm.STOP_RATIO,
                                  max_iteration=ref_sm.MAX_ITERATION)
        for objective in ("inc", "dec"):
            t0 = time.time()
            # data_list (:325) holds `data_copy.cpu().numpy()`: on a CPU run that aliases data_copy, so every saved
            # epoch shows the final cloud (on CUDA .cpu() copies).  The per-epoch clouds are therefore recorded from
            # the argument of the sampler call that follows each epoch (:328), through a recording wrapper.
            seen = []

            def recording_sampler(model_, data_, *rest, _inner=ref_common.shap_sampling_all_regions_batch):
                seen.append(data_.detach().clone().numpy())
                return _inner(model_, data_, *rest)

            ref_sm.shap_sampling_all_regions_batch = recording_sampler
            ref_sm.test_all_region(model, data, torch.tensor([LBL]), orders, region_id, d + mode + "_all/", a, objective)
            p = d + mode + "_all/allregion_%s/" % objective
            key = "%s_%s_" % (mode, objective)
            out[key + "data"] = np.stack(seen[1:])                       # seen[0] is the undisturbed cloud (:299)
            assert np.array_equal(out[key + "data"][-1], np.load(p + "data_smoothness.npy")[-1])
            out[key + "smoothness"] = np.load(p + "%s.npy" % mode)
            out[key + "phi"] = np.load(p + "region_shapley_value.npy")
            out[key + "orig_phi"] = np.load(p + "orig_shapley_value.npy")
            print(mode, objective, "epochs", out[key + "data"].shape[0], "%.1fs" % (time.time() - t0))
    class _Quiet:
        def cprint(self, text):
            pass
    orient, bounds, smooth0 = [], [], []
    a.mode = "linearity"
    for i in range(R):
        _, s0, o, b = ref_sm.get_original_region_info(data, region_id, i, _Quiet(), a)
        orient.append(torch.stack(o).numpy())
        bounds.append(np.array([float(x) for x in b], dtype=np.float32))
        smooth0.append(s0)
    out["orientations"] = np.stack(orient)           # (R,3,3): rows o1,o2,o3
    out["bounds"] = np.stack(bounds)                 # (R,6): ub1..3, lb1..3
    out["linearity_orig"] = np.array(smooth0)
    np.savez_compressed(os.path.join(HERE, "smoothness.npz"), **out)
    print("smoothness.npz written")


if __name__ == "__main__" and "smoothness" in sys.argv[1:]:
    smoothness()


def interaction_pipeline():
    """interaction_pipeline.npz: the reference's whole interaction chain on the synthetic cloud (PointNet, CPU, 12 rotation
    poses): final_gen_pair.py save_pair_random :302 -> check_adv_success :221 -> save_pair_single_region :145 ->
    save_context :45 -> save_pred_label :90, then final_point_binary_interaction_logits.save_logits :83 and
    final_cal_interactions.cal_interaction :49 (gt and pred).  Every file the chain writes is stored under its path
    relative to the experiment folder.  The dataset loaders (absent data) are replaced by a one-cloud list and
    load_model by the synthetic-weight model; the functions themselves are the reference's, unmodified."""
    import tempfile
    import final_gen_pair as ref_gp
    import final_rotate_center_enum_all as ref_r
    from tools.final_util import set_random as ref_set_random
    data, fps_idx, region_id = base_inputs(1024)
    lbl = torch.tensor([LBL])
    model, margs = load_ref_model("pointnet")
    d = tempfile.mkdtemp() + "/"
    base = d + "cloud0/"
    os.makedirs(base + "rotate_all/")
    np.save(base + "region_id.npy", region_id)
    ga = types.SimpleNamespace(angle_threshold=np.pi / 4, num_grid_enum_rotate=6)
    angles = ref_r.generate_rotate_angle(ga, torch.device("cpu")).numpy()[::18]          # 12 of the 216 poses
    np.save(base + "rotate_all/angle_tuple.npy", angles)
    np.save(base + "rotate_all/region_shapley_value.npy", np.random.RandomState(11).randn(angles.shape[0], R) * 0.1)
    a = types.SimpleNamespace(num_points=1024, num_regions=R, model="pointnet", dataset="shapenet", mode="rotate", seed=1,
                              gen_pair_seed=1, exp_folder=d, device=torch.device("cpu"), test_batch_size=1,
                              ratio=[0.0, 0.1, 1.0], num_pairs_random=5, num_save_context_max=3, softmax_type="modified",
                              interaction_batch_size=2, output_type="gt")
    for mod in (ref_gp, ref_il, ref_ci):
        mod.DataLoader = lambda *args_, **kw: [(data, lbl)]
        mod.ShapeNetDataset_Shapley_test = lambda *args_, **kw: None
        mod.load_model = lambda args_: model
        mod.folder_name_list = ["cloud0"]
        mod.selected_sample_idx = [0]
    ref_set_random(a.seed)
    ref_gp.save_pair_random(a)
    ref_gp.check_adv_success(a, disturb_fn=ref_r.rotate_xyz)
    ref_gp.save_pair_single_region(a)
    ref_gp.save_context(a)
    ref_gp.save_pred_label(a, disturb_fn=ref_r.rotate_xyz)
    ref_il.save_logits(a, ref_r.rotate_xyz)
    ref_ci.cal_interaction(a)
    a.output_type = "pred"
    ref_ci.cal_interaction(a)
    out = {}
    for root, _, files in os.walk(d):
        for f in files:
            rel = os.path.relpath(os.path.join(root, f), d)
            if f.endswith(".npy"):
                out[rel] = np.load(os.path.join(root, f))
            elif f.endswith(".pt"):
                out[rel] = torch.load(os.path.join(root, f)).numpy()
    np.savez_compressed(os.path.join(HERE, "interaction_pipeline.npz"), **out)
    print("interaction_pipeline.npz written:", len(out), "files")


if __name__ == "__main__" and "interaction_pipeline" in sys.argv[1:]:
    interaction_pipeline()


def wide_shapley(name, N, n_perm, bs, tag):
    """<tag>.npz: shap_sampling_all_regions_batch of the unmodified reference over `n_perm` seed-replayed
    permutations (logits of every masked cloud + phi): the headline call of BASELINE.json for DGCNN (100 x 33 =
    3300 clouds), DGCNN at N = 2048 and PointNet++ at 100 permutations ("widen parity")."""
    import time
    model, margs = load_ref_model(name)
    data, fps_idx, region_id = base_inputs(N)
    orders = synthetic.make_orders(1000, R)[:n_perm]
    args = types.SimpleNamespace(num_points=N, num_regions=R, shapley_batch_size=bs, num_samples=n_perm,
                                 softmax_type="modified", model=name, device=torch.device("cpu"))
    t0 = time.time()
    with torch.no_grad():
        phi, logits = ref_common.shap_sampling_all_regions_batch(model, data, torch.tensor([LBL]), region_id, orders, args)
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), shapley_phi=phi, shapley_logits=logits.numpy(),
                        shapley_nperm=np.array(n_perm), N=np.array(N))
    print(tag + ".npz written", logits.shape, "%.0fs of reference CPU time" % (time.time() - t0))


def c4_interactions(name, P=8, max_contexts=12):
    """c4_<model>.npz: BASELINE config C4 on the unmodified reference: P pairs x all 13 orders m in {0,1,2,3,6,...,27,30} x
    up to `max_contexts` contexts: compute_order_interaction_logits + compute_order_interaction per order."""
    import time
    model, margs = load_ref_model(name)
    data, fps_idx, region_id = base_inputs(1024)
    pairs, contexts = synthetic.make_pairs_and_contexts(P, R, max_contexts=max_contexts)
    iargs = types.SimpleNamespace(interaction_batch_size=25, model=name, softmax_type="modified")
    out = {"pairs": pairs, "orders_m": np.array(sorted(contexts.keys()))}
    t0 = time.time()
    for m in sorted(contexts.keys()):
        with torch.no_grad():
            il = ref_il.compute_order_interaction_logits(model, data, region_id, pairs, contexts[m], iargs)
        out["ctx_m%d" % m] = contexts[m].astype(np.int64)
        out["logits_m%d" % m] = il.numpy()
        out["inter_m%d" % m] = ref_ci.compute_order_interaction(il, torch.tensor([LBL]), iargs)
    np.savez_compressed(os.path.join(HERE, "c4_%s.npz" % name), **out)
    print("c4_%s.npz written, %.0fs of reference CPU time" % (name, time.time() - t0))


if __name__ == "__main__":
    for n in sys.argv[1:]:
        if n == "dgcnn_headline":
            wide_shapley("dgcnn", 1024, 100, 5, "dgcnn_headline")
        elif n == "dgcnn_2048_4":
            wide_shapley("dgcnn", 2048, 4, 1, "dgcnn_2048_4")
        elif n == "pointnet2_100":
            wide_shapley("pointnet2", 1024, 100, 5, "pointnet2_100")
        elif n == "c4_dgcnn":
            c4_interactions("dgcnn")
        elif n == "c4_gcnn":
            c4_interactions("gcnn")


def interactions_f64():
    """interactions_f64.npz: the float64 yardstick of every interaction golden -- the SAME masked clouds (fp32 masks,
    fp32 centre) through a float64 evaluation of the network (oracle/nets.py, which tests/test_oracle_golden.py pins to
    the reference).  |reference fp32 - float64| is the reference's own rounding noise on an interaction; the GPU tests
    hold ours to max(1e-3 * max|I|, 2 x that noise) per order and print both.  Keys: <model>_m<m> for the small sets of
    <model>.npz, c4_<model>_m<m> for c4_<model>.npz."""
    import time
    from oracle import coalition as oc
    from oracle import nets
    data, fps_idx, region_id = base_inputs(1024)
    out = {}

    def f64_interactions(name, pairs, ctx):
        sd = synthetic.make_state_dict(name)
        sd64 = {k: torch.from_numpy(v).double() if v.dtype == np.float32 else torch.from_numpy(v) for k, v in sd.items()}
        torch.set_default_dtype(torch.float64)
        try:
            real = nets.forward
            nets.forward = lambda model, x, sd_, k=20: real(model, x.double(), sd64, k)
            lg = oc.interaction_logits(name, sd, data.numpy(), region_id, pairs, ctx, R, 25)
        finally:
            nets.forward = real
            torch.set_default_dtype(torch.float32)
        return oc.interaction_from_logits(lg, LBL)

    t0 = time.time()
    pairs, contexts = synthetic.make_pairs_and_contexts(2, R, orders_m=(0, 3, 30), max_contexts=4)
    for name in ("pointnet", "dgcnn", "gcnn", "pointnet2", "pointconv"):
        for m in (0, 3, 30):
            out["%s_m%d" % (name, m)] = f64_interactions(name, pairs, contexts[m])
        print(name, "small set done, %.0fs" % (time.time() - t0), flush=True)
    for name in ("gcnn", "dgcnn"):
        g = np.load(os.path.join(HERE, "c4_%s.npz" % name))
        for m in g["orders_m"]:
            out["c4_%s_m%d" % (name, m)] = f64_interactions(name, g["pairs"], g["ctx_m%d" % m])
            print("c4", name, "m", m, "%.0fs" % (time.time() - t0), flush=True)
    np.savez_compressed(os.path.join(HERE, "interactions_f64.npz"), **out)
    print("interactions_f64.npz written")


if __name__ == "__main__" and "interactions_f64" in sys.argv[1:]:
    interactions_f64()
