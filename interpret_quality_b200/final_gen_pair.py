"""Producers of the interaction inputs, with the reference's names (final_gen_pair.py of ada-shen/Interpret_quality):
gen_context :18-43, gen_pred_label :75-88, check_adv_success :221-286, gen_pair_random :288-300.

Sampling stays on the host in numpy's global legacy stream (call tools.final_util.set_random(seed) first, like the
reference's main): the same call sequence reproduces region_pair_list.npy / ratio%d_context_list.npy bit for bit.
The 216-pose forwards of check_adv_success / gen_pred_label run on the CUDA models of this package.  The reference
walks its dataset loaders; here the clouds come from `samples` = [(data (1,N,3), lbl (1,), folder_name), ...]."""
import itertools
from math import comb

import numpy as np
import torch

from .tools.final_common import _device_of, get_reward
from .tools.final_util import load_model, mkdir


def gen_pair_random(args):
    """(num_pairs_random, 2) int64 ndarray of region pairs (i, j), j > i, drawn without replacement."""
    all_pairs = np.array([[i, j] for i in range(args.num_regions) for j in range(args.num_regions) if j > i])
    pair_idx = np.random.choice(all_pairs.shape[0], size=args.num_pairs_random, replace=False)
    return all_pairs[pair_idx]


def gen_context(region_pair_list, save_path, args, save=True):
    """For every ratio in args.ratio the contexts S of order m = int((R-2)*ratio) of every pair: all C(R-2, m)
    subsets of N \\ {i,j} if there are at most num_save_context_max of them, else num_save_context_max draws of
    np.random.choice(rest, m, replace=False).  Saves ratio%d_context_list.npy (P, ctx, m) under save_path and
    returns {ratio percent: array}."""
    out = {}
    for ratio in args.ratio:
        context_list = []
        m = int((args.num_regions - 2) * ratio)
        for region_pair in region_pair_list:
            all_S = list(range(args.num_regions))
            all_S.remove(region_pair[0])
            all_S.remove(region_pair[1])
            if comb(len(all_S), m) > args.num_save_context_max:
                context_this_pair = [np.random.choice(all_S, m, replace=False) for _ in range(args.num_save_context_max)]
            else:
                context_this_pair = list(itertools.combinations(all_S, m))
            context_list.append(context_this_pair)
        context_list = np.array(context_list)
        out[int(ratio * 100)] = context_list
        if save:
            np.save(save_path + "ratio%d_context_list.npy" % (int(ratio * 100)), context_list)
    return out


def _forward(model, clouds, args):
    """clouds (B,N,3) on the model's device -> logits (B,C) through the nn.Module contract (B,3,N)."""
    out = model(clouds.permute(0, 2, 1).contiguous())
    return out[0] if isinstance(out, tuple) else out


def gen_pred_label(model, data, lbl, disturb_fn, save_path, args):
    """Prediction at the pose stored in save_path/transform_params.npy; writes pred_labels.txt / pred_labels.npy."""
    dev = _device_of(model)
    transform_params = torch.from_numpy(np.load(save_path + "transform_params.npy").astype(np.float32)).to(dev)
    with torch.no_grad():
        logits = _forward(model, disturb_fn(data.to(dev), transform_params), args)
    pred = torch.argmax(logits, dim=1)
    with open(save_path + "pred_labels.txt", "w") as f:
        f.write("lbl: %d\npred_lbl: %d\n" % (lbl[0].item(), pred[0].item()))
    np.save(save_path + "pred_labels.npy", np.array([lbl[0].item(), pred[0].item()]))
    return int(pred[0].item())


def check_adv_success(args, disturb_fn, samples=None, model=None):
    """All enumerated poses of a cloud in ONE forward; counts the misclassified ones and saves the pose with the
    largest attacking utility (lowest reward on the ground-truth class) as <mode>_adv/pose_idx.npy and
    transform_params.npy under interaction_seed<seed>/.  Returns [(num_misclassified, pose_idx), ...]."""
    if samples is None:
        raise ValueError("check_adv_success(): pass samples=[(data, lbl, folder_name), ...]")
    if model is None:
        model = load_model(args)
    dev = _device_of(model)
    results = []
    with torch.no_grad():
        for data, lbl, folder_name in samples:
            data, lbl = data.to(dev), lbl.to(dev)
            base_folder = args.exp_folder + "%s/" % folder_name
            mode_folder = base_folder + "%s_all/" % args.mode
            interaction_folder = base_folder + "interaction_seed%d/" % args.seed
            name = "trans_vector.npy" if args.mode == "trans" else "angle_tuple.npy"
            all_transform_params = np.load(mode_folder + name)
            all_data_disturb = torch.cat([disturb_fn(data, torch.from_numpy(p).to(dev)) for p in all_transform_params], dim=0)
            logits = _forward(model, all_data_disturb, args)
            pred = torch.argmax(logits, dim=1)
            num_miscls = int((pred != lbl[0].item()).sum().item())
            v = get_reward(logits, lbl, args)
            pose_idx = int(torch.argmin(v).item())
            mkdir(interaction_folder + "%s_adv/" % args.mode)
            np.save(interaction_folder + "%s_adv/pose_idx.npy" % args.mode, pose_idx)
            np.save(interaction_folder + "%s_adv/transform_params.npy" % args.mode, all_transform_params[pose_idx])
            results.append((num_miscls, pose_idx))
    return results


# ---- folder walkers (save_pair_random :302-320, save_pair_single_region :145-218, save_context :45-71,
# ---- save_pred_label :90-123).  The reference walks a global folder_name_list / its dataset loaders; here the cloud
# ---- folder names (or `samples`) are arguments.

def _folders(args, name):
    base_folder = args.exp_folder + "%s/" % name
    interaction_folder = base_folder + "interaction_seed%d/" % args.seed
    return base_folder, interaction_folder, interaction_folder + "%s_adv_single_region/" % args.mode


def _region_folders(single_region_folder):
    """Sub-folders range_rank<rr>_region<ii>/ of a cloud's single-region folder, sorted by name like the reference."""
    import os
    return [single_region_folder + d + "/" for d in sorted(os.listdir(single_region_folder))
            if os.path.isdir(single_region_folder + d)]


def save_pair_random(args, folder_name_list):
    """region_pair_list.npy (num_pairs_random, 2) per cloud under interaction_seed<seed>/, shared by the normal and
    the adversarial pose; creates normal/ and <mode>_adv/.  One gen_pair_random draw per cloud, in list order."""
    for name in folder_name_list:
        _, interaction_folder, _ = _folders(args, name)
        mkdir(interaction_folder + "normal/")
        mkdir(interaction_folder + "%s_adv/" % args.mode)
        np.save(interaction_folder + "region_pair_list.npy", gen_pair_random(args))


def gen_pair_single_region(region, neighbor_idx, args):
    """(num_neighbors, 2) pairs (region, j) over the ball-query neighbours j != region (:127-142); an empty (0,)
    array when the region has no neighbour, like np.array([])."""
    neighbors = np.nonzero(neighbor_idx[region])[0]
    return np.array([[region, j] for j in neighbors if j != region])


def save_pair_single_region(args, samples):
    """Per cloud and region: the poses of its largest / smallest Shapley value over the enumeration written by
    tools.final_common.test (<mode>_all/region_shapley_value.npy + trans_vector.npy / angle_tuple.npy) and the pairs
    with its neighbouring regions, under <mode>_adv_single_region/range_rank<rr>_region<ii>/{normal,max_pose,min_pose}/.
    range_rank 1 is the region whose value varies most over the poses.  Returns the last region_pair_list."""
    from .final_result import BALL_QUERY_COEF, ball_query, square_distance_np
    from .tools.final_util import cal_rank
    assert args.mode == "trans" or args.mode == "rotate"
    region_pair_list = None
    for data, _, name in samples:
        data = np.asarray(data.cpu() if hasattr(data, "cpu") else data).reshape(-1, 3)
        base_folder, _, single_region_folder = _folders(args, name)
        mode_folder = base_folder + "%s_all/" % args.mode
        mkdir(single_region_folder)
        region_id = np.load(base_folder + "region_id.npy")
        values = np.load(mode_folder + "region_shapley_value.npy")                     # (poses, R)
        transform_params = np.load(mode_folder + ("trans_vector.npy" if args.mode == "trans" else "angle_tuple.npy"))
        max_pose_idx, min_pose_idx = np.argmax(values, axis=0), np.argmin(values, axis=0)
        range_rank = args.num_regions - cal_rank(values.max(axis=0) - values.min(axis=0))
        diameter = np.sqrt(np.maximum(square_distance_np(data), 0)).max()
        centers = np.zeros((args.num_regions, 3))
        for i in range(args.num_regions):
            centers[i] = data[region_id == i].mean(axis=0)
        neighbor_idx = ball_query(centers, r=BALL_QUERY_COEF * diameter)
        for region in range(args.num_regions):
            region_folder = single_region_folder + "range_rank%02d_region%02d/" % (range_rank[region], region)
            for sub in ("normal/", "max_pose/", "min_pose/"):
                mkdir(region_folder + sub)
            for sub, pose in (("max_pose/", max_pose_idx[region]), ("min_pose/", min_pose_idx[region])):
                np.save(region_folder + sub + "transform_params.npy", transform_params[pose])
                np.save(region_folder + sub + "pose_idx.npy", pose)
            region_pair_list = gen_pair_single_region(region, neighbor_idx, args)
            np.save(region_folder + "region_pair_list.npy", region_pair_list)
    return region_pair_list


def save_context(args, folder_name_list):
    """Contexts for the random pairs of every cloud, then for every single-region pair list (sorted folder order):
    the numpy stream is consumed in exactly the reference's order."""
    for name in folder_name_list:
        _, interaction_folder, single_region_folder = _folders(args, name)
        gen_context(np.load(interaction_folder + "region_pair_list.npy"), interaction_folder, args)
        for region_folder in _region_folders(single_region_folder):
            gen_context(np.load(region_folder + "region_pair_list.npy"), region_folder, args)


def save_pred_label(args, disturb_fn, samples, model=None):
    """pred_labels.{txt,npy} for the adversarial pose and for every region's max / min pose."""
    if model is None:
        model = load_model(args)
    for data, lbl, name in samples:
        _, interaction_folder, single_region_folder = _folders(args, name)
        gen_pred_label(model, data, lbl, disturb_fn, interaction_folder + "%s_adv/" % args.mode, args)
        for region_folder in _region_folders(single_region_folder):
            gen_pred_label(model, data, lbl, disturb_fn, region_folder + "max_pose/", args)
            gen_pred_label(model, data, lbl, disturb_fn, region_folder + "min_pose/", args)
