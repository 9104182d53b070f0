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
