"""Per-cloud set-up steps with the reference's signatures (final_shapley_value.py of
ada-shen/Interpret_quality): cal_region_id :20-35, cal_norm_factor :39-56,
generate_all_orders :59-72, mask_data :74-88."""
import numpy as np
import torch

from . import ops
from .tools.final_common import cal_reward


def cal_region_id(data, fps_index, result_path, save=True):
    """data (1,N,3) CUDA, fps_index (R,) -> region_id (N,) int64 ndarray; writes region_id.npy when save."""
    idx = ops.to_dev_i64(fps_index, data.device)
    region_id = ops.region_id(data.contiguous(), idx).cpu().numpy()
    if save:
        np.save(result_path + "region_id.npy", region_id)
    return region_id


def cal_norm_factor(model, data, lbl, center, result_path, args, save=True):
    """v(N) - v(empty set): two forwards, the full cloud and the all-centre cloud."""
    B = data.shape[0]
    empty = center.to(data.device).view(1, 1, 3).expand(B, args.num_points, 3).contiguous()
    v_full, _ = cal_reward(model, data, lbl, args)
    v_empty, _ = cal_reward(model, empty, lbl, args)
    norm_factor = (v_full - v_empty).cpu().item()
    if save:
        np.save(result_path + "norm_factor.npy", norm_factor)
    return norm_factor


def generate_all_orders(result_path, args, save=True):
    """num_samples_save permutations of the regions from numpy's global legacy stream (seed replay:
    call set_random(seed) first, exactly as the reference's main does)."""
    all_orders = np.stack([np.random.permutation(np.arange(0, args.num_regions, 1))
                           for _ in range(args.num_samples_save)], axis=0)
    if save:
        np.save(result_path + "all_orders.npy", all_orders)
    return all_orders


def mask_data(masked_data, center, order, region_id):
    """Single-permutation form of mask_data_batch: masked_data (R+1, N, 3) modified in place."""
    dev = masked_data.device
    order = np.asarray(order).reshape(1, -1)
    ops.mask_shapley(None, center.to(dev, torch.float32).contiguous(), ops.to_dev_i64(order, dev),
                     ops.to_dev_i64(region_id, dev), out=masked_data, in_place=True)
    return masked_data
