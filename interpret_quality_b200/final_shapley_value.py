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


def save_shapley(region_shap_value, pc_idx, count, result_path, region_id, args):
    """final_shapley_value.py:91-107: per-point and per-region Shapley estimates after `count` permutations:
    shapley/<pc_idx>_<count>.npy (N,) and region_shapley/<pc_idx>_<count>.npy (R,), float64."""
    from .tools.final_util import mkdir
    shap_value = np.zeros((args.num_points,))
    mkdir(result_path + "shapley/")
    mkdir(result_path + "region_shapley/")
    for k in range(0, args.num_regions):
        shap_value[region_id == k] = region_shap_value[k] / count
    np.save(result_path + "shapley/%s.npy" % (str(pc_idx) + '_' + str(count)), shap_value)
    np.save(result_path + "region_shapley/%s.npy" % (str(pc_idx) + '_' + str(count)), region_shap_value / count)


SAMPLE_NUMS = [100, 200, 300, 400, 500, 600, 700, 800, 900, 1000, 2000, 3000, 4000, 5000]


def shap_sampling(model, dataloader, args, folder_name_list, fps_indices=None):
    """final_shapley_value.py:111-156: the num_samples_save-permutation Shapley run of every cloud of `dataloader`
    (any iterable of (data (1,N,3), lbl (1,))), writing region_id.npy, norm_factor.npy, all_orders.npy, the
    checkpoints of save_shapley at the reference's 14 counts and region_sv_all.npy (num_samples_save, R) float64.

    The reference evaluates one permutation (33 clouds) per forward; here all permutations go through the engine in
    passes of 100, the marginal contributions v[r+1]-v[r] are formed in fp32 like the reference's tensors and the
    float64 running sums are replayed on the host in the same order (permutation by permutation), so the
    checkpoints accumulate exactly as the reference's `region_shap_value[order] += dv`.
    fps_indices defaults to the reference's file fps_<dataset>_<N>_<R>_index_final30.npy in the working directory."""
    from .tools.final_common import _device_of, shapley_partial_sums
    from .tools.final_util import mkdir
    dev = _device_of(model)
    if fps_indices is None:
        fps_indices = np.load('fps_%s_%d_%d_index_final30.npy' % (args.dataset, args.num_points, args.num_regions))
    with torch.no_grad():
        for i, (data, lbl) in enumerate(dataloader):
            result_path = args.exp_folder + '%s/' % folder_name_list[i]
            mkdir(result_path)
            data, lbl = data.to(dev), lbl.to(dev)
            region_id = cal_region_id(data, fps_indices[i], result_path, save=True)
            center = torch.mean(data, dim=1).squeeze()
            cal_norm_factor(model, data, lbl, center, result_path, args, save=True)
            all_orders = generate_all_orders(result_path, args, save=True)
            n = int(args.num_samples_save)
            R = int(args.num_regions)
            pargs = type(args)(**vars(args)) if hasattr(args, "__dict__") else args
            _, logits = shapley_partial_sums(model, data, lbl, region_id, all_orders[:n], pargs)
            y = int(lbl[0].item())
            v = ops.reward(logits, y, "normal" if args.softmax_type == "normal" else "modified").view(n, R + 1)
            dv = (v[:, 1:] - v[:, :-1]).cpu().numpy()                        # fp32, like the reference's dv
            region_sv_all = np.zeros((n, R))
            region_sv_all[np.arange(n)[:, None], all_orders[:n]] = dv        # temp[order] += dv
            region_shap_value = np.zeros((R,))
            for count in range(1, n + 1):                                    # same float64 addition order
                region_shap_value += region_sv_all[count - 1]
                if count in SAMPLE_NUMS:
                    save_shapley(region_shap_value, i, count, result_path, region_id, args)
            np.save(result_path + "region_sv_all.npy", region_sv_all)
