"""compute_order_interaction_logits with the reference's signature
(final_point_binary_interaction_logits.py:15-70 of ada-shen/Interpret_quality)."""
import numpy as np
import torch

from . import ops
from .tools.final_common import _device_of

# masked clouds evaluated per internal pass (pairs x contexts x 4 coalitions); independent of args.interaction_batch_size
ENGINE_CLOUDS_PER_PASS = 16384


def compute_order_interaction_logits(model, data_disturb, region_id, region_pair_list, context_list, args,
                                     pair_slice=None):
    """Logits of the four coalitions S+{i,j}, S+{i}, S+{j}, S of every (pair, context).

    data_disturb (1,N,3); region_id (N,) ndarray; region_pair_list (P,2); context_list (P,ctx,m) (m may
    be 0, dtype may be float64 then).  Returns a float32 CUDA tensor (P, 4*ctx, C), rows 4k..4k+3 in
    the order above.  pair_slice (start, stop) restricts the work to a shard of the pairs (used by
    distributed.py); rows of other pairs are left zero.

    The reference walks the pairs in a host loop and the contexts in batches of args.interaction_batch_size
    (final_point_binary_interaction_logits.py:37-63); here ONE mask launch expands a whole block of pairs and ONE
    forward evaluates its pairs x contexts x 4 clouds, each on its kept points only (iq_model_forward_coalitions).
    """
    dev = _device_of(model)
    R = int(getattr(args, "num_regions", 0) or int(np.max(region_id)) + 1)
    data = data_disturb.to(dev, torch.float32, non_blocking=True).reshape(-1, 3).contiguous()
    N = data.shape[0]
    pairs = np.asarray(region_pair_list).astype(np.int64).reshape(-1, 2)
    ctxs = np.asarray(context_list)
    P, ctx = pairs.shape[0], ctxs.shape[1]
    m = ctxs.shape[2] if ctxs.ndim == 3 else 0
    C = model.output_channels
    out = torch.zeros((P, 4 * ctx, C), dtype=torch.float32, device=dev)
    lo, hi = pair_slice if pair_slice is not None else (0, P)
    if hi <= lo or ctx == 0:
        return out
    ctx_d = ops.to_dev_i64(ctxs.reshape(P, ctx, m)[lo:hi], dev)
    pairs_d = ops.to_dev_i64(pairs[lo:hi], dev)
    region_d = ops.to_dev_i64(region_id, dev)
    center = ops.center(data)
    pairs_per_pass = max(1, ENGINE_CLOUDS_PER_PASS // (4 * ctx))
    masked = torch.empty((4 * ctx * min(pairs_per_pass, hi - lo), N, 3), dtype=torch.float32, device=dev)
    flat = out.view(P * 4 * ctx, C)
    with torch.no_grad():
        for p0 in range(0, hi - lo, pairs_per_pass):
            np_ = min(pairs_per_pass, hi - lo - p0)
            rows = 4 * ctx * np_
            ops.mask_interaction_pairs(data, center, pairs_d[p0:p0 + np_], ctx_d[p0:p0 + np_], region_d, R,
                                       point_major=True, out=masked[:rows])
            model.forward_point_major(masked[:rows], out=flat[(lo + p0) * 4 * ctx:(lo + p0) * 4 * ctx + rows],
                                      masked_to=center)
    return out


def save_logits_all_orders(model, data, region_id, save_path, args):
    """final_point_binary_interaction_logits.py:73-81: for every ratio in args.ratio, the logits of all (pair, context,
    4 coalitions) of the pose in `data` (normal or adversarial) -> save_path/ratio%d_all_logits.pt (P, 4*ctx, C);
    pairs and contexts are read from save_path/../ like the reference."""
    region_pair_list = np.load(save_path + "../region_pair_list.npy")
    for ratio in args.ratio:
        context_list = np.load(save_path + "../ratio%d_context_list.npy" % (int(ratio * 100)))
        all_logits = compute_order_interaction_logits(model, data, region_id, region_pair_list, context_list, args)
        torch.save(all_logits, save_path + "ratio%d_all_logits.pt" % (int(ratio * 100)))


def save_logits(args, disturb_fn, samples, model=None, selected_sample_idx=None):
    """final_point_binary_interaction_logits.py:83-135: per cloud the logits of all (pair, context, coalition) rows at
    the normal pose, at the adversarial pose (<mode>_adv/transform_params.npy) and, for the region whose Shapley value
    varies most (range_rank 01), at the normal pose with that region's pair list.  samples replaces the reference's
    dataset loader; selected_sample_idx (default: all) its global of that name.  args.gen_pair_seed names the
    interaction_seed<seed>/ folder written by final_gen_pair."""
    import os
    from .tools.final_util import load_model
    if model is None:
        model = load_model(args)
    dev = _device_of(model)
    with torch.no_grad():
        for pc_idx, (data, lbl, name) in enumerate(samples):
            if selected_sample_idx is not None and pc_idx not in selected_sample_idx:
                continue
            data = data.to(dev)
            base_folder = args.exp_folder + "%s/" % name
            interaction_folder = base_folder + "interaction_seed%d/" % args.gen_pair_seed
            single_region_folder = interaction_folder + "%s_adv_single_region/" % args.mode
            region_id = np.load(base_folder + "region_id.npy")
            save_logits_all_orders(model, data, region_id, interaction_folder + "normal/", args)
            transform_params = np.load(interaction_folder + "%s_adv/transform_params.npy" % args.mode).astype(np.float32)
            data_disturb = disturb_fn(data, torch.from_numpy(transform_params).to(dev))
            save_logits_all_orders(model, data_disturb, region_id, interaction_folder + "%s_adv/" % args.mode, args)
            for region_folder_name in sorted(os.listdir(single_region_folder)):
                if not os.path.isdir(single_region_folder + region_folder_name):
                    continue
                if int(region_folder_name[10:12]) != 1:          # range_rank<rr>_region<ii>: only the top-ranked region
                    continue
                save_logits_all_orders(model, data, region_id, single_region_folder + region_folder_name + "/normal/", args)
