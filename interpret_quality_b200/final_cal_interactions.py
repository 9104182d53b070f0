"""compute_order_interaction with the reference's signature
(final_cal_interactions.py:14-37 of ada-shen/Interpret_quality)."""
import numpy as np
import torch

from . import ops


def compute_order_interaction(all_logits, lbl, args):
    """all_logits (P, 4*ctx, C) -> (P, ctx) float64 ndarray of v[4k] + v[4k+3] - v[4k+1] - v[4k+2]."""
    y = int(lbl[0].item()) if isinstance(lbl, torch.Tensor) else int(np.asarray(lbl).reshape(-1)[0])
    soft = "normal" if getattr(args, "softmax_type", "modified") == "normal" else "modified"
    if not all_logits.is_cuda:                      # e.g. torch.load(..., map_location="cpu"); there is no CPU path
        all_logits = all_logits.to(getattr(args, "device", None) or "cuda:0")
    return ops.interaction_reduce(all_logits.contiguous(), y, soft).cpu().numpy()


def cal_interaction_all_orders(lbl, save_path, args):
    """final_cal_interactions.py:40-46: ratio%d_all_logits.pt -> ratio%d_<output_type>_interaction.npy (P, ctx) float64
    for every ratio in args.ratio; lbl is the ground-truth or the predicted label (args.output_type in {gt, pred})."""
    for ratio in args.ratio:
        all_logits = torch.load(save_path + "ratio%d_all_logits.pt" % (int(ratio * 100)))
        all_interaction = compute_order_interaction(all_logits, lbl, args)
        np.save(save_path + "ratio%d_%s_interaction.npy" % (int(ratio * 100), args.output_type), all_interaction)


def cal_interaction(args, samples, selected_sample_idx=None):
    """final_cal_interactions.py:49-99: interactions of every cloud from the logits final_point_binary_interaction_logits
    .save_logits wrote -- the normal pose against the ground truth, the adversarial pose against the ground truth or
    (args.output_type == "pred") the label predicted there (<mode>_adv/pred_labels.npy[1]), and the top-ranked region's
    pair list at the normal pose.  samples: iterable of (data, lbl, folder_name); only lbl and the name are used."""
    import os
    for pc_idx, (_, lbl, name) in enumerate(samples):
        if selected_sample_idx is not None and pc_idx not in selected_sample_idx:
            continue
        interaction_folder = args.exp_folder + "%s/interaction_seed%d/" % (name, args.gen_pair_seed)
        single_region_folder = interaction_folder + "%s_adv_single_region/" % args.mode
        cal_interaction_all_orders(lbl, interaction_folder + "normal/", args)
        adv_lbl = lbl
        if args.output_type != "gt":
            adv_lbl = torch.tensor([np.load(interaction_folder + "%s_adv/pred_labels.npy" % args.mode)[1]], dtype=torch.long)
        cal_interaction_all_orders(adv_lbl, interaction_folder + "%s_adv/" % args.mode, args)
        for region_folder_name in sorted(os.listdir(single_region_folder)):
            if not os.path.isdir(single_region_folder + region_folder_name) or int(region_folder_name[10:12]) != 1:
                continue
            cal_interaction_all_orders(lbl, single_region_folder + region_folder_name + "/normal/", args)
