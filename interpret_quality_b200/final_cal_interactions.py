"""compute_order_interaction with the reference's signature
(final_cal_interactions.py:14-37 of ada-shen/Interpret_quality)."""
import numpy as np
import torch

from . import ops


def compute_order_interaction(all_logits, lbl, args):
    """all_logits (P, 4*ctx, C) -> (P, ctx) float64 ndarray of v[4k] + v[4k+3] - v[4k+1] - v[4k+2]."""
    y = int(lbl[0].item()) if isinstance(lbl, torch.Tensor) else int(np.asarray(lbl).reshape(-1)[0])
    soft = "normal" if getattr(args, "softmax_type", "modified") == "normal" else "modified"
    return ops.interaction_reduce(all_logits.contiguous(), y, soft).cpu().numpy()


def cal_interaction_all_orders(lbl, save_path, args):
    """final_cal_interactions.py:40-46: ratio%d_all_logits.pt -> ratio%d_<output_type>_interaction.npy (P, ctx) float64
    for every ratio in args.ratio; lbl is the ground-truth or the predicted label (args.output_type in {gt, pred})."""
    import numpy as np
    import torch
    for ratio in args.ratio:
        all_logits = torch.load(save_path + "ratio%d_all_logits.pt" % (int(ratio * 100)))
        all_interaction = compute_order_interaction(all_logits, lbl, args)
        np.save(save_path + "ratio%d_%s_interaction.npy" % (int(ratio * 100), args.output_type), all_interaction)
