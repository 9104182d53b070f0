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
