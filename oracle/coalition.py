"""ORACLE (test infrastructure, not product code).

CPU restatement of the coalition loop of the reference: reward, permutation
sampled Shapley values and order-m pairwise interactions.  Masks come from
geom_oracle.c, forwards from nets.py, sums are float64 numpy like the
reference's accumulators.
"""
import math

import numpy as np
import torch

from . import geom, nets


def reward(logits, lbl, softmax_type="modified"):
    """get_reward, tools/final_common.py:11-24.  logits (B',C) tensor -> (B',) fp32 tensor.
    "normal": log_softmax[:, y]; anything else: z_y - logsumexp(z_{!=y}) = log p/(1-p)."""
    lbl = int(lbl)
    if softmax_type == "normal":
        return torch.log_softmax(logits, dim=1)[:, lbl]
    others = [c for c in range(logits.shape[1]) if c != lbl]
    return logits[:, lbl] - torch.logsumexp(logits[:, others], dim=1)


def center_of(data):
    """torch.mean(data, dim=1).squeeze(), tools/final_common.py:80.  data (1,N,3) fp32 -> (3,) fp32 ndarray."""
    return torch.mean(torch.as_tensor(data), dim=1).squeeze().numpy()


def shapley_logits(model, sd, data, region_id, orders, batch_perms, k=20):
    """Forward of every masked cloud of shap_sampling_all_regions_batch,
    tools/final_common.py:86-92: -> (len(orders)*(R+1), C) fp32 tensor."""
    data = np.asarray(data, np.float32)
    center = center_of(data)
    out = []
    for s in range(0, len(orders), batch_perms):
        masked = geom.mask_shapley(data[0], center, orders[s:s + batch_perms], region_id)
        x = torch.from_numpy(masked).permute(0, 2, 1).contiguous()      # cal_reward, tools/final_common.py:35
        out.append(nets.forward(model, x, sd, k))
    return torch.cat(out, dim=0)


def shapley_from_logits(logits, lbl, orders, num_regions, num_samples, softmax_type="modified"):
    """tools/final_common.py:93-97: dv = v[1:]-v[:-1] scattered by order into float64, / num_samples."""
    v = reward(logits, lbl, softmax_type).numpy()
    phi = np.zeros((num_regions,))
    for p, order in enumerate(orders):
        vv = v[(num_regions + 1) * p:(num_regions + 1) * (p + 1)]
        phi[order] += (vv[1:] - vv[:-1])
    return phi / num_samples


def shap_sampling_all_regions_batch(model, sd, data, lbl, region_id, load_order_list, num_regions, batch_perms,
                                    num_samples, softmax_type="modified", k=20):
    """tools/final_common.py:64-103 -> (phi (R,) float64, logits (num_samples*(R+1), C))."""
    iterations = num_samples // batch_perms
    orders = load_order_list[:iterations * batch_perms]
    logits = shapley_logits(model, sd, data, region_id, orders, batch_perms, k)
    return shapley_from_logits(logits, lbl, orders, num_regions, num_samples, softmax_type), logits


def interaction_logits(model, sd, data, region_id, pairs, contexts, num_regions, batch_ctx, k=20):
    """compute_order_interaction_logits, final_point_binary_interaction_logits.py:15-70
    -> (P, 4*ctx, C) fp32 tensor."""
    data = np.asarray(data, np.float32)
    center = center_of(data)
    ctx = contexts.shape[1]
    out = []
    for p, (ri, rj) in enumerate(pairs):
        rows = []
        for it in range(math.ceil(ctx / batch_ctx)):
            cb = np.asarray(contexts[p][it * batch_ctx:min(ctx, (it + 1) * batch_ctx)]).astype(np.int64)
            cb = cb.reshape(cb.shape[0], -1)
            x = torch.from_numpy(geom.mask_interaction(data[0], center, cb, ri, rj, region_id, num_regions))
            rows.append(nets.forward(model, x, sd, k))
        out.append(torch.cat(rows, dim=0).unsqueeze(0))
    return torch.cat(out, dim=0)


def interaction_from_logits(all_logits, lbl, softmax_type="modified"):
    """compute_order_interaction, final_cal_interactions.py:14-37 -> (P,ctx) float64:
    v[4k] + v[4k+3] - v[4k+1] - v[4k+2], fp32 arithmetic in that order, then widened."""
    P = all_logits.shape[0]
    ctx = all_logits.shape[1] // 4
    out = np.zeros((P, ctx))
    for p in range(P):
        v = reward(all_logits[p], lbl, softmax_type)
        for kk in range(ctx):
            out[p, kk] = (v[4 * kk] + v[4 * kk + 3] - v[4 * kk + 1] - v[4 * kk + 2]).item()
    return out


def interaction_float64_yardstick(model, sd, data, region_id, pairs, contexts, num_regions, lbl, softmax_type="modified", k=20):
    """(P, ctx) interactions from a FLOAT64 evaluation of the network on the same fp32 masked clouds
    interaction_logits() builds: the yardstick against which the rounding noise of an fp32 run (the reference's or
    ours) is measured in the GPU tests (tests/_gates.py)."""
    sd64 = {kk: torch.from_numpy(np.asarray(v)).double() if np.asarray(v).dtype == np.float32 else torch.from_numpy(np.asarray(v))
            for kk, v in sd.items()}
    real = nets.forward
    torch.set_default_dtype(torch.float64)
    try:
        nets.forward = lambda model_, x, sd_, k_=20: real(model_, x.double(), sd64, k_)
        lg = interaction_logits(model, sd, data, region_id, pairs, contexts, num_regions, 25, k)
    finally:
        nets.forward = real
        torch.set_default_dtype(torch.float32)
    return interaction_from_logits(lg, lbl, softmax_type)
