"""Multi-GPU sharding of the coalition path: one process per GPU, disjoint slices of the
seed-replayed permutations (or region pairs) per rank, ONE allreduce of the partial sums.

The reference has no multi-GPU path; sampled coalitions are independent, so
the shard needs no data-path collective until the final float64 sum over ranks.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of n units for `rank` of `world`; sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def allreduce_sum_(t, group=None):
    """In-place sum over ranks (no-op at world size 1, which must not need a process group)."""
    _, world = _world(group)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def shapley_values_sharded(partial_fn, load_order_list, num_samples, batch_perms, group=None):
    """phi (R,) float64 ndarray, identical on every rank.

    partial_fn(orders_slice) -> float64 tensor (R,) with the sum over the slice's permutations of the
    marginal contributions (tools.final_common.shapley_partial_sums on a GPU rank).  The permutations
    used are the first (num_samples // batch_perms) * batch_perms rows, like the reference
    (tools/final_common.py:78,87), split into contiguous per-rank blocks.
    """
    rank, world = _world(group)
    used = (int(num_samples) // int(batch_perms)) * int(batch_perms)
    lo, hi = shard_range(used, rank, world)
    part = partial_fn(load_order_list[lo:hi])
    allreduce_sum_(part, group)
    return part.detach().cpu().numpy() / num_samples


def shap_sampling_all_regions_batch(model, data_disturb, lbl, region_id, load_order_list, args, group=None):
    """Sharded drop-in of tools.final_common.shap_sampling_all_regions_batch.  Returns (phi, local_logits):
    phi is the all-rank result; the logits are this rank's rows only (rank-local)."""
    from .tools.final_common import shapley_partial_sums
    keep = {}

    def partial(orders):
        with torch.no_grad():
            phi_sum, logits = shapley_partial_sums(model, data_disturb, lbl, region_id, orders, args)
        keep["logits"] = logits
        return phi_sum

    phi = shapley_values_sharded(partial, load_order_list, args.num_samples, args.shapley_batch_size, group)
    return phi, keep["logits"]


def interactions_sharded(logits_fn, reduce_fn, num_pairs, group=None):
    """(P, ctx) float64 ndarray of interactions with the pairs sharded over ranks.

    logits_fn((lo, hi)) -> (P, 4*ctx, C) tensor with rows of pairs outside [lo, hi) left zero;
    reduce_fn(logits) -> (P, ctx) float64 tensor.  Disjoint slabs are combined by one allreduce."""
    rank, world = _world(group)
    lo, hi = shard_range(num_pairs, rank, world)
    out = reduce_fn(logits_fn((lo, hi)))
    mask = torch.zeros_like(out)
    mask[lo:hi] = 1
    out = out * mask
    allreduce_sum_(out, group)
    return out.detach().cpu().numpy()


def poses_sharded(pose_fn, n_pose, num_regions, rows, channels, device, group=None, on_pose=None):
    """Pose axis of the enumeration runner (tools/final_common.py:150-166 of the reference loops it serially).

    pose_fn(i) -> (phi (R,) float64 ndarray, logits (rows, C) float32 tensor) for pose i.  Poses are dealt
    round-robin (rank, rank+world, ...) so that neighbouring, similarly expensive poses spread over the ranks;
    every rank fills its own rows of zero-initialised (n_pose, R) float64 / (n_pose, rows, C) float32 slabs and
    one allreduce per slab makes them complete on every rank.  on_pose(i, phi) is called after each local pose
    (the runner's per-pose log line at world size 1).  Returns (shap, all_logits) tensors on `device`."""
    rank, world = _world(group)
    shap = torch.zeros((n_pose, num_regions), dtype=torch.float64, device=device)
    all_logits = torch.zeros((n_pose, rows, channels), dtype=torch.float32, device=device)
    for i in range(rank, n_pose, world):
        phi, logits = pose_fn(i)
        shap[i] = torch.as_tensor(np.asarray(phi, dtype=np.float64)).to(device)
        all_logits[i] = logits
        if on_pose is not None:
            on_pose(i, phi)
    allreduce_sum_(shap, group)
    allreduce_sum_(all_logits, group)
    return shap, all_logits
