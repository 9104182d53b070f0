"""The coalition engine behind the reference's call signatures
(tools/final_common.py of ada-shen/Interpret_quality): get_reward :11-24, cal_reward :26-43,
mask_data_batch :46-61, shap_sampling_all_regions_batch :64-103, test (pose-enumeration runner) :107-174.

Everything that computes runs in libiq_b200.so; this file only moves arguments.  Inputs may be
host tensors / numpy arrays (they are copied to the model's device here, which is what bench.py's
e2e number times) or already-resident CUDA tensors.
"""
import numpy as np
import torch

from .. import ops

# permutations evaluated per internal pass of the fused engine; independent of args.shapley_batch_size
ENGINE_PERMS_PER_PASS = 100


def _device_of(model):
    for p in model.parameters():
        return p.device
    raise RuntimeError("model has no parameters")


def get_reward(logits, lbl, args):
    """logits (B',C), lbl (1,) -> v (B',):  "normal": log_softmax[:, y];  otherwise z_y - logsumexp(z_{!=y})."""
    y = int(lbl[0].item()) if isinstance(lbl, torch.Tensor) else int(np.asarray(lbl).reshape(-1)[0])
    return ops.reward(logits.contiguous(), y, "normal" if args.softmax_type == "normal" else "modified")


def cal_reward(model, data, lbl, args):
    """data (B',N,3) point-major -> (v (B',), logits (B',C))."""
    logits = model.forward_point_major(data.contiguous())
    return get_reward(logits, lbl, args), logits


def mask_data_batch(masked_data, center, orders, region_id, args):
    """In-place, like the reference: masked_data ((R+1)*bs, N, 3) holds copies of the cloud; the points of
    regions orders[p][r:] are moved to `center` in row r of permutation p.  Returns masked_data."""
    dev = masked_data.device
    ops.mask_shapley(None, center.to(dev, torch.float32).contiguous(), ops.to_dev_i64(orders, dev),
                     ops.to_dev_i64(region_id, dev), out=masked_data, in_place=True)
    return masked_data


def shapley_partial_sums(model, data_disturb, lbl, region_id, orders, args, logits_out=None, perms_per_pass=None):
    """Sum over the given permutations of the marginal contributions, per region, as a float64 CUDA
    tensor (R,), plus the logits ((R+1)*len(orders), C).  Building block of
    shap_sampling_all_regions_batch and of the multi-GPU sharding (distributed.py)."""
    dev = _device_of(model)
    R = int(args.num_regions)
    data = data_disturb.to(dev, torch.float32, non_blocking=True).reshape(-1, 3).contiguous()
    N = data.shape[0]
    orders_d = ops.to_dev_i64(orders, dev)
    region_d = ops.to_dev_i64(region_id, dev)
    n_perm = orders_d.shape[0]
    y = int(lbl[0].item()) if isinstance(lbl, torch.Tensor) else int(np.asarray(lbl).reshape(-1)[0])
    center = ops.center(data)
    C = model.output_channels
    if logits_out is None:
        logits_out = torch.empty((n_perm * (R + 1), C), dtype=torch.float32, device=dev)
    phi_sum = torch.zeros((R,), dtype=torch.float64, device=dev)
    step = int(perms_per_pass or ENGINE_PERMS_PER_PASS)
    masked = torch.empty((min(step, max(n_perm, 1)) * (R + 1), N, 3), dtype=torch.float32, device=dev)
    v = torch.empty((masked.shape[0],), dtype=torch.float32, device=dev)
    soft = "normal" if args.softmax_type == "normal" else "modified"
    for s in range(0, n_perm, step):
        o = orders_d[s:s + step]
        rows = o.shape[0] * (R + 1)
        ops.mask_shapley(data, center, o, region_d, out=masked[:rows])
        lg = logits_out[s * (R + 1):s * (R + 1) + rows]
        model.forward_point_major(masked[:rows], out=lg, masked_to=center)
        ops.reward(lg, y, soft, out=v[:rows])
        ops.shapley_accumulate(v[:rows], o, phi_sum)
    return phi_sum, logits_out


def shap_sampling_all_regions_batch(model, data_disturb, lbl, region_id, load_order_list, args):
    """Permutation-sampled Shapley values of all regions of one (disturbed) cloud.

    data_disturb (1,N,3); lbl (1,); region_id (N,) ndarray; load_order_list (>=num_samples, R) ndarray.
    Uses the first (num_samples // shapley_batch_size) * shapley_batch_size permutations, divides by
    num_samples.  Returns (phi (R,) float64 ndarray, logits (used*(R+1), C) float32 CUDA tensor).
    """
    bs = int(args.shapley_batch_size)
    used = (int(args.num_samples) // bs) * bs
    with torch.no_grad():
        phi_sum, logits = shapley_partial_sums(model, data_disturb, lbl, region_id, load_order_list[:used], args)
    region_shap_value = phi_sum.cpu().numpy() / args.num_samples
    assert logits.shape[0] == args.num_samples * (args.num_regions + 1)
    return region_shap_value, logits


def test(args, get_transform_params_fn, disturb_fn, print_info_fn, save_info_fn, samples=None, model=None):
    """Pose-enumeration runner (tools/final_common.py:107-174 of the reference): for every cloud the Shapley
    values of its regions at the original pose and at every enumerated pose (216 translations / rotations,
    30 scales), written in the reference's files under <exp_folder>/<folder>/<mode>_all/:
    orig_shapley_value.npy (R,), region_shapley_value.npy (poses,R) float64, all_logits.pt
    (poses, num_samples*(R+1), C) float32, plus save_info_fn's pose parameters and log.txt.

    The reference iterates its ModelNet / ShapeNet loaders; datasets are not part of this package, so the
    clouds come from `samples`: an iterable of (data (1,N,3) tensor, lbl (1,) tensor, folder_name) whose
    folders already hold norm_factor.npy, region_id.npy and all_orders.npy (final_shapley_value.py writes them).
    `model` defaults to load_model(args).

    Poses are a second independent axis: under torch.distributed the poses of a cloud are
    dealt round-robin to the ranks, each rank runs whole poses, and the (poses,R) values and the logits are
    combined by one allreduce of disjoint zero-initialised slabs; rank 0 writes the files."""
    import time
    import torch.distributed as dist
    from ..distributed import poses_sharded
    from .final_util import IOStream, load_model, mkdir
    if samples is None:
        raise ValueError("test(): pass samples=[(data, lbl, folder_name), ...]; the reference's dataset loaders "
                         "are outside the scope of interpret_quality_b200")
    if model is None:
        model = load_model(args)
    dev = _device_of(model)
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)

    for data, lbl, folder_name in samples:
        data = data.to(dev)
        lbl = lbl.to(dev)
        base_folder = args.exp_folder + "%s/" % folder_name
        mode_folder = base_folder + "%s_all/" % args.mode
        io = None
        if rank == 0:
            mkdir(mode_folder)
            io = IOStream(mode_folder + "log.txt")
            io.cprint(str(args))
        norm_factor = np.load(base_folder + "norm_factor.npy")
        region_id = np.load(base_folder + "region_id.npy")
        load_order_list = np.load(base_folder + "all_orders.npy")
        if io:
            io.cprint("norm factor: %f" % norm_factor)
        t_start = time.time()
        orig_region_shap_value, _ = shap_sampling_all_regions_batch(model, data, lbl, region_id, load_order_list, args)
        if io:
            io.cprint("origin region shapley: %s" % str(orig_region_shap_value))
            np.save(mode_folder + "orig_shapley_value.npy", orig_region_shap_value)

        all_transform_params = get_transform_params_fn(args, data.device)
        n_pose = all_transform_params.size()[0]
        rows = args.num_samples * (args.num_regions + 1)
        def pose_fn(i):
            data_disturb = disturb_fn(data, all_transform_params[i])
            return shap_sampling_all_regions_batch(model, data_disturb, lbl, region_id, load_order_list, args)

        def on_pose(i, region_shap_value):
            if io and world == 1:
                print_info_fn(io, all_transform_params[i], region_shap_value, i)

        shap, all_logits = poses_sharded(pose_fn, n_pose, args.num_regions, rows, model.output_channels, dev,
                                         on_pose=on_pose)
        if rank == 0:
            region_shapley_list = shap.cpu().numpy()
            if world > 1:
                for i in range(n_pose):
                    print_info_fn(io, all_transform_params[i], region_shapley_list[i], i)
            np.save(mode_folder + "region_shapley_value.npy", region_shapley_list)
            torch.save(all_logits.cpu(), mode_folder + "all_logits.pt")
            save_info_fn(all_transform_params, mode_folder)
            io.cprint("time: %f" % (time.time() - t_start))
            io.close()
