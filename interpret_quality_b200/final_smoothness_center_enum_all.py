"""Region-smoothness enumeration with the reference's names (final_smoothness_center_enum_all.py of
ada-shen/Interpret_quality): every region of a cloud is pushed by normalised gradient
steps so that its linearity / planarity / scattering rises ("inc") or falls ("dec") by ENUM_STEP per epoch, and the
region Shapley values are sampled after every epoch.

What runs where:
* the per-region setup (principal orientations :22-45, variance bounds :75-80, original smoothness :143-159) is a
  3x3 eigen-decomposition per region, once per cloud and objective: torch on the host copy of the cloud
  (torch.linalg.eigh; the reference's torch.symeig :41 no longer exists);
* the gradient loop of an epoch, all regions, all steps: ONE launch of csrc/smoothness.cu
  (ops.region_smoothness_epoch) instead of the reference's ~10^5 tiny kernels and .item() round trips;
* the Shapley values after the epoch: tools.final_common.shap_sampling_all_regions_batch, the masked-coalition path.

Behaviour kept from the reference, on purpose: apply_distance_bound :103-120 only COUNTS the points farther than
dist_threshold from their original position (its pull-back assigns to a temporary row view and changes nothing);
args.enforce_distance_bound = True turns the pull-back on.  data_smoothness.npy holds one snapshot per epoch, as the
reference writes on CUDA (on a CPU run its `.cpu().numpy()` aliases the live cloud).

The dataset loaders of test_smoothness :353-381 are out of scope: clouds come from `samples`."""
import time

import numpy as np
import torch

from . import ops
from .tools.final_common import _device_of, shap_sampling_all_regions_batch
from .tools.final_util import IOStream, load_model, mkdir

STEP = 1e-3             # :13 step size of one gradient step
ENUM_STEP = 0.05        # :14 change of the smoothness per epoch
EPOCH = 50              # :15
VAR_THRESHOLD = 0.003   # :16 bound on the change of the variance along each principal orientation
DIST_THRESHOLD = 0.03   # :17
STOP_RATIO = 0.5        # :18
MAX_ITERATION = 100     # :19
MODES = ("linearity", "planarity", "scattering")


def set_smoothness_args(args):
    """The hyper-parameters main() :389-399 puts on args."""
    args.step, args.enum_step, args.epoch = STEP, ENUM_STEP, EPOCH
    args.var_threshold, args.dist_threshold = VAR_THRESHOLD, DIST_THRESHOLD
    args.stop_ratio, args.max_iteration = STOP_RATIO, MAX_ITERATION
    return args


def cal_principal_orientation(data_region_i_orig):
    """(S,3) -> o1,o2,o3 (3,) unit eigenvectors of the (S-1)-normalised covariance for the largest, middle and
    smallest eigenvalue (:22-45)."""
    pts = data_region_i_orig
    c = pts - pts.mean(dim=0)
    cov = (c.unsqueeze(2) * c.unsqueeze(1)).sum(dim=0) / (pts.shape[0] - 1)
    _, vec = torch.linalg.eigh(cov)
    return vec[:, 2].clone().detach(), vec[:, 1].clone().detach(), vec[:, 0].clone().detach()


def cal_variance(data_region_i, o1, o2, o3):
    """Unbiased variance of the projections on the three orientations (:48-62)."""
    return tuple(torch.var(torch.matmul(data_region_i, o)) for o in (o1, o2, o3))


def set_var_bound(var1_orig, var2_orig, var3_orig, args):
    """(:75-80) -> ub1, ub2, ub3, lb1, lb2, lb3."""
    v = (var1_orig, var2_orig, var3_orig)
    return tuple(x + args.var_threshold for x in v) + tuple(x - args.var_threshold for x in v)


def smoothness_of(var1, var2, var3, mode):
    """The mode's ratio of the sorted variances (:143-159) as a Python float."""
    s_min, s_mid, s_max = sorted([var1, var2, var3], key=float)
    if mode == "linearity":
        return ((s_max - s_mid) / s_max).item()
    if mode == "planarity":
        return ((s_mid - s_min) / s_max).item()
    if mode == "scattering":
        return (s_min / s_max).item()
    raise ValueError("mode must be one of %s" % (MODES,))


def get_original_region_info(data, region_id, region_i, io, args):
    """(:245-268) -> data_region_i_orig (S,3), smoothness_orig, (o1,o2,o3), (ub1..3, lb1..3)."""
    pts = data[:, np.asarray(region_id) == region_i, :].squeeze(0).clone().detach()
    if pts.shape[0] < 2:
        raise ValueError("region %d has %d point(s); the unbiased variance needs two" % (region_i, pts.shape[0]))
    orientations = cal_principal_orientation(pts)
    var_orig = cal_variance(pts, *orientations)
    io.cprint("var1 orig: %.8f, var2 orig: %.8f, var3 orig: %.8f" % tuple(v.item() for v in var_orig))
    bounds = set_var_bound(*var_orig, args)
    with torch.no_grad():
        smoothness_orig = smoothness_of(*var_orig, args.mode)
    io.cprint("orig %s: %.8f" % (args.mode, smoothness_orig))
    return pts, smoothness_orig, orientations, bounds


class RegionGeometry:
    """Device-resident state of the enumeration of one cloud: membership lists, the undisturbed cloud, per-region
    orientations / bounds, and the running smoothness / if_update flags."""

    def __init__(self, data, region_id, io, args, device):
        R = args.num_regions
        region_id = np.asarray(region_id)
        host = data.detach().to("cpu", torch.float32).reshape(1, -1, 3)
        members = [np.nonzero(region_id == r)[0] for r in range(R)]
        info = [get_original_region_info(host, region_id, r, io, args) for r in range(R)]
        self.num_regions = R
        self.max_region = max(len(m) for m in members)
        self.offsets = torch.tensor(np.concatenate([[0], np.cumsum([len(m) for m in members])]), dtype=torch.int32, device=device)
        self.members = torch.tensor(np.concatenate(members), dtype=torch.int32, device=device)
        self.orient = torch.stack([torch.stack(i[2]) for i in info]).to(device, torch.float32).contiguous()
        self.var_ub = torch.stack([torch.stack(i[3][:3]) for i in info]).to(device, torch.float32).contiguous()
        self.var_lb = torch.stack([torch.stack(i[3][3:]) for i in info]).to(device, torch.float32).contiguous()
        self.smoothness = torch.tensor([i[1] for i in info], dtype=torch.float64, device=device)
        self.alive = torch.ones((R,), dtype=torch.int32, device=device)
        self.data_orig = data.detach().to(device, torch.float32).reshape(-1, 3).contiguous().clone()


def update_all_regions(data_copy, geom, objective, io, args, only_region=None):
    """One epoch (:305-321): update_region for every region that is still being updated, in one launch.
    data_copy (1,N,3) CUDA tensor, modified in place.  Returns the (R,) list of smoothness values after the epoch
    (the value of the previous epoch for regions that are no longer updated)."""
    before = geom.smoothness.cpu().numpy().copy()
    was_alive = geom.alive.cpu().numpy().copy()
    alive = geom.alive
    if only_region is not None:
        alive = torch.zeros_like(geom.alive)
        alive[only_region] = 1
    flat = data_copy.view(-1, 3)
    iters, last_var, flags = ops.region_smoothness_epoch(
        flat, geom.data_orig, geom.offsets, geom.members, geom.orient, geom.var_ub, geom.var_lb, geom.smoothness, alive,
        geom.max_region, args.mode, objective, args.step, args.enum_step, args.dist_threshold, args.stop_ratio,
        args.max_iteration, clamp=bool(getattr(args, "enforce_distance_bound", False)))
    if only_region is not None:
        geom.alive[only_region] = alive[only_region]
        was_alive = np.zeros_like(was_alive)
        was_alive[only_region] = 1
    after = geom.smoothness.cpu().numpy()
    last_var, flags = last_var.cpu().numpy(), flags.cpu().numpy()
    sign = 1.0 if objective == "inc" else -1.0
    for r in range(geom.num_regions):
        if not was_alive[r]:
            continue
        io.cprint("\tregion%d orig %s: %.8f, target %s: %.8f" % (r, args.mode, before[r], args.mode,
                                                                  before[r] + sign * args.enum_step))
        if flags[r] & 1:
            io.cprint("stop: more than 50% points exceed distance bound")
        if flags[r] & 2:
            io.cprint("stop: all orientations exceed variance bound, no gradient")
        if flags[r] & 4:
            io.cprint("stop: achieve max iteration")
        io.cprint("var1: %.8f, var2: %.8f, var3: %.8f" % tuple(last_var[r]))
        io.cprint("curr smoothness: %.8f" % after[r])
    return after.tolist()


def update_region(data_copy, data_region_i_orig, region_id, region_i, objective, io, args, orientations, bounds,
                  smoothness_orig, geom=None):
    """(:184-243) one region, the reference's arguments -> (data_copy, smoothness, if_update).  Without `geom` the
    device state is rebuilt from the arguments (a one-region launch; test_all_region uses update_all_regions)."""
    dev = data_copy.device
    sel = np.nonzero(np.asarray(region_id) == region_i)[0]
    if geom is None:
        geom = RegionGeometry.__new__(RegionGeometry)
        geom.num_regions, geom.max_region = 1, len(sel)
        geom.offsets = torch.tensor([0, len(sel)], dtype=torch.int32, device=dev)
        geom.members = torch.tensor(np.concatenate([sel, np.setdiff1d(np.arange(data_copy.shape[1]), sel)]), dtype=torch.int32,
                                    device=dev)
        geom.orient = torch.stack(list(orientations)).to(dev, torch.float32).reshape(1, 3, 3).contiguous()
        geom.var_ub = torch.stack(list(bounds[:3])).to(dev, torch.float32).reshape(1, 3).contiguous()
        geom.var_lb = torch.stack(list(bounds[3:])).to(dev, torch.float32).reshape(1, 3).contiguous()
        geom.smoothness = torch.tensor([smoothness_orig], dtype=torch.float64, device=dev)
        geom.alive = torch.ones((1,), dtype=torch.int32, device=dev)
        geom.data_orig = data_copy.detach().reshape(-1, 3).clone()
        geom.data_orig[torch.as_tensor(sel, device=dev)] = data_region_i_orig.to(dev, torch.float32)
        slot = 0
    else:
        slot = region_i
    smooth = update_all_regions(data_copy, geom, objective, io, args, only_region=slot)
    return data_copy, smooth[slot], bool(geom.alive[slot].item())


def test_all_region(model, data, lbl, load_order_list, region_id, mode_folder, args, objective):
    """(:281-350) the enumeration of one cloud for one objective.  Files under mode_folder/allregion_<objective>/:
    orig_shapley_value.npy (R,), region_shapley_value.npy (epochs,R) float64, all_logits.pt (epochs, rows, C),
    <mode>.npy (epochs,R) smoothness, data_smoothness.npy (epochs,1,N,3), log.txt."""
    assert objective in ["inc", "dec"]
    t_start = time.time()
    dev = _device_of(model)
    data, lbl = data.to(dev), lbl.to(dev)
    result_path = mode_folder + "allregion_%s/" % objective
    mkdir(result_path)
    io = IOStream(result_path + "log.txt")
    io.cprint(str(args))
    data_copy = data.clone().detach().to(torch.float32).contiguous()

    orig_shap_value, _ = shap_sampling_all_regions_batch(model, data, lbl, region_id, load_order_list, args)
    io.cprint("origin shapley of this region: %s" % str(orig_shap_value))
    np.save(result_path + "orig_shapley_value.npy", orig_shap_value)
    geom = RegionGeometry(data, region_id, io, args, dev)

    data_list, smoothness_list, region_shapley_list, all_logits_list = [], [], [], []
    for i in range(args.epoch):
        io.cprint("\n************ epoch %d ***********" % i)
        smoothness_list.append(update_all_regions(data_copy, geom, objective, io, args))
        data_list.append(data_copy.cpu().numpy().copy())
        region_shap_values, all_logits_this_pose = shap_sampling_all_regions_batch(model, data_copy, lbl, region_id,
                                                                                   load_order_list, args)
        region_shapley_list.append(region_shap_values)
        all_logits_list.append(all_logits_this_pose)
        io.cprint("region shapley value: %s" % str(region_shap_values))
        if not bool(geom.alive.any().item()):
            break

    np.save(result_path + "region_shapley_value.npy", np.array(region_shapley_list))
    torch.save(torch.stack(all_logits_list, dim=0).cpu(), result_path + "all_logits.pt")
    np.save(result_path + "%s.npy" % args.mode, smoothness_list)
    np.save(result_path + "data_smoothness.npy", data_list)
    io.cprint("time: %f" % (time.time() - t_start))
    io.close()


def test_smoothness(args, samples=None, model=None):
    """(:353-381) both objectives for every cloud.  samples: iterable of (data (1,N,3), lbl (1,), folder_name) whose
    folders hold region_id.npy and all_orders.npy (final_shapley_value.py writes them)."""
    if samples is None:
        raise ValueError("test_smoothness(): pass samples=[(data, lbl, folder_name), ...]; the reference's dataset "
                         "loaders are outside the scope of interpret_quality_b200")
    if model is None:
        model = load_model(args)
    for data, lbl, folder_name in samples:
        base_folder = args.exp_folder + "%s/" % folder_name
        mode_folder = base_folder + "%s_all/" % args.mode
        region_id = np.load(base_folder + "region_id.npy")
        load_order_list = np.load(base_folder + "all_orders.npy")
        test_all_region(model, data, lbl, load_order_list, region_id, mode_folder, args, objective="inc")
        test_all_region(model, data, lbl, load_order_list, region_id, mode_folder, args, objective="dec")
