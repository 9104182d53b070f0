"""Scale poses of the pose-enumeration runner, with the reference's names
(final_scale_center_enum_all.py:14-45 of ada-shen/Interpret_quality): scale_pc :14-22,
generate_scale :25-31, print_scale_info :34-36, save_scale_info :39-40."""
import numpy as np
import torch

MODE = "scale"
SCALE_UPPER = 2.0
SCALE_LOWER = 0.5
NUM_GRID_ENUM_SCALE = 30


def scale_pc(data, scale):
    """data (B,N,3), scale scalar tensor -> scaled cloud."""
    return data * scale


def generate_scale(args, device):
    """(num_grid_enum_scale,) float32 scales, linspace(scale_lower, scale_upper)."""
    all_scale = np.linspace(start=args.scale_lower, stop=args.scale_upper, num=args.num_grid_enum_scale)
    return torch.from_numpy(all_scale).float().to(device)


def print_scale_info(io, scale, region_shapley_value, epoch):
    io.cprint("scale: %f" % scale)
    io.cprint("shapley value after %d epoch:\n%s" % (epoch, str(region_shapley_value)))


def save_scale_info(all_scale, result_path):
    np.save(result_path + "scale.npy", all_scale.cpu().numpy())
