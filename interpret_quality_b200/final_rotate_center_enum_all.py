"""Rotation poses of the pose-enumeration runner, with the reference's names
(final_rotate_center_enum_all.py:15-72 of ada-shen/Interpret_quality): rotate_xyz :15-38,
generate_rotate_angle :41-58, print_rotate_info :61-64, save_rotate_info :67-68."""
import math

import numpy as np
import torch

MODE = "rotate"
ANGLE_THRESHOLD = math.pi / 4
NUM_GRID_ENUM_ROTATE = 6


def rotate_xyz(x, angle_tuple):
    """x (B,N,3), angle_tuple (3,) = (theta_x, theta_y, theta_z) -> x R^T with R = Rx Ry Rz (float32)."""
    B = x.shape[0]
    cx, cy, cz = torch.cos(angle_tuple[0]), torch.cos(angle_tuple[1]), torch.cos(angle_tuple[2])
    sx, sy, sz = torch.sin(angle_tuple[0]), torch.sin(angle_tuple[1]), torch.sin(angle_tuple[2])
    rx = torch.tensor([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], device=x.device)
    ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], device=x.device)
    rz = torch.tensor([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], device=x.device)
    r = torch.matmul(torch.matmul(rx, ry), rz)
    return torch.matmul(x, r.expand(B, 3, 3).permute(0, 2, 1))


def generate_rotate_angle(args, device):
    """(num_grid_enum_rotate^3, 3) float32 angle tuples on a cube grid of [-angle_threshold, angle_threshold]."""
    g = np.linspace(-args.angle_threshold, args.angle_threshold, num=args.num_grid_enum_rotate)
    out = [torch.tensor([a, b, c], dtype=torch.float32) for a in g for b in g for c in g]
    return torch.stack(out, dim=0).to(device)


def print_rotate_info(io, angle_tuple, region_shapley_value, epoch):
    io.cprint("rotation angle: [%f pi, %f pi, %f pi]" % (
        angle_tuple[0].item() / np.pi, angle_tuple[1].item() / np.pi, angle_tuple[2].item() / np.pi))
    io.cprint("shapley value after %d epoch:\n%s" % (epoch, str(region_shapley_value)))


def save_rotate_info(all_rotate_angle, result_path):
    np.save(result_path + "angle_tuple.npy", all_rotate_angle.cpu().numpy())
