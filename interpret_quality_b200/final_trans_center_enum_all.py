"""Translation poses of the pose-enumeration runner, with the reference's names
(final_trans_center_enum_all.py:13-57 of ada-shen/Interpret_quality): translate_pc :13-21,
generate_trans_vector :24-43, print_trans_info :46-49, save_trans_info :52-54."""
import numpy as np
import torch

MODE = "trans"
TRANS_DIST_THRESHOLD = 0.5
NUM_GRID_ENUM_TRANS = 6


def translate_pc(data, trans):
    """data (B,N,3), trans (3,) -> translated cloud (B,N,3)."""
    return torch.add(data, trans)


def generate_trans_vector(args, device):
    """(num_grid_enum_trans^3, 3) float32 translation vectors on a cube grid, clipped to the ball of radius
    trans_dist_threshold (same float32 arithmetic as the reference: the grid value is cast first, then scaled)."""
    g = np.linspace(-args.trans_dist_threshold, args.trans_dist_threshold, num=args.num_grid_enum_trans)
    out = []
    for x in g:
        for y in g:
            for z in g:
                t = torch.tensor([x, y, z], dtype=torch.float32)
                if torch.norm(t) > args.trans_dist_threshold:
                    t = t / torch.norm(t) * args.trans_dist_threshold
                out.append(t)
    return torch.stack(out, dim=0).to(device)


def print_trans_info(io, trans, region_shapley_value, epoch):
    io.cprint("translation vector: [%f, %f, %f]" % (trans[0].item(), trans[1].item(), trans[2].item()))
    io.cprint("translation distance: %f" % torch.norm(trans).item())
    io.cprint("shapley value after %d epoch:\n%s" % (epoch, str(region_shapley_value)))


def save_trans_info(all_trans_vector, result_path):
    np.save(result_path + "trans_vector.npy", all_trans_vector.cpu().numpy())
    np.save(result_path + "trans_distance.npy", torch.norm(all_trans_vector, dim=1).cpu().numpy())
