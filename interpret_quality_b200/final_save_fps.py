"""farthest_point_sample with the reference's signature (final_save_fps.py:10-31 of
ada-shen/Interpret_quality; the in-model copies models/pointnet2.py:45-68 and
models/pointconv.py:54-77 are the same function) and save_fps :34-54."""
import numpy as np

from . import ops


def farthest_point_sample(xyz, npoint):
    """xyz (B,N,3) float32 CUDA -> (B,npoint) int64: iterative FPS from index 0, lowest index on ties."""
    return ops.fps(xyz.contiguous(), int(npoint))


def save_fps(args, data_loader):
    """FPS region centres of every cloud -> fps_<dataset>_<num_points>_<num_regions>_index_final30.npy (clouds, R) int64
    in the working directory, the file final_shapley_value.shap_sampling reads.  data_loader: iterable of
    (data (B,N,3), label) batches (the reference builds it from its dataset classes, which are out of scope)."""
    rows = []
    for data, _ in data_loader:
        rows.append(farthest_point_sample(data.to(args.device), args.num_regions).cpu().numpy())
    fps_index_all = np.concatenate(rows)
    np.save("fps_%s_%d_%d_index_final30.npy" % (args.dataset, args.num_points, args.num_regions), fps_index_all)
    return fps_index_all
