"""farthest_point_sample with the reference's signature (final_save_fps.py:10-31 of
ada-shen/Interpret_quality; the in-model copies models/pointnet2.py:45-68 and
models/pointconv.py:54-77 are the same function)."""
from . import ops


def farthest_point_sample(xyz, npoint):
    """xyz (B,N,3) float32 CUDA -> (B,npoint) int64: iterative FPS from index 0, lowest index on ties."""
    return ops.fps(xyz.contiguous(), int(npoint))
