"""Pose-sharded enumeration runner on N GPUs (tools.final_common.test): the poses of one cloud
are dealt round-robin to the ranks and combined by one allreduce per slab.  Prints forwards/s and, on rank 0, checks the
sharded result against a single-rank run of the same poses.   torchrun --nproc-per-node N scripts/pose_shard_probe.py"""
import os, sys, tempfile, time, types
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from interpret_quality_b200 import ops, synthetic
from interpret_quality_b200 import final_scale_center_enum_all as sc
from interpret_quality_b200.distributed import poses_sharded
from interpret_quality_b200.tools import final_common, final_util

R, LBL, N = 32, 3, 1024
args = types.SimpleNamespace(model="dgcnn", k=20, dataset="shapenet", feature_transform=True, device=dev, num_points=N,
                             num_regions=R, shapley_batch_size=5, num_samples=100, softmax_type="modified", mode="scale",
                             scale_lower=0.5, scale_upper=2.0, num_grid_enum_scale=32)
model = final_util.build_model(args, synthetic.make_state_dict("dgcnn"))
data = torch.from_numpy(synthetic.make_cloud(N)).to(dev)
rid = ops.region_id(data, ops.fps(data, R)[0].contiguous()).cpu().numpy()
orders = synthetic.make_orders(1000, R)
scales = sc.generate_scale(args, dev)
n_pose = scales.shape[0]
lbl = torch.tensor([LBL])
rows = args.num_samples * (R + 1)


def pose_fn(i):
    return final_common.shap_sampling_all_regions_batch(model, sc.scale_pc(data, scales[i]), lbl, rid, orders, args)


for it in range(2):                                     # warm-up pass, then the timed one
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    shap, logits = poses_sharded(pose_fn, n_pose, R, rows, model.output_channels, dev)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
if rank == 0:
    print("pose-shard x%d: %d poses x %d forwards in %.3f s -> %.0f forwards/s (wall clock incl. the two allreduces)"
          % (world, n_pose, rows, dt, n_pose * rows / dt))
    # single-rank replay of three poses: the sharded slabs must hold exactly these values
    for i in (0, n_pose // 2, n_pose - 1):
        phi, lg = pose_fn(i)
        assert np.array_equal(shap[i].cpu().numpy(), phi), "pose %d: sharded phi differs" % i
        assert torch.equal(logits[i], lg), "pose %d: sharded logits differ" % i
    print("sharded == single-rank on poses 0, %d, %d (bitwise)" % (n_pose // 2, n_pose - 1))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
