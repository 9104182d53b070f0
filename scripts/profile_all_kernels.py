"""Runs every kernel family of libiq_b200 once inside a cudaProfiler range, for
   ncu --set full --profile-from-start off ... python scripts/profile_all_kernels.py [families]
families (default all): coalition, dgcnn, gcnn, pointnet, pointnet2, pointconv.  Sizes are the benchmark's where that is
cheap (coalition kernels: 100 permutations / 300 pairs) and one small batch per model (every kernel of the forward
appears with its production tile shapes; batch = one chunk lane)."""
import os, sys, types
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from interpret_quality_b200 import ops, synthetic
from interpret_quality_b200.tools import final_util

fam = sys.argv[1:] or ["coalition", "dgcnn", "gcnn", "pointnet", "pointnet2", "pointconv"]
dev = torch.device("cuda:0")
R, N, LBL = 32, 1024, 3
data = torch.from_numpy(synthetic.make_cloud(N)).to(dev)
fps_idx = ops.fps(data, R)
rid = ops.region_id(data, fps_idx[0].contiguous())
cen = ops.center(data.reshape(-1, 3))
orders = torch.from_numpy(synthetic.make_orders(1000, R)[:100].copy()).to(dev)
pairs_np, ctxs = synthetic.make_pairs_and_contexts(8, R, orders_m=(15,))
pairs = torch.from_numpy(pairs_np).to(dev)
ctx = torch.from_numpy(ctxs[15].astype(np.int64)).to(dev)
masked = ops.mask_shapley(data.reshape(-1, 3), cen, orders, rid)
logits = torch.randn((3300, 10), device=dev)
ilog = torch.randn((300, 400, 10), device=dev)


def model_of(name):
    a = types.SimpleNamespace(model=name, k=20, dataset="shapenet", feature_transform=True, device=dev, num_points=N,
                              num_regions=R, softmax_type="modified")
    m = final_util.build_model(a, synthetic.make_state_dict(name))
    m.set_lanes(1)
    return m


# one chunk lane per model and a PLAIN forward (every cloud at N points: one launch per kernel of the forward, with the
# production tile shapes); IQ_PROFILE_COLLAPSED=1 profiles the collapsed forward instead (8 size groups: 8x the launches)
BATCH = {"dgcnn": 148, "gcnn": 148, "pointnet": 330, "pointnet2": 33, "pointconv": 66}
COLLAPSED = bool(os.environ.get("IQ_PROFILE_COLLAPSED"))
models = {n: model_of(n) for n in fam if n != "coalition"}
for n, m in models.items():                                    # warm-up outside the profiled range (workspace, handles)
    m.forward_point_major(masked[:BATCH[n]], masked_to=cen if COLLAPSED else None)
torch.cuda.synchronize()
torch.cuda.profiler.start()
if "coalition" in fam:
    ops.fps(data, R)
    ops.region_id(data, fps_idx[0].contiguous())
    ops.center(data.reshape(-1, 3))
    ops.mask_shapley(data.reshape(-1, 3), cen, orders, rid, out=masked)
    ops.mask_interaction_pairs(data.reshape(-1, 3), cen, pairs, ctx, rid, R)
    phi = torch.zeros(R, dtype=torch.float64, device=dev)
    v = ops.reward(logits, LBL)
    ops.shapley_accumulate(v, orders, phi)
    ops.interaction_reduce(ilog, LBL)
    if "dgcnn" in models:            # the collapse kernels (count, compact, scatter) + one 640-point size group of DGCNN
        models["dgcnn"].forward_point_major(masked[16:17].repeat(64, 1, 1).contiguous(), masked_to=cen)
for n, m in models.items():
    m.forward_point_major(masked[:BATCH[n]], masked_to=cen if COLLAPSED else None)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled families:", fam)
