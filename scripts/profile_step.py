"""ONE step of the headline workload (DGCNN, 100 permutations x 33 clouds through shapley_partial_sums: mask -> collapse ->
forward -> reward -> Shapley sums) inside a cudaProfiler range, after one warm-up step outside it:
   ncu --profile-from-start off --metrics ... python scripts/profile_step.py [model] [perms]"""
import os, sys, types
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from interpret_quality_b200 import ops, synthetic
from interpret_quality_b200.tools import final_common, final_util

name = sys.argv[1] if len(sys.argv) > 1 else "dgcnn"
perms = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda:0")
R, N, LBL = 32, 1024, 3
a = types.SimpleNamespace(model=name, k=20, dataset="shapenet", feature_transform=True, device=dev, num_points=N,
                          num_regions=R, shapley_batch_size=5, num_samples=perms, softmax_type="modified")
model = final_util.build_model(a, synthetic.make_state_dict(name))
model.set_lanes(1)                                           # kernels in program order on one stream
data = torch.from_numpy(synthetic.make_cloud(N)).to(dev)
rid = ops.region_id(data, ops.fps(data, R)[0].contiguous())
orders = torch.from_numpy(synthetic.make_orders(1000, R)[:perms].copy()).to(dev)
lbl = torch.tensor([LBL])
with torch.no_grad():
    final_common.shapley_partial_sums(model, data, lbl, rid, orders, a)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    final_common.shapley_partial_sums(model, data, lbl, rid, orders, a)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("profiled one step:", name, perms, "permutations; evaluated", model.last_buckets())
