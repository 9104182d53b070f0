"""Throughput of the other BASELINE.json configurations (C2-C5) through the public Python API, one GPU.
Writes a markdown table.   python scripts/bench_configs.py [out.md]
CUDA events around the whole call, 1 warm-up + median of 3; inputs are host numpy / pinned tensors (end to end)."""
import os, sys, types
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from interpret_quality_b200 import ops, synthetic
from interpret_quality_b200.tools import final_common, final_util
from interpret_quality_b200 import final_point_binary_interaction_logits as fpb, final_cal_interactions as fci

dev = torch.device("cuda:0")
R, LBL = 32, 3
rows = []


def timed(fn, reps=3):
    fn()
    ms = []
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def setup(model, N):
    data = synthetic.make_cloud(N)
    d = torch.from_numpy(data).to(dev)
    fidx = ops.fps(d, R)
    rid = ops.region_id(d, fidx[0].contiguous()).cpu().numpy()
    a = types.SimpleNamespace(model=model, k=20, dataset="shapenet", feature_transform=True, device=dev, num_points=N,
                              num_regions=R, shapley_batch_size=5, num_samples=100, softmax_type="modified")
    m = final_util.build_model(a, synthetic.make_state_dict(model))
    return torch.from_numpy(data).pin_memory(), rid, a, m


def shapley(model, N, nperm, tag):
    data, rid, a, m = setup(model, N)
    a.num_samples = nperm
    orders = synthetic.make_orders(1000, R)
    ms = timed(lambda: final_common.shap_sampling_all_regions_batch(m, data, torch.tensor([LBL]), rid, orders, a))
    fw = nperm * (R + 1)
    rows.append("| %s | %s Shapley, N=%d, %d permutations | %d | %.1f | %.0f |" % (tag, model, N, nperm, fw, ms, fw / ms * 1e3))


def interactions(model, P, tag):
    data, rid, a, m = setup(model, 1024)
    total_fw, total_ms = 0, 0.0
    for mm in (0, 1, 2, 3, 6, 9, 12, 15, 18, 21, 24, 27, 30):
        pairs, ctxs = synthetic.make_pairs_and_contexts(P, R, orders_m=(mm,))
        c = ctxs[mm]
        def run():
            lg = fpb.compute_order_interaction_logits(m, data, rid, pairs, c, a)
            return fci.compute_order_interaction(lg, torch.tensor([LBL]), a)
        total_ms += timed(run, reps=2)
        total_fw += P * c.shape[1] * 4
    rows.append("| %s | %s interactions, 13 orders, %d pairs, <=100 contexts | %d | %.1f | %.0f |" % (tag, model, P, total_fw, total_ms, total_fw / total_ms * 1e3))


def sweep(model, tag):
    data, rid, a, m = setup(model, 1024)
    d = data.to(dev)
    cen = ops.center(d.reshape(-1, 3))
    ridd = torch.from_numpy(rid).to(dev)
    for B in (64, 256, 1024, 4096):
        nperm = (B + R) // (R + 1)
        orders = torch.from_numpy(synthetic.make_orders(1000, R)[:nperm].copy()).to(dev)
        masked = ops.mask_shapley(d.reshape(-1, 3), cen, orders, ridd)[:B].contiguous()
        ms = timed(lambda: m.forward_point_major(masked))
        rows.append("| %s | %s forward of %d masked clouds | %d | %.2f | %.0f |" % (tag, model, B, B, ms, B / ms * 1e3))


shapley("pointnet", 1024, 100, "C1")
shapley("pointnet2", 1024, 1000, "C2")
shapley("dgcnn", 2048, 100, "C3 (one GPU)")
interactions("dgcnn", 8, "C4")
interactions("gcnn", 8, "C4")
sweep("pointconv", "C5")
text = "\n".join(["# Other BASELINE.json configurations, one B200, end to end through the Python API", "",
                  "| config | workload | masked forwards | ms | forwards/s |", "|---|---|---:|---:|---:|"] + rows) + "\n"
print(text)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
