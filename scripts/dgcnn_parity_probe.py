"""Per-cloud parity of the DGCNN forward against the reference's wider golden sample, both GEMM engines."""
import os, sys, types
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from interpret_quality_b200 import synthetic, ops
from interpret_quality_b200.tools import final_util
R = 32
g = np.load(os.path.join(ROOT, "tests/golden/dgcnn_more.npz"))
g0 = np.load(os.path.join(ROOT, "tests/golden/dgcnn.npz"))
geo = np.load(os.path.join(ROOT, "tests/golden/geometry.npz"))
ref = np.concatenate([g0["shapley_logits"], g["shapley_logits"]], 0)
dev = "cuda:0"
a = types.SimpleNamespace(model="dgcnn", k=20, dataset="shapenet", device=dev)
model = final_util.build_model(a, synthetic.make_state_dict("dgcnn"))
data = synthetic.make_cloud(1024)
orders = synthetic.make_orders(16, R)
d = torch.from_numpy(data[0]).to(dev)
center = torch.from_numpy(np.asarray(geo["center"])).to(dev)
masked = ops.mask_shapley(d, center, torch.from_numpy(orders).to(dev), torch.from_numpy(geo["region_id_1024"]).to(dev))
scale = np.abs(ref).max()
for eng in ("3xtf32", "fp32"):
    model.set_engine(eng)
    out = model.forward_point_major(masked).cpu().numpy()
    err = np.abs(out - ref).max(1) / scale
    bad = np.where(err > 1e-4)[0]
    print(eng, "max %.2e  median %.2e  clouds>1e-4: %d of %d" % (err.max(), np.median(err), len(bad), len(err)), bad.tolist()[:40])
