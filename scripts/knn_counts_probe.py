"""Candidate-list sizes of the fused kNN on REAL DGCNN features of masked clouds (layer 1-3 outputs, computed by the
CPU oracle): how tight is the block-maxima threshold on the actual workload?"""
import os, sys
import numpy as np, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from interpret_quality_b200 import ops, synthetic
from oracle import coalition, geom, nets
R = 32
data = synthetic.make_cloud(1024)
rid = geom.region_id(data[0], geom.fps(data, R)[0])
masked = geom.mask_shapley(data[0], coalition.center_of(data), synthetic.make_orders(1, R), rid)   # (33,N,3)
sd = synthetic.make_state_dict("dgcnn")
sd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
x = torch.from_numpy(masked).permute(0, 2, 1).contiguous()
with torch.no_grad():
    idx = nets.knn_indices(x, 20)
    h = x
    for i in range(1, 4):
        if i > 1:
            idx = nets.knn_indices(h, 20)
        e = nets.edge_features(h, idx)
        e = F.leaky_relu(nets._bn(nets._conv(e, sd, "conv%d.0" % i), sd, "bn%d" % i), 0.2)
        h = e.max(dim=-1)[0]
        feat = h.permute(0, 2, 1).contiguous().cuda()
        _, cnt = ops.knn_features(feat, 20, return_counts=True)
        c = cnt.cpu().numpy()
        print("layer %d C=%d: candidates mean %.1f median %d p90 %d max %d  >32: %.1f%%  >64: %.2f%%" % (
            i, feat.shape[2], c.mean(), np.median(c), np.percentile(c, 90), c.max(), 100 * (c > 32).mean(), 100 * (c > 64).mean()))
        per_cloud = c.mean(1)
        print("   per masked-cloud row (0 = everything masked ... 32 = nothing masked):", np.round(per_cloud[::4], 1))
