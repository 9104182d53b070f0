"""Times the fused kNN kernel alone on one 32-cloud chunk, with the epilogue math or the MMAs switched off
(IQ_KNN_DBG), to see which side bounds it."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from interpret_quality_b200 import _lib, ops
rs = np.random.RandomState(0)
for C in (64, 128):
    x = torch.from_numpy(rs.normal(size=(32, 1024, C)).astype(np.float32)).cuda()
    for dbg in (0, 1, 2, 3, 5):
        os.environ["IQ_KNN_DBG"] = str(dbg)
        _lib.load().iq_debug_reload_env()
        for _ in range(2):
            ops.knn_features(x, 20)
        _lib.profile_enable(True)
        for _ in range(5):
            ops.knn_features(x, 20)
        rep = _lib.profile_report()
        _lib.profile_enable(False)
        print("C=%d dbg=%d" % (C, dbg), {k: round(v[0] / v[1] * 1e3, 1) for k, v in rep.items()}, flush=True)
os.environ["IQ_KNN_DBG"] = "0"
