"""Times the tcgen05 STORE GEMM alone (EdgeConv-4 shape) with the epilogue switched off (IQ_TC_DBG=1)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from interpret_quality_b200 import _lib, ops
rs = np.random.RandomState(0)
for (M, N, K) in ((32768, 512, 128), (32768, 128, 64), (32768, 1024, 512)):
    x = torch.from_numpy(rs.normal(size=(M, K)).astype(np.float32)).cuda()
    w = torch.from_numpy(rs.normal(size=(N, K)).astype(np.float32)).cuda()
    b = torch.zeros(N, device="cuda")
    for dbg in (0, 1):
        os.environ["IQ_TC_DBG"] = str(dbg)
        _lib.load().iq_debug_reload_env()
        for _ in range(2):
            ops.linear(x, w, b, act=0, engine=1)
        _lib.profile_enable(True)
        for _ in range(5):
            ops.linear(x, w, b, act=0, engine=1)
        rep = _lib.profile_report()
        _lib.profile_enable(False)
        print((M, N, K), "dbg=%d" % dbg, {k: round(v[0] / v[1] * 1e3, 1) for k, v in rep.items()}, flush=True)
os.environ["IQ_TC_DBG"] = "0"
