"""Bisecting probe of csrc/chain_tc.cu: runs one configuration with IQ_CHAIN_DBG and reports whether the kernel completes.
   IQ_CHAIN_DBG=<bits> python scripts/chain_probe.py C1 C2 C3 K [tiles]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from interpret_quality_b200 import ops
C1, C2, C3, K = (int(x) for x in sys.argv[1:5])
tiles = int(sys.argv[5]) if len(sys.argv) > 5 else 160
g = torch.Generator().manual_seed(1)
clouds, nsrc = 1, 200
S = 128 * tiles // K
U = torch.randn((clouds * nsrc, C1), generator=g).cuda()
V = torch.randn((clouds * S, C1), generator=g).cuda() * 0.5
b1 = torch.zeros((C1,)).cuda()
idx = torch.randint(0, nsrc, (clouds * S * K,), generator=g, dtype=torch.int32).cuda()
W2 = (torch.randn((C2, C1), generator=g) / C1 ** 0.5).cuda(); b2 = torch.zeros((C2,)).cuda()
W3 = (torch.randn((C3, C2), generator=g) / C2 ** 0.5).cuda(); b3 = torch.zeros((C3,)).cuda()
try:
    out = ops.grouped_mlp_max(U, V, b1, idx, clouds, S, K, W2, b2, W3, b3)
    torch.cuda.synchronize()
    print("dbg=%s %s tiles=%d: completed, out mean %.4f" % (os.environ.get("IQ_CHAIN_DBG", "0"), sys.argv[1:5], tiles, float(out.mean())))
except Exception as e:
    print("dbg=%s %s tiles=%d: FAILED %s" % (os.environ.get("IQ_CHAIN_DBG", "0"), sys.argv[1:5], tiles, str(e).splitlines()[0]))
