"""HBM-side kernels of the path (regions, masking, reward, reductions) timed alone at sizes where they can be measured,
against the copy bandwidth of MEASURED_PEAKS.json.  Writes a markdown table (profiles/<tag>_hbm_kernels.md).

    python scripts/bench_kernels.py [out.md]

Timing: the library's own CUDA events around each launch on the launching stream (iq_profile_enable: no host time in
the bracket), 3 warm-ups, mean of 10, a 256 MiB buffer rewritten between repeats (L2 flush)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from interpret_quality_b200 import _lib, ops, synthetic

dev = "cuda:0"
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(reps):
        flush.fill_(1)
        fn()
    rep = _lib.profile_report()
    _lib.profile_enable(False)
    assert len(rep) == 1, rep
    (total_ms, n), = rep.values()
    return total_ms / n


rows = []
def add(name, what, nbytes, ms, note=""):
    gbs = nbytes / (ms * 1e-3) / 1e9
    rows.append("| %s | %s | %.1f MB | %.3f | %.0f | %.2f | %s |" % (name, what, nbytes / 1e6, ms, gbs, gbs / peak, note))


N, R = 1024, 32
data = torch.from_numpy(synthetic.make_cloud(N)).to(dev)
fidx = ops.fps(data, R)
rid = ops.region_id(data, fidx[0].contiguous())
cen = ops.center(data.reshape(-1, 3))
for nperm in (100, 1000):
    orders = torch.from_numpy(synthetic.make_orders(1000, R)[:nperm].copy()).to(dev)
    out = torch.empty((nperm * (R + 1), N, 3), dtype=torch.float32, device=dev)
    ms = timed(lambda: ops.mask_shapley(data.reshape(-1, 3), cen, orders, rid, out=out))
    add("mask_shapley_kernel", "%d permutations x 33 clouds" % nperm, out.numel() * 4, ms, "12*N B written per masked cloud")
ctx = 8192
contexts = torch.from_numpy(np.stack([np.random.RandomState(i).permutation(R)[:16] for i in range(ctx)]).astype(np.int64)).to(dev)
outi = torch.empty((4 * ctx, 3, N), dtype=torch.float32, device=dev)
ms = timed(lambda: ops.mask_interaction(data.reshape(-1, 3), cen, contexts, 1, 2, rid, R, out=outi))
add("mask_interaction_kernel", "%d contexts (m=16) x 4 clouds" % ctx, outi.numel() * 4, ms, "12*N B written per masked cloud")

B = 3300 * 400
logits = torch.randn(B, 10, device=dev)
v = torch.empty(B, device=dev)
ms = timed(lambda: ops.reward(logits, 3, "modified", out=v))
add("reward_kernel", "%d clouds x 10 logits" % B, B * 44, ms, "44 B per cloud; 10 libm expf per row: ALU bound")
nperm = 4000
ordp = torch.from_numpy(np.stack([np.random.RandomState(i).permutation(R) for i in range(nperm)]).astype(np.int64)).to(dev)
vv = torch.randn(nperm * (R + 1), device=dev)
phi = torch.zeros(R, dtype=torch.float64, device=dev)
ms = timed(lambda: ops.shapley_accumulate(vv, ordp, phi))
add("shapley_accumulate_kernel", "%d permutations" % nperm, nperm * ((R + 1) * 4 + R * 8), ms, "ordered float64 sums (same addition order as the reference): serial per region")
P, cx = 300, 1024
lg = torch.randn(P, 4 * cx, 10, device=dev)
ms = timed(lambda: ops.interaction_reduce(lg, 3, "modified"))
add("interaction_reduce_kernel", "%d pairs x %d contexts" % (P, cx), P * cx * (160 + 8), ms, "160 B read + 8 B written per context")

big = torch.from_numpy(np.random.RandomState(0).normal(size=(1 << 22, 3)).astype(np.float32)).to(dev)
fi = torch.arange(R, device=dev, dtype=torch.int64)
ms = timed(lambda: ops.region_id(big, fi))
add("region_id_kernel", "%d points, 32 centres" % big.shape[0], big.shape[0] * 20, ms, "12 B read + 8 B written per point")

for (Bf, npnt) in ((16, 32), (3300, 512)):
    xyz = torch.from_numpy(np.random.RandomState(1).normal(size=(Bf, N, 3)).astype(np.float32)).to(dev)
    ms = timed(lambda: ops.fps(xyz, npnt))
    add("fps_kernel", "%d clouds, %d of 1024 points" % (Bf, npnt), Bf * (N * 12 + npnt * 8), ms,
        "serial chain: %d dependent rounds, %.2f us per round" % (npnt, ms * 1e3 / npnt))

hdr = ["# HBM-side kernels alone (B200, CUDA events, L2 flushed between repeats; peak = %.0f GB/s measured copy bandwidth)" % peak, "",
       "| kernel | workload | algorithmic bytes | ms | GB/s | fraction of peak | note |", "|---|---|---:|---:|---:|---:|---|"]
text = "\n".join(hdr + rows) + "\n"
print(text)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
