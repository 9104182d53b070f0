"""Does running two chunk streams side by side hide the launch gaps / pipeline fill of the per-chunk kernels?
Two model instances (own workspaces) on two CUDA streams, half of the 3300 masked clouds each, against one instance
on one stream.  Host-side experiment only; nothing in the library changes."""
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from interpret_quality_b200 import synthetic  # noqa: E402
from interpret_quality_b200.tools import final_util  # noqa: E402

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "dgcnn"
a = types.SimpleNamespace(model=name, k=20, dataset="shapenet", feature_transform=True, device=dev)
sd = synthetic.make_state_dict(name)
m0, m1, m2 = (final_util.build_model(a, sd) for _ in range(3))
rng = np.random.RandomState(0)
base = torch.from_numpy(synthetic.make_cloud(1024)).to(dev)
B = 3256                                                # 22 chunks of 148
x = base.repeat(B, 1, 1) + 0.01 * torch.randn(B, 1024, 3, device=dev)
mask = torch.rand(B, 1024, 1, device=dev) < 0.5
x = torch.where(mask, x.mean(dim=1, keepdim=True), x).contiguous()
out = torch.empty(B, 10, device=dev)
half = B // 2
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def one():
    m0.forward_point_major(x, out=out)


def two():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        m1.forward_point_major(x[:half], out=out[:half])
    with torch.cuda.stream(s2):
        m2.forward_point_major(x[half:], out=out[half:])
    cur.wait_stream(s1); cur.wait_stream(s2)


for fn in (one, two, one, two):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("%s %s: %.2f ms per %d clouds -> %.0f forwards/s" % (name, fn.__name__, ms, B, B / ms * 1e3))
