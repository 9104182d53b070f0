"""GPU probe: kernel epochs vs the oracle for every mode / objective -- per-epoch max coordinate / smoothness
differences, step-count agreement, and the time of one epoch launch against the oracle's CPU loop."""
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from interpret_quality_b200 import final_smoothness_center_enum_all as sm  # noqa: E402
from interpret_quality_b200 import ops, synthetic  # noqa: E402
from oracle import smoothness as osm  # noqa: E402


class Quiet:
    def cprint(self, text):
        pass


def main():
    dev = torch.device("cuda:0")
    data = synthetic.make_cloud(1024)
    rid = np.load(os.path.join(ROOT, "tests", "golden", "geometry.npz"))["region_id_1024"]
    E = 6
    for mode in sm.MODES:
        for objective in ("inc", "dec"):
            a = sm.set_smoothness_args(types.SimpleNamespace(num_regions=32, mode=mode))
            t = torch.from_numpy(data).to(dev)
            geom = sm.RegionGeometry(t, rid, Quiet(), a, dev)
            cur = t.clone().view(-1, 3)
            t0 = time.time()
            want_c, want_s, want_it = osm.run_epochs(data, rid, 32, mode, objective, E)
            t_cpu = time.time() - t0
            ms = []
            for e in range(want_c.shape[0]):
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                iters, last_var, flags = ops.region_smoothness_epoch(
                    cur, geom.data_orig, geom.offsets, geom.members, geom.orient, geom.var_ub, geom.var_lb, geom.smoothness,
                    geom.alive, geom.max_region, mode, objective, a.step, a.enum_step, a.dist_threshold, a.stop_ratio,
                    a.max_iteration)
                ev1.record()
                torch.cuda.synchronize()
                ms.append(ev0.elapsed_time(ev1))
                it = iters.cpu().numpy()
                dc = np.abs(cur.cpu().numpy() - want_c[e, 0]).max()
                ds = np.abs(geom.smoothness.cpu().numpy() - want_s[e]).max()
                print("%-10s %s epoch %d: |dx| %.2e |ds| %.2e steps equal %2d/32 maxdiff %d total %d  %.3f ms"
                      % (mode, objective, e, dc, ds, int((it == want_it[e]).sum()), int(np.abs(it - want_it[e]).max()),
                         int(it.sum()), ms[-1]))
            print("  oracle CPU loop %.2f s for %d epochs; kernel %.3f ms total" % (t_cpu, want_c.shape[0], sum(ms)))


if __name__ == "__main__":
    main()
