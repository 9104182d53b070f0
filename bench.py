#!/usr/bin/env python
"""Benchmark of the coalition-evaluation hot path (BASELINE.json metric: masked-coalition forwards/sec,
DGCNN k=20, 1024 points, 32 regions, at 1/2/4/8 B200).

One step = one shap_sampling_all_regions_batch call: 100 seed-replayed permutations x 33 masked
clouds = 3300 forwards through mask -> forward -> reward -> Shapley sums (tools/final_common.py:64-103
of the reference).  Weak scaling: every rank evaluates its own 100-permutation slice of the 1000 saved
permutations and the per-region float64 sums are combined by one NCCL allreduce per step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--model dgcnn]

Prints ONE JSON line on rank 0 (see the contract in the task description / DESIGN.md section Measurement).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R, LBL = 32, 3
METRIC = "masked-coalition forwards/sec"
UNIT = "forwards/s"

# algorithmic FLOPs per forward of the as-written reference models (SURVEY.md section 6, 2*MAC)
MODEL_GFLOP = {"dgcnn": 5.326, "gcnn": 4.789, "pointnet": 0.879, "pointnet2": 7.842, "pointconv": 2.414}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="dgcnn", choices=["dgcnn", "gcnn", "pointnet", "pointnet2", "pointconv"])
    ap.add_argument("--points", type=int, default=1024)
    ap.add_argument("--perms", type=int, default=100, help="permutations per step (NUM_SAMPLES of the reference)")
    ap.add_argument("--chunk", type=int, default=0, help="clouds per internal pass (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    fn = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(fn):
        p = json.load(open(fn))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def config_of(a, n_gpus):
    return {"workload": "%s_k20_shapley_%dperm_x33clouds_N%d_R32" % (a.model, a.perms, a.points),
            "model_class": {"dgcnn": "DGCNN_cls", "gcnn": "GCNN_cls", "pointnet": "PointNetCls", "pointnet2": "PointNet2ClsMsg",
                            "pointconv": "PointConvDensityClsSsg"}[a.model],
            "num_points": a.points, "num_regions": R, "permutations_per_step_per_gpu": a.perms,
            "forwards_per_step_per_gpu": a.perms * (R + 1), "parallelism": "perm-shard x%d" % n_gpus,
            "l2_policy": "256 MiB buffer rewritten between timed steps (flush); per-step working set also exceeds L2",
            "weights": "seeded trained-like random init (interpret_quality_b200/synthetic.py)",
            "chunk_lanes": "library default (2 chunks in flight, PointNet 3); the per-kernel roofline pass runs 1 lane so "
                           "that every kernel is timed alone"}


# ------------------------------------------------------------------------------------------------ CPU arm
def oracle_inputs(a):
    from interpret_quality_b200 import synthetic
    from oracle import geom
    data = synthetic.make_cloud(a.points)
    rid = geom.region_id(data[0], geom.fps(data, R)[0])
    return data, rid, synthetic.make_orders(1000, R), synthetic.make_state_dict(a.model)


def cpu_time_forwards(a, n_perm, batch_perms, repeats=1):
    """Seconds per call of the oracle's shap_sampling_all_regions_batch on the host cores."""
    import torch
    from oracle import coalition, geom
    geom.build()
    torch.set_num_threads(os.cpu_count() or 1)              # torchrun pins OMP_NUM_THREADS=1; use every host core
    data, rid, orders, sd = oracle_inputs(a)
    t = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        coalition.shap_sampling_all_regions_batch(a.model, sd, data, LBL, rid, orders, R, batch_perms, n_perm)
        t.append(time.perf_counter() - t0)
    return t, torch.get_num_threads()


def run_reference(a):
    """--impl reference: the reference's CPU path.  The reference is pure Python/PyTorch and does not exist
    on the GPU box, so this times its restatement oracle/ (kind "port") with all host threads; each step is
    a bounded sample of the workload: 2 permutations x 33 clouds = 66 forwards."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    n_perm = 2
    secs, cores = cpu_time_forwards(a, n_perm, 2, repeats=a.warmup + a.steps)
    timed = secs[a.warmup:]
    total = sum(timed)
    fwd = n_perm * (R + 1) * len(timed)
    value = fwd / total
    sample = "%d steps x %d permutations x 33 clouds (%d forwards) of the same workload, oracle port, %d threads" % (
        len(timed), n_perm, fwd, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of(a, a.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "host_cpus": os.cpu_count()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ GPU arm
# tcgen05 kind::tf32 issue rate measured on this pool's B200 by scripts/microbench/umma_rate.cu (64 clk per 128x128x8
# MMA, SS and TS mode): the ceiling of the EXECUTED tf32 FLOPs of the 3xTF32 kernels
TF32_TFLOPS_MEASURED = 1100.0


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the round's ncu --set full capture of one 148-cloud chunk
# (profiles/r1_v4_full_summary.md; cold-cache kernel replay).  Only quoted for the default workload it was captured on.
NCU_TRAFFIC_BYTES = {"tc_conv5_pool": 295.3e6, "tc_edge_pq": 156.7e6, "tc_gram_knn_c64": 89.7e6, "knn_xyz": 1.9e6}


def kernel_work(a):
    """Algorithmic work per forward (one masked cloud) of the kernel families, for the roofline leg (DESIGN.md
    section 4).  "tensor": (logical FLOPs = 2*MAC of the fp32 product the kernel evaluates, tcgen05 MMAs executed per
    logical MAC: 3 for 3xTF32, 6 for the two-sweep Gram).  "hbm": compulsory bytes (every operand read once, every
    result written once; gathers that hit L2 are not counted)."""
    N, k = a.points, 20
    mac = lambda *terms: 2.0 * sum(r * ci * co for r, ci, co in terms)
    T = lambda flops, mult=3: ("tensor", flops, mult)
    H = lambda nbytes: ("hbm", nbytes, 1)
    w = {"mask_shapley": H(12.0 * N), "reward": H(44.0), "shapley_accumulate": H(4.0 + 8.0 * R / (R + 1))}
    if a.model in ("dgcnn", "gcnn"):
        w["tc_conv5_pool"] = T(mac((N, 512, 1024)))
        w["sgemm_conv5_pool"] = T(mac((N, 512, 1024)), 1)
        couts = (64, 64, 128, 256)
        if a.model == "dgcnn":
            w["sgemm_edge_pq"] = T(mac((N, 3, 128), (N, 64, 128), (N, 64, 256)), 1)
            w["tc_edge_pq"] = T(mac((N, 128, 512)))
            w["sgemm_gram"] = T(mac((N, N, 64), (N, N, 64), (N, N, 128)), 1)
            w["tc_gram_knn_c64"] = T(mac((N, N, 64)) * 2, 6)                  # two layers with 64-wide features
            w["tc_gram_knn_c128"] = T(mac((N, N, 128)), 6)
            w["topk_rows"] = H(3.0 * (4.0 * N * N + 4.0 * N * k))
            # masks (2 bits per column pair) + the feature rows once + neighbour lists, three layers
            w["knn_rerank"] = H(3.0 * (N * N / 4.0 + 4.0 * N * k) + 4.0 * N * (64 + 64 + 128))
        else:
            w["sgemm_edge_pq"] = T(mac((N, 3, 128)), 1)
            w["tc_edge_pq"] = T(mac((N, 64, 128), (N, 64, 256), (N, 128, 512)))
        # P|Q rows read once, neighbour lists, fp32 output + its tf32 hi/lo split + the squared norm
        w["gather_max"] = H(sum(4.0 * N * 2 * c + 4.0 * N * k + 3 * 4.0 * N * c + 4.0 * N for c in couts))
        w["knn_xyz"] = H(12.0 * N + 4.0 * N * k)
    elif a.model == "pointnet":
        w["tc_conv_pool"] = T(mac((N, 128, 1024)) * 3)
        w["tc_conv"] = T(mac((N, 64, 128)) * 3 + mac((N, 64, 64)))
    elif a.model == "pointnet2":
        w["tc_sa_mlp2"] = T(mac((8192, 32, 32), (16384, 64, 64), (65536, 64, 96), (4096, 64, 64),
                                (8192, 128, 128), (16384, 128, 128)))
        w["tc_sa_mlp3_pool"] = T(mac((8192, 32, 64), (16384, 64, 128), (65536, 96, 128), (4096, 64, 128),
                                     (8192, 128, 256), (16384, 128, 256)))
    elif a.model == "pointconv":
        w["tc_sa_mlp2"] = T(mac((16384, 64, 64), (8192, 128, 128), (128, 256, 512)))
        w["tc_sa_mlp3"] = T(mac((16384, 64, 128), (8192, 128, 256), (128, 512, 1024)))
        w["tc_sa_linear"] = T(mac((512, 2048, 128), (128, 4096, 256)))
    return w


def run_b200(a):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the iq_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from interpret_quality_b200 import _lib, build, ops, synthetic
    from interpret_quality_b200.tools import final_common, final_util
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    _lib.load()

    N = a.points
    data_np = synthetic.make_cloud(N)
    data_host = torch.from_numpy(data_np).pin_memory()
    data_dev = data_host.to(dev)
    fps_idx = ops.fps(data_dev, R)
    rid_dev = ops.region_id(data_dev, fps_idx[0].contiguous())
    rid_np = rid_dev.cpu().numpy()
    all_orders = synthetic.make_orders(1000, R)
    lo = (rank * a.perms) % 1000
    if lo + a.perms > 1000:
        lo = 0
    orders_np = np.ascontiguousarray(all_orders[lo:lo + a.perms])
    orders_dev = torch.from_numpy(orders_np).to(dev)
    lbl = torch.tensor([LBL])
    margs = types.SimpleNamespace(model=a.model, k=20, dataset="shapenet", feature_transform=True, device=dev,
                                  num_points=N, num_regions=R, shapley_batch_size=5, num_samples=a.perms,
                                  softmax_type="modified")
    model = final_util.build_model(margs, synthetic.make_state_dict(a.model))
    if a.chunk:
        model.set_chunk(a.chunk)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    fwd_per_step = a.perms * (R + 1)

    def local_step():
        with torch.no_grad():
            return final_common.shapley_partial_sums(model, data_dev, lbl, rid_dev, orders_dev, margs)[0]

    def resident_step():
        phi_sum = local_step()
        if world > 1:
            dist.all_reduce(phi_sum)                         # the one collective of the path: 32 float64 sums
        return phi_sum

    def e2e_step():
        # host buffers in, host result out: H2D of cloud / region ids / permutations and D2H of phi inside
        if world > 1:                                        # every rank passes its own slice, one allreduce inside
            phi_sum, _ = final_common.shapley_partial_sums(model, data_host, lbl, rid_np, orders_np, margs)
            dist.all_reduce(phi_sum)
            return phi_sum.cpu().numpy() / (a.perms * world)
        return final_common.shap_sampling_all_regions_batch(model, data_host, lbl, rid_np, orders_np, margs)[0]

    def timed(step_fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            step_fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        launches0 = _lib.launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for s, e in ev:
            flush.fill_(1)                                   # untimed L2 flush between timed steps
            s.record()
            step_fn()
            e.record()
        torch.cuda.synchronize()
        launches = _lib.launch_count() - launches0
        clocks = sampler.stop() if sampler else None
        if world > 1:
            dist.barrier()
        total_ms = sum(s.elapsed_time(e) for s, e in ev)
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches, clocks

    # correctness gate before timing: phi of the first 4 permutations against the reference's golden vector
    gate = None
    gfile = os.path.join(ROOT, "tests", "golden", "%s.npz" % a.model)
    if rank == 0 and N == 1024 and os.path.exists(gfile):
        g = np.load(gfile)
        ga = types.SimpleNamespace(**vars(margs))
        ga.shapley_batch_size, ga.num_samples = 2, int(g["shapley_nperm"])
        grid = np.load(os.path.join(ROOT, "tests", "golden", "geometry.npz"))["region_id_1024"]
        assert np.array_equal(grid, rid_np), "region ids differ from the reference's golden vector"
        phi, lg = final_common.shap_sampling_all_regions_batch(model, data_host, lbl, rid_np, all_orders, ga)
        e_phi = float(np.abs(phi - g["shapley_phi"]).max() / np.abs(g["shapley_phi"]).max())
        e_lg = float(np.abs(lg.cpu().numpy() - g["shapley_logits"]).max() / np.abs(g["shapley_logits"]).max())
        gate = {"phi_rel_err": e_phi, "logits_rel_err": e_lg, "tolerance": 1e-3}
        if not (e_phi <= 1e-3 and e_lg <= 1e-3):
            raise RuntimeError("parity gate failed: %s" % gate)

    total_ms, launches, clocks = timed(resident_step, a.steps, max(a.warmup, 3), ClockSampler(local))
    value = world * fwd_per_step * a.steps / (total_ms * 1e-3)
    e2e_ms, _, _ = timed(e2e_step, a.steps, 1)
    e2e_value = world * fwd_per_step * a.steps / (e2e_ms * 1e-3)
    h2d = int(data_host.numel() * 4 + orders_np.nbytes + rid_np.nbytes)
    d2h = int(R * 8)

    # per-kernel timing of one more step (CUDA events around every launch, on the launching stream)
    roofline, breakdown = None, None
    if rank == 0:
        lanes = model.get_lanes()
        model.set_lanes(1)                                   # one chunk at a time: every kernel is timed alone on its stream
        local_step()                                         # (re-sizes the workspace outside the profiled step)
        _lib.profile_enable(True)
        local_step()                                         # rank-local: no collective outside the timed region
        rep = _lib.profile_report()
        _lib.profile_enable(False)
        model.set_lanes(lanes)
        pk = peaks()
        work = kernel_work(a)
        tot = sum(ms for ms, _ in rep.values())
        breakdown = {k: {"ms": round(ms, 3), "launches": n, "share": round(ms / tot, 4)} for k, (ms, n) in
                     sorted(rep.items(), key=lambda kv: -kv[1][0])}
        kernels = []
        for name, (ms, n) in sorted(rep.items(), key=lambda kv: -kv[1][0]):
            if name not in work:
                continue
            bound, per_fwd, mult = work[name]
            per_launch = per_fwd * fwd_per_step / n
            dur = ms * 1e-3 / n
            if bound == "tensor":
                ach, peak, unit = per_launch / dur / 1e12, pk["bf16_tflops_sustained"], "TFLOP/s"
            else:
                ach, peak, unit = per_launch / dur / 1e9, pk["hbm_gbs"], "GB/s"
            default_workload = a.model == "dgcnn" and a.points == 1024 and a.perms == 100 and not a.chunk
            k = {"kernel": name, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                 "traffic": NCU_TRAFFIC_BYTES.get(name) if default_workload else None, "avg_launch_ms": dur * 1e3, "launches_per_step": n, "share_of_step": ms / tot,
                 "algorithmic_per_launch": per_launch, "peak_source": pk["source"]}
            if bound == "tensor":
                # the kernels compute fp32 products as `mult` tf32 MMAs each: the executed rate is what the tensor pipe
                # sees, measured against the tf32 issue rate of this GPU (scripts/microbench/umma_rate.cu)
                k["mmas_per_logical_mac"] = mult
                k["executed_tflops"] = ach * mult
                k["executed_frac_of_tf32_peak"] = ach * mult / TF32_TFLOPS_MEASURED if mult > 1 else None
            kernels.append(k)
        if kernels:
            roofline = dict(kernels[0])
            roofline["note"] = ("dominant kernel of the step by device time; peak = MEASURED_PEAKS.json (dense bf16 sustained "
                                "for tensor kernels, copy bandwidth for the others); achieved = algorithmic work / "
                                "CUDA-event duration; tensor kernels evaluate exact-fp32-grade products as 3 (Gram: 6) tf32 "
                                "MMAs per MAC, see executed_tflops; every kernel of the step is listed under `kernels`")
        breakdown = {"by_kernel": breakdown, "kernels": kernels}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        n_perm = 4
        secs, cores = cpu_time_forwards(a, n_perm, 2)
        cpu = {"value": n_perm * (R + 1) / secs[0], "unit": UNIT, "cores": cores, "kind": "port",
               "host_cpus": os.cpu_count(),
               "sample": "%d permutations x 33 clouds (%d forwards) of the same workload, oracle port of the "
                         "reference's torch-CPU path, %d threads" % (n_perm, n_perm * (R + 1), cores)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": total_ms / a.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_of(a, world), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms / a.steps},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "parity_gate": gate,
                "breakdown": breakdown,
                "as_written_tflops": (value * MODEL_GFLOP[a.model] / 1e3) if a.points == 1024 else None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
