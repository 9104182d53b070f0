#!/usr/bin/env python
"""Benchmark of the coalition-evaluation hot path (BASELINE.json metric: masked-coalition forwards/sec,
DGCNN k=20, 1024 points, 32 regions, at 1/2/4/8 B200).

One step of the headline workload = one shap_sampling_all_regions_batch call: 100 seed-replayed
permutations x 33 masked clouds = 3300 forwards through mask -> forward -> reward -> Shapley sums
(tools/final_common.py:64-103 of the reference).  Weak scaling (default): every rank evaluates its own
100-permutation slice of the 1000 saved permutations and the per-region float64 sums are combined by one NCCL
allreduce per step.  The same line also carries, measured in the same run at the same number of GPUs,
  "strong":  the reference's FIXED-size call (100 permutations in total) split over the ranks,
  "configs": the other BASELINE.json configurations C1..C5 (short runs; --config Cn makes one of them the headline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config headline|C1..C5]
                    [--scaling weak|strong] [--no-extras]

Prints ONE JSON line on rank 0 (see the contract in the task description / DESIGN.md section Measurement).
"""
import argparse
import contextlib
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
ORDERS_M = (0, 1, 2, 3, 6, 9, 12, 15, 18, 21, 24, 27, 30)

# algorithmic FLOPs per forward of the as-written reference models (SURVEY.md section 6, 2*MAC)
MODEL_GFLOP = {"dgcnn": 5.326, "gcnn": 4.789, "pointnet": 0.879, "pointnet2": 7.842, "pointconv": 2.414}
MODEL_CLASS = {"dgcnn": "DGCNN_cls", "gcnn": "GCNN_cls", "pointnet": "PointNetCls", "pointnet2": "PointNet2ClsMsg",
               "pointconv": "PointConvDensityClsSsg"}

# BASELINE.json `configs`, in order.  kind: shapley (perms per step) | interactions | sweep
CONFIGS = {
    "headline": {"kind": "shapley", "model": "dgcnn", "points": 1024, "perms": 100},
    "C1": {"kind": "shapley", "model": "pointnet", "points": 1024, "perms": 100},
    "C2": {"kind": "shapley", "model": "pointnet2", "points": 1024, "perms": 1000, "split": "strong"},
    "C3": {"kind": "shapley", "model": "dgcnn", "points": 2048, "perms": 100},
    "C4": {"kind": "interactions", "model": "dgcnn", "points": 1024, "pairs": 8},
    "C4g": {"kind": "interactions", "model": "gcnn", "points": 1024, "pairs": 8},
    "C5": {"kind": "sweep", "model": "pointconv", "points": 1024, "batches": (64, 256, 1024, 4096)},
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="headline", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: the step's permutations / pairs are a fixed total split over the ranks")
    ap.add_argument("--model", default=None, choices=sorted(MODEL_GFLOP), help="override the config's model")
    ap.add_argument("--points", type=int, default=None)
    ap.add_argument("--perms", type=int, default=None, help="permutations per step (NUM_SAMPLES of the reference)")
    ap.add_argument("--pairs", type=int, default=None, help="interaction configs: region pairs per step")
    ap.add_argument("--chunk", type=int, default=0, help="clouds per internal pass (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the strong-scaling and C1..C5 legs of the line")
    a = ap.parse_args()
    c = dict(CONFIGS[a.config])
    for k in ("model", "points", "perms", "pairs"):
        if getattr(a, k) is not None:
            c[k] = getattr(a, k)
    if a.scaling == "strong":
        c["split"] = "strong"
    a.cfg = c
    return a


def peaks():
    fn = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(fn):
        p = json.load(open(fn))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "sm_max_mhz": float(p.get("sm_max_mhz", 1965.0)), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0, "source": "fallback"}


def workload_name(c, split=None):
    if c["kind"] == "shapley":
        return "%s_k20_shapley_%dperm_x33clouds_N%d_R32" % (c["model"], c["perms"], c["points"])
    if c["kind"] == "interactions":
        return "%s_k20_interactions_13orders_%dpairs_100ctx_N%d_R32" % (c["model"], c["pairs"], c["points"])
    return "%s_forward_sweep_%s_masked_clouds_N%d_R32" % (c["model"], "-".join(map(str, c["batches"])), c["points"])


def config_of(c, n_gpus, split):
    shard = {"shapley": "perm-shard", "interactions": "pair-shard", "sweep": "replicas"}[c["kind"]]
    d = {"workload": workload_name(c), "model_class": MODEL_CLASS[c["model"]], "num_points": c["points"], "num_regions": R,
         "parallelism": "%s x%d" % (shard, n_gpus),
         "split": "%s: %s" % (split, "every rank runs the whole per-GPU workload" if split == "weak" else
                              "the workload is a fixed total split over the ranks"),
         "l2_policy": "256 MiB buffer rewritten between timed steps (flush); per-step working set also exceeds L2",
         "weights": "seeded trained-like random init (interpret_quality_b200/synthetic.py)",
         "coalition_collapse": "DGCNN / GCNN / PointNet evaluate each masked cloud on its kept points + copies of the "
                               "masking location (exact; iq_model_forward_coalitions); IQ_NO_COLLAPSE=1 switches it off",
         "chunk_lanes": "library default (2 chunks in flight, PointNet 3); the per-kernel roofline pass runs 1 lane so "
                        "that every kernel is timed alone"}
    if c["kind"] == "shapley":
        d["permutations_per_step"] = c["perms"]
        d["forwards_per_step"] = c["perms"] * (R + 1)
    return d


# ------------------------------------------------------------------------------------------------ CPU arm
def reference_root():
    """The UNMODIFIED reference, when it travelled with the repo (baseline/install_reference.sh -> baseline/_ref/) or
    is named by $INTERPRET_QUALITY_REF; None otherwise (then the oracle port is timed)."""
    for cand in (os.environ.get("INTERPRET_QUALITY_REF"), os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.exists(os.path.join(cand, "tools", "final_common.py")):
            return cand
    return None


def oracle_inputs(model, points):
    from interpret_quality_b200 import synthetic
    from oracle import geom
    data = synthetic.make_cloud(points)
    rid = geom.region_id(data[0], geom.fps(data, R)[0])
    return data, rid, synthetic.make_orders(1000, R), synthetic.make_state_dict(model)


def cpu_time_forwards(model, points, n_perm, batch_perms, repeats=1):
    """Seconds per call of shap_sampling_all_regions_batch on the host cores: the unmodified reference
    (tools/final_common.py:64-103, torch CPU) when available, else its oracle port.  Returns (secs, threads, kind)."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)              # torchrun pins OMP_NUM_THREADS=1; use every host core
    ref = reference_root()
    if ref is not None:
        from interpret_quality_b200 import synthetic
        sys.path.insert(0, ref)
        with contextlib.redirect_stdout(sys.stderr):        # the reference prints progress lines
            from tools import final_common as ref_common
            from tools import final_util as ref_util
            margs = types.SimpleNamespace(model=model, k=20, dataset="shapenet", feature_transform=True,
                                          device=torch.device("cpu"))
            cls = {"pointnet2": ref_util.PointNet2ClsMsg, "pointnet": ref_util.PointNetCls, "dgcnn": ref_util.DGCNN_cls,
                   "gcnn": ref_util.GCNN_cls, "pointconv": ref_util.PointConvDensityClsSsg}[model]
            net = cls(margs)
            net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synthetic.make_state_dict(model).items()})
            net.eval()
            data, rid, orders, _ = oracle_inputs(model, points)
            args = types.SimpleNamespace(num_points=points, num_regions=R, shapley_batch_size=batch_perms,
                                         num_samples=n_perm, softmax_type="modified", model=model,
                                         device=torch.device("cpu"))
            t = []
            for _ in range(repeats):
                t0 = time.perf_counter()
                with torch.no_grad():
                    ref_common.shap_sampling_all_regions_batch(net, torch.from_numpy(data), torch.tensor([LBL]), rid,
                                                               orders, args)
                t.append(time.perf_counter() - t0)
        return t, torch.get_num_threads(), "reference"
    from oracle import coalition, geom
    geom.build()
    data, rid, orders, sd = oracle_inputs(model, points)
    t = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        coalition.shap_sampling_all_regions_batch(model, sd, data, LBL, rid, orders, R, batch_perms, n_perm)
        t.append(time.perf_counter() - t0)
    return t, torch.get_num_threads(), "port"


def run_reference(a):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores, all threads.
    Each step is a bounded sample of the workload: 2 permutations x 33 clouds = 66 forwards of the config's model
    through shap_sampling_all_regions_batch (interaction / sweep configs: the same model's Shapley call -- the
    per-forward cost is what the metric counts)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = a.cfg
    n_perm = 2
    secs, cores, kind = cpu_time_forwards(c["model"], c["points"], n_perm, 2, repeats=a.warmup + a.steps)
    timed = secs[a.warmup:]
    total = sum(timed)
    fwd = n_perm * (R + 1) * len(timed)
    value = fwd / total
    what = ("the UNMODIFIED reference (baseline/_ref, tools/final_common.py:64-103, torch CPU)" if kind == "reference"
            else "oracle port of the reference's torch-CPU path (baseline/_ref absent)")
    sample = "%d steps x %d permutations x 33 clouds (%d forwards) of the same workload, %s, %d threads" % (
        len(timed), n_perm, fwd, what, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True,
            "scaling": a.cfg.get("split", a.scaling), "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(c, a.gpus, a.cfg.get("split", a.scaling)),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
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
def measure_tf32_peak(dev):
    """Dense TF32 throughput of this GPU the way MEASURED_PEAKS.json measures bf16: torch.matmul on fp32 8192^3 with
    allow_tf32, best of 10 (burst) and back to back for ~1.5 s (sustained).  The ceiling of the EXECUTED tf32 FLOPs of
    the 3xTF32 kernels.  (cuBLAS is used here as the yardstick only; nothing on the hot path calls it.)"""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        x = torch.randn((n, n), device=dev)
        y = torch.randn((n, n), device=dev)
        z = torch.empty((n, n), device=dev)
        flop = 2.0 * n ** 3
        for _ in range(3):
            torch.matmul(x, y, out=z)
        torch.cuda.synchronize(dev)
        best = 0.0
        for _ in range(10):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); torch.matmul(x, y, out=z); e.record()
            torch.cuda.synchronize(dev)
            best = max(best, flop / (s.elapsed_time(e) * 1e-3) / 1e12)
        reps = max(10, int(1.5 / (flop / (best * 1e12))))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            torch.matmul(x, y, out=z)
        e.record()
        torch.cuda.synchronize(dev)
        sustained = reps * flop / (s.elapsed_time(e) * 1e-3) / 1e12
        del x, y, z
        return {"tf32_tflops": best, "tf32_tflops_sustained": sustained,
                "how": "torch.matmul fp32 8192^3, allow_tf32: best of 10 and %d back to back, CUDA events, this run" % reps}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def traffic_table(workload, f16=0):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch (average over the launches of one step) from the
    committed ncu --set full capture profiles/r2_traffic.json; only quoted for the workload it was captured on."""
    fn = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(fn):
        return {}
    t = json.load(open(fn))
    if t.get("workload") != workload:
        return {}
    kernels = dict(t.get("kernels", {}))
    if t.get("f16_paths", 0) != f16:                         # captured with other operand formats: not this run's bytes
        for stale in ("tc_conv5_pool", "gather_max", "tc_edge_pq", "tc_gram_knn_c64", "tc_gram_knn_c128"):
            kernels.pop(stale, None)
    return kernels


def kernel_work(model, k, buckets, points=1024, f16=0):
    """Algorithmic work of one step per kernel family for the roofline leg, from the clouds the
    step really evaluated: buckets = {points per evaluated cloud: clouds} (a collapsed coalition cloud has fewer
    points than the cloud it stands for, csrc/collapse.cu).  "tensor": (logical FLOPs = 2*MAC of the fp32 product the
    kernel evaluates, tcgen05 MMAs executed per logical MAC: 3 for 3xTF32, 6 for the two-sweep Gram).  "hbm":
    compulsory bytes (every operand read once, every result written once; gathers that hit L2 are not counted)."""
    clouds = float(sum(buckets.values()))
    rows = float(sum(n * c for n, c in buckets.items()))
    sq = float(sum(n * n * c for n, c in buckets.items()))
    T = lambda flops, mult=3, kind="tf32": ("tensor", flops, mult, kind)
    H = lambda nbytes: ("hbm", nbytes, 1, None)
    f16_conv5, f16_store, f16_gram = bool(f16 & 1), bool(f16 & 2), bool(f16 & 4)   # csrc/edgeconv_model.cu, iq_f16_paths()
    w = {"reward": H(44.0 * clouds), "shapley_accumulate": H((4.0 + 8.0 * R / (R + 1)) * clouds),
         # coalition expansion at the full cloud size, then the collapse: count reads it, compact reads it and writes the rows kept
         "mask_shapley": H(12.0 * points * clouds), "collapse_count": H(12.0 * points * clouds),
         "collapse_compact": H(12.0 * points * clouds + 12.0 * rows)}
    if model in ("dgcnn", "gcnn"):
        w["tc_conv5_pool"] = T(2.0 * rows * 512 * 1024, 3, "f16" if f16_conv5 else "tf32")
        w["sgemm_conv5_pool"] = T(2.0 * rows * 512 * 1024, 1, "fp32-simt")
        couts = (64, 64, 128, 256)
        if model == "dgcnn":
            w["sgemm_edge_pq"] = T(2.0 * rows * (3 * 128 + 64 * 128 + 64 * 256), 1, "fp32-simt")
            w["tc_edge_pq"] = T(2.0 * rows * 128 * 512, 3, "f16" if f16_store else "tf32")
            w["sgemm_gram"] = T(2.0 * sq * (64 + 64 + 128), 1, "fp32-simt")
            w["tc_gram_knn_c64"] = T(2.0 * sq * 64 * 2, 6, "f16" if f16_gram else "tf32")   # two layers with 64-wide features
            w["tc_gram_knn_c128"] = T(2.0 * sq * 128, 6, "f16" if f16_gram else "tf32")
            w["topk_rows"] = H(3.0 * (4.0 * sq + 4.0 * rows * k))
            # masks (2 bits per column pair) + the feature rows once + neighbour lists, three layers
            w["knn_rerank"] = H(3.0 * (sq / 4.0 + 4.0 * rows * k) + 4.0 * rows * (64 + 64 + 128))
        else:
            w["sgemm_edge_pq"] = T(2.0 * rows * 3 * 128, 1, "fp32-simt")
            w["tc_edge_pq"] = T(2.0 * rows * (64 * 128 + 64 * 256 + 128 * 512), 3, "f16" if f16_store else "tf32")
        # P|Q rows read once, neighbour lists, the squared norm, and the output in the formats its consumers read
        # (edgeconv_model.cu): fp32 (SIMT products + exact re-rank: DGCNN layers 1-3), the tf32 pair (Gram kNN and tf32
        # tcgen05 products), the fp16 pair (kind::f16 products)
        out_bytes = []
        for l, c in enumerate(couts):
            f32 = 4.0 if (model == "dgcnn" and l < 3) else 0.0
            tf32 = 8.0 if (not f16_conv5 or (l < 3 and ((model == "dgcnn" and not f16_gram) or not f16_store))) else 0.0
            out_bytes.append((f32 + tf32 + (4.0 if f16_conv5 else 0.0)) * rows * c)
        w["gather_max"] = H(sum(4.0 * rows * 2 * c + 4.0 * rows * k + 4.0 * rows for c in couts) + sum(out_bytes))
        w["knn_xyz"] = H(12.0 * rows + 4.0 * rows * k)
    elif model == "pointnet":
        w["tc_conv_pool"] = T(2.0 * rows * 128 * 1024 * 3)
        w["tc_conv"] = T(2.0 * rows * (64 * 128 * 3 + 64 * 64))
    elif model == "pointnet2":
        mac = lambda *terms: 2.0 * clouds * sum(r * ci * co for r, ci, co in terms)
        l2 = ((8192, 32, 32), (16384, 64, 64), (65536, 64, 96), (4096, 64, 64), (8192, 128, 128), (16384, 128, 128))
        l3 = ((8192, 32, 64), (16384, 64, 128), (65536, 96, 128), (4096, 64, 128), (8192, 128, 256), (16384, 128, 256))
        w["tc_sa_chain"] = T(mac(*(l2 + l3)))                  # layers 2 + 3 of every scale in one kernel (chain_tc.cu)
        w["tc_sa_mlp12"] = T(mac(*l2))                         # IQ_TC_NO_CHAIN route
        w["tc_sa_mlp3_pool"] = T(mac(*l3))
        w["tc_sa_point"] = T(mac((512, 324, 320)))             # sa2: [features ; xyz] W1cat^T per source point
        w["tc_sa3"] = T(mac((128, 644, 256), (128, 256, 512)))
        w["tc_sa3_pool"] = T(mac((128, 512, 1024)))
    elif model == "pointconv":
        mac = lambda *terms: 2.0 * clouds * sum(r * ci * co for r, ci, co in terms)
        w["tc_sa_mlp2"] = T(mac((16384, 64, 64), (8192, 128, 128), (128, 256, 512)))
        w["tc_sa_mlp3"] = T(mac((16384, 64, 128), (8192, 128, 256), (128, 512, 1024)))
        w["tc_sa_linear"] = T(mac((512, 2048, 128), (128, 4096, 256)))
        w["tc_sa_point"] = T(mac((512, 132, 128), (128, 260, 256)))
        w["tc_sa3_linear"] = T(mac((1, 16384, 1024)))
    return w


FP32_FMA_LANES_PER_SM, SM_COUNT = 128, 148                 # B200: fp32 FMA peak = 148 SMs x 128 lanes x 2 FLOP x SM clock


def kernel_rooflines(rep, work, traffic, pk, tf32):
    """Per-kernel roofline entries of one profiled step, largest device time first.  rep: {kernel: (ms per step, launches)}
    from iq_profile_report; work: kernel_work(); traffic: traffic_table(); pk: peaks(); tf32: measure_tf32_peak()."""
    tot = sum(ms for ms, _ in rep.values())
    kernels = []
    for name, (ms, n) in sorted(rep.items(), key=lambda kv: -kv[1][0]):
        if name not in work or n <= 0 or ms <= 0.0:
            continue
        bound, per_step, mult, mma_kind = work[name]
        per_launch = per_step / n
        dur = ms * 1e-3 / n
        if bound == "tensor":
            ach, peak, unit = per_launch / dur / 1e12, pk["bf16_tflops_sustained"], "TFLOP/s"
        else:
            ach, peak, unit = per_launch / dur / 1e9, pk["hbm_gbs"], "GB/s"
        tr = traffic.get(name)
        k = {"kernel": name, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
             "traffic": tr["dram_bytes_per_launch"] if tr else None,
             "traffic_source": ("profiles/r2_traffic.json: ncu --set full, dram bytes averaged over the %d launches "
                                "of one step" % tr["launches"]) if tr else None,
             "avg_launch_ms": dur * 1e3, "launches_per_step": n, "share_of_step": ms / tot,
             "algorithmic_per_launch": per_launch, "peak_source": pk["source"]}
        if bound == "tensor":
            # fp32-grade products cost `mult` MMAs each: the executed rate is what the tensor pipe sees
            k["mmas_per_logical_mac"] = mult
            k["mma_kind"] = mma_kind
            k["executed_tflops"] = ach * mult
            k["executed_frac_of_tf32_peak"] = k["executed_frac_of_tf32_sustained"] = None
            if mma_kind == "f16":
                # two-term fp16 operands: the MMAs run at the bf16 / fp16 rate, the contract's own yardstick
                k["executed_frac_of_bf16_sustained"] = ach * mult / pk["bf16_tflops_sustained"]
            elif mma_kind == "fp32-simt":
                # CUDA-core GEMM (exact fp32 upstream of a dynamic kNN): its own ceiling is the fp32 FMA rate, not the tensor pipe
                fp32_peak = SM_COUNT * FP32_FMA_LANES_PER_SM * 2 * pk.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
                k["frac_of_fp32_fma_peak"] = ach / fp32_peak
                k["fp32_fma_peak_tflops"] = fp32_peak
            elif mult > 1 and tf32:
                # cuBLAS dense TF32 of this run: best-of-10 ("burst") and back-to-back ("sustained", power-capped clocks)
                k["executed_frac_of_tf32_peak"] = ach * mult / tf32["tf32_tflops"]
                k["executed_frac_of_tf32_sustained"] = ach * mult / tf32["tf32_tflops_sustained"]
        kernels.append(k)
    return kernels


def pick_roofline(kernels):
    """The line's `roofline` object from the per-kernel list (sorted by device time, largest first): the dominant kernel's
    entry, plus -- when that is not a tensor-core kernel -- the tensor-core kernel with the largest share beside it."""
    if not kernels:
        return None
    roofline = dict(kernels[0])
    if roofline["kernel"] == "knn_rerank":
        roofline["bound_detail"] = ("not an HBM kernel: ~25 candidate rows per point are gathered out of L2 (ncu: 80 % L2 hit, 7 % "
                                    "DRAM, 60 % issue-slot utilisation, profiles/r2_ncu_summary.md); the fraction is its "
                                    "compulsory bytes against the copy bandwidth, as the contract asks")
    # the tensor-core kernel with the largest share (round 1's dominant kernel, the one VERDICT.md names), for comparison
    top_tc = next((k for k in kernels if k["bound"] == "tensor" and k.get("mma_kind") in ("f16", "tf32")), None)
    if top_tc is not None and top_tc["kernel"] != roofline["kernel"]:
        roofline["top_tensor_kernel"] = {q: top_tc.get(q) for q in (
            "kernel", "frac", "achieved", "peak", "unit", "share_of_step", "mma_kind", "mmas_per_logical_mac",
            "executed_frac_of_bf16_sustained", "executed_frac_of_tf32_sustained")}
    roofline["note"] = ("dominant kernel of the step by device time; peak = MEASURED_PEAKS.json (dense bf16 sustained for tensor "
                        "kernels, copy bandwidth for the others); achieved = algorithmic work of the clouds AS EVALUATED "
                        "(collapsed coalition clouds, see evaluated_clouds_by_points) / CUDA-event duration, averaged over the "
                        "step's launches; tensor kernels evaluate exact-fp32-grade products as 3 (Gram: 6) MMAs per MAC -- tf32 "
                        "pairs, or two-term fp16 splits on kind::f16 (mma_kind) -- see executed_tflops; every kernel of the step "
                        "is listed under `kernels`")
    return roofline


class Rig:
    """Process-wide state of the GPU arm: device, ranks, timing helper."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device: the iq_b200 hot path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        from interpret_quality_b200 import _lib, build
        if self.rank == 0:
            build.build()
        if self.world > 1:
            dist.barrier()
        _lib.load()
        self.lib = _lib
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)

    def timed(self, step_fn, steps, warmup, sampler=None):
        """(total ms over `steps` steps, max over ranks; kernels launched; clocks).  CUDA events per step on the current
        stream, an untimed L2 flush before every step, barrier + synchronize on both sides."""
        torch, dist = self.torch, self.dist
        for _ in range(warmup):
            step_fn()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        launches0 = self.lib.launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for s, e in ev:
            self.flush.fill_(1)                              # untimed L2 flush between timed steps
            s.record()
            step_fn()
            e.record()
        torch.cuda.synchronize()
        launches = self.lib.launch_count() - launches0
        clocks = sampler.stop() if sampler else None
        if self.world > 1:
            dist.barrier()
        total_ms = sum(s.elapsed_time(e) for s, e in ev)
        t = torch.tensor([total_ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches, clocks


def build_inputs(rig, c):
    torch = rig.torch
    from interpret_quality_b200 import ops, synthetic
    from interpret_quality_b200.tools import final_util
    N = c["points"]
    data_np = synthetic.make_cloud(N)
    data_host = torch.from_numpy(data_np).pin_memory()
    data_dev = data_host.to(rig.dev)
    fps_idx = ops.fps(data_dev, R)
    rid_dev = ops.region_id(data_dev, fps_idx[0].contiguous())
    margs = types.SimpleNamespace(model=c["model"], k=20, dataset="shapenet", feature_transform=True, device=rig.dev,
                                  num_points=N, num_regions=R, shapley_batch_size=5, num_samples=c.get("perms", 100),
                                  softmax_type="modified", interaction_batch_size=25)
    model = final_util.build_model(margs, synthetic.make_state_dict(c["model"]))
    return types.SimpleNamespace(N=N, data_host=data_host, data_dev=data_dev, rid_dev=rid_dev,
                                 rid_np=rid_dev.cpu().numpy(), margs=margs, model=model, lbl=torch.tensor([LBL]))


def shapley_steps(rig, c, inp, split):
    """(resident_step, e2e_step, local_step, forwards per step over all ranks, h2d bytes, d2h bytes)."""
    torch, dist = rig.torch, rig.dist
    from interpret_quality_b200 import synthetic
    from interpret_quality_b200.distributed import shard_range
    from interpret_quality_b200.tools import final_common
    perms = c["perms"]
    all_orders = synthetic.make_orders(1000, R)
    if split == "strong":                                    # a fixed total of `perms` permutations over the ranks
        lo, hi = shard_range(perms, rig.rank, rig.world)
        total_perms = perms
    else:                                                    # every rank its own `perms` permutations
        lo = (rig.rank * perms) % 1000
        if lo + perms > 1000:
            lo = 0
        hi = lo + perms
        total_perms = perms * rig.world
    orders_np = np.ascontiguousarray(all_orders[lo:hi])
    orders_dev = torch.from_numpy(orders_np).to(rig.dev)
    m, margs = inp.model, inp.margs

    def local_step():
        with torch.no_grad():
            return final_common.shapley_partial_sums(m, inp.data_dev, inp.lbl, inp.rid_dev, orders_dev, margs)[0]

    def resident_step():
        phi_sum = local_step()
        if rig.world > 1:
            dist.all_reduce(phi_sum)                         # the one collective of the path: 32 float64 sums
        return phi_sum

    def e2e_step():
        # host buffers in, host result out: H2D of cloud / region ids / permutations and D2H of phi inside
        if rig.world > 1:                                    # every rank passes its own slice, one allreduce inside
            with torch.no_grad():
                phi_sum, _ = final_common.shapley_partial_sums(m, inp.data_host, inp.lbl, inp.rid_np, orders_np, margs)
            dist.all_reduce(phi_sum)
            return phi_sum.cpu().numpy() / total_perms
        margs.num_samples = orders_np.shape[0]
        return final_common.shap_sampling_all_regions_batch(m, inp.data_host, inp.lbl, inp.rid_np, orders_np, margs)[0]

    h2d = int(inp.data_host.numel() * 4 + orders_np.nbytes + inp.rid_np.nbytes)
    return resident_step, e2e_step, local_step, total_perms * (R + 1), h2d, int(R * 8)


def interaction_steps(rig, c, inp, split):
    """C4: all 13 orders x <=100 contexts x 4 coalitions for this rank's pairs; one allreduce of the (P, ctx) float64
    interaction slabs per order (pair-shard)."""
    torch, dist = rig.torch, rig.dist
    from interpret_quality_b200 import final_cal_interactions as fci
    from interpret_quality_b200 import final_point_binary_interaction_logits as fpb
    from interpret_quality_b200 import ops, synthetic
    from interpret_quality_b200.distributed import shard_range
    P_total = c["pairs"] * (1 if split == "strong" else rig.world)
    pairs, ctxs = synthetic.make_pairs_and_contexts(P_total, R, orders_m=ORDERS_M)
    lo, hi = shard_range(P_total, rig.rank, rig.world)
    fw = sum(P_total * ctxs[mm].shape[1] * 4 for mm in ORDERS_M)
    m, margs = inp.model, inp.margs
    soft = "modified"

    def step(data):
        outs = []
        for mm in ORDERS_M:
            lg = fpb.compute_order_interaction_logits(m, data, inp.rid_np, pairs, ctxs[mm], margs, pair_slice=(lo, hi))
            inter = ops.interaction_reduce(lg, LBL, soft)                  # (P_total, ctx) float64; other ranks' rows are
            if rig.world > 1:                                              # built from zero logits: masked before the sum
                keep = torch.zeros_like(inter)
                keep[lo:hi] = 1
                inter = inter * keep
                dist.all_reduce(inter)
            outs.append(inter)
        return outs

    def resident_step():
        return step(inp.data_dev)

    def e2e_step():
        return [o.cpu().numpy() for o in step(inp.data_host)]

    h2d = int(inp.data_host.numel() * 4 + inp.rid_np.nbytes + pairs[lo:hi].nbytes +
              sum(ctxs[mm][lo:hi].astype(np.int64).nbytes for mm in ORDERS_M))
    d2h = int(sum(P_total * ctxs[mm].shape[1] * 8 for mm in ORDERS_M))
    return resident_step, e2e_step, resident_step, fw, h2d, d2h


def sweep_steps(rig, c, inp, split):
    """C5: forwards of B masked clouds for B in `batches` (one step = the whole sweep); ranks are independent replicas."""
    torch = rig.torch
    from interpret_quality_b200 import ops, synthetic
    d = inp.data_dev.reshape(-1, 3)
    cen = ops.center(d)
    Bmax = max(c["batches"])
    nperm = (Bmax + R) // (R + 1)
    orders = torch.from_numpy(synthetic.make_orders(1000, R)[:nperm].copy()).to(rig.dev)
    masked = ops.mask_shapley(d, cen, orders, inp.rid_dev)
    masked_host = masked.cpu().pin_memory()
    m = inp.model

    def resident_step():
        return [m.forward_point_major(masked[:B], masked_to=cen) for B in c["batches"]]

    def e2e_step():
        out = []
        for B in c["batches"]:
            x = masked_host[:B].to(rig.dev, non_blocking=True)
            out.append(m.forward_point_major(x, masked_to=cen).cpu())
        return out

    fw = sum(c["batches"]) * rig.world
    return resident_step, e2e_step, resident_step, fw, int(sum(c["batches"]) * inp.N * 12), int(sum(c["batches"]) * 40)


STEPS = {"shapley": shapley_steps, "interactions": interaction_steps, "sweep": sweep_steps}


def quick_config(rig, name, split=None):
    """A short measurement of one BASELINE config for the `configs` leg: 3 warm-up + 3 timed steps (e2e: 1 + 2)."""
    c = dict(CONFIGS[name])
    split = split or c.get("split", "weak")
    inp = build_inputs(rig, c)
    resident, e2e, _, fw, h2d, d2h = STEPS[c["kind"]](rig, c, inp, split)
    ms, _, _ = rig.timed(resident, 3, 3)
    e_ms, _, _ = rig.timed(e2e, 2, 3 if c["kind"] != "shapley" or c["perms"] <= 100 else 1)
    out = {"workload": workload_name(c), "split": split, "forwards_per_step": fw, "value": fw * 3 / (ms * 1e-3),
           "e2e": fw * 2 / (e_ms * 1e-3), "unit": UNIT, "ms_per_step": ms / 3, "steps": 3, "warmup": 3,
           "rows_evaluated_fraction": inp.model.last_row_fraction(),
           "as_written_tflops": fw * 3 / (ms * 1e-3) * MODEL_GFLOP[c["model"]] / 1e3 if c["points"] == 1024 else None}
    if c["kind"] == "sweep":                                 # per batch size, inputs resident
        from interpret_quality_b200 import ops
        cen = ops.center(inp.data_dev.reshape(-1, 3))
        per = {}
        for B in c["batches"]:
            sub = dict(c, batches=(B,))
            r1, _, _, fw1, _, _ = sweep_steps(rig, sub, inp, split)
            ms1, _, _ = rig.timed(r1, 3, 3)
            per[str(B)] = fw1 * 3 / (ms1 * 1e-3)
        out["forwards_per_s_by_batch"] = per
        out["note"] = "one step = forwards of %s masked clouds back to back; per-batch values: 3 timed steps each" % (
            list(c["batches"]),)
    del inp
    rig.torch.cuda.empty_cache()
    return out


def run_b200(a):
    rig = Rig()
    torch, dist = rig.torch, rig.dist
    from interpret_quality_b200 import synthetic
    from interpret_quality_b200.tools import final_common
    c = a.cfg
    split = c.get("split", a.scaling)
    inp = build_inputs(rig, c)
    model, margs = inp.model, inp.margs
    if a.chunk:
        model.set_chunk(a.chunk)
    resident_step, e2e_step, local_step, fwd_per_step, h2d, d2h = STEPS[c["kind"]](rig, c, inp, split)

    # correctness gate before timing: phi of the first 4 permutations against the reference's golden vector
    gate = None
    gfile = os.path.join(ROOT, "tests", "golden", "%s.npz" % c["model"])
    if rig.rank == 0 and c["points"] == 1024 and os.path.exists(gfile):
        g = np.load(gfile)
        ga = types.SimpleNamespace(**vars(margs))
        ga.shapley_batch_size, ga.num_samples = 2, int(g["shapley_nperm"])
        grid = np.load(os.path.join(ROOT, "tests", "golden", "geometry.npz"))["region_id_1024"]
        assert np.array_equal(grid, inp.rid_np), "region ids differ from the reference's golden vector"
        phi, lg = final_common.shap_sampling_all_regions_batch(model, inp.data_host, inp.lbl, inp.rid_np,
                                                               synthetic.make_orders(1000, R), ga)
        e_phi = float(np.abs(phi - g["shapley_phi"]).max() / np.abs(g["shapley_phi"]).max())
        e_lg = float(np.abs(lg.cpu().numpy() - g["shapley_logits"]).max() / np.abs(g["shapley_logits"]).max())
        gate = {"phi_rel_err": e_phi, "logits_rel_err": e_lg, "tolerance": 1e-3,
                "rows_evaluated_fraction": model.last_row_fraction()}
        if not (e_phi <= 1e-3 and e_lg <= 1e-3):
            raise RuntimeError("parity gate failed: %s" % gate)

    warm = max(a.warmup, 3)
    total_ms, launches, clocks = rig.timed(resident_step, a.steps, warm, ClockSampler(rig.local))
    value = fwd_per_step * a.steps / (total_ms * 1e-3)
    e2e_ms, _, _ = rig.timed(e2e_step, a.steps, 3)
    e2e_value = fwd_per_step * a.steps / (e2e_ms * 1e-3)
    row_fraction = model.last_row_fraction()

    # per-kernel timing of one more step (CUDA events around every launch, on the launching stream)
    roofline, breakdown, tf32 = None, None, None
    if rig.rank == 0:
        lanes = model.get_lanes()
        model.set_lanes(1)                                   # one chunk at a time: every kernel is timed alone on its stream
        local_step()                                         # (re-sizes the workspace outside the profiled step)
        rig.lib.profile_enable(True)
        local_step()                                         # rank-local: no collective outside the timed region
        rep = rig.lib.profile_report()
        rig.lib.profile_enable(False)
        buckets = model.last_buckets()
        model.set_lanes(lanes)
        # after the profiled step: 1.5 s of dense matmul heats the part into its power cap and would slow that step
        tf32 = measure_tf32_peak(rig.dev)
        pk = peaks()
        breakdown = {"by_kernel": None, "kernels": [], "evaluated_clouds_by_points": buckets}
        tot = sum(ms for ms, _ in rep.values())
        breakdown["by_kernel"] = {k: {"ms": round(ms, 3), "launches": n, "share": round(ms / tot, 4)} for k, (ms, n) in
                                  sorted(rep.items(), key=lambda kv: -kv[1][0])}
        if c["kind"] == "shapley":                            # one forward call per profiled step: the buckets describe it
            work = kernel_work(c["model"], 20, buckets, c["points"], rig.lib.f16_paths())
            traffic = traffic_table(workload_name(c), rig.lib.f16_paths())
            kernels = kernel_rooflines(rep, work, traffic, pk, tf32)
            breakdown["kernels"] = kernels
            roofline = pick_roofline(kernels)

    cpu = None
    if rig.rank == 0 and rig.world == 1 and not a.no_cpu_baseline:
        n_perm = 4
        secs, cores, kind = cpu_time_forwards(c["model"], c["points"], n_perm, 2)
        cpu = {"value": n_perm * (R + 1) / secs[0], "unit": UNIT, "cores": cores, "kind": kind,
               "host_cpus": os.cpu_count(),
               "sample": "%d permutations x 33 clouds (%d forwards) of the same model and cloud through "
                         "shap_sampling_all_regions_batch, %s, %d threads" % (
                             n_perm, n_perm * (R + 1),
                             "the UNMODIFIED reference (baseline/_ref, torch CPU)" if kind == "reference"
                             else "oracle port of the reference's torch-CPU path", cores)}

    # the other legs of the line: the fixed-size call split over the ranks, and the other BASELINE configs
    strong, configs = None, None
    if not a.no_extras and a.config == "headline":
        del inp, model
        torch.cuda.empty_cache()
        if split == "weak":
            sc = dict(c)
            sinp = build_inputs(rig, sc)
            s_res, s_e2e, _, s_fw, _, _ = shapley_steps(rig, sc, sinp, "strong")
            s_ms, _, _ = rig.timed(s_res, a.steps, 3)
            strong = {"value": s_fw * a.steps / (s_ms * 1e-3), "unit": UNIT, "ms_per_step": s_ms / a.steps,
                      "steps": a.steps, "permutations_total": sc["perms"], "forwards_per_step": s_fw,
                      "note": "the reference's fixed-size call (tools/final_common.py:64-103, num_samples = %d) with its "
                              "permutations split over the %d rank(s), one allreduce" % (sc["perms"], rig.world)}
            del sinp
            torch.cuda.empty_cache()
        configs = {}
        for name in ("C1", "C2", "C3", "C4", "C4g", "C5"):
            try:
                configs[name] = quick_config(rig, name)
            except Exception as exc:                         # a failed extra must not lose the headline
                configs[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}

    if rig.rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": rig.world, "steps": a.steps,
                "warmup": warm, "ms_per_step": total_ms / a.steps, "higher_is_better": True,
                "scaling": split, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_of(c, rig.world, split), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms / a.steps},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "parity_gate": gate,
                "rows_evaluated_fraction": row_fraction, "tf32_peak": tf32, "strong": strong, "configs": configs,
                "breakdown": breakdown,
                "as_written_tflops": (value * MODEL_GFLOP[c["model"]] / 1e3) if c["points"] == 1024 else None,
                # SURVEY.md section 8(d), path level: forwards/s x FLOPs of the model AS WRITTEN by the reference / measured bf16
                # sustained peak (the implementation executes fewer: EdgeConv restructuring, coalition collapse)
                "as_written_frac_of_bf16_sustained": (value * MODEL_GFLOP[c["model"]] / 1e3 / peaks()["bf16_tflops_sustained"] / rig.world)
                if c["points"] == 1024 else None}
        print(json.dumps(line), flush=True)
    if rig.world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
