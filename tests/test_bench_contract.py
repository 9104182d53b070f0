"""bench.py's JSON contract on the CPU side: the reference arm (which runs without a GPU) prints one line with the keys
the driver reads, and the config table covers BASELINE.json's configs."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_valid_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C1", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "masked-coalition forwards/sec" and d["unit"] == "forwards/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("pointnet_") and "model" not in d["config"]
    # the unmodified reference is used whenever it travelled with the repo
    if os.path.exists(os.path.join(ROOT, "baseline", "_ref", "tools", "final_common.py")):
        assert d["cpu_baseline"]["kind"] == "reference"


def test_config_table_covers_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert len(base["configs"]) == 5
    for name in ("headline", "C1", "C2", "C3", "C4", "C5"):
        assert name in bench.CONFIGS
    assert bench.CONFIGS["headline"] == {"kind": "shapley", "model": "dgcnn", "points": 1024, "perms": 100}
    assert bench.CONFIGS["C2"]["model"] == "pointnet2" and bench.CONFIGS["C2"]["perms"] == 1000
    assert bench.CONFIGS["C3"]["points"] == 2048 and bench.CONFIGS["C5"]["model"] == "pointconv"
    # work model: a collapsed batch does less work than the same clouds at full size, kernel by kernel
    full = bench.kernel_work("dgcnn", 20, {1024: 100})
    half = bench.kernel_work("dgcnn", 20, {512: 100})
    for k in full:
        assert half[k][1] <= full[k][1]
    assert abs(half["tc_gram_knn_c64"][1] / full["tc_gram_knn_c64"][1] - 0.25) < 1e-9
    assert abs(half["tc_conv5_pool"][1] / full["tc_conv5_pool"][1] - 0.5) < 1e-9


def test_roofline_object_from_a_recorded_kernel_list():
    """bench.pick_roofline on the per-kernel list of a recorded B200 line (profiles/r2_v7_bench_1gpu.json): the dominant
    kernel's entry with the contract's keys, and the largest tensor-core kernel beside it when the dominant one is not."""
    sys.path.insert(0, ROOT)
    import bench
    line = json.loads(open(os.path.join(ROOT, "profiles", "r2_v7_bench_1gpu.json")).read().strip().splitlines()[-1])
    kernels = line["breakdown"]["kernels"]
    rf = bench.pick_roofline(kernels)
    assert rf["kernel"] == kernels[0]["kernel"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in rf
    assert rf["bound"] in ("hbm", "tensor") and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    if rf["bound"] != "tensor":
        top = rf["top_tensor_kernel"]
        assert top["kernel"] == "tc_conv5_pool" and top["mma_kind"] == "f16" and 0.25 < top["frac"] < 0.45
    assert bench.pick_roofline([]) is None
    json.dumps(rf)
    # kernel_work labels the formats the library reports (iq_f16_paths bit mask)
    w7 = bench.kernel_work("dgcnn", 20, {1024: 10}, 1024, 7)
    w0 = bench.kernel_work("dgcnn", 20, {1024: 10}, 1024, 0)
    assert w7["tc_conv5_pool"][3] == "f16" and w0["tc_conv5_pool"][3] == "tf32" and w7["sgemm_edge_pq"][3] == "fp32-simt"
    assert w7["gather_max"][1] < w0["gather_max"][1]                  # no tf32 pair left to write


def test_kernel_rooflines_reproduce_a_recorded_line():
    """bench.kernel_rooflines fed with the recorded per-kernel times and evaluated-cloud buckets of a B200 run reproduces
    that run's roofline fractions (the arithmetic of the roofline leg, checked without a GPU)."""
    sys.path.insert(0, ROOT)
    import bench
    line = json.loads(open(os.path.join(ROOT, "profiles", "r2_v7_bench_1gpu.json")).read().strip().splitlines()[-1])
    bd = line["breakdown"]
    rep = {k: (v["ms"], v["launches"]) for k, v in bd["by_kernel"].items()}
    buckets = {int(n): c for n, c in bd["evaluated_clouds_by_points"].items()}
    work = bench.kernel_work("dgcnn", 20, buckets, 1024, 7)
    pk = bench.peaks()
    got = bench.kernel_rooflines(rep, work, {}, pk, line["tf32_peak"])
    want = {k["kernel"]: k for k in bd["kernels"]}
    assert [k["kernel"] for k in got][:3] == [k["kernel"] for k in bd["kernels"]][:3]
    for k in got:
        w = want[k["kernel"]]
        ms = rep[k["kernel"]][0]                                       # recorded rounded to 1 us: that much slack on the ratio
        assert abs(k["frac"] - w["frac"]) <= (6e-4 / ms + 1e-3) * w["frac"] + 1e-9, k["kernel"]
        assert k["bound"] == w["bound"] and k["launches_per_step"] == w["launches_per_step"]
    conv5 = next(k for k in got if k["kernel"] == "tc_conv5_pool")
    assert conv5["mma_kind"] == "f16" and abs(conv5["executed_frac_of_bf16_sustained"] - 3 * conv5["frac"]) < 1e-12
    simt = next(k for k in got if k["kernel"] == "sgemm_edge_pq")
    assert simt["mma_kind"] == "fp32-simt" and 0.2 < simt["frac_of_fp32_fma_peak"] < 0.6
    json.dumps(got)
    assert bench.kernel_rooflines({}, work, {}, pk, None) == []
