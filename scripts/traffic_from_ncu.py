"""profiles/r2_traffic.json from the ncu CSV of scripts/r2_profile.sh (dram__bytes_read.sum + dram__bytes_write.sum and
gpu__time_duration.sum of every launch of `bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline`): per kernel
family the DRAM bytes per launch averaged over all captured launches of the full-size steps, which is what bench.py
quotes as roofline.traffic next to its algorithmic bytes per launch (same averaging).
   python scripts/traffic_from_ncu.py gpurun_out/<tag>_traffic.csv profiles/r2_traffic.json"""
import csv, io, json, re, sys
from collections import defaultdict

FAMILY = [(r"gemm_tc_kernel<128, 3, 1>", "tc_conv5_pool"), (r"gemm_tc_kernel<128, 3, 0>", "tc_edge_pq"),
          (r"gram_knn_kernel<128, 5, 64>", "tc_gram_knn_c64"), (r"gram_knn_kernel<128, 5, 128>", "tc_gram_knn_c128"),
          (r"knn_rerank_mask_kernel", "knn_rerank"), (r"gather_max_smem_kernel", "gather_max"), (r"knn_xyz_kernel", "knn_xyz"),
          (r"sgemm_kernel", "sgemm_edge_pq"), (r"mask_shapley_kernel", "mask_shapley"), (r"collapse_count_kernel", "collapse_count"),
          (r"collapse_compact_kernel", "collapse_compact"), (r"reward_kernel", "reward"),
          (r"shapley_accumulate_kernel", "shapley_accumulate")]
src, dst = sys.argv[1], sys.argv[2]
lines = [ln for ln in open(src) if not ln.startswith("==")]
per = defaultdict(dict)
names = {}
for r in csv.DictReader(io.StringIO("".join(lines))):
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if r["Metric Name"].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    else:
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(unit, 1)
    per[r["ID"]][r["Metric Name"]] = v
    names[r["ID"]] = r["Kernel Name"]
agg = defaultdict(lambda: [0, 0.0, 0.0])
for i, m in per.items():
    fam = next((f for rx, f in FAMILY if re.search(rx, names[i])), None)
    if fam is None:
        continue
    a = agg[fam]
    a[0] += 1
    a[1] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    a[2] += m.get("gpu__time_duration.sum", 0.0)
out = {"workload": "dgcnn_k20_shapley_100perm_x33clouds_N1024_R32",
       "command": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                  "python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline",
       "note": "averages over every launch of the run (parity gate of 4 permutations, warm-ups, timed and profiled steps); "
               "kernel replay: cold caches, serialised",
       "kernels": {k: {"launches": a[0], "dram_bytes_per_launch": a[1] / a[0], "avg_us": a[2] / a[0]} for k, a in agg.items()}}
json.dump(out, open(dst, "w"), indent=1)
for k, v in sorted(out["kernels"].items(), key=lambda kv: -kv[1]["avg_us"] * kv[1]["launches"]):
    print("%-20s launches %5d  dram %8.1f MB/launch  %8.1f us/launch" % (k, v["launches"], v["dram_bytes_per_launch"] / 1e6, v["avg_us"]))
