"""profiles/r2_traffic.json from the ncu CSV of scripts/r2_profile.sh (dram__bytes_read.sum + dram__bytes_write.sum and
gpu__time_duration.sum of every launch of ONE headline step, scripts/profile_step.py): per kernel family the DRAM bytes
per launch averaged over the step's launches, which is what bench.py quotes as roofline.traffic next to its algorithmic
bytes per launch (same averaging), plus a launch-list table with each family's share of the step.
   python scripts/traffic_from_ncu.py gpurun_out/<tag>_step.csv profiles/r2_traffic.json > profiles/r2_launches.md"""
import csv, io, json, re, sys
from collections import defaultdict

FAMILY = [(r"gemm_tc_kernel<128, 3, 1[,>]", "tc_conv5_pool"), (r"gemm_tc_kernel<128, 3, 0[,>]", "tc_edge_pq"),
          (r"gram_knn_kernel<128, 5, 64[,>]", "tc_gram_knn_c64"), (r"gram_knn_kernel<128, 5, 128[,>]", "tc_gram_knn_c128"),
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
is_f16 = lambda rx: any(re.search(rx + r", (true|1|\(bool\)1)>", n) for n in names.values())
f16_paths = (1 if is_f16(r"gemm_tc_kernel<128, 3, 1") else 0) | (2 if is_f16(r"gemm_tc_kernel<128, 3, 0") else 0) | \
            (4 if is_f16(r"gram_knn_kernel<128, 5, 64") else 0)
out = {"workload": "dgcnn_k20_shapley_100perm_x33clouds_N1024_R32",
       "f16_paths": f16_paths,
       "command": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                  "--profile-from-start off python scripts/profile_step.py",
       "note": "one step (100 permutations x 33 clouds, collapsed), averages over the step's launches of each kernel; kernel "
               "replay: cold caches, serialised -- compare shares, not absolutes",
       "kernels": {k: {"launches": a[0], "dram_bytes_per_launch": a[1] / a[0], "avg_us": a[2] / a[0]} for k, a in agg.items()}}
json.dump(out, open(dst, "w"), indent=1)
tot = sum(v["avg_us"] * v["launches"] for v in out["kernels"].values())
print("# ncu launch list of one headline step (DGCNN, 100 permutations x 33 clouds, collapsed)\n")
print("`%s`\n" % out["command"])
print("| kernel | launches | total us | avg us | share | dram MB / launch |")
print("|---|---:|---:|---:|---:|---:|")
for k, v in sorted(out["kernels"].items(), key=lambda kv: -kv[1]["avg_us"] * kv[1]["launches"]):
    print("| %s | %d | %.1f | %.1f | %.1f%% | %.1f |" % (k, v["launches"], v["avg_us"] * v["launches"], v["avg_us"],
                                                     100 * v["avg_us"] * v["launches"] / tot, v["dram_bytes_per_launch"] / 1e6))
print("\ntotal %.1f us over %d launches (cold-cache, serialised kernel replay: compare shares, not absolutes)"
      % (tot, sum(v["launches"] for v in out["kernels"].values())))
