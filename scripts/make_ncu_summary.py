"""profiles/r2_ncu_summary.md from the two `scripts/summarize_ncu.py table` outputs of scripts/r2_profile.sh.
   python scripts/make_ncu_summary.py <hot_table.md> <models_table.md>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
clean = lambda t: t.replace("void <unnamed>::", "").replace("<unnamed>::", "").replace("void unnamed>::", "").replace("unnamed>::", "").replace("void ", "")
hot = clean(open(sys.argv[1]).read())
rows = []
for ln in clean(open(sys.argv[2]).read()).splitlines():        # the lighter capture has no DRAM byte counters: drop those columns
    c = [x.strip() for x in ln.strip().strip("|").split("|")]
    if len(c) >= 12:
        rows.append("| " + " | ".join(c[:6] + c[8:11]) + " |")
models = "\n".join(rows)
doc = f"""# ncu evidence, round 2 (B200, `--clock-control none`, kernel replay: cold caches, serialised)

Captured by `scripts/r2_profile.sh` under gpurun on the final code of the round (reports exported to CSV on the box;
`scripts/summarize_ncu.py table` keeps the longest instance of every kernel).  Use these for utilisation and traffic, not
for absolute time; the step shares are in `profiles/r2_launches.md`, the timed numbers in `profiles/r2_v4_bench_*.json`.

## 1. Hot path: coalition kernels at the benchmark's sizes + one 148-cloud DGCNN chunk (N = 1024, k = 20) + the collapse kernels

`ncu --set full --profile-from-start off python scripts/profile_all_kernels.py coalition dgcnn`
(region FPS 1 x 1024 -> 32, region ids, centre, `mask_shapley` 100 permutations = 3300 clouds, `mask_interaction` 8 pairs x
100 contexts x 4 = 3200 clouds, reward 3300, Shapley sums, `interaction_reduce` 300 pairs x 100 contexts)

{hot}
Reading it:

* **tensor kernels** (`gemm_tc_kernel<128,3,1>` = conv5 + BN + LReLU + max/avg pool, `<128,3,0>` = EdgeConv-4 P|Q STORE,
  `gram_knn_kernel`): 72 % / 55 % / 55-63 % tensor-pipe active.  conv5 moves ~635 MB of DRAM per 148-cloud launch for 620 MB
  of algorithmic operand bytes (148 x 1024 x 512 x 4 B x hi,lo): no wasted re-reads, the re-use is served by L2 (84-88 % hit).
* **`mask_shapley`** (40.5 MB written, 13 us = 3.1 TB/s) and **`mask_interaction`** (39.3 MB, 13.5 us): the DRAM write counter
  shows 0.1 MB -- at the benchmark's size the whole output is absorbed by the 126 MB L2 and written back after the kernel, so
  the kernel is bound by launch ramp + the L2 write path, not by HBM; at 1000 permutations (405 MB) it runs at 0.82 of the
  copy bandwidth (`profiles/r1_hbm_kernels.md`).
* **region FPS** (13.7 us): one CTA, 32 dependent argmax rounds of 0.43 us -- the serial chain SURVEY section 8d predicts;
  **region ids** 5.8 us, **reward** 6.3 us, **Shapley sums** 9.6 us, **interaction_reduce** 11 us (4.9 MB read): launch-latency
  sized kernels (one wave or less), a few KB to a few MB each.  None of them is visible in the step (0.1 % together).
* **`knn_rerank_mask_kernel`**: 60-64 % issue-slot utilisation, 77-81 % L2 hit, 7-8 % DRAM: instruction-bound on candidate
  decoding + distance evaluation out of L2 (this capture: with the REDUX selection, 195 -> 175 us at C = 128, 160 -> 131 us
  at C = 64); **`knn_xyz_kernel`**: 71 % issue: selection network, not memory.
* **`gather_max_smem_kernel`**: 47 % of DRAM peak (3.8 TB/s: P|Q rows in, fp32 + tf32 hi/lo rows out) with the k = 20 gathers
  served from shared memory.
* **collapse kernels**: count 4.7 us, compact 7.8 us, scatter 4.1 us on 64 clouds (launch-latency sized; 13 / 32 / 7 us on the
  3300 clouds of a step by CUDA events).

## 2. The other models (one chunk lane each: GCNN 148 clouds, PointNet 330, PointNet++ 33, PointConv 66)

`ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy
--section ComputeWorkloadAnalysis --profile-from-start off python scripts/profile_all_kernels.py gcnn pointnet pointnet2 pointconv`

{models}

* `sa_chain_kernel<C2,C3>` (PointNet++ grouped MLP, layers 1-2-3 + group max in one kernel): 0.4-1.7 % of DRAM peak, 94-98 % L2
  hit -- the grouped activations no longer touch HBM; 42-44 % tensor-pipe active at the large scales under ncu's serialised
  replay (by CUDA events the kernel executes 454 TFLOP/s of tf32 MMAs, 0.76 of the run's sustained cuBLAS TF32 rate).
  Round 1's pair `gemm_tc_kernel<*,*,3>` (gathered-A STORE, 29 % DRAM) + POOL is what `IQ_TC_NO_CHAIN=1` still runs.
* `gemm_tc_kernel<128,6,4>` (PointNet conv3 + max pool with the weight tile parked in TMEM): 78 % tensor-pipe active.
* `sgemm_kernel<1>` at 2.8 ms (grid 8) is PointConv's 16384 -> 1024 Linear on 66 clouds: eight CTAs.  It now splits K sixteen
  ways for batches of <= 1024 clouds (`splitk_reduce_kernel`); the capture predates that change.
* `aggregate_kernel` (PointConv): 44 % of DRAM peak; `fps_kernel` / `ball_query_kernel` / `knn_point_kernel` / `density_kernel`:
  instruction- or latency-bound geometry, 0.1 % DRAM.
"""
open(os.path.join(ROOT, "profiles", "r2_ncu_summary.md"), "w").write(doc)
print("profiles/r2_ncu_summary.md written")
