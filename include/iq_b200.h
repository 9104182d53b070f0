/*
 * iq_b200 -- C ABI of the B200-native coalition-evaluation path of
 * ada-shen/Interpret_quality (FPS regions -> coalition masking -> masked forward
 * -> Shapley / interaction reduction).
 *
 * Conventions
 *   - every pointer named *_dev is device memory on the current CUDA device,
 *     owned by the caller (torch's caching allocator in the Python host);
 *     nothing here allocates on the forward path except iq_model_create;
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream); every
 *     call only enqueues work on it, there are no hidden synchronisations;
 *   - tensors are dense, row-major, float32 / int64 exactly like the reference's
 *     torch tensors and numpy arrays;
 *   - return value 0 = ok, negative = error, message via iq_last_error();
 *   - there is no CPU fallback: without a CUDA device every compute call fails.
 *
 * Each entry names the reference function (file:line in ada-shen/Interpret_quality)
 * whose behaviour it reproduces.  INTEGRATION.md shows the ctypes binding.
 */
#ifndef IQ_B200_H
#define IQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct iq_model iq_model;

/* library state */
int iq_version(void);
const char *iq_last_error(void);
/* number of CUDA kernels launched by this library since load (bench.py's gpu_launches) */
uint64_t iq_launch_count(void);

/* The diagnostic environment switches (IQ_LANES, IQ_TC_DBG, IQ_KNN_DBG, IQ_TC_NO_ATM, IQ_TC_NO_GATHER, IQ_NO_COLLAPSE)
 * are read once per process; this makes the next launch read them again (scripts/*_probe.py flip them in-process). */
int iq_debug_reload_env(void);

/* Which DGCNN / GCNN products currently run on kind::f16 tensor-core MMAs over two-term fp16 operand splits instead of
 * 3xTF32 (same 22-bit operand accuracy, twice the MMA rate, half the operand bytes): bit 0 conv5 + pooling, bit 1 the
 * tcgen05 EdgeConv products, bit 2 the Gram kNN nomination.  Library defaults, overridden by IQ_F16_CONV5 /
 * IQ_F16_STORE / IQ_F16_GRAM = 0 | 1 (bench.py labels its roofline lines with it). */
int iq_f16_paths(void);
/* Host half of that operand format (no device needed; CPU tests): the two-term fp16 split the library applies to a weight
 * matrix -- hi = fp16(s w), lo = fp16(s w - hi) as IEEE binary16 bit patterns, s the power of two that brings max|w| into
 * [2^9, 2^10) -- so that (hi + lo) / s reproduces w to 2^-22 relative. */
int iq_split_f16_host(const float *w_host, int64_t n, uint16_t *hi_host, uint16_t *lo_host, float *scale_host);

/* Per-kernel timing for bench.py's roofline leg: while enabled, every kernel launch of this library is
 * bracketed by CUDA events on its stream.  iq_profile_report synchronises the device and returns the
 * number of distinct kernel names, filling up to `cap` (name, total milliseconds, launches) triples. */
int iq_profile_enable(int on);
int iq_profile_report(const char **names, double *ms, long long *counts, int cap);

/* ---- regions ------------------------------------------------------------------------- */

/* farthest_point_sample(xyz, npoint): final_save_fps.py:10-31 (same loop in
 * models/pointnet2.py:45-68, models/pointconv.py:54-77).
 * xyz (B,N,3) f32 -> idx (B,npoint) i64; start index 0, lowest index on ties. */
int iq_fps(const float *xyz_dev, int64_t B, int64_t N, int64_t npoint, int64_t *idx_dev, void *stream);

/* square_distance(src, dst) for 3-d points: tools/final_util.py:134-147.
 * src (B,N,3), dst (B,M,3) -> out (B,N,M) f32. */
int iq_square_distance3(const float *src_dev, const float *dst_dev, int64_t B, int64_t N, int64_t M, float *out_dev,
                        void *stream);

/* cal_region_id(data, fps_index): final_shapley_value.py:20-35.
 * xyz (N,3) f32, fps_index (R) i64 -> region_id (N) i64 (argmin, first minimum). */
int iq_region_id(const float *xyz_dev, const int64_t *fps_index_dev, int64_t N, int64_t R, int64_t *region_id_dev,
                 void *stream);

/* torch.mean(data, dim=1).squeeze(): tools/final_common.py:80.  xyz (N,3) -> center (3). */
int iq_center(const float *xyz_dev, int64_t N, float *center_dev, void *stream);

/* ---- coalition masking --------------------------------------------------------------- */

/* mask_data_batch(masked_data, center, orders, region_id, args): tools/final_common.py:46-61
 * (single permutation: mask_data, final_shapley_value.py:74-88).
 * in_place = 1: masked_dev ((R+1)*bs, N, 3) already holds bs*(R+1) copies of the cloud and only the
 *               masked entries are overwritten (the reference's mutate-and-return contract);
 *               data_dev may be NULL.
 * in_place = 0: the expansion data.expand(...).clone() of tools/final_common.py:88 is fused in and
 *               every element of masked_dev is written from data_dev (N,3). */
int iq_mask_shapley(const float *data_dev, const float *center_dev, const int64_t *orders_dev,
                    const int64_t *region_id_dev, int64_t bs, int64_t R, int64_t N, float *masked_dev, int in_place,
                    void *stream);

/* the 4-clouds-per-context block of compute_order_interaction_logits:
 * final_point_binary_interaction_logits.py:42-56.  data (N,3), contexts (ctx,m) i64 ->
 * out (4*ctx, 3, N) channel-first (point_major = 0, the reference layout) or (4*ctx, N, 3). */
int iq_mask_interaction(const float *data_dev, const float *center_dev, const int64_t *contexts_dev, int64_t ctx,
                        int64_t m, int64_t region_i, int64_t region_j, const int64_t *region_id_dev, int64_t R,
                        int64_t N, int point_major, float *out_dev, void *stream);

/* The same block for ALL pairs of compute_order_interaction_logits at once (the reference loops over them on the host,
 * final_point_binary_interaction_logits.py:37-41): pairs (P,2) i64, contexts (P,ctx,m) i64 ->
 * out (P*ctx*4, 3, N) or (P*ctx*4, N, 3); cloud (p*ctx + k)*4 + v is variant v of context k of pair p. */
int iq_mask_interaction_pairs(const float *data_dev, const float *center_dev, const int64_t *pairs_dev,
                              const int64_t *contexts_dev, int64_t P, int64_t ctx, int64_t m,
                              const int64_t *region_id_dev, int64_t R, int64_t N, int point_major, float *out_dev,
                              void *stream);

/* ---- reward and reductions ----------------------------------------------------------- */

/* get_reward(logits, lbl, args): tools/final_common.py:11-24.
 * softmax_normal = 1 -> "normal" (log_softmax[:, lbl]); 0 -> "modified" (and any other string). */
int iq_reward(const float *logits_dev, int64_t B, int64_t C, int64_t lbl, int softmax_normal, float *v_dev,
              void *stream);

/* the per-permutation loop of shap_sampling_all_regions_batch: tools/final_common.py:93-96.
 * phi_sum[orders[p][r]] += (double)(v[p][r+1] - v[p][r]) for p = 0..bs-1 in order; the caller divides by
 * num_samples (tools/final_common.py:97).  v ((R+1)*bs) f32, orders (bs,R) i64, phi_sum (R) f64. */
int iq_shapley_accumulate(const float *v_dev, const int64_t *orders_dev, int64_t bs, int64_t R, double *phi_sum_dev,
                          void *stream);

/* compute_order_interaction(all_logits, lbl, args): final_cal_interactions.py:14-37.
 * logits (P, 4*ctx, C) f32 -> out (P, ctx) f64 = v[4k] + v[4k+3] - v[4k+1] - v[4k+2]. */
int iq_interaction_reduce(const float *logits_dev, int64_t P, int64_t ctx, int64_t C, int64_t lbl, int softmax_normal,
                          double *out_dev, void *stream);

/* ---- masked forward ------------------------------------------------------------------ */

/* load_model(args): tools/final_util.py:236-262.  kind in {"pointnet","pointnet2","pointconv","dgcnn","gcnn"}
 * selects PointNetCls / PointNet2ClsMsg / PointConvDensityClsSsg / DGCNN_cls / GCNN_cls (models/*.py).
 * The weights arrive as the checkpoint's state dict: n_tensors float32 host arrays with the
 * reference's key names ("module." prefix already stripped).  k = args.k (DGCNN/GCNN).
 * Returns NULL on error. */
iq_model *iq_model_create(const char *kind, int n_tensors, const char *const *names, const float *const *host_data,
                          const int64_t *numel, int k, int num_classes);
void iq_model_destroy(iq_model *m);

/* clouds processed per internal pass (tuning knob; results do not depend on it) */
int iq_model_set_chunk(iq_model *m, int chunk);
int iq_model_get_chunk(const iq_model *m);

/* chunks in flight (1..4): chunks are dealt round-robin over the caller's stream and lanes-1 internal side streams that
 * fork from and join back into it, each lane with its own slice of the workspace (iq_model_workspace_bytes accounts for
 * it).  Default 2; results do not depend on it.  The environment variable IQ_LANES overrides it (diagnostics). */
int iq_model_set_lanes(iq_model *m, int lanes);
int iq_model_get_lanes(const iq_model *m);

/* bytes of scratch iq_model_forward needs for B clouds of N points */
int64_t iq_model_workspace_bytes(iq_model *m, int64_t B, int64_t N);

/* model(data) in eval mode under no_grad: the forward of models/*.py.
 * x_dev: (B,3,N) when point_major = 0 (the nn.Module contract) or (B,N,3) when 1.
 * logits_dev (B, num_classes).  PointNet only (may be NULL): trans_feat_dev (B,64,64),
 * crt_points_dev (B,1024) i64 -- the other two members of PointNetCls.forward's tuple. */
int iq_model_forward(iq_model *m, const float *x_dev, int point_major, int64_t B, int64_t N, float *logits_dev,
                     void *workspace_dev, int64_t workspace_bytes, float *trans_feat_dev, int64_t *crt_points_dev,
                     void *stream);

/* The same forward for a batch of COALITION clouds: clouds produced by the reference's masking rule
 * (tools/final_common.py:56-60 mask_data_batch, final_shapley_value.py:74-88 mask_data,
 * final_point_binary_interaction_logits.py:42-56), i.e. every point of an absent region sits on masked_to_dev
 * (3 floats on the device: `center`, tools/final_common.py:80).  Coincident points have identical rows in every
 * layer, so DGCNN / GCNN / PointNet evaluate each cloud on its kept points plus a few copies of that location
 * (grouped by compacted size in steps of 128 points) -- the logits are those of iq_model_forward up to the
 * summation order of the average pool.  PointNet++ / PointConv (FPS, ball query and density depend on the
 * multiplicities) and clouds without coincident points run exactly as in iq_model_forward.  One host round trip
 * (per-cloud kept counts) inside the call; the stream is synchronised once. */
int iq_model_forward_coalitions(iq_model *m, const float *x_dev, int point_major, int64_t B, int64_t N,
                                const float *masked_to_dev, float *logits_dev, void *workspace_dev,
                                int64_t workspace_bytes, void *stream);
/* rows evaluated / (B*N) of the last iq_model_forward_coalitions call of this model (1 when nothing collapsed) */
double iq_model_last_row_fraction(const iq_model *m);
/* Host half of iq_model_forward_coalitions, exported for tests (HOST pointers, no device work): from the kept-point counts
 * kept[b] of B clouds of N points (N a multiple of 128) and the number of copies of the masking location a cloud must keep
 * (k for DGCNN / GCNN, 1 for PointNet): the clouds ordered by compacted size, largest first -- src[s] = cloud at sorted
 * position s, size[s] = its compacted number of points, extra[s] = weight the average pool owes its last point beyond 1,
 * count[t] (t = 0 .. N/128) = clouds evaluated at 128*t points. */
int iq_collapse_plan(const int32_t *kept, int64_t B, int64_t N, int copies, int32_t *src, int32_t *size, float *extra,
                     int64_t *count);

/* how the last forward of this model was evaluated: counts[t-1] = clouds run at 128*t points (a plain forward has all
 * its clouds in the last entry).  Fills up to `cap` entries, returns the number of entries (ceil(N/128)); bench.py
 * derives the work each kernel really did from it. */
int iq_model_last_buckets(const iq_model *m, int64_t *counts, int cap);

/* ---- building blocks of the forward pass, exported for unit tests ---------------------------- */

/* query_ball_point(radius, nsample, xyz, new_xyz): models/pointnet2.py:70-91 (copy: models/pointconv.py:80-100).
 * xyz (B,N,3), new_xyz (B,S,3) -> idx (B,S,nsample) i32: first nsample indices in index order with d^2 <= r^2
 * (fp32 compare against float(radius^2)), padded with the first hit. */
int iq_ball_query(const float *xyz_dev, const float *new_xyz_dev, int64_t B, int64_t N, int64_t S, double radius,
                  int nsample, int32_t *idx_dev, void *stream);

/* knn(x, k) for 3-d input: models/dgcnn.py:12-18.  xyz (B,N,3) point-major -> idx (B,N,k) i32, the k
 * largest of -|xj|^2 + 2 xi.xj - |xi|^2 per row, unordered, lowest index on ties at the boundary. */
int iq_knn_xyz(const float *xyz_dev, int64_t B, int64_t N, int k, int32_t *idx_dev, void *stream);

/* knn(x, k) on features: models/dgcnn.py:12-18 as used by DGCNN_cls.forward :94-104.  x (B*N, C) point-major,
 * C in {64,128}, N a multiple of 128, k <= 20 -> idx (B,N,k) i32, the SET of the k nearest points of the same cloud under
 * (squared distance, index), in no particular order (the consumer takes a max over it); the distance is sum_c (x_i[c]-x_j[c])^2 evaluated directly.  cand_count_dev (B*N) i32 or
 * NULL receives the size of each row's candidate list (diagnostics: > 64 means the exhaustive path decided the row). */
int iq_knn_features(const float *x_dev, int64_t B, int64_t N, int64_t C, int k, int32_t *idx_dev,
                    int32_t *cand_count_dev, void *stream);

/* torch.topk(keys, k, dim=-1, largest, sorted=False)[1] on a (rows, N) matrix with row stride ld. */
int iq_topk_rows(const float *keys_dev, int64_t rows, int64_t N, int64_t ld, int k, int largest, int32_t *idx_dev,
                 void *stream);

/* y = act(x W^T + b): x (M,K), W (N,K), b (N) or NULL, act 0 none / 1 relu / 2 leaky_relu(0.2).
 * engine 0 = exact fp32 SIMT GEMM, 1 = tcgen05 3xTF32 GEMM, 2 = tcgen05 kind::f16 on two-term fp16 splits of the
 * operands (K % 8 == 0; iq_linear: N % 128 == 0, iq_linear_pool: points % 128 == 0) -- the form conv5 runs in. */
int iq_linear(const float *x_dev, const float *w_dev, const float *b_dev, int64_t M, int64_t N, int64_t K, int act,
              int engine, float *y_dev, void *stream);

/* conv1d(k=1) + BN + activation + global max / average pooling over the points of each cloud, the pattern of
 * models/pointnet.py:35,83 and models/dgcnn.py:108-111: x (clouds*points, K), W (N,K) ->
 * out_max (clouds,N), out_mean (clouds,N) or NULL, out_arg (clouds,N) i64 point index of the max or NULL. */
int iq_linear_pool(const float *x_dev, const float *w_dev, const float *b_dev, int64_t clouds, int64_t points,
                   int64_t N, int64_t K, int act, int engine, float *out_max_dev, float *out_mean_dev,
                   int64_t *out_arg_dev, void *stream);

/* One grouped shared MLP of a PointNet++ set-abstraction scale (models/pointnet2.py:215-232) as a single kernel
 * (csrc/chain_tc.cu), exported for unit tests:
 *   out[g][:] = max over the K rows r of group g of relu(relu(relu(U[cloud*nsrc + idx[r]] - V[g] + b1) W2^T + b2) W3^T + b3)
 * U (clouds*nsrc, C1), V (clouds*S, C1), idx (clouds*S*K) i32, W2 (C2, C1), W3 (C3, C2), out (clouds*S, C3);
 * (C2, C3) in {(32,64), (64,128), (96,128), (128,256)}, K in {16, 32, 64, 128}, clouds*S*K a multiple of 128. */
int iq_grouped_mlp_max(const float *U_dev, const float *V_dev, const float *b1_dev, const int32_t *idx_dev, int64_t clouds,
                       int64_t S, int64_t K, int64_t nsrc, int64_t C1, const float *W2_dev, const float *b2_dev, int64_t C2,
                       const float *W3_dev, const float *b3_dev, int64_t C3, float *out_dev, void *stream);

/* One epoch of the geometry ascent / descent of final_smoothness_center_enum_all.py: update_region :184-243 for every
 * region whose alive flag is set (the loop over regions at :305-321), in one launch, a CTA per region.
 *   data (N,3) in/out, data_orig (N,3): the cloud being pushed and the undisturbed cloud;
 *   region_offsets (R+1) i32, region_members (N) i32: the points of region r are region_members[offsets[r]..offsets[r+1]),
 *     in ascending point order (the order of data[:, region_id == r, :]); every region needs >= 2 points;
 *   orient (R,3,3): rows o1,o2,o3 of cal_principal_orientation :22-45; var_ub / var_lb (R,3): set_var_bound :75-80;
 *   smoothness (R) f64 in/out: the value of the last epoch in, the value measured before the last step out;
 *   alive (R) i32 in/out: if_update; iters (R) i32 out; last_var (R,3) f32 out: var1..3 of the last step (:240);
 *   stop_flags (R) i32 out: bit 0 too many points beyond dist_threshold, bit 1 no gradient, bit 2 max_iteration;
 *   mode 0 linearity / 1 planarity / 2 scattering; rising 1 = "inc", 0 = "dec";
 *   clamp 0 = the reference's behaviour (points beyond dist_threshold are counted, its pull-back :118 is a no-op on a
 *   temporary view), 1 = pull them back onto the dist_threshold sphere as :103-120 describes. */
int iq_region_smoothness_epoch(float *data_dev, const float *data_orig_dev, const int32_t *region_offsets_dev,
                               const int32_t *region_members_dev, const float *orient_dev, const float *var_ub_dev,
                               const float *var_lb_dev, double *smoothness_dev, int32_t *alive_dev, int32_t *iters_dev,
                               float *last_var_dev, int32_t *stop_flags_dev, int64_t N, int64_t R, int64_t max_region,
                               int mode, int rising, double step, double enum_step, double dist_threshold,
                               double stop_ratio, int max_iteration, int clamp, void *stream);

/* GEMM engine of the masked forward: 1 = tcgen05 on two-term operand splits (default: fp16 pairs on kind::f16 for the
 * DGCNN / GCNN products iq_f16_paths() reports, 3xTF32 pairs elsewhere), 0 = exact fp32 SIMT everywhere. */
int iq_model_set_engine(iq_model *m, int engine);

#ifdef __cplusplus
}
#endif
#endif /* IQ_B200_H */
