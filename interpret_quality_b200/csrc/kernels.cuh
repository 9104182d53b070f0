// Internal launch functions shared between translation units of libiq_b200.
#pragma once
#include "common.cuh"

namespace iq {

// coalition.cu
int launch_mask_shapley(const float *data, const float *center, const int64_t *orders, const int64_t *region_id,
                        int64_t bs, int64_t R, int64_t N, float *out, bool in_place, cudaStream_t st);
int launch_mask_interaction(const float *data, const float *center, const int64_t *contexts, int64_t ctx, int64_t m,
                            int64_t region_i, int64_t region_j, const int64_t *region_id, int64_t R, int64_t N,
                            int point_major, float *out, cudaStream_t st);
int launch_mask_interaction_pairs(const float *data, const float *center, const int64_t *pairs, const int64_t *contexts,
                                  int64_t P, int64_t ctx, int64_t m, const int64_t *region_id, int64_t R, int64_t N,
                                  int point_major, float *out, cudaStream_t st);
int launch_reward(const float *logits, int64_t B, int64_t C, int64_t lbl, int softmax_normal, float *v, cudaStream_t st);
int launch_shapley_accumulate(const float *v, const int64_t *orders, int64_t bs, int64_t R, double *phi_sum,
                              cudaStream_t st);
int launch_interaction_reduce(const float *logits, int64_t P, int64_t ctx, int64_t C, int64_t lbl, int softmax_normal,
                              double *out, cudaStream_t st);

// collapse.cu -- coalition collapse: coincident masked points -> compacted clouds (see the file header)
int launch_collapse_count(const float *x, int point_major, int64_t B, int64_t N, const float *loc, int32_t *kept,
                          cudaStream_t st);
int launch_collapse_compact(const float *x, int point_major, int64_t B, int64_t N, const float *loc, const int32_t *src,
                            const int32_t *row_off, const int32_t *size, float *out, cudaStream_t st);
int launch_scatter_rows(const float *in, const int32_t *dst, int64_t rows, int C, float *out, cudaStream_t st);

// geometry.cu
int launch_fps(const float *xyz, int64_t B, int64_t N, int64_t npoint, int64_t *idx64, int32_t *idx32, float *new_xyz,
               cudaStream_t st);
int launch_region_id(const float *xyz, const int64_t *fps_index, int64_t N, int64_t R, int64_t *region_id,
                     cudaStream_t st);
int launch_square_distance3(const float *src, const float *dst, int64_t B, int64_t N, int64_t M, float *out,
                            cudaStream_t st);
int launch_center(const float *xyz, int64_t N, float *center, cudaStream_t st);

// sgemm.cu
struct GemmDesc {
    const float *A = nullptr;      // (M,K) row-major, leading dimension lda
    const float *B = nullptr;      // (N,K) row-major, leading dimension ldb   ->  C = A * B^T
    float *C = nullptr;            // (M,N) row-major, leading dimension ldc (may be null in pool mode)
    float *C_hi = nullptr, *C_lo = nullptr;   // optional tf32 hi/lo split of the output (same ldc) for a tcgen05 consumer
    int64_t lda = 0, ldb = 0, ldc = 0;
    int64_t strideA = 0, strideB = 0, strideC = 0, strideBias = 0;   // per batch element
    int M = 0, N = 0, K = 0, batch = 1;
    float alpha = 1.0f;
    const float *bias = nullptr;   // per output column n (length N), may be null
    const float *row_bias = nullptr;  // optional (M/row_group, N) addend indexed by m / row_group
    int row_group = 1;
    int64_t ld_row_bias = 0;
    int act = ACT_NONE;
    const char *tag = "sgemm";    // label for the profiling report
    // pooling epilogue: per (m-tile of 128 rows, n) partial max / argmax / sum over the tile's rows
    float *pool_max = nullptr;     // (ceil(M/128), N)
    int32_t *pool_arg = nullptr;   // optional, row index (within the whole M) of the max, lowest on ties
    float *pool_sum = nullptr;     // optional
};
int launch_sgemm(const GemmDesc &g, cudaStream_t st);
int launch_pool_finish(const float *pmax, const int32_t *parg, const float *psum, int64_t groups, int tiles_per_group,
                       int rows_per_group, int N, float *out_max, int64_t ld_max, int64_t *out_arg, float *out_mean,
                       int64_t ld_mean, cudaStream_t st);

// gemm_tc.cu -- tcgen05 3xTF32 GEMM on pre-split (hi, lo) operands
struct TcGemm {
    const float *A_hi = nullptr, *A_lo = nullptr;   // (rows, K) K-major, leading dimension lda
    const float *B_hi = nullptr, *B_lo = nullptr;   // (rows, K) K-major, leading dimension ldb
    // Alternative operand format (when Ah_hi is set the four pointers above are ignored): two-term fp16 splits
    // (common.cuh, split_f16) of A * sa and B * sb, same leading dimensions counted in halves (multiples of 8); the
    // caller folds 1 / (sa * sb) into alpha.  kind::f16 MMAs run at twice the tf32 rate on half the operand bytes.
    // Plain STORE (128-column tiles) and POOL with groups of >= 128 columns only.
    const __half *Ah_hi = nullptr, *Ah_lo = nullptr, *Bh_hi = nullptr, *Bh_lo = nullptr;
    int64_t lda = 0, ldb = 0;
    int K = 0;
    int mode = 0;                 // 0 = STORE: C = act(alpha * A B^T + bias[col]);  1 = POOL (see gemm_tc.cu)
    // STORE
    int M = 0, N = 0;             // A rows, output columns
    int rows_per_batch = 0;       // > 0: batched Gram, B rows taken from the batch the A rows belong to
    float *C = nullptr;           // fp32 output and / or its tf32 hi/lo split (for a tcgen05 consumer)
    float *C_hi = nullptr, *C_lo = nullptr;
    int64_t ldc = 0;
    // POOL: A = weights (cout, K), B = activations (clouds * points, K); reduces over each run of `points`
    // consecutive rows of B (points of a cloud, or the neighbours of one centroid)
    int clouds = 0, points = 0, cout = 0;
    float *out_max = nullptr, *out_mean = nullptr;
    int64_t *out_arg = nullptr;
    int64_t ld_out = 0;
    // POOL, optional: per cloud, the LAST column stands for 1 + pool_extra[cloud] identical points (collapse.cu):
    // mean = (sum + extra * y_last) / (points + extra); the max is unaffected
    const float *pool_extra = nullptr;
    float alpha = 1.0f;
    const float *bias = nullptr;  // STORE: per output column; POOL: per output channel
    int act = ACT_NONE;
    int four_terms = 0;           // STORE: add the Alo*Blo term (4xTF32): the accuracy of an fp32 FMA chain
    const char *tag = "gemm_tc";
    // STORE only, optional: produce the A rows on the fly from a grouping instead of reading A_hi / A_lo,
    //   A[row][c] = act(U[cloud*nsrc + idx[row]][c] - V[row / K][c] + bias[c]),  cloud = row / (K * S)
    // (the first layer of a grouped MLP: launch_group_sub_act fused into its consumer); needs N <= 128
    struct Gather {
        const float *U = nullptr, *V = nullptr, *bias = nullptr;
        const int32_t *idx = nullptr;
        int64_t ldu = 0, ldv = 0;
        int K = 0, S = 0, nsrc = 0, act = ACT_NONE;
    } gather;
};
bool tc_gemm_supported(const TcGemm &g);
int launch_gemm_tc(const TcGemm &g, cudaStream_t st);
int launch_split_tf32(const float *x, int64_t rows, int cols, int64_t ldx, float *hi, float *lo, int64_t ldo,
                      cudaStream_t st);
int launch_split_f16(const float *x, int64_t rows, int cols, int64_t ldx, float scale, __half *hi, __half *lo, int64_t ldo,
                     cudaStream_t st);

// chain_tc.cu -- one grouped shared MLP of a set-abstraction scale in one kernel:
//   out[g] = max over the K rows of group g of relu(relu(relu(U[idx] - V + b1) W2^T + b2) W3^T + b3)
struct SaChain {
    const float *U = nullptr, *V = nullptr, *b1 = nullptr;   // layer 1 as in TcGemm::Gather (U per source point, V per centroid)
    const int32_t *idx = nullptr;                            // (rows) source point of every grouped row
    int64_t ldu = 0, ldv = 0;
    int64_t rows = 0;                                        // clouds * S * K, multiple of 128
    int K = 0, S = 0, nsrc = 0;                              // neighbours per centroid, centroids / source points per cloud
    int C1 = 0, C2 = 0, C3 = 0;                              // layer widths
    const float *W2_hi = nullptr, *W2_lo = nullptr, *b2 = nullptr;   // (C2, C1) tf32 split, leading dimension ldw2
    const float *W3_hi = nullptr, *W3_lo = nullptr, *b3 = nullptr;   // (C3, C2)
    int64_t ldw2 = 0, ldw3 = 0;
    float *out = nullptr;                                    // (rows / K, ld_out)
    int64_t ld_out = 0;
    const char *tag = "tc_sa_chain";
};
bool sa_chain_supported(const SaChain &g);
int launch_sa_chain(const SaChain &g, cudaStream_t st);

// knn_tc.cu -- fused tcgen05 Gram + candidate selection + exact re-rank (feature-space kNN of DGCNN)
constexpr int KNN_CAND_CAP = 64;      // most candidates per row the re-rank handles
bool knn_features_tc_supported(int64_t N, int C, int k);
// x (clouds*N, C) fp32 with leading dimension ld, x_hi / x_lo its tf32 split (same ld); nxx (rows, nxx_parts): the
// squared norm of row j is |sum_p nxx[j][p]| (one negated value, or positive partial sums);
// scratch: masks (rows, 2, N/32) u32, cnt (rows) i32; out idx (rows, k): the SET of the k nearest by (distance, index)
// h16 (optional): nominate on kind::f16 MMAs over the two-term fp16 split of scale * x instead of the tf32 pair
// (x_hi / x_lo are then not read); the exact re-rank reads the fp32 rows either way
struct KnnOperands16 {
    const __half *hi = nullptr, *lo = nullptr;
    int64_t ld = 0;               // leading dimension in halves, multiple of 8
    float scale = 1.0f;
};
int launch_knn_features_tc(const float *x, const float *x_hi, const float *x_lo, int64_t ld, int C, const float *nxx,
                           int nxx_parts, int64_t clouds, int64_t N, int k, uint32_t *masks, int32_t *cnt, int32_t *idx,
                           cudaStream_t st, const KnnOperands16 *h16 = nullptr);

// graph.cu
int launch_knn_xyz(const float *xyz, int point_major, int64_t B, int64_t N, int k, int32_t *idx, cudaStream_t st);
int launch_knn_point(const float *xyz, const float *new_xyz, int64_t B, int64_t N, int64_t S, int k, int32_t *idx,
                     cudaStream_t st);
int launch_topk_rows(const float *keys, int64_t rows, int64_t N, int64_t ld, int k, int largest, int32_t *idx,
                     cudaStream_t st);
int launch_knn_rerank(const float *x, int64_t ld, int C, const int32_t *cand, int64_t rows, int64_t N, int k,
                      int32_t *idx, cudaStream_t st);
int launch_sqnorm_rows(const float *x, int64_t rows, int C, int64_t ld, float *out, cudaStream_t st);
// neg_sqnorm (optional): -|out_i|^2 per row; when sq_parts is given the kernel may instead leave *sq_parts positive
// partial sums per row, (rows, *sq_parts), whose total is |out_i|^2 (buffer of rows * Cout/32 floats)
// h16 (optional): also leave the rows as a two-term fp16 split of out * scale (split_f16) for a kind::f16 consumer;
// `out` may then be null (no fp32 consumer)
struct GatherOut16 {
    __half *hi = nullptr, *lo = nullptr;
    int64_t ld = 0;               // leading dimension in halves, multiple of 8
    float scale = 1.0f;
};
int launch_gather_max(const float *PQ, int64_t ldpq, const int32_t *idx, int64_t B, int64_t N, int k, int Cout,
                      int act, float *out, int64_t ldo, float *neg_sqnorm, int *sq_parts, float *out_hi, float *out_lo,
                      const GatherOut16 *h16, cudaStream_t st);
int launch_xyz_to_point_major(const float *x_cf, int64_t B, int64_t N, float *x_pm, cudaStream_t st);

// grouping.cu
int launch_ball_query(const float *xyz, const float *new_xyz, int64_t B, int64_t N, int64_t S, double radius, int K,
                      int32_t *idx, cudaStream_t st);
int launch_group_sub_act(const float *U, int64_t ldu, const float *V, int64_t ldv, const float *bias,
                         const int32_t *idx, int64_t B, int S, int K, int Nsrc, int C, int act, float *out,
                         float *out_hi, float *out_lo, int64_t ldo, cudaStream_t st);
int launch_group_max(const float *in, int64_t ld, int64_t groups, int K, int C, float *out, int64_t ldo, cudaStream_t st);
int launch_copy_cols(const float *src, int64_t lds, int64_t rows, int cols, float *dst, int64_t ldd, int pad_to,
                     cudaStream_t st);

// smoothness.cu
int launch_region_smoothness_epoch(float *data, const float *data_orig, const int32_t *offsets, const int32_t *members,
                                   const float *orient, const float *var_ub, const float *var_lb, double *smooth,
                                   int32_t *alive, int32_t *iters, float *last_var, int32_t *stop_flags, int64_t R,
                                   int64_t max_region, int mode, int rising, double step, double enum_step, double dist_thr,
                                   double stop_ratio, int max_iteration, int clamp, cudaStream_t st);

}  // namespace iq
