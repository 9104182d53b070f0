// extern "C" surface of libiq_b200 (declared in include/iq_b200.h).
#include <string.h>

#include "../../include/iq_b200.h"
#include "model.cuh"

using namespace iq;

struct iq_model {
    std::unique_ptr<Model> impl;
};

extern "C" {

int iq_version(void) { return 100; }
const char *iq_last_error(void) { return last_error(); }
uint64_t iq_launch_count(void) { return g_launch_count.load(std::memory_order_relaxed); }

int iq_debug_reload_env(void)
{
    env_forget();
    return 0;
}

int iq_f16_paths(void) { return f16_paths(); }

int iq_split_f16_host(const float *w, int64_t n, uint16_t *hi, uint16_t *lo, float *scale)
{
    IQ_CHECK(w && hi && lo && scale && n >= 0, "iq_split_f16_host: bad argument");
    std::vector<float> v(w, w + n);
    std::vector<__half> h, l;
    *scale = split_f16_host(v, h, l);
    if (n) {
        memcpy(hi, h.data(), sizeof(uint16_t) * (size_t)n);
        memcpy(lo, l.data(), sizeof(uint16_t) * (size_t)n);
    }
    return 0;
}

int iq_profile_enable(int on)
{
    profile_enable(on != 0);
    return 0;
}

int iq_profile_report(const char **names, double *ms, long long *counts, int cap)
{
    return profile_report(names, ms, counts, cap);
}

int iq_fps(const float *xyz, int64_t B, int64_t N, int64_t npoint, int64_t *idx, void *stream)
{
    IQ_CHECK(xyz && idx, "iq_fps: null pointer");
    return launch_fps(xyz, B, N, npoint, idx, nullptr, nullptr, as_stream(stream));
}

int iq_square_distance3(const float *src, const float *dst, int64_t B, int64_t N, int64_t M, float *out, void *stream)
{
    IQ_CHECK(src && dst && out, "iq_square_distance3: null pointer");
    return launch_square_distance3(src, dst, B, N, M, out, as_stream(stream));
}

int iq_region_id(const float *xyz, const int64_t *fps_index, int64_t N, int64_t R, int64_t *region_id, void *stream)
{
    IQ_CHECK(xyz && fps_index && region_id, "iq_region_id: null pointer");
    return launch_region_id(xyz, fps_index, N, R, region_id, as_stream(stream));
}

int iq_center(const float *xyz, int64_t N, float *center, void *stream)
{
    IQ_CHECK(xyz && center, "iq_center: null pointer");
    return launch_center(xyz, N, center, as_stream(stream));
}

int iq_mask_shapley(const float *data, const float *center, const int64_t *orders, const int64_t *region_id, int64_t bs,
                    int64_t R, int64_t N, float *masked, int in_place, void *stream)
{
    IQ_CHECK(center && orders && region_id && masked, "iq_mask_shapley: null pointer");
    IQ_CHECK(in_place || data, "iq_mask_shapley: data is required unless in_place");
    return launch_mask_shapley(data, center, orders, region_id, bs, R, N, masked, in_place != 0, as_stream(stream));
}

int iq_mask_interaction(const float *data, const float *center, const int64_t *contexts, int64_t ctx, int64_t m,
                        int64_t region_i, int64_t region_j, const int64_t *region_id, int64_t R, int64_t N,
                        int point_major, float *out, void *stream)
{
    IQ_CHECK(data && center && region_id && out, "iq_mask_interaction: null pointer");
    IQ_CHECK(contexts || m == 0 || ctx == 0, "iq_mask_interaction: null contexts");
    return launch_mask_interaction(data, center, contexts, ctx, m, region_i, region_j, region_id, R, N, point_major, out,
                                   as_stream(stream));
}

int iq_mask_interaction_pairs(const float *data, const float *center, const int64_t *pairs, const int64_t *contexts,
                              int64_t P, int64_t ctx, int64_t m, const int64_t *region_id, int64_t R, int64_t N,
                              int point_major, float *out, void *stream)
{
    IQ_CHECK(data && center && region_id && out, "iq_mask_interaction_pairs: null pointer");
    IQ_CHECK(pairs || P == 0, "iq_mask_interaction_pairs: null pairs");
    IQ_CHECK(contexts || m == 0 || ctx == 0 || P == 0, "iq_mask_interaction_pairs: null contexts");
    return launch_mask_interaction_pairs(data, center, pairs, contexts, P, ctx, m, region_id, R, N, point_major, out,
                                         as_stream(stream));
}

int iq_reward(const float *logits, int64_t B, int64_t C, int64_t lbl, int softmax_normal, float *v, void *stream)
{
    IQ_CHECK(logits && v, "iq_reward: null pointer");
    return launch_reward(logits, B, C, lbl, softmax_normal, v, as_stream(stream));
}

int iq_shapley_accumulate(const float *v, const int64_t *orders, int64_t bs, int64_t R, double *phi_sum, void *stream)
{
    IQ_CHECK(v && orders && phi_sum, "iq_shapley_accumulate: null pointer");
    return launch_shapley_accumulate(v, orders, bs, R, phi_sum, as_stream(stream));
}

int iq_interaction_reduce(const float *logits, int64_t P, int64_t ctx, int64_t C, int64_t lbl, int softmax_normal,
                          double *out, void *stream)
{
    IQ_CHECK(logits && out, "iq_interaction_reduce: null pointer");
    return launch_interaction_reduce(logits, P, ctx, C, lbl, softmax_normal, out, as_stream(stream));
}

iq_model *iq_model_create(const char *kind, int n_tensors, const char *const *names, const float *const *host_data,
                          const int64_t *numel, int k, int num_classes)
{
    if (!kind || (n_tensors > 0 && (!names || !host_data || !numel))) {
        set_error("iq_model_create: null argument");
        return nullptr;
    }
    int dev_count = 0;
    if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) {
        set_error("iq_model_create: no CUDA device (this library has no CPU fallback)");
        return nullptr;
    }
    StateDict sd;
    for (int i = 0; i < n_tensors; ++i) {
        std::string key = names[i];
        if (key.compare(0, 7, "module.") == 0) key = key.substr(7);   // tools/final_util.py:253-257
        sd[key] = HostTensor{host_data[i], numel[i]};
    }
    std::string err;
    Model *impl = nullptr;
    const std::string kd = kind;
    if (kd == "dgcnn") impl = create_edgeconv_model(sd, true, k, num_classes, err);
    else if (kd == "gcnn" || kd == "gcnn_adv") impl = create_edgeconv_model(sd, false, k, num_classes, err);
    else if (kd == "pointnet") impl = create_pointnet_model(sd, num_classes, err);
    else if (kd == "pointnet2") impl = create_pointnet2_model(sd, num_classes, err);
    else if (kd == "pointconv") impl = create_pointconv_model(sd, num_classes, err);
    else err = "unknown model kind '" + kd + "'";
    if (!impl) {
        set_error("iq_model_create: " + err);
        return nullptr;
    }
    iq_model *m = new iq_model();
    m->impl.reset(impl);
    return m;
}

void iq_model_destroy(iq_model *m) { delete m; }

int iq_model_set_chunk(iq_model *m, int chunk)
{
    IQ_CHECK(m && chunk >= 1, "iq_model_set_chunk: bad argument");
    m->impl->chunk = chunk;
    return 0;
}

int iq_model_get_chunk(const iq_model *m) { return m ? m->impl->chunk : -1; }

int iq_model_set_lanes(iq_model *m, int lanes)
{
    IQ_CHECK(m && lanes >= 1 && lanes <= Model::MAX_LANES, "iq_model_set_lanes: lanes must be 1..4");
    m->impl->lanes = lanes;
    return 0;
}

int iq_model_get_lanes(const iq_model *m) { return m ? m->impl->lanes : -1; }

int64_t iq_model_workspace_bytes(iq_model *m, int64_t B, int64_t N)
{
    if (!m) { set_error("iq_model_workspace_bytes: null model"); return -1; }
    return m->impl->workspace_bytes(B, N);
}

int iq_model_forward(iq_model *m, const float *x, int point_major, int64_t B, int64_t N, float *logits, void *ws,
                     int64_t ws_bytes, float *trans_feat, int64_t *crt_points, void *stream)
{
    IQ_CHECK(m, "iq_model_forward: null model");
    return m->impl->forward(x, point_major, B, N, logits, ws, ws_bytes, trans_feat, crt_points, nullptr, as_stream(stream));
}

int iq_model_forward_coalitions(iq_model *m, const float *x, int point_major, int64_t B, int64_t N, const float *masked_to,
                                float *logits, void *ws, int64_t ws_bytes, void *stream)
{
    IQ_CHECK(m, "iq_model_forward_coalitions: null model");
    return m->impl->forward(x, point_major, B, N, logits, ws, ws_bytes, nullptr, nullptr, masked_to, as_stream(stream));
}

double iq_model_last_row_fraction(const iq_model *m) { return m ? m->impl->last_row_fraction : -1.0; }

int iq_model_last_buckets(const iq_model *m, int64_t *counts, int cap)
{
    IQ_CHECK(m && (counts || cap == 0), "iq_model_last_buckets: null argument");
    const int n = (int)m->impl->last_buckets.size();
    for (int i = 0; i < n && i < cap; ++i) counts[i] = m->impl->last_buckets[i];
    return n;
}

int iq_ball_query(const float *xyz, const float *new_xyz, int64_t B, int64_t N, int64_t S, double radius, int nsample,
                  int32_t *idx, void *stream)
{
    IQ_CHECK(xyz && new_xyz && idx, "iq_ball_query: null pointer");
    return launch_ball_query(xyz, new_xyz, B, N, S, radius, nsample, idx, as_stream(stream));
}

int iq_knn_xyz(const float *xyz, int64_t B, int64_t N, int k, int32_t *idx, void *stream)
{
    IQ_CHECK(xyz && idx, "iq_knn_xyz: null pointer");
    return launch_knn_xyz(xyz, 1, B, N, k, idx, as_stream(stream));
}

int iq_knn_features(const float *x, int64_t B, int64_t N, int64_t C, int k, int32_t *idx, int32_t *cand_count,
                    void *stream)
{
    IQ_CHECK(x && idx, "iq_knn_features: null pointer");
    IQ_CHECK(knn_features_tc_supported(N, (int)C, k), "iq_knn_features: unsupported shape (N % 128, C in {64,128}, k <= 20)");
    // unit-test path: scratch is allocated here; the model keeps all of it in its workspace
    cudaStream_t st = as_stream(stream);
    const int64_t rows = B * N;
    if (rows == 0) return 0;
    char *buf = nullptr;
    const size_t nf = (size_t)rows * C;
    const size_t bytes = sizeof(float) * (2 * nf + rows) + sizeof(uint32_t) * rows * 2 * (N / 32) + sizeof(int32_t) * rows;
    IQ_CUDA(cudaMalloc(&buf, bytes));
    float *hi = reinterpret_cast<float *>(buf), *lo = hi + nf, *nxx = lo + nf;
    int32_t *cnt = reinterpret_cast<int32_t *>(nxx + rows);
    uint32_t *cand = reinterpret_cast<uint32_t *>(cnt + rows);
    int rc = 0;
    KnnOperands16 h16;
    const bool f16 = (f16_paths() & F16_GRAM) != 0;      // the operand format the DGCNN forward nominates with
    if (f16) {                                            // the fp16 pair of 8 x reuses the tf32 scratch
        h16.hi = reinterpret_cast<__half *>(hi); h16.lo = reinterpret_cast<__half *>(lo); h16.ld = C; h16.scale = 8.0f;
        rc = launch_split_f16(x, rows, (int)C, C, 8.0f, reinterpret_cast<__half *>(hi), reinterpret_cast<__half *>(lo), C, st);
    } else {
        rc = launch_split_tf32(x, rows, (int)C, C, hi, lo, C, st);
    }
    if (!rc) rc = launch_sqnorm_rows(x, rows, (int)C, C, nxx, st);
    if (!rc) rc = launch_knn_features_tc(x, hi, lo, C, (int)C, nxx, 1, B, N, k, cand, cnt, idx, st, f16 ? &h16 : nullptr);
    if (!rc && cand_count && cudaMemcpyAsync(cand_count, cnt, sizeof(int32_t) * rows, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        rc = -2;
    cudaStreamSynchronize(st);
    cudaFree(buf);
    return rc;
}

int iq_region_smoothness_epoch(float *data, const float *data_orig, const int32_t *offsets, const int32_t *members,
                               const float *orient, const float *var_ub, const float *var_lb, double *smooth, int32_t *alive,
                               int32_t *iters, float *last_var, int32_t *stop_flags, int64_t N, int64_t R, int64_t max_region,
                               int mode, int rising, double step, double enum_step, double dist_threshold, double stop_ratio,
                               int max_iteration, int clamp, void *stream)
{
    if (R == 0) return 0;
    IQ_CHECK(data && data_orig && offsets && members && orient && var_ub && var_lb && smooth && alive && iters && last_var &&
                 stop_flags,
             "iq_region_smoothness_epoch: null pointer");
    IQ_CHECK(N >= 2 && max_region <= N, "iq_region_smoothness_epoch: max_region exceeds the cloud");
    return launch_region_smoothness_epoch(data, data_orig, offsets, members, orient, var_ub, var_lb, smooth, alive, iters,
                                          last_var, stop_flags, R, max_region, mode, rising, step, enum_step, dist_threshold,
                                          stop_ratio, max_iteration, clamp, as_stream(stream));
}

int iq_topk_rows(const float *keys, int64_t rows, int64_t N, int64_t ld, int k, int largest, int32_t *idx, void *stream)
{
    IQ_CHECK(keys && idx, "iq_topk_rows: null pointer");
    return launch_topk_rows(keys, rows, N, ld, k, largest, idx, as_stream(stream));
}

int iq_linear(const float *x, const float *w, const float *b, int64_t M, int64_t N, int64_t K, int act, int engine,
              float *y, void *stream)
{
    IQ_CHECK(x && w && y, "iq_linear: null pointer");
    if (engine == 1) {
        // unit-test path: split the operands into scratch, then the tcgen05 kernel
        float *buf = nullptr;
        const size_t na = (size_t)M * K, nb = (size_t)N * K;
        IQ_CUDA(cudaMalloc(&buf, sizeof(float) * 2 * (na + nb)));
        cudaStream_t st = as_stream(stream);
        int rc = launch_split_tf32(x, M, (int)K, K, buf, buf + na, K, st);
        if (!rc) rc = launch_split_tf32(w, N, (int)K, K, buf + 2 * na, buf + 2 * na + nb, K, st);
        TcGemm t;
        t.A_hi = buf; t.A_lo = buf + na; t.lda = K; t.B_hi = buf + 2 * na; t.B_lo = buf + 2 * na + nb; t.ldb = K;
        t.K = (int)K; t.M = (int)M; t.N = (int)N; t.C = y; t.ldc = N; t.bias = b; t.act = act;
        if (!rc) rc = launch_gemm_tc(t, st);
        cudaStreamSynchronize(st);
        cudaFree(buf);
        return rc;
    }
    if (engine == 2) {
        // unit-test path of the kind::f16 form: two-term fp16 splits of x * 8 and w * 256 (powers of two: exact)
        IQ_CHECK(K % 8 == 0, "iq_linear: the fp16 engine needs K % 8 == 0");
        __half *buf = nullptr;
        const size_t na = (size_t)M * K, nb = (size_t)N * K;
        IQ_CUDA(cudaMalloc(&buf, sizeof(__half) * 2 * (na + nb)));
        cudaStream_t st = as_stream(stream);
        int rc = launch_split_f16(x, M, (int)K, K, 8.0f, buf, buf + na, K, st);
        if (!rc) rc = launch_split_f16(w, N, (int)K, K, 256.0f, buf + 2 * na, buf + 2 * na + nb, K, st);
        TcGemm t;
        t.Ah_hi = buf; t.Ah_lo = buf + na; t.lda = K; t.Bh_hi = buf + 2 * na; t.Bh_lo = buf + 2 * na + nb; t.ldb = K;
        t.K = (int)K; t.M = (int)M; t.N = (int)N; t.C = y; t.ldc = N; t.bias = b; t.act = act; t.alpha = 1.0f / 2048.0f;
        if (!rc) rc = launch_gemm_tc(t, st);
        cudaStreamSynchronize(st);
        cudaFree(buf);
        return rc;
    }
    IQ_CHECK(engine == 0, "iq_linear: unknown engine");
    GemmDesc g;
    g.A = x; g.lda = K; g.B = w; g.ldb = K; g.C = y; g.ldc = N;
    g.M = (int)M; g.N = (int)N; g.K = (int)K; g.bias = b; g.act = act;
    return launch_sgemm(g, as_stream(stream));
}

int iq_linear_pool(const float *x, const float *w, const float *b, int64_t clouds, int64_t points, int64_t N, int64_t K,
                   int act, int engine, float *out_max, float *out_mean, int64_t *out_arg, void *stream)
{
    IQ_CHECK(x && w && out_max, "iq_linear_pool: null pointer");
    cudaStream_t st = as_stream(stream);
    const int64_t M = clouds * points;
    if (engine == 1) {
        float *buf = nullptr;
        const size_t na = (size_t)M * K, nb = (size_t)N * K;
        IQ_CUDA(cudaMalloc(&buf, sizeof(float) * 2 * (na + nb)));
        int rc = launch_split_tf32(x, M, (int)K, K, buf, buf + na, K, st);
        if (!rc) rc = launch_split_tf32(w, N, (int)K, K, buf + 2 * na, buf + 2 * na + nb, K, st);
        TcGemm t;
        t.mode = 1;
        t.A_hi = buf + 2 * na; t.A_lo = buf + 2 * na + nb; t.lda = K; t.B_hi = buf; t.B_lo = buf + na; t.ldb = K;
        t.K = (int)K; t.clouds = (int)clouds; t.points = (int)points; t.cout = (int)N;
        t.out_max = out_max; t.out_mean = out_mean; t.out_arg = out_arg; t.ld_out = N; t.bias = b; t.act = act;
        if (!rc) rc = launch_gemm_tc(t, st);
        cudaStreamSynchronize(st);
        cudaFree(buf);
        return rc;
    }
    if (engine == 2) {
        IQ_CHECK(K % 8 == 0, "iq_linear_pool: the fp16 engine needs K % 8 == 0");
        __half *buf = nullptr;
        const size_t na = (size_t)M * K, nb = (size_t)N * K;
        IQ_CUDA(cudaMalloc(&buf, sizeof(__half) * 2 * (na + nb)));
        int rc = launch_split_f16(x, M, (int)K, K, 8.0f, buf, buf + na, K, st);
        if (!rc) rc = launch_split_f16(w, N, (int)K, K, 256.0f, buf + 2 * na, buf + 2 * na + nb, K, st);
        TcGemm t;
        t.mode = 1;
        t.Ah_hi = buf + 2 * na; t.Ah_lo = buf + 2 * na + nb; t.lda = K; t.Bh_hi = buf; t.Bh_lo = buf + na; t.ldb = K;
        t.K = (int)K; t.clouds = (int)clouds; t.points = (int)points; t.cout = (int)N;
        t.out_max = out_max; t.out_mean = out_mean; t.out_arg = out_arg; t.ld_out = N; t.bias = b; t.act = act;
        t.alpha = 1.0f / 2048.0f;
        if (!rc) rc = launch_gemm_tc(t, st);
        cudaStreamSynchronize(st);
        cudaFree(buf);
        return rc;
    }
    IQ_CHECK(engine == 0, "iq_linear_pool: unknown engine");
    IQ_CHECK(points % 128 == 0, "iq_linear_pool: points must be a multiple of 128");
    const int tiles = (int)(points / 128);
    float *pm = nullptr;
    IQ_CUDA(cudaMalloc(&pm, sizeof(float) * 3 * (size_t)clouds * tiles * N));
    float *ps = pm + (size_t)clouds * tiles * N;
    int32_t *pa = reinterpret_cast<int32_t *>(ps + (size_t)clouds * tiles * N);
    GemmDesc g;
    g.A = x; g.lda = K; g.B = w; g.ldb = K; g.M = (int)M; g.N = (int)N; g.K = (int)K; g.bias = b; g.act = act;
    g.pool_max = pm; g.pool_sum = ps; g.pool_arg = pa;
    int rc = launch_sgemm(g, st);
    if (!rc) rc = launch_pool_finish(pm, pa, ps, clouds, tiles, (int)points, (int)N, out_max, N, out_arg, out_mean, N, st);
    cudaStreamSynchronize(st);
    cudaFree(pm);
    return rc;
}

int iq_grouped_mlp_max(const float *U, const float *V, const float *b1, const int32_t *idx, int64_t clouds, int64_t S,
                       int64_t K, int64_t nsrc, int64_t C1, const float *W2, const float *b2, int64_t C2, const float *W3,
                       const float *b3, int64_t C3, float *out, void *stream)
{
    IQ_CHECK(U && V && b1 && idx && W2 && b2 && W3 && b3 && out, "iq_grouped_mlp_max: null pointer");
    // unit-test path of chain_tc.cu: the weights are split into scratch here; the models keep them split
    cudaStream_t st = as_stream(stream);
    float *buf = nullptr;
    const size_t n2 = (size_t)C2 * C1, n3 = (size_t)C3 * C2;
    IQ_CUDA(cudaMalloc(&buf, sizeof(float) * 2 * (n2 + n3)));
    int rc = launch_split_tf32(W2, C2, (int)C1, C1, buf, buf + n2, C1, st);
    if (!rc) rc = launch_split_tf32(W3, C3, (int)C2, C2, buf + 2 * n2, buf + 2 * n2 + n3, C2, st);
    SaChain ch;
    ch.U = U; ch.ldu = C1; ch.V = V; ch.ldv = C1; ch.b1 = b1; ch.idx = idx; ch.rows = clouds * S * K; ch.K = (int)K;
    ch.S = (int)S; ch.nsrc = (int)nsrc; ch.C1 = (int)C1; ch.C2 = (int)C2; ch.C3 = (int)C3;
    ch.W2_hi = buf; ch.W2_lo = buf + n2; ch.b2 = b2; ch.ldw2 = C1;
    ch.W3_hi = buf + 2 * n2; ch.W3_lo = buf + 2 * n2 + n3; ch.b3 = b3; ch.ldw3 = C2;
    ch.out = out; ch.ld_out = C3;
    if (!rc && !sa_chain_supported(ch)) { set_error("iq_grouped_mlp_max: unsupported shape"); rc = -1; }
    if (!rc) rc = launch_sa_chain(ch, st);
    cudaStreamSynchronize(st);
    cudaFree(buf);
    return rc;
}

int iq_collapse_plan(const int32_t *kept, int64_t B, int64_t N, int copies, int32_t *src, int32_t *size, float *extra,
                     int64_t *count)
{
    IQ_CHECK(kept && src && size && extra && count, "iq_collapse_plan: null pointer");
    IQ_CHECK(N >= 128 && N % 128 == 0 && copies >= 1 && B >= 0, "iq_collapse_plan: N must be a positive multiple of 128");
    for (int64_t b = 0; b < B; ++b) IQ_CHECK(kept[b] >= 0 && kept[b] <= N, "iq_collapse_plan: kept count out of range");
    collapse_plan(kept, B, N, copies, src, size, extra, count);
    return 0;
}

int iq_model_set_engine(iq_model *m, int engine)
{
    IQ_CHECK(m && (engine == 0 || engine == 1), "iq_model_set_engine: bad argument");
    m->impl->engine = engine;
    return 0;
}

}  // extern "C"
