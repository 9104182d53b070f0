// DGCNN_cls / GCNN_cls forward pass (eval mode) on the folded weights.
//
// Reference behaviour restated (never copied): models/dgcnn.py:51-120 (dynamic
// graph recomputed in feature space before every EdgeConv) and :123-194 (graph
// fixed on the input coordinates).  State-dict keys are the reference's
// checkpoint names (conv%d.0.weight, bn%d.*, conv5.0.weight, linear%d.*).
//
// Per chunk of clouds (activations are point-major, (cloud*N + point, channel)):
//   layer l:  idx  = top-k of 2 X X^T - |x_j|^2         (knn_xyz / batched GEMM + topk_rows)
//             PQ   = X [s*Wa ; s*(Wb-Wa)]^T + [0 ; t]    (GEMM)
//             X_l  = lrelu(max_j P[idx] + Q)             (gather_max) -> column slice of the 512-wide concat
//   conv5 + BN + lrelu with fused max / mean pooling over the points, then the 3-layer head.
#include <stdlib.h>

#include "model.cuh"

namespace iq {

namespace {

// Activations that feed kind::f16 products are stored as two-term fp16 splits of x * 8 (common.cuh, split_f16): full
// 22-bit accuracy down to |x| = 2^-6, absolute error 4e-9 below, saturation only beyond |x| = 8188.
constexpr float H_ACT_SCALE = 8.0f;

struct EdgeLayer {
    float *wcat_hi = nullptr, *wcat_lo = nullptr;   // tf32 split of wcat
    __half *wcat_h_hi = nullptr, *wcat_h_lo = nullptr;   // two-term fp16 split of wcat * h_scale
    float h_scale = 1.0f;
    float *wcat = nullptr;       // (2*cout, cin): rows [0,cout) = s*Wa, rows [cout,2cout) = s*(Wb - Wa)
    float *bcat = nullptr;       // (2*cout): [0 ; t]
    int cin = 0, cout = 0, col = 0;
};

class EdgeConvModel : public Model {
public:
    bool dynamic = true;
    int k = 20;
    EdgeLayer layers[4];
    Dense conv5, lin1, lin2, lin3;
    const char *kind() const override { return dynamic ? "dgcnn" : "gcnn"; }

protected:
    int pooled_dim() const override { return 2048; }
    // A k-nearest-neighbour list sees at most k of the coincident copies (collapse.cu); the average pool carries the
    // rest as a weight, which only the tcgen05 pooling epilogue implements.
    int collapse_copies() const override { return engine == 1 ? k : -1; }   // (forward() only collapses when N % 128 == 0)

    int run_head(Workspace &ws, const float *g, int64_t B, float *logits, cudaStream_t st) override
    {
        float *h1 = ws.take<float>(B * 512);
        float *h2 = ws.take<float>(B * 256);
        IQ_CHECK(ws.ok(), "dgcnn/gcnn: workspace too small");
        if (ws.dry) return 0;
        const Dense *head[3] = {&lin1, &lin2, &lin3};
        const float *hin[3] = {g, h1, h2};
        float *hout[3] = {h1, h2, logits};
        for (int i = 0; i < 3; ++i) {
            GemmDesc h;
            h.A = hin[i]; h.lda = head[i]->cin; h.B = head[i]->w; h.ldb = head[i]->cin;
            h.C = hout[i]; h.ldc = head[i]->cout; h.M = (int)B; h.N = head[i]->cout; h.K = head[i]->cin;
            h.bias = head[i]->b; h.act = i < 2 ? ACT_LRELU : ACT_NONE; h.tag = "sgemm_head";
            if (int rc = launch_sgemm(h, st)) return rc;
        }
        return 0;
    }

    int run_body(Workspace &ws, const float *x, int point_major, int64_t Bc, int64_t N, float *g, float *,
                 int64_t *, cudaStream_t st) override
    {
        IQ_CHECK(N <= 2048, "dgcnn/gcnn: num_points must be <= 2048");
        IQ_CHECK(k <= N, "dgcnn/gcnn: k exceeds num_points");
        // The tensor-core kernels tile a cloud in 128-point blocks.  Any other number of points (the reference takes N from
        // the tensor's shape, models/dgcnn.py:12-18) runs the exact fp32 route: SIMT Gram + exact top-k, SIMT products, the
        // generic gather, and conv5 written out and pooled row by row -- the same results as the fp32 engine, at its speed.
        const bool aligned = N % 128 == 0;
        const int64_t rows = Bc * N;
        float *xyz = ws.take<float>(rows * 3);
        int32_t *idx = ws.take<int32_t>(rows * k);
        const bool tc = engine == 1 && aligned;
        // feature-space kNN fused on tcgen05 (knn_tc.cu): the N x N key matrix is never written
        const bool tc_knn = dynamic && tc && knn_features_tc_supported(N, 64, k) && knn_features_tc_supported(N, 128, k);
        uint32_t *cand = tc_knn ? ws.take<uint32_t>(rows * 2 * (N / 32)) : nullptr;
        int32_t *cnt = tc_knn ? ws.take<int32_t>(rows) : nullptr;
        float *feat = ws.take<float>(rows * 512);
        float *nxx = ws.take<float>(rows * 4);          // squared norms of the layer just written, in up to 4 partial sums
        float *pq = ws.take<float>(rows * 512);
        float *feat_hi = tc ? ws.take<float>(rows * 512) : nullptr;
        float *feat_lo = tc ? ws.take<float>(rows * 512) : nullptr;
        // conv5 (and, as an experiment, the tcgen05 EdgeConv products) on kind::f16 MMAs: two-term fp16 operands carry the
        // 22 significand bits of the tf32 pair at twice the MMA rate and half the bytes (gemm_tc.cu).  The buffers are
        // taken whatever the switches say, so the workspace size does not depend on the environment.
        __half *feat_h_hi = tc ? ws.take<__half>(rows * 512) : nullptr;
        __half *feat_h_lo = tc ? ws.take<__half>(rows * 512) : nullptr;
        const int h_paths = f16_paths();
        const bool h_conv5 = tc && conv5.w_h_hi && (h_paths & F16_CONV5);
        const bool h_store = h_conv5 && (h_paths & F16_STORE);
        const bool h_gram = h_conv5 && tc_knn && (h_paths & F16_GRAM);   // the Gram kNN nominates on kind::f16 too
        float *dist = (dynamic && !tc_knn) ? ws.take<float>(Bc * N * N) : nullptr;
        const int tiles = (int)(N / 128);
        float *pmax = aligned ? ws.take<float>(Bc * tiles * 1024) : nullptr;
        float *psum = aligned ? ws.take<float>(Bc * tiles * 1024) : nullptr;
        float *c5out = aligned ? nullptr : ws.take<float>(rows * 1024);
        IQ_CHECK(ws.ok(), "dgcnn/gcnn: workspace too small");
        if (ws.dry) return 0;

        const float *pts = x;
        if (!point_major) {
            if (int rc = launch_xyz_to_point_major(x, Bc, N, xyz, st)) return rc;
            pts = xyz;
        }
        if (int rc = launch_knn_xyz(pts, 1, Bc, N, k, idx, st)) return rc;
        int nxx_parts = 1;
        for (int l = 0; l < 4; ++l) {
            const EdgeLayer &L = layers[l];
            const float *in = l == 0 ? pts : feat + layers[l - 1].col;
            const int64_t ldin = l == 0 ? 3 : 512;
            // A masked region is hundreds of coincident points whose identical kNN rows flip together, so noise in
            // a dynamic-graph decision shows up as 1e-3-sized jumps of the logits ("precision policy").
            // Features upstream of a dynamic kNN therefore stay on the exact fp32 engine, and the tcgen05 Gram keys
            // only nominate candidates; the k neighbours are decided on directly evaluated distances (knn_tc.cu).
            // Downstream-only products (last EdgeConv, conv5) and GCNN run on tcgen05 3xTF32.
            if (l > 0 && tc_knn) {
                KnnOperands16 g16;
                g16.hi = feat_h_hi + layers[l - 1].col; g16.lo = feat_h_lo + layers[l - 1].col; g16.ld = 512; g16.scale = H_ACT_SCALE;
                if (int rc = launch_knn_features_tc(in, feat_hi + layers[l - 1].col, feat_lo + layers[l - 1].col, 512,
                                                    L.cin, nxx, nxx_parts, Bc, N, k, cand, cnt, idx, st, h_gram ? &g16 : nullptr))
                    return rc;
            } else if (l > 0 && dynamic) {
                GemmDesc d;
                d.A = in; d.lda = ldin; d.strideA = N * ldin;
                d.B = in; d.ldb = ldin; d.strideB = N * ldin;
                d.C = dist; d.ldc = N; d.strideC = N * N;
                d.M = (int)N; d.N = (int)N; d.K = L.cin; d.batch = (int)Bc;
                d.alpha = 2.0f; d.bias = nxx; d.strideBias = N; d.tag = "sgemm_gram";
                if (int rc = launch_sgemm(d, st)) return rc;
                if (int rc = launch_topk_rows(dist, rows, N, N, k, 1, idx, st)) return rc;
            }
            const bool tc_all = env_int("IQ_TC_ALL", 0) != 0;            // experiment: EdgeConv 2-3 P|Q on tcgen05 too
            if (l > 0 && tc && (!dynamic || l == 3 || tc_all)) {
                TcGemm p;
                if (h_store && !(dynamic && l < 3)) {
                    p.Ah_hi = feat_h_hi + layers[l - 1].col; p.Ah_lo = feat_h_lo + layers[l - 1].col;
                    p.Bh_hi = L.wcat_h_hi; p.Bh_lo = L.wcat_h_lo;
                    p.alpha = 1.0f / (H_ACT_SCALE * L.h_scale);
                }
                p.A_hi = feat_hi + layers[l - 1].col; p.A_lo = feat_lo + layers[l - 1].col; p.lda = 512;
                p.B_hi = L.wcat_hi; p.B_lo = L.wcat_lo; p.ldb = L.cin;
                p.K = L.cin; p.M = (int)rows; p.N = 2 * L.cout; p.C = pq; p.ldc = 2 * L.cout; p.bias = L.bcat;
                p.tag = "tc_edge_pq";
                p.four_terms = (dynamic && l < 3) ? 1 : 0;      // upstream of a dynamic kNN (IQ_TC_ALL experiment): 4xTF32 (never fp16)
                if (int rc = launch_gemm_tc(p, st)) return rc;
            } else {
                GemmDesc p;
                p.A = in; p.lda = ldin; p.B = L.wcat; p.ldb = L.cin; p.C = pq; p.ldc = 2 * L.cout;
                p.M = (int)rows; p.N = 2 * L.cout; p.K = L.cin; p.bias = L.bcat; p.tag = "sgemm_edge_pq";
                if (int rc = launch_sgemm(p, st)) return rc;
            }
            // who reads this layer's output: fp32 -- the SIMT products, the exact re-rank and the fp32 conv5; the tf32 pair --
            // the Gram kNN and the tf32 tcgen05 products; the fp16 pair -- the kind::f16 products
            const bool need_f32 = !tc || (dynamic && l < 3);
            const bool need_tf32 = tc && (!h_conv5 || (l < 3 && ((tc_knn && !h_gram) || tc_all || !h_store)));
            GatherOut16 h16;
            if (h_conv5) { h16.hi = feat_h_hi + L.col; h16.lo = feat_h_lo + L.col; h16.ld = 512; h16.scale = H_ACT_SCALE; }
            if (int rc = launch_gather_max(pq, 2 * L.cout, idx, Bc, N, k, L.cout, ACT_LRELU, need_f32 ? feat + L.col : nullptr, 512,
                                           (dynamic && l < 3) ? nxx : nullptr, tc_knn ? &nxx_parts : nullptr,
                                           need_tf32 ? feat_hi + L.col : nullptr,
                                           need_tf32 ? feat_lo + L.col : nullptr, h_conv5 ? &h16 : nullptr, st))
                return rc;
        }
        if (tc) {
            TcGemm c5;
            c5.mode = 1;
            c5.A_hi = conv5.w_hi; c5.A_lo = conv5.w_lo; c5.lda = 512;
            c5.B_hi = feat_hi; c5.B_lo = feat_lo; c5.ldb = 512; c5.K = 512;
            if (h_conv5) {
                c5.Ah_hi = conv5.w_h_hi; c5.Ah_lo = conv5.w_h_lo; c5.Bh_hi = feat_h_hi; c5.Bh_lo = feat_h_lo;
                c5.alpha = 1.0f / (H_ACT_SCALE * conv5.h_scale);
            }
            c5.clouds = (int)Bc; c5.points = (int)N; c5.cout = 1024;
            c5.out_max = g; c5.out_mean = g + 1024; c5.ld_out = 2048; c5.bias = conv5.b; c5.act = ACT_LRELU;
            c5.pool_extra = cur_.pool_extra;
            c5.tag = "tc_conv5_pool";
            return launch_gemm_tc(c5, st);
        }
        GemmDesc c5;
        c5.A = feat; c5.lda = 512; c5.B = conv5.w; c5.ldb = 512; c5.M = (int)rows; c5.N = 1024; c5.K = 512;
        c5.bias = conv5.b; c5.act = ACT_LRELU; c5.tag = "sgemm_conv5_pool";
        if (!aligned) {                                  // (B, N, 1024) written once, max / mean over each cloud's N rows
            c5.C = c5out; c5.ldc = 1024;
            if (int rc = launch_sgemm(c5, st)) return rc;
            return launch_pool_finish(c5out, nullptr, c5out, Bc, (int)N, (int)N, 1024, g, 2048, nullptr, g + 1024, 2048, st);
        }
        c5.pool_max = pmax; c5.pool_sum = psum;
        if (int rc = launch_sgemm(c5, st)) return rc;
        if (int rc = launch_pool_finish(pmax, nullptr, psum, Bc, tiles, (int)N, 1024, g, 2048, nullptr, g + 1024, 2048, st))
            return rc;
        return 0;
    }
};

}  // namespace

Model *create_edgeconv_model(const StateDict &sd, bool dynamic_graph, int k, int num_classes, std::string &err)
{
    std::unique_ptr<EdgeConvModel> m(new EdgeConvModel());
    m->dynamic = dynamic_graph;
    m->k = k;
    m->num_classes = num_classes;
    m->chunk = 148;              // 148 clouds x 8 row tiles = 8 full waves of the persistent tensor-core kernels
    const int cin[4] = {3, 64, 64, 128}, cout[4] = {64, 64, 128, 256}, col[4] = {0, 64, 128, 256};
    for (int l = 0; l < 4; ++l) {
        std::vector<float> w, b;
        const std::string n = std::to_string(l + 1);
        if (!fold_dense(sd, "conv" + n + ".0.weight", "", "bn" + n, cout[l], 2 * cin[l], w, b, err)) return nullptr;
        // W = [Wa | Wb] over the 2*cin input channels [x_j - x_i ; x_i]
        std::vector<float> wcat((size_t)2 * cout[l] * cin[l]), bcat((size_t)2 * cout[l], 0.0f);
        for (int o = 0; o < cout[l]; ++o)
            for (int i = 0; i < cin[l]; ++i) {
                const float wa = w[(size_t)o * 2 * cin[l] + i], wb = w[(size_t)o * 2 * cin[l] + cin[l] + i];
                wcat[(size_t)o * cin[l] + i] = wa;
                wcat[(size_t)(cout[l] + o) * cin[l] + i] = (float)((double)wb - (double)wa);
            }
        for (int o = 0; o < cout[l]; ++o) bcat[cout[l] + o] = b[o];
        EdgeLayer &L = m->layers[l];
        L.cin = cin[l]; L.cout = cout[l]; L.col = col[l];
        std::vector<float> whi, wlo;
        split_tf32_host(wcat, whi, wlo);
        std::vector<__half> hhi, hlo;
        L.h_scale = split_f16_host(wcat, hhi, hlo);
        if (m->arena_.upload(wcat, &L.wcat) || m->arena_.upload(bcat, &L.bcat) || m->arena_.upload(whi, &L.wcat_hi) ||
            m->arena_.upload(wlo, &L.wcat_lo) ||
            m->arena_.upload_bytes(hhi.data(), hhi.size() * sizeof(__half), reinterpret_cast<void **>(&L.wcat_h_hi)) ||
            m->arena_.upload_bytes(hlo.data(), hlo.size() * sizeof(__half), reinterpret_cast<void **>(&L.wcat_h_lo))) {
            err = last_error();
            return nullptr;
        }
    }
    struct { Dense *d; const char *w; const char *b; const char *bn; int co, ci; } dense[4] = {
        {&m->conv5, "conv5.0.weight", "", "bn5", 1024, 512},
        {&m->lin1, "linear1.weight", "", "bn6", 512, 2048},
        {&m->lin2, "linear2.weight", "linear2.bias", "bn7", 256, 512},
        {&m->lin3, "linear3.weight", "linear3.bias", "", num_classes, 256}};
    for (auto &e : dense) {
        std::vector<float> w, b;
        if (!fold_dense(sd, e.w, e.b, e.bn, e.co, e.ci, w, b, err)) return nullptr;
        e.d->cout = e.co; e.d->cin = e.ci;
        std::vector<float> whi, wlo;
        split_tf32_host(w, whi, wlo);
        if (m->arena_.upload(w, &e.d->w) || m->arena_.upload(b, &e.d->b) || m->arena_.upload(whi, &e.d->w_hi) ||
            m->arena_.upload(wlo, &e.d->w_lo)) {
            err = last_error();
            return nullptr;
        }
        if (e.d == &m->conv5) {
            std::vector<__half> hhi, hlo;
            e.d->h_scale = split_f16_host(w, hhi, hlo);
            if (m->arena_.upload_bytes(hhi.data(), hhi.size() * sizeof(__half), reinterpret_cast<void **>(&e.d->w_h_hi)) ||
                m->arena_.upload_bytes(hlo.data(), hlo.size() * sizeof(__half), reinterpret_cast<void **>(&e.d->w_h_lo))) {
                err = last_error();
                return nullptr;
            }
        }
    }
    return m.release();
}

}  // namespace iq
