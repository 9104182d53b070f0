// PointNet2ClsMsg forward pass (eval mode) on the folded weights.
//
// Reference behaviour restated (never copied): models/pointnet2.py:180-240 (multi-scale set abstraction: FPS,
// three ball queries, three shared MLPs, max over the group), :138-178 with group_all (sa3), :244-276 (classifier).
// Conv2d layers carry a bias; grouped input channel order is [features ; xyz - centre] in the MSG layers and
// [xyz ; features] in the group-all layer.
//
// Per chunk of clouds:
//   sa:  new_xyz = FPS(xyz)                                              geometry.cu (exact argmax recipe)
//        U = [f ; x] W1cat^T per point, V = c W1x^T per centroid         fp32 SIMT GEMMs (first layers of the 3 scales at once)
//        per scale: idx = ball_query;  H1 = relu(U[idx] - V + b1)        grouping.cu, tf32 hi/lo
//                   H2 = relu(H1 W2^T + b2)                              tcgen05 STORE -> hi/lo
//                   out = max_group relu(H2 W3^T + b3)                   tcgen05 POOL over the K neighbours
//   sa3: 643 -> 256 -> 512 -> 1024 on the 128 remaining points, max over them (fp32 SIMT + pooling epilogue)
// FPS and ball queries depend on coordinates only, so the MLPs can all run in 3xTF32 without touching any
// discrete decision.
#include <stdlib.h>

#include "model.cuh"

namespace iq {

namespace {

struct Scale {
    double radius = 0;
    int K = 0, c1 = 0, c2 = 0, c3 = 0, col1 = 0, col3 = 0;   // col1: column of this scale in U/V, col3: in the concat output
    float *b1 = nullptr;
    Dense l2, l3;
};

struct SaMsg {
    int npoint = 0, cin_feat = 0, c1_total = 0, cout_total = 0;
    float *w1cat = nullptr;      // (c1_total, ldw): first layers of the three scales over [features ; xyz ; pad]
    float *w1cat_hi = nullptr, *w1cat_lo = nullptr;   // its tf32 split (tcgen05 path of the per-point product)
    float *w1x = nullptr;        // (c1_total, 3): their xyz columns (per-centroid term)
    int ldw = 0;
    Scale sc[3];
};

class PointNet2Model : public Model {
public:
    SaMsg sa1, sa2;
    Dense s3a, s3b, s3c, fc1, fc2, fc3;
    const char *kind() const override { return "pointnet2"; }

protected:
    int pooled_dim() const override { return 1024; }

    static int sgemm(const float *A, int64_t lda, const float *W, int64_t ldw, const float *bias, float *C, int64_t ldc,
                     int64_t M, int N, int K, int act, const char *tag, cudaStream_t st)
    {
        GemmDesc g;
        g.A = A; g.lda = lda; g.B = W; g.ldb = ldw; g.C = C; g.ldc = ldc; g.M = (int)M; g.N = N; g.K = K;
        g.bias = bias; g.act = act; g.tag = tag;
        return launch_sgemm(g, st);
    }

    int run_head(Workspace &ws, const float *g, int64_t B, float *logits, cudaStream_t st) override
    {
        float *f512 = ws.take<float>(B * 512);
        float *f256 = ws.take<float>(B * 256);
        IQ_CHECK(ws.ok(), "pointnet2: workspace too small");
        if (ws.dry) return 0;
        if (int rc = sgemm(g, 1024, fc1.w, 1024, fc1.b, f512, 512, B, 512, 1024, ACT_RELU, "sgemm_head", st)) return rc;
        if (int rc = sgemm(f512, 512, fc2.w, 512, fc2.b, f256, 256, B, 256, 512, ACT_RELU, "sgemm_head", st)) return rc;
        return sgemm(f256, 256, fc3.w, 256, fc3.b, logits, num_classes, B, num_classes, 256, ACT_NONE, "sgemm_head", st);
    }

    // one multi-scale set abstraction: src (Bc*Nsrc rows) -> out (Bc*npoint, ld_out) columns [out_col, out_col+cout_total)
    int run_sa(const SaMsg &sa, const float *src_xyz, const float *src_in, int64_t ld_in, int kin, int64_t Bc, int Nsrc,
               float *new_xyz, float *U, float *V, int32_t *gidx, float *h1hi, float *h1lo, float *h2hi, float *h2lo,
               float *out, int64_t ld_out, int out_col, cudaStream_t st)
    {
        const int S = sa.npoint;
        if (int rc = launch_fps(src_xyz, Bc, Nsrc, S, nullptr, nullptr, new_xyz, st)) return rc;
        // U = [features ; xyz] W1cat^T per source point.  sa1 has K = 3 (SIMT); sa2's 324 -> 320 product runs on tcgen05:
        // the source rows are split into tf32 hi / lo in the (then idle) grouped-activation scratch first
        TcGemm pt;
        pt.A_hi = h1hi; pt.A_lo = h1lo; pt.lda = ld_in; pt.B_hi = sa.w1cat_hi; pt.B_lo = sa.w1cat_lo; pt.ldb = sa.ldw;
        pt.K = kin; pt.M = (int)(Bc * Nsrc); pt.N = sa.c1_total; pt.C = U; pt.ldc = sa.c1_total; pt.tag = "tc_sa_point";
        pt.four_terms = 1;           // U - V cancels downstream (relu(U[idx] - V + b1)): keep the Alo*Blo term, fp32-chain accuracy
        if (engine == 1 && kin >= 32 && h1lo && tc_gemm_supported(pt)) {
            if (int rc = launch_split_tf32(src_in, Bc * Nsrc, kin, ld_in, h1hi, h1lo, ld_in, st)) return rc;
            if (int rc = launch_gemm_tc(pt, st)) return rc;
        } else if (int rc = sgemm(src_in, ld_in, sa.w1cat, sa.ldw, nullptr, U, sa.c1_total, Bc * Nsrc, sa.c1_total, kin,
                                  ACT_NONE, "sgemm_sa_point", st))
            return rc;
        if (int rc = sgemm(new_xyz, 3, sa.w1x, 3, nullptr, V, sa.c1_total, Bc * S, sa.c1_total, 3, ACT_NONE,
                           "sgemm_sa_centroid", st))
            return rc;
        for (int s = 0; s < 3; ++s) {
            const Scale &sc = sa.sc[s];
            if (int rc = launch_ball_query(src_xyz, new_xyz, Bc, Nsrc, S, sc.radius, sc.K, gidx, st)) return rc;
            const int64_t rows = Bc * S * sc.K;
            if (engine == 1) {
                // layers 1 and 2 in one kernel: H1 = relu(U[idx] - V + b1) is produced tile by tile inside the GEMM that
                // consumes it (gemm_tc.cu, gathered-A variant) and never written to HBM.  IQ_TC_NO_GATHER (tests) takes the
                // two-kernel route through H1; the results are bitwise the same.
                // all three layers and the max over the group in one kernel (chain_tc.cu): H1 is gathered into the
                // operand ring and H2 stays in TMEM as the A operand of layer 3, so neither touches HBM.
                // IQ_TC_NO_CHAIN (tests, A/B runs) takes the two-kernel route below through H2 in HBM.
                SaChain ch;
                ch.U = U + sc.col1; ch.ldu = sa.c1_total; ch.V = V + sc.col1; ch.ldv = sa.c1_total; ch.b1 = sc.b1;
                ch.idx = gidx; ch.rows = rows; ch.K = sc.K; ch.S = S; ch.nsrc = Nsrc; ch.C1 = sc.c1; ch.C2 = sc.c2; ch.C3 = sc.c3;
                ch.W2_hi = sc.l2.w_hi; ch.W2_lo = sc.l2.w_lo; ch.b2 = sc.l2.b; ch.ldw2 = sc.c1;
                ch.W3_hi = sc.l3.w_hi; ch.W3_lo = sc.l3.w_lo; ch.b3 = sc.l3.b; ch.ldw3 = sc.c2;
                ch.out = out + out_col + sc.col3; ch.ld_out = ld_out;
                if (sa_chain_supported(ch) && !env_int("IQ_TC_NO_CHAIN", 0) && !env_int("IQ_TC_NO_GATHER", 0)) {
                    if (int rc = launch_sa_chain(ch, st)) return rc;
                    continue;
                }
                TcGemm a;
                if (env_int("IQ_TC_NO_GATHER", 0)) {
                    if (int rc = launch_group_sub_act(U + sc.col1, sa.c1_total, V + sc.col1, sa.c1_total, sc.b1, gidx, Bc, S,
                                                      sc.K, Nsrc, sc.c1, ACT_RELU, nullptr, h1hi, h1lo, sc.c1, st))
                        return rc;
                    a.A_hi = h1hi; a.A_lo = h1lo; a.lda = sc.c1;
                } else {
                    a.gather.U = U + sc.col1; a.gather.ldu = sa.c1_total; a.gather.V = V + sc.col1; a.gather.ldv = sa.c1_total;
                    a.gather.bias = sc.b1; a.gather.idx = gidx; a.gather.K = sc.K; a.gather.S = S; a.gather.nsrc = Nsrc;
                    a.gather.act = ACT_RELU;
                }
                a.B_hi = sc.l2.w_hi; a.B_lo = sc.l2.w_lo; a.ldb = sc.c1;
                a.K = sc.c1; a.M = (int)rows; a.N = sc.c2; a.C_hi = h2hi; a.C_lo = h2lo; a.ldc = sc.c2;
                a.bias = sc.l2.b; a.act = ACT_RELU; a.tag = "tc_sa_mlp12";
                if (int rc = launch_gemm_tc(a, st)) return rc;
                TcGemm b;
                b.mode = 1;
                b.A_hi = sc.l3.w_hi; b.A_lo = sc.l3.w_lo; b.lda = sc.c2; b.B_hi = h2hi; b.B_lo = h2lo; b.ldb = sc.c2;
                b.K = sc.c2; b.clouds = (int)(Bc * S); b.points = sc.K; b.cout = sc.c3;
                b.out_max = out + out_col + sc.col3; b.ld_out = ld_out; b.bias = sc.l3.b; b.act = ACT_RELU;
                b.tag = "tc_sa_mlp3_pool";
                if (int rc = launch_gemm_tc(b, st)) return rc;
            } else {
                // exact fp32 path: H1, H2 in fp32 (h1hi / h2hi reused as plain buffers), pooling by tiles of 128 rows
                if (int rc = launch_group_sub_act(U + sc.col1, sa.c1_total, V + sc.col1, sa.c1_total, sc.b1, gidx, Bc, S,
                                                  sc.K, Nsrc, sc.c1, ACT_RELU, h1hi, nullptr, nullptr, sc.c1, st))
                    return rc;
                if (int rc = sgemm(h1hi, sc.c1, sc.l2.w, sc.c1, sc.l2.b, h2hi, sc.c2, rows, sc.c2, sc.c1, ACT_RELU,
                                   "sgemm_sa_mlp2", st))
                    return rc;
                if (int rc = sgemm(h2hi, sc.c2, sc.l3.w, sc.c2, sc.l3.b, h1hi, sc.c3, rows, sc.c3, sc.c2, ACT_RELU,
                                   "sgemm_sa_mlp3", st))
                    return rc;
                if (int rc = launch_group_max(h1hi, sc.c3, Bc * S, sc.K, sc.c3, out + out_col + sc.col3, ld_out, st))
                    return rc;
            }
        }
        return 0;
    }

    int run_body(Workspace &ws, const float *x, int point_major, int64_t Bc, int64_t N, float *pooled, float *,
                 int64_t *, cudaStream_t st) override
    {
        IQ_CHECK(N >= 512 && N <= 2048, "pointnet2: num_points must be in [512,2048]");
        const int S1 = 512, S2 = 128;
        float *xyz = ws.take<float>(Bc * N * 3);
        float *xyz1 = ws.take<float>(Bc * S1 * 3);
        float *xyz2 = ws.take<float>(Bc * S2 * 3);
        float *U = ws.take<float>(std::max<int64_t>(Bc * N * sa1.c1_total, Bc * S1 * sa2.c1_total));
        float *V = ws.take<float>(std::max<int64_t>(Bc * S1 * sa1.c1_total, Bc * S2 * sa2.c1_total));
        int32_t *gidx = ws.take<int32_t>(Bc * S1 * 128);
        // largest grouped activations: sa1 scale 3 (65536 rows x 64 / 96) vs sa2 scale 3 (16384 x 128 / 128)
        const int64_t h1n = std::max<int64_t>(Bc * S1 * 128 * 128, Bc * S2 * 128 * 256);
        const int64_t h2n = std::max<int64_t>(Bc * S1 * 128 * 96, Bc * S2 * 128 * 128);
        float *h1hi = ws.take<float>(h1n);
        float *h1lo = engine == 1 ? ws.take<float>(h1n) : nullptr;
        float *h2hi = ws.take<float>(h2n);
        float *h2lo = engine == 1 ? ws.take<float>(h2n) : nullptr;
        float *l1 = ws.take<float>(Bc * S1 * 324);               // [320 features | xyz | pad]
        float *l2 = ws.take<float>(Bc * S2 * 644);               // [xyz | 640 features | pad]
        float *a256 = ws.take<float>(Bc * S2 * 256);
        float *a512 = ws.take<float>(Bc * S2 * 512);
        float *a256lo = engine == 1 ? ws.take<float>(Bc * S2 * 256) : nullptr;
        float *a512lo = engine == 1 ? ws.take<float>(Bc * S2 * 512) : nullptr;
        float *pmax = ws.take<float>(Bc * 1024);
        IQ_CHECK(ws.ok(), "pointnet2: workspace too small");
        if (ws.dry) return 0;

        const float *pts = x;
        if (!point_major) {
            if (int rc = launch_xyz_to_point_major(x, Bc, N, xyz, st)) return rc;
            pts = xyz;
        }
        if (int rc = run_sa(sa1, pts, pts, 3, 3, Bc, (int)N, xyz1, U, V, gidx, h1hi, h1lo, h2hi, h2lo, l1, 324, 0, st))
            return rc;
        if (int rc = launch_copy_cols(xyz1, 3, Bc * S1, 3, l1 + 320, 324, 4, st)) return rc;
        if (int rc = run_sa(sa2, xyz1, l1, 324, 324, Bc, S1, xyz2, U, V, gidx, h1hi, h1lo, h2hi, h2lo, l2, 644, 3, st))
            return rc;
        if (int rc = launch_copy_cols(xyz2, 3, Bc * S2, 3, l2, 644, 3, st)) return rc;
        if (int rc = launch_copy_cols(xyz2, 3, Bc * S2, 0, l2 + 643, 644, 1, st)) return rc;     // zero the pad column
        if (engine == 1 && (Bc * S2) % 128 == 0) {
            // group-all layer on tcgen05 (3xTF32): 644 -> 256 -> 512 as hi/lo STORE products, 512 -> 1024 with the max over
            // the cloud's 128 points in the POOL epilogue (round 1 ran these three on the fp32 SIMT GEMM: 30 ms per 3300 clouds)
            if (int rc = launch_split_tf32(l2, Bc * S2, 644, 644, h1hi, h1lo, 644, st)) return rc;
            TcGemm a;
            a.A_hi = h1hi; a.A_lo = h1lo; a.lda = 644; a.B_hi = s3a.w_hi; a.B_lo = s3a.w_lo; a.ldb = 644; a.K = 644;
            a.M = (int)(Bc * S2); a.N = 256; a.C_hi = a256; a.C_lo = a256lo; a.ldc = 256; a.bias = s3a.b; a.act = ACT_RELU;
            a.tag = "tc_sa3";
            if (int rc = launch_gemm_tc(a, st)) return rc;
            TcGemm b;
            b.A_hi = a256; b.A_lo = a256lo; b.lda = 256; b.B_hi = s3b.w_hi; b.B_lo = s3b.w_lo; b.ldb = 256; b.K = 256;
            b.M = (int)(Bc * S2); b.N = 512; b.C_hi = a512; b.C_lo = a512lo; b.ldc = 512; b.bias = s3b.b; b.act = ACT_RELU;
            b.tag = "tc_sa3";
            if (int rc = launch_gemm_tc(b, st)) return rc;
            TcGemm c;
            c.mode = 1;
            c.A_hi = s3c.w_hi; c.A_lo = s3c.w_lo; c.lda = 512; c.B_hi = a512; c.B_lo = a512lo; c.ldb = 512; c.K = 512;
            c.clouds = (int)Bc; c.points = S2; c.cout = 1024; c.out_max = pooled; c.ld_out = 1024; c.bias = s3c.b;
            c.act = ACT_RELU; c.tag = "tc_sa3_pool";
            return launch_gemm_tc(c, st);
        }
        if (int rc = sgemm(l2, 644, s3a.w, 644, s3a.b, a256, 256, Bc * S2, 256, 644, ACT_RELU, "sgemm_sa3", st)) return rc;
        if (int rc = sgemm(a256, 256, s3b.w, 256, s3b.b, a512, 512, Bc * S2, 512, 256, ACT_RELU, "sgemm_sa3", st)) return rc;
        GemmDesc g;
        g.A = a512; g.lda = 512; g.B = s3c.w; g.ldb = 512; g.M = (int)(Bc * S2); g.N = 1024; g.K = 512; g.bias = s3c.b;
        g.act = ACT_RELU; g.pool_max = pmax; g.tag = "sgemm_sa3_pool";
        if (int rc = launch_sgemm(g, st)) return rc;
        return launch_pool_finish(pmax, nullptr, nullptr, Bc, 1, S2, 1024, pooled, 1024, nullptr, nullptr, 0, st);
    }
};

bool up(PointNet2Model *m, const std::vector<float> &h, float **d, std::string &err)
{
    if (m->arena_.upload(h, d)) { err = last_error(); return false; }
    return true;
}

bool make_dense(PointNet2Model *m, const StateDict &sd, const std::string &conv, const std::string &bn, int co, int ci,
                Dense &d, std::string &err, int pad_cin = 0)
{
    std::vector<float> w, b;
    if (!fold_dense(sd, conv + ".weight", conv + ".bias", bn, co, ci, w, b, err)) return false;
    if (pad_cin > ci) {
        std::vector<float> wp((size_t)co * pad_cin, 0.0f);
        for (int o = 0; o < co; ++o)
            for (int i = 0; i < ci; ++i) wp[(size_t)o * pad_cin + i] = w[(size_t)o * ci + i];
        w.swap(wp);
        ci = pad_cin;
    }
    std::vector<float> hi, lo;
    split_tf32_host(w, hi, lo);
    d.cout = co; d.cin = ci;
    return up(m, w, &d.w, err) && up(m, b, &d.b, err) && up(m, hi, &d.w_hi, err) && up(m, lo, &d.w_lo, err);
}

bool make_sa(PointNet2Model *m, const StateDict &sd, const std::string &p, int npoint, int cin_feat,
             const double (&radii)[3], const int (&Ks)[3], const int (&mlps)[3][3], SaMsg &sa, std::string &err)
{
    sa.npoint = npoint;
    sa.cin_feat = cin_feat;
    const int cin = cin_feat + 3;
    sa.ldw = cin_feat ? cin + 1 : 3;                             // [features ; xyz ; pad] (K multiple of 4) or xyz only
    sa.c1_total = mlps[0][0] + mlps[1][0] + mlps[2][0];
    sa.cout_total = mlps[0][2] + mlps[1][2] + mlps[2][2];
    std::vector<float> w1cat((size_t)sa.c1_total * sa.ldw, 0.0f), w1x((size_t)sa.c1_total * 3);
    int col1 = 0, col3 = 0;
    for (int s = 0; s < 3; ++s) {
        Scale &sc = sa.sc[s];
        sc.radius = radii[s]; sc.K = Ks[s]; sc.c1 = mlps[s][0]; sc.c2 = mlps[s][1]; sc.c3 = mlps[s][2];
        sc.col1 = col1; sc.col3 = col3;
        const std::string cb = p + ".conv_blocks." + std::to_string(s) + ".", bb = p + ".bn_blocks." + std::to_string(s) + ".";
        std::vector<float> w, b;
        if (!fold_dense(sd, cb + "0.weight", cb + "0.bias", bb + "0", sc.c1, cin, w, b, err)) return false;
        for (int o = 0; o < sc.c1; ++o) {
            for (int i = 0; i < cin; ++i) w1cat[(size_t)(col1 + o) * sa.ldw + i] = w[(size_t)o * cin + i];
            for (int i = 0; i < 3; ++i) w1x[(size_t)(col1 + o) * 3 + i] = w[(size_t)o * cin + cin_feat + i];
        }
        if (!up(m, b, &sc.b1, err)) return false;
        if (!make_dense(m, sd, cb + "1", bb + "1", sc.c2, sc.c1, sc.l2, err)) return false;
        if (!make_dense(m, sd, cb + "2", bb + "2", sc.c3, sc.c2, sc.l3, err)) return false;
        col1 += sc.c1;
        col3 += sc.c3;
    }
    std::vector<float> whi, wlo;
    split_tf32_host(w1cat, whi, wlo);
    return up(m, w1cat, &sa.w1cat, err) && up(m, w1x, &sa.w1x, err) && up(m, whi, &sa.w1cat_hi, err) &&
           up(m, wlo, &sa.w1cat_lo, err);
}

}  // namespace

Model *create_pointnet2_model(const StateDict &sd, int num_classes, std::string &err)
{
    std::unique_ptr<PointNet2Model> m(new PointNet2Model());
    m->num_classes = num_classes;
    m->chunk = 165;
    PointNet2Model *p = m.get();
    const double r1[3] = {0.1, 0.2, 0.4}, r2[3] = {0.2, 0.4, 0.8};
    const int k1[3] = {16, 32, 128}, k2[3] = {32, 64, 128};
    const int m1[3][3] = {{32, 32, 64}, {64, 64, 128}, {64, 96, 128}};
    const int m2[3][3] = {{64, 64, 128}, {128, 128, 256}, {128, 128, 256}};
    const bool ok = make_sa(p, sd, "sa1", 512, 0, r1, k1, m1, m->sa1, err) &&
                    make_sa(p, sd, "sa2", 128, 320, r2, k2, m2, m->sa2, err) &&
                    make_dense(p, sd, "sa3.mlp_convs.0", "sa3.mlp_bns.0", 256, 643, m->s3a, err, 644) &&
                    make_dense(p, sd, "sa3.mlp_convs.1", "sa3.mlp_bns.1", 512, 256, m->s3b, err) &&
                    make_dense(p, sd, "sa3.mlp_convs.2", "sa3.mlp_bns.2", 1024, 512, m->s3c, err) &&
                    make_dense(p, sd, "fc1", "bn1", 512, 1024, m->fc1, err) &&
                    make_dense(p, sd, "fc2", "bn2", 256, 512, m->fc2, err) &&
                    make_dense(p, sd, "fc3", "", num_classes, 256, m->fc3, err);
    return ok ? m.release() : nullptr;
}

}  // namespace iq
