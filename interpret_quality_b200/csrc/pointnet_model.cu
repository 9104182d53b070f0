// PointNetCls forward pass (eval mode, feature_transform=True) on the folded weights.
//
// Reference behaviour restated (never copied): models/pointnet.py:11-47 (STNkd),
// :49-89 (PointNetfeat), :91-115 (PointNetCls).  Quirks kept: no ReLU after
// feat.bn3 (:82), BN applied after dropout on fc2 (:112, dropout is identity in
// eval), forward returns (logits, trans_feat, crt_points).
//
// Every shared MLP is a GEMM over point-major activations with BN folded in;
// the three 128->1024 layers use the pooling epilogue, so the (B, 1024, N)
// tensors are never written.  The T-Net outputs are produced already transposed
// (rows of fc3 permuted on the host, identity folded into the bias) so they can
// be used directly as the K-major B operand of the batched point transform.
#include "model.cuh"

namespace iq {

namespace {

__global__ void transpose_square_kernel(const float *__restrict__ in, int n, float *__restrict__ out)
{
    const int b = blockIdx.x;
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
        const int r = t / n, c = t % n;
        out[(int64_t)b * n * n + c * n + r] = in[(int64_t)b * n * n + t];
    }
}

struct TNet {
    Dense c1, c2, c3, f1, f2, f3;   // f3 emits the transposed matrix with the identity added
    int k = 3;
};

// (B, Nin, 3) or (B, 3, Nin) -> point-major (B, Npad, 3); rows >= Nin repeat the cloud's first point
__global__ void pad_cloud_kernel(const float *__restrict__ x, int point_major, int Nin, int Npad, float *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = blockIdx.y;
    if (i >= Npad) return;
    const int j = i < Nin ? i : 0;
    float *o = out + (b * Npad + i) * 3;
    if (point_major) {
        const float *p = x + (b * Nin + j) * 3;
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
    } else {
        const float *p = x + b * 3 * Nin + j;
        o[0] = p[0]; o[1] = p[Nin]; o[2] = p[2 * (int64_t)Nin];
    }
}

class PointNetModel : public Model {
public:
    TNet stn, fstn;
    Dense conv1, conv2, conv3, fc1, fc2, fc3;
    const char *kind() const override { return "pointnet"; }

protected:
    static int dense(const Dense &d, const float *in, int64_t ldin, float *out, int64_t M, int act, cudaStream_t st,
                     float *out_hi = nullptr, float *out_lo = nullptr)
    {
        GemmDesc g;
        g.A = in; g.lda = ldin; g.B = d.w; g.ldb = d.cin; g.C = out; g.C_hi = out_hi; g.C_lo = out_lo; g.ldc = d.cout;
        g.M = (int)M; g.N = d.cout; g.K = d.cin; g.bias = d.b; g.act = act;
        return launch_sgemm(g, st);
    }
    // tcgen05: (rows, cin) hi/lo -> act(x W^T + b) as hi/lo (rows, cout)
    static int dense_tc(const Dense &d, const float *in_hi, const float *in_lo, float *out_hi, float *out_lo, int64_t M,
                        int act, cudaStream_t st)
    {
        TcGemm t;
        t.A_hi = in_hi; t.A_lo = in_lo; t.lda = d.cin; t.B_hi = d.w_hi; t.B_lo = d.w_lo; t.ldb = d.cin; t.K = d.cin;
        t.M = (int)M; t.N = d.cout; t.C_hi = out_hi; t.C_lo = out_lo; t.ldc = d.cout; t.bias = d.b; t.act = act;
        t.tag = "tc_conv";
        return launch_gemm_tc(t, st);
    }
    // tcgen05: conv + BN + act + max over the points of each cloud (+ argmax)
    static int dense_pool_tc(const Dense &d, const float *in_hi, const float *in_lo, int64_t Bc, int64_t N, int act,
                             float *out, int64_t *out_arg, cudaStream_t st)
    {
        TcGemm t;
        t.mode = 1;
        t.A_hi = d.w_hi; t.A_lo = d.w_lo; t.lda = d.cin; t.B_hi = in_hi; t.B_lo = in_lo; t.ldb = d.cin; t.K = d.cin;
        t.clouds = (int)Bc; t.points = (int)N; t.cout = d.cout; t.out_max = out; t.out_arg = out_arg; t.ld_out = d.cout;
        t.bias = d.b; t.act = act; t.tag = "tc_conv_pool";
        return launch_gemm_tc(t, st);
    }
    static int dense_pool(const Dense &d, const float *in, int64_t ldin, int64_t Bc, int64_t N, int act, float *pmax,
                          int32_t *parg, float *out, int64_t *out_arg, cudaStream_t st)
    {
        GemmDesc g;
        g.A = in; g.lda = ldin; g.B = d.w; g.ldb = d.cin; g.M = (int)(Bc * N); g.N = d.cout; g.K = d.cin;
        g.bias = d.b; g.act = act; g.pool_max = pmax; g.pool_arg = parg;
        if (int rc = launch_sgemm(g, st)) return rc;
        return launch_pool_finish(pmax, parg, nullptr, Bc, (int)(N / 128), (int)N, d.cout, out, d.cout, out_arg, nullptr,
                                  0, st);
    }
    // T-Net: in (rows, k) -> tmat (Bc, k*k) holding the transposed transform
    // in: fp32 rows (SIMT first layer when in_hi == nullptr) or a tf32 hi/lo pair
    int run_tnet(const TNet &t, const float *in, const float *in_hi, const float *in_lo, int64_t ldin, int64_t Bc,
                 int64_t N, float *a64, float *a64lo, float *a128, float *a128lo, float *pmax, float *g1024, float *f512,
                 float *f256, float *tmat, cudaStream_t st)
    {
        const int64_t rows = Bc * N;
        if (engine == 1) {
            if (in_hi) { if (int rc = dense_tc(t.c1, in_hi, in_lo, a64, a64lo, rows, ACT_RELU, st)) return rc; }
            else { if (int rc = dense(t.c1, in, ldin, nullptr, rows, ACT_RELU, st, a64, a64lo)) return rc; }
            if (int rc = dense_tc(t.c2, a64, a64lo, a128, a128lo, rows, ACT_RELU, st)) return rc;
            if (int rc = dense_pool_tc(t.c3, a128, a128lo, Bc, N, ACT_RELU, g1024, nullptr, st)) return rc;
        } else {
            if (int rc = dense(t.c1, in, ldin, a64, rows, ACT_RELU, st)) return rc;
            if (int rc = dense(t.c2, a64, 64, a128, rows, ACT_RELU, st)) return rc;
            if (int rc = dense_pool(t.c3, a128, 128, Bc, N, ACT_RELU, pmax, nullptr, g1024, nullptr, st)) return rc;
        }
        if (int rc = dense(t.f1, g1024, 1024, f512, Bc, ACT_RELU, st)) return rc;
        if (int rc = dense(t.f2, f512, 512, f256, Bc, ACT_RELU, st)) return rc;
        return dense(t.f3, f256, 256, tmat, Bc, ACT_NONE, st);
    }
    // out[b] (N, k) = in[b] (N, k) * T[b], with tmat[b] = T[b]^T stored (k, k) row-major
    static int apply_transform(const float *in, const float *tmat, int k, int64_t Bc, int64_t N, float *out,
                               cudaStream_t st, float *out_hi = nullptr, float *out_lo = nullptr)
    {
        GemmDesc g;
        g.A = in; g.lda = k; g.strideA = N * k; g.B = tmat; g.ldb = k; g.strideB = (int64_t)k * k;
        g.C = out; g.C_hi = out_hi; g.C_lo = out_lo; g.ldc = k; g.strideC = N * k; g.M = (int)N; g.N = k; g.K = k;
        g.batch = (int)Bc; g.tag = "sgemm_transform";
        return launch_sgemm(g, st);
    }

    int pooled_dim() const override { return 1024; }
    // point-wise layers + max pooling only: one copy of the collapsed location carries all of them (collapse.cu)
    int collapse_copies() const override { return 1; }

    int run_head(Workspace &ws, const float *g, int64_t B, float *logits, cudaStream_t st) override
    {
        float *f512 = ws.take<float>(B * 512);
        float *f256 = ws.take<float>(B * 256);
        IQ_CHECK(ws.ok(), "pointnet: workspace too small");
        if (ws.dry) return 0;
        if (int rc = dense(fc1, g, 1024, f512, B, ACT_RELU, st)) return rc;
        if (int rc = dense(fc2, f512, 512, f256, B, ACT_RELU, st)) return rc;
        return dense(fc3, f256, 256, logits, B, ACT_NONE, st);
    }

    int run_body(Workspace &ws, const float *x, int point_major, int64_t Bc, int64_t N, float *pooled,
                 float *aux_trans_feat, int64_t *aux_crt, cudaStream_t st) override
    {
        // The pooled GEMMs tile a cloud in 128-point blocks.  PointNet only ever takes the max over the points, so any other
        // number of points is padded with copies of each cloud's first point: the pooled maxima and -- ties go to the lowest
        // index -- the critical-point indices are those of the unpadded cloud (models/pointnet.py:77-88 takes N from the shape).
        const int64_t Nin = N;
        N = round_up(N, 128);
        const int64_t rows = Bc * N;
        const int tiles = (int)(N / 128);
        float *xyz = ws.take<float>(rows * 3);
        float *xt = ws.take<float>(rows * 3);
        float *a64 = ws.take<float>(rows * 64);
        float *a128 = ws.take<float>(rows * 128);
        float *h64 = ws.take<float>(rows * 64);
        float *h64t = ws.take<float>(rows * 64);
        const bool tc = engine == 1;
        float *a64lo = tc ? ws.take<float>(rows * 64) : nullptr;
        float *a128lo = tc ? ws.take<float>(rows * 128) : nullptr;
        float *h64hi = tc ? ws.take<float>(rows * 64) : nullptr;
        float *h64lo = tc ? ws.take<float>(rows * 64) : nullptr;
        float *pmax = ws.take<float>(Bc * tiles * 1024);
        int32_t *parg = ws.take<int32_t>(Bc * tiles * 1024);
        float *g1024 = ws.take<float>(Bc * 1024);
        float *f512 = ws.take<float>(Bc * 512);
        float *f256 = ws.take<float>(Bc * 256);
        float *t9 = ws.take<float>(Bc * 9);
        float *t4096 = ws.take<float>(Bc * 4096);
        IQ_CHECK(ws.ok(), "pointnet: workspace too small");
        if (ws.dry) return 0;

        const float *pts = x;
        if (Nin != N) {
            pad_cloud_kernel<<<dim3((unsigned)ceil_div(N, 256), (unsigned)Bc), 256, 0, st>>>(x, point_major, (int)Nin, (int)N, xyz);
            IQ_COUNT_LAUNCH();
            IQ_LAUNCH_CHECK();
            pts = xyz;
        } else if (!point_major) {
            if (int rc = launch_xyz_to_point_major(x, Bc, N, xyz, st)) return rc;
            pts = xyz;
        }
        if (int rc = run_tnet(stn, pts, nullptr, nullptr, 3, Bc, N, a64, a64lo, a128, a128lo, pmax, g1024, f512, f256, t9, st))
            return rc;
        if (int rc = apply_transform(pts, t9, 3, Bc, N, xt, st)) return rc;
        if (int rc = dense(conv1, xt, 3, h64, rows, ACT_RELU, st, h64hi, h64lo)) return rc;
        if (int rc = run_tnet(fstn, h64, h64hi, h64lo, 64, Bc, N, a64, a64lo, a128, a128lo, pmax, g1024, f512, f256, t4096, st))
            return rc;
        if (aux_trans_feat) {
            transpose_square_kernel<<<(unsigned)Bc, 256, 0, st>>>(t4096, 64, aux_trans_feat);
            IQ_COUNT_LAUNCH();
            IQ_LAUNCH_CHECK();
        }
        if (tc) {
            // the point transform writes its output already split; conv2 / conv3 run on tcgen05
            if (int rc = apply_transform(h64, t4096, 64, Bc, N, nullptr, st, h64t, a64lo)) return rc;
            if (int rc = dense_tc(conv2, h64t, a64lo, a128, a128lo, rows, ACT_RELU, st)) return rc;
            return dense_pool_tc(conv3, a128, a128lo, Bc, N, ACT_NONE, pooled, aux_crt, st);
        }
        if (int rc = apply_transform(h64, t4096, 64, Bc, N, h64t, st)) return rc;
        if (int rc = dense(conv2, h64t, 64, a128, rows, ACT_RELU, st)) return rc;
        return dense_pool(conv3, a128, 128, Bc, N, ACT_NONE, pmax, parg, pooled, aux_crt, st);
    }
};

bool make_dense(PointNetModel *m, const StateDict &sd, const std::string &conv, const std::string &bn, int co, int ci,
                Dense &d, std::string &err, int transpose_k = 0)
{
    std::vector<float> w, b;
    if (!fold_dense(sd, conv + ".weight", conv + ".bias", bn, co, ci, w, b, err)) return false;
    if (transpose_k) {
        // rows reordered so the output is T^T, and the identity of models/pointnet.py:42-45 folded into the bias
        const int k = transpose_k;
        std::vector<float> w2(w.size()), b2(b.size());
        for (int r = 0; r < k; ++r)
            for (int c = 0; c < k; ++c) {
                const int src = r * k + c, dst = c * k + r;
                for (int i = 0; i < ci; ++i) w2[(size_t)dst * ci + i] = w[(size_t)src * ci + i];
                b2[dst] = b[src] + (r == c ? 1.0f : 0.0f);
            }
        w.swap(w2);
        b.swap(b2);
    }
    d.cout = co; d.cin = ci;
    std::vector<float> hi, lo;
    split_tf32_host(w, hi, lo);
    if (m->arena_.upload(w, &d.w) || m->arena_.upload(b, &d.b) || m->arena_.upload(hi, &d.w_hi) ||
        m->arena_.upload(lo, &d.w_lo)) {
        err = last_error();
        return false;
    }
    return true;
}

bool make_tnet(PointNetModel *m, const StateDict &sd, const std::string &p, int k, TNet &t, std::string &err)
{
    t.k = k;
    return make_dense(m, sd, p + "conv1", p + "bn1", 64, k, t.c1, err) &&
           make_dense(m, sd, p + "conv2", p + "bn2", 128, 64, t.c2, err) &&
           make_dense(m, sd, p + "conv3", p + "bn3", 1024, 128, t.c3, err) &&
           make_dense(m, sd, p + "fc1", p + "bn4", 512, 1024, t.f1, err) &&
           make_dense(m, sd, p + "fc2", p + "bn5", 256, 512, t.f2, err) &&
           make_dense(m, sd, p + "fc3", "", k * k, 256, t.f3, err, k);
}

}  // namespace

Model *create_pointnet_model(const StateDict &sd, int num_classes, std::string &err)
{
    std::unique_ptr<PointNetModel> m(new PointNetModel());
    m->num_classes = num_classes;
    m->chunk = 660;
    m->lanes = 3;                // five chunks per 3300-cloud step, short kernels: 3 lanes measured +5 % over 2
    PointNetModel *p = m.get();
    const bool ok = make_tnet(p, sd, "feat.stn.", 3, m->stn, err) && make_tnet(p, sd, "feat.fstn.", 64, m->fstn, err) &&
                    make_dense(p, sd, "feat.conv1", "feat.bn1", 64, 3, m->conv1, err) &&
                    make_dense(p, sd, "feat.conv2", "feat.bn2", 128, 64, m->conv2, err) &&
                    make_dense(p, sd, "feat.conv3", "feat.bn3", 1024, 128, m->conv3, err) &&
                    make_dense(p, sd, "fc1", "bn1", 512, 1024, m->fc1, err) &&
                    make_dense(p, sd, "fc2", "bn2", 256, 512, m->fc2, err) &&
                    make_dense(p, sd, "fc3", "", num_classes, 256, m->fc3, err);
    return ok ? m.release() : nullptr;
}

}  // namespace iq
