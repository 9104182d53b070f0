// PointConvDensityClsSsg forward pass (eval mode) on the folded weights.
//
// Reference behaviour restated (never copied): models/pointconv.py:199-209 (compute_density), :103-114 (knn_point),
// :117-170 (sample_and_group / sample_and_group_all with density), :212-235 (DensityNet -- ReLU after every layer,
// the sigmoid branch is unreachable and that is kept), :238-265 (WeightNet), :324-391
// (PointConvDensitySetAbstraction), :394-424 (classifier).
//
// Per set-abstraction layer and chunk of clouds:
//   inverse density  1 / mean_j exp(-d_ij / 2bw^2) / (2.5 bw)        density_kernel (exact K=3 distance recipe)
//   new_xyz          FPS (exact)  |  mean of the points (group-all)
//   idx              K nearest points of each centroid                knn_point_kernel (exact distances, warp top-k)
//   H1               relu(U[idx] - V + b1),  U = [x ; f] W1^T, V = c W1x^T    fp32 SIMT GEMMs + grouping.cu
//   H2, H3           shared MLP layers 2 and 3                        tcgen05 3xTF32
//   Wd               DensityNet(inv / max inv) * WeightNet(x_j - c)   per-row small nets on CUDA cores
//   agg              sum_j H3[j][c] * Wd[j][w]  -> (centroid, 16 C)   aggregate_kernel
//   out              relu(bn(Linear(agg)))                             tcgen05 STORE (sa1, sa2) / head GEMM (sa3)
// FPS and kNN depend on coordinates only, so no discrete decision sees the 3xTF32 arithmetic.
#include <stdlib.h>

#include "model.cuh"

namespace iq {

namespace {

constexpr unsigned FULL = 0xffffffffu;

struct SmallNets {                 // BN-folded DensityNet (1-16-8-1) and WeightNet (3-8-8-16)
    float d1w[16], d1b[16], d2w[8 * 16], d2b[8], d3w[8], d3b[1];
    float w1w[8 * 3], w1b[8], w2w[8 * 8], w2b[8], w3w[16 * 8], w3b[16];
};

__device__ __forceinline__ float sqdist3_exact(float sx, float sy, float sz, float ss, float dx, float dy, float dz, float dd)
{
    float dot = __fmul_rn(sx, dx);
    dot = __fmaf_rn(sy, dy, dot);
    dot = __fmaf_rn(sz, dz, dot);
    float t = __fmul_rn(-2.0f, dot);
    t = __fadd_rn(t, ss);
    return __fadd_rn(t, dd);
}
__device__ __forceinline__ float sqn3(float x, float y, float z)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// inverse KDE density of every point of a cloud (compute_density, models/pointconv.py:199-209, then 1/density);
// one thread per point, cloud staged in shared memory.  The density only feeds DensityNet (a float path, no discrete
// decision), so the N x N Gaussian runs on the SFU: exp(-d / 2bw^2) = ex2(d * c), one multiply by 1/(2.5 bw N) at the
// end (the exact div.rn / expf form measured 4x slower: 15.5 ms per 3300 clouds).
__global__ void __launch_bounds__(256)
density_kernel(const float *__restrict__ xyz, int N, float two_bw2, float norm, float *__restrict__ inv_density)
{
    extern __shared__ float4 pts[];
    const int b = blockIdx.y;
    const float *p = xyz + (int64_t)b * N * 3;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
        pts[i] = make_float4(x, y, z, sqn3(x, y, z));
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float4 q = pts[i];
    const float c = -1.4426950408889634f / two_bw2;               // log2(e) / (2 bw^2), negated
    float acc0 = 0.0f, acc1 = 0.0f;
    for (int j = 0; j < N; j += 2) {
        const float4 c0 = pts[j], c1 = pts[min(j + 1, N - 1)];
        const float d0 = sqdist3_exact(q.x, q.y, q.z, q.w, c0.x, c0.y, c0.z, c0.w);
        const float d1 = sqdist3_exact(q.x, q.y, q.z, q.w, c1.x, c1.y, c1.z, c1.w);
        acc0 += exp2f(d0 * c);
        if (j + 1 < N) acc1 += exp2f(d1 * c);
    }
    const float density = (acc0 + acc1) / (norm * (float)N);
    inv_density[(int64_t)b * N + i] = 1.0f / density;
}

// per-cloud mean of the points (sample_and_group_all, models/pointconv.py:160)
__global__ void __launch_bounds__(128) cloud_mean_kernel(const float *__restrict__ xyz, int N, float *__restrict__ mean)
{
    __shared__ double part[3][128];
    const int b = blockIdx.x;
    double a[3] = {0, 0, 0};
    for (int i = threadIdx.x; i < N; i += 128)
        for (int c = 0; c < 3; ++c) a[c] += (double)xyz[((int64_t)b * N + i) * 3 + c];
    for (int c = 0; c < 3; ++c) part[c][threadIdx.x] = a[c];
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s)
            for (int c = 0; c < 3; ++c) part[c][threadIdx.x] += part[c][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x < 3) mean[b * 3 + threadIdx.x] = (float)(part[threadIdx.x][0] / (double)N);
}

__global__ void iota_mod_kernel(int32_t *out, int64_t n, int mod)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = (int32_t)(t % mod);
}

// DensityNet + WeightNet for every grouped row; one warp per centroid (K rows), lanes stride over the rows.
// Wd[row][w] = density_scale(row) * weight(row)[w]
__global__ void __launch_bounds__(256)
small_nets_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz, const float *__restrict__ inv_density,
                  const int32_t *__restrict__ idx, int64_t centroids, int S, int K, int Nsrc, const SmallNets nets,
                  float *__restrict__ Wd)
{
    const int lane = threadIdx.x & 31;
    const int64_t cen = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (cen >= centroids) return;
    const int64_t cloud = cen / S;
    const float cx = new_xyz[cen * 3], cy = new_xyz[cen * 3 + 1], cz = new_xyz[cen * 3 + 2];
    float mx = -INFINITY;
    for (int j = lane; j < K; j += 32) mx = fmaxf(mx, inv_density[cloud * Nsrc + idx[cen * K + j]]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
    for (int j = lane; j < K; j += 32) {
        const int64_t src = cloud * Nsrc + idx[cen * K + j];
        const float ds0 = __fdiv_rn(inv_density[src], mx);
        float h1[16], h2[8];
#pragma unroll
        for (int o = 0; o < 16; ++o) h1[o] = fmaxf(fmaf(nets.d1w[o], ds0, nets.d1b[o]), 0.0f);
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            float a = nets.d2b[o];
#pragma unroll
            for (int i = 0; i < 16; ++i) a = fmaf(nets.d2w[o * 16 + i], h1[i], a);
            h2[o] = fmaxf(a, 0.0f);
        }
        float ds = nets.d3b[0];
#pragma unroll
        for (int i = 0; i < 8; ++i) ds = fmaf(nets.d3w[i], h2[i], ds);
        ds = fmaxf(ds, 0.0f);
        const float gx = xyz[src * 3] - cx, gy = xyz[src * 3 + 1] - cy, gz = xyz[src * 3 + 2] - cz;
        float a1[8], a2[8];
#pragma unroll
        for (int o = 0; o < 8; ++o)
            a1[o] = fmaxf(fmaf(nets.w1w[o * 3 + 2], gz, fmaf(nets.w1w[o * 3 + 1], gy, fmaf(nets.w1w[o * 3], gx, nets.w1b[o]))), 0.0f);
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            float a = nets.w2b[o];
#pragma unroll
            for (int i = 0; i < 8; ++i) a = fmaf(nets.w2w[o * 8 + i], a1[i], a);
            a2[o] = fmaxf(a, 0.0f);
        }
        float *out = Wd + (cen * K + j) * 16;
#pragma unroll
        for (int o = 0; o < 16; ++o) {
            float a = nets.w3b[o];
#pragma unroll
            for (int i = 0; i < 8; ++i) a = fmaf(nets.w3w[o * 8 + i], a2[i], a);
            out[o] = fmaxf(a, 0.0f) * ds;
        }
    }
}

// agg[cen][c*16 + w] = sum_j H3[cen*K + j][c] * Wd[cen*K + j][w]  (models/pointconv.py:378-380, the per-centroid
// (C x K)(K x 16) product).  One warp per (centroid, block of 128 channels): a lane owns 4 channels x 16 weights = 64
// accumulators, reads its 4 channels of a neighbour row as one float4 (512 B per warp, coalesced) and the 16 weights of
// that neighbour as four broadcast float4 loads, so the loop is 64 FMAs per 5 loads; the result leaves as 256
// contiguous bytes per lane.  (The first version used one thread per channel with 16 scalar shared-memory reads per
// FMA group and ran at 0.15 of the FMA peak.)
__global__ void __launch_bounds__(256, 2)
aggregate_kernel(const float *__restrict__ H3, const float *__restrict__ Wd, int64_t cents, int K, int C,
                 float *__restrict__ agg, float *__restrict__ agg_hi, float *__restrict__ agg_lo, int64_t ld)
{
    extern __shared__ float4 wd_s[];                               // per warp: K x 16 weights of its centroid
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int blocks_per_cen = C >> 7;
    const int64_t unit = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (unit >= cents * blocks_per_cen) return;
    const int64_t cen = unit / blocks_per_cen;
    const int cb = (int)(unit - cen * blocks_per_cen);
    const float4 *h = reinterpret_cast<const float4 *>(H3 + cen * K * (int64_t)C + cb * 128) + lane;
    float4 *w = wd_s + wib * (K * 4);
    {
        const float4 *src = reinterpret_cast<const float4 *>(Wd + cen * K * 16);
        for (int t = lane; t < K * 4; t += 32) w[t] = __ldg(src + t);
        __syncwarp();
    }
    float acc[4][16];
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int q = 0; q < 16; ++q) acc[e][q] = 0.0f;
    const int64_t hs = C >> 2;
    for (int j0 = 0; j0 < K; j0 += 4) {                              // K is a multiple of 4; four rows in flight
        float4 hv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) hv[u] = __ldg(h + (int64_t)(j0 + u) * hs);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float he[4] = {hv[u].x, hv[u].y, hv[u].z, hv[u].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 wv = w[(j0 + u) * 4 + q];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    acc[e][4 * q] = fmaf(he[e], wv.x, acc[e][4 * q]);
                    acc[e][4 * q + 1] = fmaf(he[e], wv.y, acc[e][4 * q + 1]);
                    acc[e][4 * q + 2] = fmaf(he[e], wv.z, acc[e][4 * q + 2]);
                    acc[e][4 * q + 3] = fmaf(he[e], wv.w, acc[e][4 * q + 3]);
                }
            }
        }
    }
    const int64_t o = cen * ld + (int64_t)(cb * 128 + lane * 4) * 16;
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int q = 0; q < 16; q += 4) {
            const float *a = &acc[e][q];
            if (agg) *reinterpret_cast<float4 *>(agg + o + e * 16 + q) = make_float4(a[0], a[1], a[2], a[3]);
            if (agg_hi) {
                float h4[4], l4[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    h4[t] = __uint_as_float((__float_as_uint(a[t]) + 0x1000u) & 0xffffe000u);
                    const float d = a[t] - h4[t];
                    l4[t] = __uint_as_float((__float_as_uint(d) + 0x1000u) & 0xffffe000u);
                }
                *reinterpret_cast<float4 *>(agg_hi + o + e * 16 + q) = make_float4(h4[0], h4[1], h4[2], h4[3]);
                *reinterpret_cast<float4 *>(agg_lo + o + e * 16 + q) = make_float4(l4[0], l4[1], l4[2], l4[3]);
            }
        }
}

struct SaLayer {
    int npoint = 0, K = 0, cin_feat = 0, c1 = 0, c2 = 0, c3 = 0, ldw = 0;
    double bandwidth = 0;
    bool group_all = false;
    float *w1 = nullptr;          // (c1, ldw) over our point-feature rows [features ; xyz ; pad] (reference order is xyz first)
    float *w1_hi = nullptr, *w1_lo = nullptr;   // its tf32 split (tcgen05 path of the per-point product)
    float *w1x = nullptr;         // (c1, 3)
    float *b1 = nullptr;
    Dense l2, l3, lin;            // lin: Linear(16*c3 -> c3) + bn_linear
    SmallNets nets;
};

// out[m][n] = relu(sum_s part[s][m][n] + bias[n]): the fixed-order reduction of a split-K product
__global__ void splitk_reduce_kernel(const float *__restrict__ part, int S, int64_t MN, int N, const float *__restrict__ bias,
                                     float *__restrict__ out)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= MN) return;
    float acc = part[t];
    for (int s = 1; s < S; ++s) acc += part[(int64_t)s * MN + t];
    out[t] = fmaxf(acc + bias[t % N], 0.0f);
}

class PointConvModel : public Model {
public:
    SaLayer sa[3];
    Dense fc1, fc2, fc3;
    const char *kind() const override { return "pointconv"; }

protected:
    int pooled_dim() const override { return 16 * 1024; }

    static int sgemm(const float *A, int64_t lda, const float *W, int64_t ldw, const float *bias, float *C, int64_t ldc,
                     int64_t M, int N, int K, int act, const char *tag, cudaStream_t st)
    {
        GemmDesc g;
        g.A = A; g.lda = lda; g.B = W; g.ldb = ldw; g.C = C; g.ldc = ldc; g.M = (int)M; g.N = N; g.K = K;
        g.bias = bias; g.act = act; g.tag = tag;
        return launch_sgemm(g, st);
    }

    int run_head(Workspace &ws, const float *agg3, int64_t B, float *logits, cudaStream_t st) override
    {
        float *f1024 = ws.take<float>(B * 1024);
        float *f512 = ws.take<float>(B * 512);
        float *f256 = ws.take<float>(B * 256);
        // Few clouds leave the 16384 -> 1024 Linear with ceil(B / 128) * 8 CTAs (2.8 ms for 66 clouds, ncu): split K sixteen ways
        // over the batch dimension of the SIMT GEMM and add the partial products in a fixed order
        constexpr int SPLIT = 16;
        const bool splitk = B <= 1024;
        float *part = splitk ? ws.take<float>((int64_t)SPLIT * B * 1024) : nullptr;
        IQ_CHECK(ws.ok(), "pointconv: workspace too small");
        if (ws.dry) return 0;
        const Dense &lin = sa[2].lin;
        if (splitk) {
            GemmDesc g;
            g.A = agg3; g.lda = 16384; g.strideA = 16384 / SPLIT; g.B = lin.w; g.ldb = 16384; g.strideB = 16384 / SPLIT;
            g.C = part; g.ldc = 1024; g.strideC = B * 1024; g.M = (int)B; g.N = 1024; g.K = 16384 / SPLIT; g.batch = SPLIT;
            g.tag = "sgemm_sa3_linear";
            if (int rc = launch_sgemm(g, st)) return rc;
            splitk_reduce_kernel<<<(unsigned)ceil_div(B * 1024, 256), 256, 0, st>>>(part, SPLIT, B * 1024, 1024, lin.b, f1024);
            IQ_COUNT_LAUNCH();
            IQ_LAUNCH_CHECK();
        } else
        // Linear(16 * 1024 -> 1024) + BN + ReLU of the group-all layer stays on the fp32 SIMT GEMM.  Measured on tcgen05
        // (gemm_tc STORE, K = 16384): the logits sit 1.1e-4 of scale from float64 with 3 and with 4 split terms alike -- so it is
        // the length of the TMEM accumulation, not the operand split -- against 2e-5 for the FMA chain; the 10 ms per 5440
        // clouds it would save are not worth leaving the fp32-noise band the tests hold every model to.
        if (int rc = sgemm(agg3, 16384, lin.w, 16384, lin.b, f1024, 1024, B, 1024, 16384, ACT_RELU, "sgemm_sa3_linear", st))
            return rc;
        if (int rc = sgemm(f1024, 1024, fc1.w, 1024, fc1.b, f512, 512, B, 512, 1024, ACT_RELU, "sgemm_head", st)) return rc;
        if (int rc = sgemm(f512, 512, fc2.w, 512, fc2.b, f256, 256, B, 256, 512, ACT_RELU, "sgemm_head", st)) return rc;
        return sgemm(f256, 256, fc3.w, 256, fc3.b, logits, num_classes, B, num_classes, 256, ACT_NONE, "sgemm_head", st);
    }

    struct Scratch {
        float *inv, *U, *V, *h1hi, *h1lo, *h2hi, *h2lo, *h3, *wd, *agghi, *agglo;
        int32_t *idx;
    };

    // one PointConvDensitySetAbstraction: (src_xyz, src_in) -> new_xyz and either out (hi/lo/fp32, ld_out) or agg_out
    int run_sa(const SaLayer &L, const float *src_xyz, const float *src_in, int64_t ld_in, int kin, int64_t Bc, int Nsrc,
               float *new_xyz, const Scratch &s, float *out, int64_t ld_out, float *agg_out, cudaStream_t st)
    {
        const int S = L.group_all ? 1 : L.npoint, K = L.group_all ? Nsrc : L.K;
        const int64_t cents = Bc * S, rows = cents * K;
        {   // inverse density of the source points
            ProfileScope _ps("density", st);
            const float two_bw2 = (float)(2.0 * L.bandwidth * L.bandwidth), norm = (float)(2.5 * L.bandwidth);
            dim3 grid((unsigned)ceil_div(Nsrc, 256), (unsigned)Bc);
            density_kernel<<<grid, 256, sizeof(float4) * (size_t)Nsrc, st>>>(src_xyz, Nsrc, two_bw2, norm, s.inv);
            IQ_COUNT_LAUNCH();
            IQ_LAUNCH_CHECK();
        }
        if (L.group_all) {
            ProfileScope _ps("group_all_setup", st);
            cloud_mean_kernel<<<(unsigned)Bc, 128, 0, st>>>(src_xyz, Nsrc, new_xyz);
            IQ_COUNT_LAUNCH();
            iota_mod_kernel<<<(unsigned)ceil_div(rows, 256), 256, 0, st>>>(s.idx, rows, Nsrc);
            IQ_COUNT_LAUNCH();
            IQ_LAUNCH_CHECK();
        } else {
            if (int rc = launch_fps(src_xyz, Bc, Nsrc, S, nullptr, nullptr, new_xyz, st)) return rc;
            if (int rc = launch_knn_point(src_xyz, new_xyz, Bc, Nsrc, S, K, s.idx, st)) return rc;
        }
        // U = [features ; xyz] W1^T per source point: K = 3 in sa1 (SIMT), 132 / 260 in sa2 / sa3 -> tcgen05 on the tf32 split
        // of the source rows (written into the grouped-activation scratch, idle at this point)
        TcGemm pt;
        pt.A_hi = s.h1hi; pt.A_lo = s.h1lo; pt.lda = ld_in; pt.B_hi = L.w1_hi; pt.B_lo = L.w1_lo; pt.ldb = L.ldw; pt.K = kin;
        pt.M = (int)(Bc * Nsrc); pt.N = L.c1; pt.C = s.U; pt.ldc = L.c1; pt.tag = "tc_sa_point";
        pt.four_terms = 1;           // U - V cancels downstream (relu(U[idx] - V + b1)): keep the Alo*Blo term, fp32-chain accuracy
        if (engine == 1 && kin >= 32 && s.h1lo && tc_gemm_supported(pt)) {
            if (int rc = launch_split_tf32(src_in, Bc * Nsrc, kin, ld_in, s.h1hi, s.h1lo, ld_in, st)) return rc;
            if (int rc = launch_gemm_tc(pt, st)) return rc;
        } else if (int rc = sgemm(src_in, ld_in, L.w1, L.ldw, nullptr, s.U, L.c1, Bc * Nsrc, L.c1, kin, ACT_NONE,
                                  "sgemm_sa_point", st))
            return rc;
        if (int rc = sgemm(new_xyz, 3, L.w1x, 3, nullptr, s.V, L.c1, cents, L.c1, 3, ACT_NONE, "sgemm_sa_centroid", st))
            return rc;
        const bool tc = engine == 1;
        const bool fuse12 = tc && L.c2 <= 128 && rows % 128 == 0 && !env_int("IQ_TC_NO_GATHER", 0);   // layers 1 + 2 in one kernel (gemm_tc.cu, gathered A)
        if (!fuse12)
            if (int rc = launch_group_sub_act(s.U, L.c1, s.V, L.c1, L.b1, s.idx, Bc, S, K, Nsrc, L.c1, ACT_RELU,
                                              tc ? nullptr : s.h1hi, tc ? s.h1hi : nullptr, tc ? s.h1lo : nullptr, L.c1, st))
                return rc;
        if (tc) {
            TcGemm a;
            if (fuse12) {
                a.gather.U = s.U; a.gather.ldu = L.c1; a.gather.V = s.V; a.gather.ldv = L.c1; a.gather.bias = L.b1;
                a.gather.idx = s.idx; a.gather.K = K; a.gather.S = S; a.gather.nsrc = Nsrc; a.gather.act = ACT_RELU;
            }
            a.A_hi = s.h1hi; a.A_lo = s.h1lo; a.lda = L.c1; a.B_hi = L.l2.w_hi; a.B_lo = L.l2.w_lo; a.ldb = L.c1;
            a.K = L.c1; a.M = (int)rows; a.N = L.c2; a.C_hi = s.h2hi; a.C_lo = s.h2lo; a.ldc = L.c2; a.bias = L.l2.b;
            a.act = ACT_RELU; a.tag = "tc_sa_mlp2";
            if (int rc = launch_gemm_tc(a, st)) return rc;
            TcGemm b;
            b.A_hi = s.h2hi; b.A_lo = s.h2lo; b.lda = L.c2; b.B_hi = L.l3.w_hi; b.B_lo = L.l3.w_lo; b.ldb = L.c2;
            b.K = L.c2; b.M = (int)rows; b.N = L.c3; b.C = s.h3; b.ldc = L.c3; b.bias = L.l3.b; b.act = ACT_RELU;
            b.tag = "tc_sa_mlp3";
            if (int rc = launch_gemm_tc(b, st)) return rc;
        } else {
            if (int rc = sgemm(s.h1hi, L.c1, L.l2.w, L.c1, L.l2.b, s.h2hi, L.c2, rows, L.c2, L.c1, ACT_RELU, "sgemm_sa_mlp2", st))
                return rc;
            if (int rc = sgemm(s.h2hi, L.c2, L.l3.w, L.c2, L.l3.b, s.h3, L.c3, rows, L.c3, L.c2, ACT_RELU, "sgemm_sa_mlp3", st))
                return rc;
        }
        {
            ProfileScope _ps("small_nets", st);
            small_nets_kernel<<<(unsigned)ceil_div(cents * 32, 256), 256, 0, st>>>(src_xyz, new_xyz, s.inv, s.idx, cents, S, K,
                                                                                  Nsrc, L.nets, s.wd);
            IQ_COUNT_LAUNCH();
            IQ_LAUNCH_CHECK();
        }
        {
            ProfileScope _ps("aggregate", st);
            const bool to_lin = agg_out == nullptr;                   // sa1 / sa2: feed the Linear right away
            IQ_CHECK(L.c3 % 128 == 0, "pointconv: aggregate needs a channel count that is a multiple of 128");
            const int64_t warps = cents * (L.c3 / 128);
            IQ_CHECK(K % 4 == 0 && K <= 512, "pointconv: aggregate needs a neighbour count that is a multiple of 4");
            const size_t agg_smem = sizeof(float) * 16 * (size_t)K * 8;
            if (agg_smem > 48 * 1024)
                if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(&aggregate_kernel), (int)agg_smem)) return rc;
            aggregate_kernel<<<(unsigned)ceil_div(warps * 32, 256), 256, agg_smem, st>>>(
                s.h3, s.wd, cents, K, L.c3, (to_lin && tc) ? nullptr : (to_lin ? s.agghi : agg_out),
                (to_lin && tc) ? s.agghi : nullptr, (to_lin && tc) ? s.agglo : nullptr, 16 * (int64_t)L.c3);
            IQ_COUNT_LAUNCH();
            IQ_LAUNCH_CHECK();
        }
        if (agg_out) return 0;
        if (tc) {
            TcGemm l;
            l.A_hi = s.agghi; l.A_lo = s.agglo; l.lda = 16 * L.c3; l.B_hi = L.lin.w_hi; l.B_lo = L.lin.w_lo;
            l.ldb = 16 * L.c3; l.K = 16 * L.c3; l.M = (int)cents; l.N = L.c3; l.C = out; l.ldc = ld_out; l.bias = L.lin.b;
            l.act = ACT_RELU; l.tag = "tc_sa_linear";
            return launch_gemm_tc(l, st);
        }
        return sgemm(s.agghi, 16 * L.c3, L.lin.w, 16 * L.c3, L.lin.b, out, ld_out, cents, L.c3, 16 * L.c3, ACT_RELU,
                     "sgemm_sa_linear", st);
    }

    int run_body(Workspace &ws, const float *x, int point_major, int64_t Bc, int64_t N, float *pooled, float *,
                 int64_t *, cudaStream_t st) override
    {
        IQ_CHECK(N >= 512 && N <= 2048 && N % 128 == 0, "pointconv: num_points must be a multiple of 128 in [512,2048]");
        const int S1 = 512, S2 = 128;
        const bool tc = engine == 1;
        Scratch s;
        float *xyz = ws.take<float>(Bc * N * 3);
        float *xyz1 = ws.take<float>(Bc * S1 * 3);
        float *xyz2 = ws.take<float>(Bc * S2 * 3);
        float *xyz3 = ws.take<float>(Bc * 3);
        s.inv = ws.take<float>(Bc * N);
        s.U = ws.take<float>(std::max<int64_t>(Bc * N * 64, std::max<int64_t>(Bc * S1 * 128, Bc * S2 * 256)));
        s.V = ws.take<float>(std::max<int64_t>(Bc * S1 * 64, std::max<int64_t>(Bc * S2 * 128, Bc * 256)));
        s.idx = ws.take<int32_t>(std::max<int64_t>(Bc * S1 * 32, Bc * S2 * 64));
        const int64_t rows_max_c1 = std::max<int64_t>(Bc * S1 * 32 * 64, std::max<int64_t>(Bc * S2 * 64 * 128, Bc * S2 * 256));
        const int64_t rows_max_c2 = std::max<int64_t>(Bc * S1 * 32 * 64, std::max<int64_t>(Bc * S2 * 64 * 128, Bc * S2 * 512));
        const int64_t rows_max_c3 = std::max<int64_t>(Bc * S1 * 32 * 128, std::max<int64_t>(Bc * S2 * 64 * 256, Bc * S2 * 1024));
        s.h1hi = ws.take<float>(rows_max_c1);
        s.h1lo = tc ? ws.take<float>(rows_max_c1) : nullptr;
        s.h2hi = ws.take<float>(rows_max_c2);
        s.h2lo = tc ? ws.take<float>(rows_max_c2) : nullptr;
        s.h3 = ws.take<float>(rows_max_c3);
        s.wd = ws.take<float>(std::max<int64_t>(Bc * S1 * 32, Bc * S2 * 64) * 16);
        const int64_t aggn = std::max<int64_t>(Bc * S1 * 2048, Bc * S2 * 4096);
        s.agghi = ws.take<float>(aggn);
        s.agglo = tc ? ws.take<float>(aggn) : nullptr;
        float *l1 = ws.take<float>(Bc * S1 * 132);                // [128 features | xyz | pad]
        float *l2 = ws.take<float>(Bc * S2 * 260);                // [256 features | xyz | pad]
        IQ_CHECK(ws.ok(), "pointconv: workspace too small");
        if (ws.dry) return 0;

        const float *pts = x;
        if (!point_major) {
            if (int rc = launch_xyz_to_point_major(x, Bc, N, xyz, st)) return rc;
            pts = xyz;
        }
        if (int rc = run_sa(sa[0], pts, pts, 3, 3, Bc, (int)N, xyz1, s, l1, 132, nullptr, st)) return rc;
        if (int rc = launch_copy_cols(xyz1, 3, Bc * S1, 3, l1 + 128, 132, 4, st)) return rc;
        if (int rc = run_sa(sa[1], xyz1, l1, 132, 132, Bc, S1, xyz2, s, l2, 260, nullptr, st)) return rc;
        if (int rc = launch_copy_cols(xyz2, 3, Bc * S2, 3, l2 + 256, 260, 4, st)) return rc;
        return run_sa(sa[2], xyz2, l2, 260, 260, Bc, S2, xyz3, s, nullptr, 0, pooled, st);
    }
};

bool up(PointConvModel *m, const std::vector<float> &h, float **d, std::string &err)
{
    if (m->arena_.upload(h, d)) { err = last_error(); return false; }
    return true;
}

bool make_dense(PointConvModel *m, const StateDict &sd, const std::string &conv, const std::string &bn, int co, int ci,
                Dense &d, std::string &err)
{
    std::vector<float> w, b, hi, lo;
    if (!fold_dense(sd, conv + ".weight", conv + ".bias", bn, co, ci, w, b, err)) return false;
    split_tf32_host(w, hi, lo);
    d.cout = co; d.cin = ci;
    return up(m, w, &d.w, err) && up(m, b, &d.b, err) && up(m, hi, &d.w_hi, err) && up(m, lo, &d.w_lo, err);
}

bool fold_small(const StateDict &sd, const std::string &p, int j, int co, int ci, float *w, float *b, std::string &err)
{
    std::vector<float> wv, bv;
    const std::string n = std::to_string(j);
    if (!fold_dense(sd, p + ".mlp_convs." + n + ".weight", p + ".mlp_convs." + n + ".bias", p + ".mlp_bns." + n, co, ci, wv, bv, err))
        return false;
    for (int i = 0; i < co * ci; ++i) w[i] = wv[i];
    for (int i = 0; i < co; ++i) b[i] = bv[i];
    return true;
}

bool make_sa(PointConvModel *m, const StateDict &sd, const std::string &p, int npoint, int K, int cin_feat, const int (&mlp)[3],
             double bw, bool group_all, SaLayer &L, std::string &err)
{
    L.npoint = npoint; L.K = K; L.cin_feat = cin_feat; L.c1 = mlp[0]; L.c2 = mlp[1]; L.c3 = mlp[2];
    L.bandwidth = bw; L.group_all = group_all;
    const int cin = cin_feat + 3;
    L.ldw = cin_feat ? cin + 1 : 3;
    std::vector<float> w, b;
    if (!fold_dense(sd, p + ".mlp_convs.0.weight", p + ".mlp_convs.0.bias", p + ".mlp_bns.0", L.c1, cin, w, b, err)) return false;
    std::vector<float> w1((size_t)L.c1 * L.ldw, 0.0f), w1x((size_t)L.c1 * 3);
    for (int o = 0; o < L.c1; ++o) {
        for (int i = 0; i < cin_feat; ++i) w1[(size_t)o * L.ldw + i] = w[(size_t)o * cin + 3 + i];
        for (int i = 0; i < 3; ++i) {                                  // xyz is first in the checkpoint, last in our rows
            w1[(size_t)o * L.ldw + cin_feat + i] = w[(size_t)o * cin + i];
            w1x[(size_t)o * 3 + i] = w[(size_t)o * cin + i];
        }
    }
    std::vector<float> w1hi, w1lo;
    split_tf32_host(w1, w1hi, w1lo);
    if (!(up(m, w1, &L.w1, err) && up(m, w1x, &L.w1x, err) && up(m, b, &L.b1, err) && up(m, w1hi, &L.w1_hi, err) &&
          up(m, w1lo, &L.w1_lo, err)))
        return false;
    if (!make_dense(m, sd, p + ".mlp_convs.1", p + ".mlp_bns.1", L.c2, L.c1, L.l2, err)) return false;
    if (!make_dense(m, sd, p + ".mlp_convs.2", p + ".mlp_bns.2", L.c3, L.c2, L.l3, err)) return false;
    if (!make_dense(m, sd, p + ".linear", p + ".bn_linear", L.c3, 16 * L.c3, L.lin, err)) return false;
    SmallNets &n = L.nets;
    return fold_small(sd, p + ".densitynet", 0, 16, 1, n.d1w, n.d1b, err) &&
           fold_small(sd, p + ".densitynet", 1, 8, 16, n.d2w, n.d2b, err) &&
           fold_small(sd, p + ".densitynet", 2, 1, 8, n.d3w, n.d3b, err) &&
           fold_small(sd, p + ".weightnet", 0, 8, 3, n.w1w, n.w1b, err) &&
           fold_small(sd, p + ".weightnet", 1, 8, 8, n.w2w, n.w2b, err) &&
           fold_small(sd, p + ".weightnet", 2, 16, 8, n.w3w, n.w3b, err);
}

}  // namespace

Model *create_pointconv_model(const StateDict &sd, int num_classes, std::string &err)
{
    std::unique_ptr<PointConvModel> m(new PointConvModel());
    m->num_classes = num_classes;
    m->chunk = 330;
    PointConvModel *p = m.get();
    const int m1[3] = {64, 64, 128}, m2[3] = {128, 128, 256}, m3[3] = {256, 512, 1024};
    const bool ok = make_sa(p, sd, "sa1", 512, 32, 0, m1, 0.1, false, m->sa[0], err) &&
                    make_sa(p, sd, "sa2", 128, 64, 128, m2, 0.2, false, m->sa[1], err) &&
                    make_sa(p, sd, "sa3", 1, 0, 256, m3, 0.4, true, m->sa[2], err) &&
                    make_dense(p, sd, "fc1", "bn1", 512, 1024, m->fc1, err) &&
                    make_dense(p, sd, "fc2", "bn2", 256, 512, m->fc2, err) &&
                    make_dense(p, sd, "fc3", "", num_classes, 256, m->fc3, err);
    return ok ? m.release() : nullptr;
}

}  // namespace iq
