// Neighbourhood construction shared by PointNet++ (MSG) and PointConv.
//
// Reference behaviour restated (never copied):
//   query_ball_point      models/pointnet2.py:70-91  first K indices in index order with d^2 <= r^2, padded with the first
//   index_points + "grouped - centre" + first shared-MLP layer      models/pointnet2.py:221-236, models/pointconv.py:128-137
//
// The first 1x1 conv of every grouped MLP is linear in its input [f_j ; x_j - c_i], so it is evaluated once per
// point (U = W [f ; x]) and once per centroid (V = W_x c) by dense GEMMs, and the grouped activation is
//   H1[i][j] = act(U[idx[i][j]] - V[i] + b)
// produced by the gather kernel below, already split into tf32 hi/lo for the tcgen05 layers that follow.
#include "common.cuh"
#include "kernels.cuh"

namespace iq {

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// one warp per centroid: ordered ballot compaction over the points, 32 at a time
__global__ void __launch_bounds__(256)
ball_query_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz, int N, int S, float r2, int K,
                  int32_t *__restrict__ idx)
{
    extern __shared__ float4 pts[];                               // N x (x, y, z, |p|^2)
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *p = xyz + (int64_t)b * N * 3;
    for (int i = threadIdx.x; i < N; i += 256) {
        const float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
        pts[i] = make_float4(x, y, z, __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    }
    __syncthreads();
    const int per_cta = 64;
    for (int s = blockIdx.x * per_cta + warp; s < min(S, (int)(blockIdx.x + 1) * per_cta); s += 8) {
        const float *c = new_xyz + ((int64_t)b * S + s) * 3;
        const float cx = c[0], cy = c[1], cz = c[2];
        const float cc = __fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz));
        int32_t *o = idx + ((int64_t)b * S + s) * K;
        int cnt = 0, first = N;
        for (int j0 = 0; j0 < N && cnt < K; j0 += 32) {
            const int j = j0 + lane;
            bool in = false;
            if (j < N) {
                const float4 q = pts[j];
                float dot = __fmul_rn(cx, q.x);
                dot = __fmaf_rn(cy, q.y, dot);
                dot = __fmaf_rn(cz, q.z, dot);
                float t = __fmul_rn(-2.0f, dot);
                t = __fadd_rn(t, cc);
                t = __fadd_rn(t, q.w);
                in = !(t > r2);
            }
            const unsigned bal = __ballot_sync(FULL, in);
            if (bal && first == N) first = j0 + __ffs(bal) - 1;
            const int pos = cnt + __popc(bal & lanemask_lt());
            if (in && pos < K) o[pos] = j;
            cnt += __popc(bal);
        }
        cnt = min(cnt, K);
        for (int t = cnt + lane; t < K; t += 32) o[t] = first;
    }
}

// H1[(row), c] = act(U[cloud_base + idx[row]][c] - V[row / K][c] + bias[c]),  4 channels per thread
__global__ void __launch_bounds__(256)
group_sub_act_kernel(const float *__restrict__ U, int64_t ldu, const float *__restrict__ V, int64_t ldv,
                     const float *__restrict__ bias, const int32_t *__restrict__ idx, int64_t rows, int K, int S,
                     int Nsrc, int C, int act, float *__restrict__ out, float *__restrict__ out_hi,
                     float *__restrict__ out_lo, int64_t ldo)
{
    const int cq = C >> 2;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * cq) return;
    const int64_t row = t / cq;
    const int c = (int)(t - row * cq) << 2;
    const int64_t cen = row / K;                                    // (cloud, centroid)
    const int64_t cloud = cen / S;
    const int64_t src = cloud * Nsrc + idx[row];
    const float4 u = *reinterpret_cast<const float4 *>(U + src * ldu + c);
    const float4 v = *reinterpret_cast<const float4 *>(V + cen * ldv + c);
    const float4 b = *reinterpret_cast<const float4 *>(bias + c);
    float r[4] = {apply_act((u.x - v.x) + b.x, act), apply_act((u.y - v.y) + b.y, act),
                  apply_act((u.z - v.z) + b.z, act), apply_act((u.w - v.w) + b.w, act)};
    if (out) *reinterpret_cast<float4 *>(out + row * ldo + c) = make_float4(r[0], r[1], r[2], r[3]);
    if (out_hi) {
        float h[4], l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            h[q] = __uint_as_float((__float_as_uint(r[q]) + 0x1000u) & 0xffffe000u);
            const float d = r[q] - h[q];
            l[q] = __uint_as_float((__float_as_uint(d) + 0x1000u) & 0xffffe000u);
        }
        *reinterpret_cast<float4 *>(out_hi + row * ldo + c) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4 *>(out_lo + row * ldo + c) = make_float4(l[0], l[1], l[2], l[3]);
    }
}

// out[g][c] = max over the K consecutive rows of group g (torch.max(new_points, 2) in the reference)
__global__ void group_max_kernel(const float *__restrict__ in, int64_t ld, int64_t groups, int K, int C,
                                 float *__restrict__ out, int64_t ldo)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= groups * C) return;
    const int64_t g = t / C;
    const int c = (int)(t - g * C);
    float m = -INFINITY;
    for (int j = 0; j < K; ++j) m = fmaxf(m, in[(g * K + j) * ld + c]);
    out[g * ldo + c] = m;
}

__global__ void copy_cols_kernel(const float *__restrict__ src, int64_t lds, int64_t rows, int cols,
                                 float *__restrict__ dst, int64_t ldd, int pad_to)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * pad_to) return;
    const int64_t r = t / pad_to;
    const int c = (int)(t - r * pad_to);
    dst[r * ldd + c] = c < cols ? src[r * lds + c] : 0.0f;
}

}  // namespace

int launch_ball_query(const float *xyz, const float *new_xyz, int64_t B, int64_t N, int64_t S, double radius, int K,
                      int32_t *idx, cudaStream_t st)
{
    ProfileScope _ps("ball_query", st);
    IQ_CHECK(N >= 1 && N <= 2048, "ball_query: num_points must be in [1,2048]");
    IQ_CHECK(K >= 1 && B <= 65535, "ball_query: bad K or batch");
    if (B * S == 0) return 0;
    const float r2 = (float)(radius * radius);       // torch compares the fp32 distances with the scalar in fp32
    dim3 grid((unsigned)ceil_div(S, 64), (unsigned)B);
    ball_query_kernel<<<grid, 256, sizeof(float4) * (size_t)N, st>>>(xyz, new_xyz, (int)N, (int)S, r2, K, idx);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

int launch_group_sub_act(const float *U, int64_t ldu, const float *V, int64_t ldv, const float *bias,
                         const int32_t *idx, int64_t B, int S, int K, int Nsrc, int C, int act, float *out,
                         float *out_hi, float *out_lo, int64_t ldo, cudaStream_t st)
{
    ProfileScope _ps("group_sub_act", st);
    IQ_CHECK(C % 4 == 0 && ldu % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0, "group_sub_act: widths must be multiples of 4");
    const int64_t rows = B * S * K;
    if (rows == 0) return 0;
    const int64_t n = rows * (C / 4);
    group_sub_act_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(U, ldu, V, ldv, bias, idx, rows, K, S, Nsrc, C, act,
                                                                    out, out_hi, out_lo, ldo);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

int launch_group_max(const float *in, int64_t ld, int64_t groups, int K, int C, float *out, int64_t ldo, cudaStream_t st)
{
    ProfileScope _ps("group_max", st);
    if (groups == 0) return 0;
    group_max_kernel<<<(unsigned)ceil_div(groups * C, 256), 256, 0, st>>>(in, ld, groups, K, C, out, ldo);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

int launch_copy_cols(const float *src, int64_t lds, int64_t rows, int cols, float *dst, int64_t ldd, int pad_to,
                     cudaStream_t st)
{
    ProfileScope _ps("copy_cols", st);
    if (rows == 0) return 0;
    copy_cols_kernel<<<(unsigned)ceil_div(rows * pad_to, 256), 256, 0, st>>>(src, lds, rows, cols, dst, ldd, pad_to);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

}  // namespace iq
