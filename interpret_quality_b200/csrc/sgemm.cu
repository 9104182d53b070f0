// Exact-fp32 SIMT GEMM  C = act(alpha * A * B^T + bias)  with fused epilogues.
//
// This is the fp32 (CUDA-core) contraction used for the small / odd-shaped
// products of the masked forward pass (K = 3, 6, heads with M = batch) and as
// the reference precision path for every shared-MLP 1x1 conv
// (nn.Conv1d/Conv2d(k=1) + eval BatchNorm + activation in the reference's
// models/*.py; BN is folded into the weights on the host, see models.cu).
// The tensor-core path for the large products lives in gemm_tc.cu.
//
// Tile 128x128x8, 256 threads, 8x8 outputs per thread split as 2x2 blocks of
// 4x4 so that shared-memory float4 reads are conflict free; register-staged
// double buffering, one __syncthreads per k-tile.
//
// The inner product runs on packed fp32 FMAs (fma.rn.f32x2 -> FFMA2, sm_100): a
// scalar FFMA issues every second cycle per scheduler on this part, so 64 of
// them per k-step cap the kernel at half the fp32 rate; 32 FFMA2 (a[i] broadcast
// x two adjacent columns of b) do the same work.  Each half of an FFMA2 is an
// IEEE fp32 fma and every output keeps its own k-ordered chain, so the results
// are bit-identical to the scalar form (IQ_SGEMM_F2=0 runs that form).
//
// Epilogues:
//   store          C[m][n] = act(alpha*acc + bias[n] + row_bias[m / row_group][n])
//   pool           per 128-row tile: column max (+ lowest-index argmax) and column
//                  sum of the activated values -> global max / average pooling over
//                  the points of a cloud without materialising the (B, C, N) tensor
//                  (torch.max(x, 2), adaptive_{max,avg}_pool1d in the reference).
#include "common.cuh"
#include "kernels.cuh"

namespace iq {

namespace {

constexpr int BM = 128, BN = 128, BK = 8, LDS = BM + 4;
constexpr int SGEMM_F2_DEFAULT = 0;

template <bool kVec>
__device__ __forceinline__ void load_tile(const float *__restrict__ P, int64_t ld, int rows, int K, int r0, int k0,
                                          int tid, float (&reg)[4])
{
    const int row = tid >> 1, kq = (tid & 1) * 4;
    const int gr = r0 + row;
    if (kVec) {
        if (gr < rows && k0 + kq < K) {
            const float4 v = *reinterpret_cast<const float4 *>(P + (int64_t)gr * ld + k0 + kq);
            reg[0] = v.x; reg[1] = v.y; reg[2] = v.z; reg[3] = v.w;
        } else {
            reg[0] = reg[1] = reg[2] = reg[3] = 0.0f;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + kq + j;
            reg[j] = (gr < rows && k < K) ? P[(int64_t)gr * ld + k] : 0.0f;
        }
    }
}

__device__ __forceinline__ void store_tile(float (*S)[LDS], int tid, const float (&reg)[4])
{
    const int row = tid >> 1, kq = (tid & 1) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) S[kq + j][row] = reg[j];
}

template <bool kVec, bool F2>
__global__ void __launch_bounds__(256, 2)
sgemm_kernel(const GemmDesc g)
{
    __shared__ __align__(16) float As[2][BK][LDS];
    __shared__ __align__(16) float Bs[2][BK][LDS];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN, bz = blockIdx.z;
    const float *A = g.A + (int64_t)bz * g.strideA;
    const float *B = g.B + (int64_t)bz * g.strideB;

    float acc[8][8];
    unsigned long long acc2[8][4];                     // F2: columns (2j, 2j+1) of row i packed as one f32x2 accumulator
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc2[i][j] = 0ull;
    }

    float ra[4], rb[4];
    load_tile<kVec>(A, g.lda, g.M, g.K, m0, 0, tid, ra);
    load_tile<kVec>(B, g.ldb, g.N, g.K, n0, 0, tid, rb);
    store_tile(As[0], tid, ra);
    store_tile(Bs[0], tid, rb);
    __syncthreads();

    const int nk = (g.K + BK - 1) / BK;
    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
            load_tile<kVec>(A, g.lda, g.M, g.K, m0, (kt + 1) * BK, tid, ra);
            load_tile<kVec>(B, g.ldb, g.N, g.K, n0, (kt + 1) * BK, tid, rb);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[cur][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[cur][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[cur][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[cur][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            if (F2) {
                const unsigned long long bp[4] = {pack_f2(b[0], b[1]), pack_f2(b[2], b[3]), pack_f2(b[4], b[5]), pack_f2(b[6], b[7])};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const unsigned long long ap = pack_f2(a[i], a[i]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc2[i][j] = fma_f2(ap, bp[j], acc2[i][j]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
        if (kt + 1 < nk) {
            store_tile(As[cur ^ 1], tid, ra);
            store_tile(Bs[cur ^ 1], tid, rb);
        }
        __syncthreads();
    }

    if (F2) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) unpack_f2(acc2[i][j], acc[i][2 * j], acc[i][2 * j + 1]);
    }

    // ---- epilogue
    const float *bias = g.bias ? g.bias + (int64_t)bz * g.strideBias : nullptr;
    float bcol[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
        bcol[j] = (bias && n < g.N) ? bias[n] : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        const float *rbias = (g.row_bias && m < g.M) ? g.row_bias + (int64_t)(m / g.row_group) * g.ld_row_bias : nullptr;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            float v = fmaf(g.alpha, acc[i][j], bcol[j]);
            if (rbias && n < g.N) v += rbias[n];
            acc[i][j] = apply_act(v, g.act);
        }
    }

    if (g.C) {
        float *C = g.C + (int64_t)bz * g.strideC;
        const bool vec_ok = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
            if (m >= g.M) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = n0 + h * 64 + tx * 4;
                float *dst = C + (int64_t)m * g.ldc + n;
                if (vec_ok && n + 3 < g.N) {
                    *reinterpret_cast<float4 *>(dst) =
                        make_float4(acc[i][h * 4], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (n + j < g.N) dst[j] = acc[i][h * 4 + j];
                }
            }
        }
    }

    if (g.C_hi) {
        float *Ch = g.C_hi + (int64_t)bz * g.strideC, *Cl = g.C_lo + (int64_t)bz * g.strideC;
        const bool vec_ok = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(Ch) & 15) == 0) &&
                            ((reinterpret_cast<uintptr_t>(Cl) & 15) == 0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
            if (m >= g.M) continue;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int n = n0 + hh * 64 + tx * 4;
                float hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float v = acc[i][hh * 4 + j];
                    hi[j] = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
                    const float l = v - hi[j];
                    lo[j] = __uint_as_float((__float_as_uint(l) + 0x1000u) & 0xffffe000u);
                }
                float *dh = Ch + (int64_t)m * g.ldc + n, *dl = Cl + (int64_t)m * g.ldc + n;
                if (vec_ok && n + 3 < g.N) {
                    *reinterpret_cast<float4 *>(dh) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4 *>(dl) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (n + j < g.N) { dh[j] = hi[j]; dl[j] = lo[j]; }
                }
            }
        }
    }

    if (g.pool_max) {
        // host guarantees M % 128 == 0, so every row of the tile is valid
        float cmax[8], csum[8];
        int carg[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { cmax[j] = -INFINITY; csum[j] = 0.0f; carg[j] = 0; }
#pragma unroll
        for (int i = 0; i < 8; ++i) {                              // rows visited in ascending order
            const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (acc[i][j] > cmax[j]) { cmax[j] = acc[i][j]; carg[j] = m; }
                csum[j] += acc[i][j];
            }
        }
        float *red_v = &As[0][0][0];                               // 16 x 128 floats each, aliasing the operand tiles
        float *red_s = &Bs[0][0][0];
        __shared__ int red_a[16][BN];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            red_v[ty * BN + c] = cmax[j];
            red_s[ty * BN + c] = csum[j];
            red_a[ty][c] = carg[j];
        }
        __syncthreads();
        if (tid < BN && n0 + tid < g.N) {
            float mv = red_v[tid], sv = red_s[tid];
            int ma = red_a[0][tid];
            for (int t = 1; t < 16; ++t) {
                const float v = red_v[t * BN + tid];
                const int a = red_a[t][tid];
                if (v > mv || (v == mv && a < ma)) { mv = v; ma = a; }
                sv += red_s[t * BN + tid];
            }
            const int64_t o = ((int64_t)bz * gridDim.y + blockIdx.y) * g.N + n0 + tid;
            g.pool_max[o] = mv;
            if (g.pool_arg) g.pool_arg[o] = ma;
            if (g.pool_sum) g.pool_sum[o] = sv;
        }
    }
}

__global__ void pool_finish_kernel(const float *__restrict__ pmax, const int32_t *__restrict__ parg,
                                   const float *__restrict__ psum, int tiles, int rows_per_group, int N,
                                   float *__restrict__ out_max, int64_t ld_max, int64_t *__restrict__ out_arg,
                                   float *__restrict__ out_mean, int64_t ld_mean)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gidx = blockIdx.y;
    if (n >= N) return;
    const int64_t base = gidx * tiles * N + n;
    float mv = pmax[base];
    int ma = parg ? parg[base] : 0;
    float sv = psum ? psum[base] : 0.0f;
    for (int t = 1; t < tiles; ++t) {
        const float v = pmax[base + (int64_t)t * N];
        if (v > mv) { mv = v; if (parg) ma = parg[base + (int64_t)t * N]; }
        if (psum) sv += psum[base + (int64_t)t * N];
    }
    out_max[gidx * ld_max + n] = mv;
    if (out_arg) out_arg[gidx * N + n] = (int64_t)ma - gidx * rows_per_group;
    if (out_mean) out_mean[gidx * ld_mean + n] = sv / (float)rows_per_group;
}

}  // namespace

int launch_sgemm(const GemmDesc &g, cudaStream_t st)
{
    ProfileScope _ps(g.tag, st);
    IQ_CHECK(g.M >= 0 && g.N >= 0 && g.K >= 1 && g.batch >= 1, "sgemm: bad shape");
    if (g.M == 0 || g.N == 0) return 0;
    IQ_CHECK(g.C || g.C_hi || g.pool_max, "sgemm: no output requested");
    if (g.pool_max) IQ_CHECK(g.M % BM == 0, "sgemm: pooling epilogue needs M % 128 == 0");
    IQ_CHECK(g.batch <= 65535 && ceil_div(g.M, BM) <= 65535, "sgemm: grid too large");
    const bool vec = (g.K % 4 == 0) && (g.lda % 4 == 0) && (g.ldb % 4 == 0) && (g.strideA % 4 == 0) &&
                     (g.strideB % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0);
    dim3 grid((unsigned)ceil_div(g.N, BN), (unsigned)ceil_div(g.M, BM), (unsigned)g.batch);
    const bool f2 = env_int("IQ_SGEMM_F2", SGEMM_F2_DEFAULT) != 0;
    if (vec && f2) sgemm_kernel<true, true><<<grid, 256, 0, st>>>(g);
    else if (vec) sgemm_kernel<true, false><<<grid, 256, 0, st>>>(g);
    else if (f2) sgemm_kernel<false, true><<<grid, 256, 0, st>>>(g);
    else sgemm_kernel<false, false><<<grid, 256, 0, st>>>(g);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

int launch_pool_finish(const float *pmax, const int32_t *parg, const float *psum, int64_t groups, int tiles_per_group,
                       int rows_per_group, int N, float *out_max, int64_t ld_max, int64_t *out_arg, float *out_mean,
                       int64_t ld_mean, cudaStream_t st)
{
    ProfileScope _ps("pool_finish", st);
    if (groups == 0) return 0;
    IQ_CHECK(groups <= 65535, "pool_finish: too many groups");
    dim3 grid((unsigned)ceil_div(N, 128), (unsigned)groups);
    pool_finish_kernel<<<grid, 128, 0, st>>>(pmax, parg, psum, tiles_per_group, rows_per_group, N, out_max, ld_max,
                                             out_arg, out_mean, ld_mean);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

}  // namespace iq
