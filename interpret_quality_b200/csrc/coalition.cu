// Coalition masking, reward and marginal-contribution reductions.
//
// Reference behaviour restated (never copied):
//   mask_data_batch            tools/final_common.py:46-61
//   mask_data                  final_shapley_value.py:74-88
//   interaction mask block     final_point_binary_interaction_logits.py:42-56
//   get_reward                 tools/final_common.py:11-24
//   Shapley accumulation       tools/final_common.py:93-97
//   interaction score          final_cal_interactions.py:28-36
//
// All of these are HBM-bound byte movers or tiny reductions: coalesced 16-byte
// stores, staging of the per-permutation lookup in shared memory, grids sized
// from the number of permutations / contexts.
#include "common.cuh"
#include "kernels.cuh"

namespace iq {

// ------------------------------------------------------------------ Shapley masks
// One CTA handles one permutation p and a group of consecutive rows.  The
// rank-of-region table is built once in shared memory; a point is masked in
// row r iff rank[region_id[point]] >= r.  Output is point-major (rows, N, 3),
// written with float4 stores over the flat 3N-float row.
template <bool kInPlace>
__global__ void __launch_bounds__(256)
mask_shapley_kernel(const float *__restrict__ data, const float *__restrict__ center,
                    const int64_t *__restrict__ orders, const int64_t *__restrict__ region_id, int R, int N,
                    int rows_per_cta, float *__restrict__ out)
{
    extern __shared__ unsigned char smem[];
    unsigned char *qpt = smem;                                   // N bytes: rank of each point's region (255 = never)
    __shared__ unsigned char rank[256];
    const int p = blockIdx.x;
    const int row0 = blockIdx.y * rows_per_cta;
    for (int r = threadIdx.x; r < 256; r += blockDim.x) rank[r] = 255;
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const int64_t reg = orders[(int64_t)p * R + r];
        if (reg >= 0 && reg < 256) rank[reg] = (unsigned char)r;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int64_t reg = region_id[i];
        qpt[i] = (reg >= 0 && reg < 256) ? rank[reg] : 255;
    }
    __syncthreads();
    const float c0 = center[0], c1 = center[1], c2 = center[2];
    const int nflt = 3 * N;
    const int nvec = nflt >> 2;                                  // host guarantees 3N % 4 == 0
    const int row_end = min(row0 + rows_per_cta, R + 1);
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        const int f = v << 2;
        float cen[4];
        unsigned char q[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int pt = (f + t) / 3, ch = (f + t) - pt * 3;
            cen[t] = ch == 0 ? c0 : (ch == 1 ? c1 : c2);
            q[t] = qpt[pt];
        }
        if (kInPlace) {
            // reference semantics: only the masked entries are rewritten
            for (int r = row0; r < row_end; ++r) {
                float *o = out + ((int64_t)p * (R + 1) + r) * nflt + f;
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (q[t] != 255 && q[t] >= r) o[t] = cen[t];
            }
        } else {
            const float4 d = reinterpret_cast<const float4 *>(data)[v];
            const float dv[4] = {d.x, d.y, d.z, d.w};
            for (int r = row0; r < row_end; ++r) {
                float4 o;
                o.x = (q[0] != 255 && q[0] >= r) ? cen[0] : dv[0];
                o.y = (q[1] != 255 && q[1] >= r) ? cen[1] : dv[1];
                o.z = (q[2] != 255 && q[2] >= r) ? cen[2] : dv[2];
                o.w = (q[3] != 255 && q[3] >= r) ? cen[3] : dv[3];
                __stcs(reinterpret_cast<float4 *>(out + ((int64_t)p * (R + 1) + r) * nflt) + v, o);
            }
        }
    }
}

int launch_mask_shapley(const float *data, const float *center, const int64_t *orders, const int64_t *region_id,
                        int64_t bs, int64_t R, int64_t N, float *out, bool in_place, cudaStream_t st)
{
    ProfileScope _ps("mask_shapley", st);
    IQ_CHECK(R >= 1 && R <= 255, "mask_shapley: num_regions must be in [1,255]");
    IQ_CHECK(N >= 4 && (3 * N) % 4 == 0, "mask_shapley: num_points must be a multiple of 4");
    IQ_CHECK(N <= 48 * 1024, "mask_shapley: num_points too large for the shared-memory lookup");
    if (bs == 0) return 0;
    // enough CTAs to cover 148 SMs a few times over, each streaming >= 3 rows
    int groups = 1;
    while (bs * groups < 148 * 4 && groups * 3 < (R + 1)) ++groups;
    const int rows_per_cta = (int)ceil_div(R + 1, groups);
    dim3 grid((unsigned)bs, (unsigned)ceil_div(R + 1, rows_per_cta));
    if (in_place)
        mask_shapley_kernel<true><<<grid, 256, N, st>>>(data, center, orders, region_id, (int)R, (int)N, rows_per_cta, out);
    else
        mask_shapley_kernel<false><<<grid, 256, N, st>>>(data, center, orders, region_id, (int)R, (int)N, rows_per_cta, out);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------ interaction masks
// One CTA per context k: clouds 4k..4k+3 keep S+{i,j}, S+{i}, S+{j}, S.  The
// arithmetic is the reference's multiply form, data*mask + center*(1-mask),
// so that -0.0 inputs come out exactly as they do there.  Output is
// channel-first (4*ctx, 3, N) like the reference, or point-major (4*ctx, N, 3)
// when the caller feeds the forward pass directly.
// With `pairs` ((P,2) on the device) the launch covers every pair at once: context k belongs to pair k / ctx_per_pair
// (the reference loops over the pairs on the host, final_point_binary_interaction_logits.py:37).
__global__ void __launch_bounds__(256)
mask_interaction_kernel(const float *__restrict__ data, const float *__restrict__ center,
                        const int64_t *__restrict__ contexts, int m, int region_i, int region_j,
                        const int64_t *__restrict__ pairs, int ctx_per_pair,
                        const int64_t *__restrict__ region_id, int R, int N, int point_major, float *__restrict__ out)
{
    __shared__ unsigned char inS[256];
    const int k = blockIdx.x;
    if (pairs) {
        const int64_t p = k / ctx_per_pair;
        region_i = (int)pairs[2 * p];
        region_j = (int)pairs[2 * p + 1];
    }
    for (int r = threadIdx.x; r < 256; r += blockDim.x) inS[r] = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < m; t += blockDim.x) {
        const int64_t r = contexts[(int64_t)k * m + t];
        if (r >= 0 && r < R) inS[r] = 1;
    }
    __syncthreads();
    const float cen[3] = {center[0], center[1], center[2]};
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int64_t reg = region_id[i];
        const bool s = (reg >= 0 && reg < 256) ? inS[reg] : false;
        const bool is_i = reg == region_i, is_j = reg == region_j;
        const float x[3] = {data[3 * i], data[3 * i + 1], data[3 * i + 2]};
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const bool keep = s || (is_i && (v == 0 || v == 1)) || (is_j && (v == 0 || v == 2));
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float y = __fmul_rn(x[c], keep ? 1.0f : 0.0f);
                y = __fadd_rn(y, keep ? 0.0f : cen[c]);
                const int64_t cloud = (int64_t)k * 4 + v;
                if (point_major) out[(cloud * N + i) * 3 + c] = y;
                else out[(cloud * 3 + c) * N + i] = y;
            }
        }
    }
}

int launch_mask_interaction(const float *data, const float *center, const int64_t *contexts, int64_t ctx, int64_t m,
                            int64_t region_i, int64_t region_j, const int64_t *region_id, int64_t R, int64_t N,
                            int point_major, float *out, cudaStream_t st)
{
    ProfileScope _ps("mask_interaction", st);
    IQ_CHECK(R >= 1 && R <= 255, "mask_interaction: num_regions must be in [1,255]");
    if (ctx == 0) return 0;
    mask_interaction_kernel<<<(unsigned)ctx, 256, 0, st>>>(data, center, contexts, (int)m, (int)region_i, (int)region_j,
                                                          nullptr, 1, region_id, (int)R, (int)N, point_major, out);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

int launch_mask_interaction_pairs(const float *data, const float *center, const int64_t *pairs, const int64_t *contexts,
                                  int64_t P, int64_t ctx, int64_t m, const int64_t *region_id, int64_t R, int64_t N,
                                  int point_major, float *out, cudaStream_t st)
{
    ProfileScope _ps("mask_interaction", st);
    IQ_CHECK(R >= 1 && R <= 255, "mask_interaction: num_regions must be in [1,255]");
    IQ_CHECK(P * ctx < ((int64_t)1 << 31), "mask_interaction: too many contexts for one launch");
    if (P * ctx == 0) return 0;
    mask_interaction_kernel<<<(unsigned)(P * ctx), 256, 0, st>>>(data, center, contexts, (int)m, 0, 0, pairs, (int)ctx,
                                                                region_id, (int)R, (int)N, point_major, out);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------ reward
// modified: z_y - logsumexp(z_{!=y})  (= log p/(1-p));  normal: log_softmax(z)_y.
// libm expf / logf on purpose: interactions are differences of four rewards of magnitude ~10 that nearly cancel, and the
// SFU forms (ex2.approx / lg2.approx) moved them by 1e-5 of their scale.  With 10 exact exponentials per 40 bytes these
// kernels are ALU bound (0.15-0.3 of the copy bandwidth when timed alone); they are <0.05 % of a step.
__device__ __forceinline__ float reward_of_row(const float *z, int C, int lbl, int softmax_normal)
{
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c)
        if (softmax_normal || c != lbl) mx = fmaxf(mx, z[c]);
    float s = 0.0f;
    for (int c = 0; c < C; ++c)
        if (softmax_normal || c != lbl) s += expf(z[c] - mx);
    // log_softmax subtracts the max first; logsumexp adds it back last
    return softmax_normal ? (z[lbl] - mx) - logf(s) : z[lbl] - (logf(s) + mx);
}

// A row of logits is 40 bytes (160 for the 4 clouds of a context): read one row per thread, a warp touches 10-40 lines per
// load instruction and L1 throughput, not HBM, bounds the kernel.  The block's rows are contiguous, so they are copied
// with coalesced loads into shared memory (row stride odd: conflict-free) and evaluated from there.
__device__ __forceinline__ void stage_rows(const float *__restrict__ src, int64_t first_row, int64_t total_rows, int row_len,
                                           float *smem)
{
    const int rows = (int)min((int64_t)blockDim.x, total_rows - first_row);
    const float *base = src + first_row * row_len;
    const int n = rows * row_len, stride = row_len | 1;
    if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0) {        // 16-byte loads, 4x the bytes in flight
        for (int t = threadIdx.x; t < (n >> 2); t += blockDim.x) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(base) + t);
            const float e[4] = {v.x, v.y, v.z, v.w};
            int r = (4 * t) / row_len, c = 4 * t - r * row_len;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                smem[r * stride + c] = e[j];
                if (++c == row_len) { c = 0; ++r; }
            }
        }
    } else {
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            const int r = t / row_len, c = t - r * row_len;
            smem[r * stride + c] = __ldg(base + t);
        }
    }
    __syncthreads();
}

__global__ void reward_kernel(const float *__restrict__ logits, int64_t B, int C, int lbl, int softmax_normal,
                              float *__restrict__ v)
{
    extern __shared__ float rows_s[];
    const int64_t first = blockIdx.x * (int64_t)blockDim.x;
    stage_rows(logits, first, B, C, rows_s);
    const int64_t b = first + threadIdx.x;
    if (b < B) v[b] = reward_of_row(rows_s + threadIdx.x * (C | 1), C, lbl, softmax_normal);
}

int launch_reward(const float *logits, int64_t B, int64_t C, int64_t lbl, int softmax_normal, float *v, cudaStream_t st)
{
    ProfileScope _ps("reward", st);
    IQ_CHECK(C >= 2 && C <= 256 && lbl >= 0 && lbl < C, "reward: label out of range");
    if (B == 0) return 0;
    const size_t smem = sizeof(float) * 128 * (size_t)(C | 1);
    if (smem > 48 * 1024)
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(&reward_kernel), (int)smem)) return rc;
    reward_kernel<<<(unsigned)ceil_div(B, 128), 128, smem, st>>>(logits, B, (int)C, (int)lbl, softmax_normal, v);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------ Shapley accumulation
// phi_sum[order[p][r]] += double(v[p][r+1] - v[p][r]), permutations visited in
// ascending p for every region, i.e. the same float64 addition sequence as the
// reference's numpy accumulator.  One CTA; the work is tiny.
__global__ void __launch_bounds__(1024)
shapley_accumulate_kernel(const float *__restrict__ v, const int64_t *__restrict__ orders, int bs, int R,
                          double *__restrict__ phi_sum)
{
    extern __shared__ float dv_by_region[];                       // [tile][R]
    const int tile = (int)blockDim.x / R;
    for (int p0 = 0; p0 < bs; p0 += tile) {
        const int np = min(tile, bs - p0);
        for (int t = threadIdx.x; t < np * R; t += blockDim.x) dv_by_region[t] = 0.0f;
        __syncthreads();
        for (int t = threadIdx.x; t < np * R; t += blockDim.x) {
            const int p = p0 + t / R, r = t % R;
            const float d = __fsub_rn(v[(int64_t)p * (R + 1) + r + 1], v[(int64_t)p * (R + 1) + r]);
            const int64_t reg = orders[(int64_t)p * R + r];
            if (reg >= 0 && reg < R) dv_by_region[(p - p0) * R + (int)reg] = d;
        }
        __syncthreads();
        if ((int)threadIdx.x < R) {
            double acc = phi_sum[threadIdx.x];
            for (int q = 0; q < np; ++q) acc += (double)dv_by_region[q * R + threadIdx.x];
            phi_sum[threadIdx.x] = acc;
        }
        __syncthreads();
    }
}

int launch_shapley_accumulate(const float *v, const int64_t *orders, int64_t bs, int64_t R, double *phi_sum,
                              cudaStream_t st)
{
    ProfileScope _ps("shapley_accumulate", st);
    IQ_CHECK(R >= 1 && R <= 255, "shapley_accumulate: num_regions must be in [1,255]");
    if (bs == 0) return 0;
    const int threads = (int)((1024 / R) * R);
    const size_t smem = sizeof(float) * (size_t)threads;
    shapley_accumulate_kernel<<<1, threads, smem, st>>>(v, orders, (int)bs, (int)R, phi_sum);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------ interaction score
// out[p][k] = double((v[4k] + v[4k+3]) - v[4k+1] - v[4k+2]) with the fp32
// operation order of final_cal_interactions.py:33.
__global__ void interaction_reduce_kernel(const float *__restrict__ logits, int64_t total, int C, int lbl,
                                          int softmax_normal, double *__restrict__ out)
{
    extern __shared__ float rows_s[];
    const int64_t first = blockIdx.x * (int64_t)blockDim.x;
    stage_rows(logits, first, total, 4 * C, rows_s);
    const int64_t t = first + threadIdx.x;
    if (t >= total) return;
    const float *z = rows_s + threadIdx.x * ((4 * C) | 1);
    const float v0 = reward_of_row(z, C, lbl, softmax_normal);
    const float v1 = reward_of_row(z + C, C, lbl, softmax_normal);
    const float v2 = reward_of_row(z + 2 * C, C, lbl, softmax_normal);
    const float v3 = reward_of_row(z + 3 * C, C, lbl, softmax_normal);
    out[t] = (double)__fsub_rn(__fsub_rn(__fadd_rn(v0, v3), v1), v2);
}

int launch_interaction_reduce(const float *logits, int64_t P, int64_t ctx, int64_t C, int64_t lbl, int softmax_normal,
                              double *out, cudaStream_t st)
{
    ProfileScope _ps("interaction_reduce", st);
    IQ_CHECK(C >= 2 && C <= 256 && lbl >= 0 && lbl < C, "interaction_reduce: label out of range");
    const int64_t total = P * ctx;
    if (total == 0) return 0;
    const int threads = C <= 64 ? 128 : 32;
    const size_t smem = sizeof(float) * threads * (size_t)((4 * C) | 1);
    if (smem > 48 * 1024)
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(&interaction_reduce_kernel), (int)smem)) return rc;
    interaction_reduce_kernel<<<(unsigned)ceil_div(total, threads), threads, smem, st>>>(logits, total, (int)C, (int)lbl,
                                                                                        softmax_normal, out);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

}  // namespace iq
