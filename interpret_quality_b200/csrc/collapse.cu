// Coalition collapse: the masking rule of the reference (tools/final_common.py:56-60,
// final_point_binary_interaction_logits.py:42-56) moves every point of an absent region onto ONE
// location (`center`), so a masked cloud is its kept points plus M coincident copies of that location.
// Coincident points have identical rows in every layer of a point-wise / k-nearest-neighbour network:
//   * max pooling and the max over a neighbourhood ignore multiplicity;
//   * a k-nearest-neighbour list only sees min(M, k) of the copies (the k+1-th copy can never be chosen
//     before the k-th), so a cloud with min(M, k) <= m <= M copies has the same neighbour SETS;
//   * average pooling weights the copy row by M (carried as `extra` = M - m for the pooling epilogue).
// These kernels count the kept points of every cloud, and rewrite the clouds -- grouped by their
// compacted size n = round_up(U + min(M, copies), 128) -- as U kept points (original order) followed by
// n - U copies of the location, so the forward runs on n instead of N points per cloud, exactly.
#include "common.cuh"
#include "kernels.cuh"

namespace iq {

namespace {

__device__ __forceinline__ void load_point(const float *__restrict__ x, int point_major, int64_t cloud, int N, int i,
                                           float &a, float &b, float &c)
{
    if (point_major) {
        const float *p = x + (cloud * N + i) * 3;
        a = p[0]; b = p[1]; c = p[2];
    } else {
        const float *p = x + cloud * 3 * N + i;
        a = p[0]; b = p[N]; c = p[2 * (int64_t)N];
    }
}

// kept[b] = number of points of cloud b that differ from `loc` (float compare: -0.0 == +0.0)
__global__ void __launch_bounds__(256)
collapse_count_kernel(const float *__restrict__ x, int point_major, int N, const float *__restrict__ loc,
                      int32_t *__restrict__ kept)
{
    __shared__ int warp_cnt[8];
    const float l0 = loc[0], l1 = loc[1], l2 = loc[2];
    int n = 0;
    for (int i = threadIdx.x; i < N; i += 256) {
        float a, b, c;
        load_point(x, point_major, blockIdx.x, N, i, a, b, c);
        n += !(a == l0 && b == l1 && c == l2);
    }
    n = __reduce_add_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += warp_cnt[w];
        kept[blockIdx.x] = t;
    }
}

// One CTA per sorted position s: cloud src[s] -> rows [row_off[s], row_off[s] + size[s]) of the point-major output:
// its kept points in their original order, then copies of `loc`.
__global__ void __launch_bounds__(256)
collapse_compact_kernel(const float *__restrict__ x, int point_major, int N, const float *__restrict__ loc,
                        const int32_t *__restrict__ src, const int32_t *__restrict__ row_off,
                        const int32_t *__restrict__ size, float *__restrict__ out)
{
    __shared__ int warp_cnt[8];
    __shared__ int base_s;
    const int s = blockIdx.x;
    const int64_t cloud = src[s];
    float *o = out + (int64_t)row_off[s] * 3;
    const int n = size[s];
    const float l0 = loc[0], l1 = loc[1], l2 = loc[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    for (int i0 = 0; i0 < N; i0 += 256) {
        const int i = i0 + threadIdx.x;
        float a = 0.0f, b = 0.0f, c = 0.0f;
        bool keep = false;
        if (i < N) {
            load_point(x, point_major, cloud, N, i, a, b, c);
            keep = !(a == l0 && b == l1 && c == l2);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int pos = base_s + __popc(bal & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) pos += warp_cnt[w];
        if (keep && pos < n) { o[3 * pos] = a; o[3 * pos + 1] = b; o[3 * pos + 2] = c; }
        __syncthreads();
        if (threadIdx.x == 255) base_s = pos + (keep ? 1 : 0);     // last thread: its position + its own point = running total
        __syncthreads();
    }
    for (int i = base_s + threadIdx.x; i < n; i += 256) { o[3 * i] = l0; o[3 * i + 1] = l1; o[3 * i + 2] = l2; }
}

// out[dst[s]][:] = in[s][:]  (rows of C floats): the logits back in the caller's cloud order
__global__ void scatter_rows_kernel(const float *__restrict__ in, const int32_t *__restrict__ dst, int64_t rows, int C,
                                    float *__restrict__ out)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * C) return;
    const int64_t s = t / C;
    const int c = (int)(t - s * C);
    out[(int64_t)dst[s] * C + c] = in[t];
}

}  // namespace

int launch_collapse_count(const float *x, int point_major, int64_t B, int64_t N, const float *loc, int32_t *kept,
                          cudaStream_t st)
{
    ProfileScope _ps("collapse_count", st);
    if (B == 0) return 0;
    collapse_count_kernel<<<(unsigned)B, 256, 0, st>>>(x, point_major, (int)N, loc, kept);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

int launch_collapse_compact(const float *x, int point_major, int64_t B, int64_t N, const float *loc, const int32_t *src,
                            const int32_t *row_off, const int32_t *size, float *out, cudaStream_t st)
{
    ProfileScope _ps("collapse_compact", st);
    if (B == 0) return 0;
    collapse_compact_kernel<<<(unsigned)B, 256, 0, st>>>(x, point_major, (int)N, loc, src, row_off, size, out);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

int launch_scatter_rows(const float *in, const int32_t *dst, int64_t rows, int C, float *out, cudaStream_t st)
{
    ProfileScope _ps("scatter_rows", st);
    if (rows == 0) return 0;
    scatter_rows_kernel<<<(unsigned)ceil_div(rows * C, 256), 256, 0, st>>>(in, dst, rows, C, out);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

}  // namespace iq
