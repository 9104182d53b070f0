// kNN graph construction and EdgeConv aggregation for DGCNN / GCNN.
//
// Reference behaviour restated (never copied):
//   knn                 models/dgcnn.py:12-18   top-k of  -|xj|^2 + 2 xi.xj - |xi|^2
//   get_graph_feature   models/dgcnn.py:21-47   [x_j - x_i ; x_i] per edge
//   conv + BN + LeakyReLU + max over k   models/dgcnn.py:92-105
//
// B200-first restructuring (DESIGN.md section "EdgeConv"): because the 1x1 conv is
// linear and LeakyReLU is monotone,
//     max_j lrelu(s*(Wa (x_j - x_i) + Wb x_i) + t) = lrelu(max_j P_j + Q_i),
//     P = (s*Wa) x,  Q = (s*(Wb - Wa)) x + t,
// so the (B, 2C, N, k) edge tensor is never materialised: one dense per-point GEMM
// (sgemm.cu / gemm_tc.cu) followed by the gather-max kernel below.
//
// Top-k is exact: the k-th largest key of a row is found per warp (lane maxima ->
// lower bound -> compaction of the few keys above it in shared memory -> rank
// select; bitwise radix descent as the general fallback) and ties at the boundary
// are broken by lowest index.  Masked clouds collapse hundreds of points onto one
// coordinate, so huge tie groups are the common case, not the corner case.
#include "common.cuh"
#include "kernels.cuh"

namespace iq {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int TOPK_SCRATCH = 64;

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ---- warp top-k ---------------------------------------------------------------------------------
// k largest of the 32*V keys of a warp; lane l holds candidates j = v*32 + l.  Writes k candidate
// indices (unordered) to out[0..k).  scratch: TOPK_SCRATCH floats per warp.
//
// Instruction budget matters (ncu: the first version spent ~1900 warp instructions per row, all in
// ballot/popc chains), so membership is tracked in per-lane bit masks (bit v <=> key[v]), totals use
// REDUX (__reduce_*_sync), positions come from one 5-step warp scan, and the ordered ballot walk is
// only used to break ties at the boundary by lowest index.
template <int V> struct LaneMask { typedef uint32_t type; };
template <> struct LaneMask<64> { typedef uint64_t type; };
__device__ __forceinline__ int mask_popc(uint32_t m) { return __popc(m); }
__device__ __forceinline__ int mask_popc(uint64_t m) { return __popcll(m); }
__device__ __forceinline__ int mask_ffs(uint32_t m) { return __ffs(m) - 1; }
__device__ __forceinline__ int mask_ffs(uint64_t m) { return __ffsll((long long)m) - 1; }

__device__ __forceinline__ int warp_exclusive_scan(int x, int lane)
{
    int incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += y;
    }
    return incl - x;
}

// value of rank `r` (0 = largest) among the 32 per-lane values `u` (ordered-uint keys), by REDUX extraction
// from whichever end is closer
__device__ __forceinline__ uint32_t warp_select_rank(uint32_t u, int r, int lane)
{
    uint32_t res = 0;
    if (r < 16) {
        for (int it = 0; it <= r; ++it) {
            res = __reduce_max_sync(FULL, u);
            const unsigned b = __ballot_sync(FULL, u == res);
            if (lane == __ffs(b) - 1) u = 0u;                       // remove one instance (0 is below every real key)
        }
    } else {
        for (int it = 0; it <= 31 - r; ++it) {
            res = __reduce_min_sync(FULL, u);
            const unsigned b = __ballot_sync(FULL, u == res);
            if (lane == __ffs(b) - 1) u = 0xffffffffu;
        }
    }
    return res;
}

// descending bitonic sort of one 32-bit key per lane
__device__ __forceinline__ uint32_t warp_sort_desc(uint32_t u, int lane)
{
#pragma unroll
    for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
        for (int jj = kk >> 1; jj > 0; jj >>= 1) {
            const uint32_t o = __shfl_xor_sync(FULL, u, jj);
            const bool keep_max = ((lane & kk) == 0) == ((lane & jj) == 0);
            u = keep_max ? max(u, o) : min(u, o);
        }
    }
    return u;
}

template <int V>
__device__ __forceinline__ void warp_topk(const float (&key)[V], int k, int lane, int32_t *out, float *scratch)
{
    typedef typename LaneMask<V>::type mask_t;
    float T = 0.0f;
    bool have = false;
    if (k <= 32) {
        // 64 block maxima (two per lane), each half sorted over the lanes by a 15-step bitonic network; the k-th
        // largest of the union of two descending lists a, b is min_i max(a[i], b[k-1-i]): at least k keys are >= T0
        float ma = key[0], mb = key[V / 2];
#pragma unroll
        for (int v = 1; v < V / 2; ++v) { ma = fmaxf(ma, key[v]); mb = fmaxf(mb, key[V / 2 + v]); }
        const uint32_t sa = warp_sort_desc(ordered_u32(ma), lane);
        const uint32_t sb = warp_sort_desc(ordered_u32(mb), lane);
        const uint32_t bo = __shfl_sync(FULL, sb, (k - 1 - lane) & 31);
        const float T0 = from_ordered_u32(__reduce_min_sync(FULL, lane < k ? max(sa, bo) : 0xffffffffu));
        mask_t mg = 0;
#pragma unroll
        for (int v = 0; v < V; ++v) mg |= (mask_t)(key[v] > T0 ? 1 : 0) << v;
        const int g = mask_popc(mg);
        const int G = __reduce_add_sync(FULL, g);
        if (G < k) {
            T = T0;                                                // fewer than k above T0, at least k at or above it
            have = true;
        } else if (G <= 32) {
            // the k-th largest key is among the G keys above T0: one per lane, sorted, rank k-1
            const int off = warp_exclusive_scan(g, lane);
            int w = off;
#pragma unroll
            for (int v = 0; v < V; ++v)
                if ((mg >> v) & 1) scratch[w++] = key[v];
            __syncwarp();
            const uint32_t c = warp_sort_desc(lane < G ? ordered_u32(scratch[lane]) : 0u, lane);
            T = from_ordered_u32(__shfl_sync(FULL, c, k - 1));
            have = true;
            __syncwarp();
        }
    }
    if (!have) {                                                   // exact radix descent on the ordered bit pattern
        uint32_t tu = 0;
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t cand = tu | (1u << bit);
            int c = 0;
#pragma unroll
            for (int v = 0; v < V; ++v) c += ordered_u32(key[v]) >= cand ? 1 : 0;
            if (__reduce_add_sync(FULL, c) >= k) tu = cand;
        }
        T = from_ordered_u32(tu);
    }
    // ---- emission: everything above T, then the lowest-index keys equal to T
    mask_t mgt = 0, meq = 0;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        mgt |= (mask_t)(key[v] > T ? 1 : 0) << v;
        meq |= (mask_t)(key[v] == T ? 1 : 0) << v;
    }
    const int cgt = mask_popc(mgt), ceq = mask_popc(meq);
    const int G2 = __reduce_add_sync(FULL, cgt);
    const int E = __reduce_add_sync(FULL, ceq);
    const int need = k - G2;
    int w = warp_exclusive_scan(cgt, lane);
    while (mgt) {
        const int v = mask_ffs(mgt);
        mgt &= mgt - 1;
        out[w++] = v * 32 + lane;
    }
    if (E == need) {                                               // no tie at the boundary: take them all
        w = G2 + warp_exclusive_scan(ceq, lane);
        while (meq) {
            const int v = mask_ffs(meq);
            meq &= meq - 1;
            out[w++] = v * 32 + lane;
        }
    } else {                                                       // ties: first `need` in index order (v major, lane minor)
        int base = 0;
#pragma unroll 1
        for (int v = 0; v < V && base < need; ++v) {
            const bool eq = (meq >> v) & 1;
            const unsigned b = __ballot_sync(FULL, eq);
            const int slot = base + __popc(b & lanemask_lt());
            if (eq && slot < need) out[G2 + slot] = v * 32 + lane;
            base += __popc(b);
        }
    }
}

// ---- layer-1 kNN on raw coordinates: distances on the fly, exact fp32 recipe
template <int V>
__global__ void __launch_bounds__(256)
knn_xyz_kernel(const float *__restrict__ xyz, int point_major, int N, int k, int rows_per_cta, int32_t *__restrict__ idx)
{
    extern __shared__ float4 pts[];                               // N x (x, y, z, |p|^2)
    __shared__ float scratch[8][TOPK_SCRATCH];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *p = xyz + (int64_t)b * N * 3;
    for (int i = tid; i < N; i += 256) {
        float x, y, z;
        if (point_major) { x = p[3 * i]; y = p[3 * i + 1]; z = p[3 * i + 2]; }
        else { x = p[i]; y = p[N + i]; z = p[2 * N + i]; }
        const float xx = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
        pts[i] = make_float4(x, y, z, xx);
    }
    __syncthreads();
    const int row_begin = blockIdx.x * rows_per_cta;
    const int row_end = min(row_begin + rows_per_cta, N);
    for (int i = row_begin + warp; i < row_end; i += 8) {
        const float4 q = pts[i];
        float key[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int j = v * 32 + lane;
            if (j < N) {
                const float4 c = pts[j];
                float dot = __fmul_rn(q.x, c.x);
                dot = __fmaf_rn(q.y, c.y, dot);
                dot = __fmaf_rn(q.z, c.z, dot);
                const float inner = __fmul_rn(-2.0f, dot);
                key[v] = __fsub_rn(__fsub_rn(-c.w, inner), q.w);
            } else {
                key[v] = -INFINITY;
            }
        }
        warp_topk<V>(key, k, lane, idx + ((int64_t)b * N + i) * k, scratch[warp]);
    }
}

// ---- knn_point (models/pointconv.py:103-114): K nearest source points of every centroid, exact
// square_distance(new_xyz, xyz) recipe, torch.topk(largest=False, sorted=False)
template <int V>
__global__ void __launch_bounds__(256)
knn_point_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz, int N, int S, int k,
                 int32_t *__restrict__ idx)
{
    extern __shared__ float4 pts[];
    __shared__ float scratch[8][TOPK_SCRATCH];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *p = xyz + (int64_t)b * N * 3;
    for (int i = tid; i < N; i += 256) {
        const float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
        pts[i] = make_float4(x, y, z, __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    }
    __syncthreads();
    for (int s = blockIdx.x * 64 + warp; s < min(S, (int)(blockIdx.x + 1) * 64); s += 8) {
        const float *c = new_xyz + ((int64_t)b * S + s) * 3;
        const float cx = c[0], cy = c[1], cz = c[2];
        const float cc = __fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz));
        float key[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int j = v * 32 + lane;
            if (j < N) {
                const float4 q = pts[j];
                float dot = __fmul_rn(cx, q.x);
                dot = __fmaf_rn(cy, q.y, dot);
                dot = __fmaf_rn(cz, q.z, dot);
                float t = __fmul_rn(-2.0f, dot);
                t = __fadd_rn(t, cc);
                key[v] = -__fadd_rn(t, q.w);
            } else {
                key[v] = -INFINITY;
            }
        }
        warp_topk<V>(key, k, lane, idx + ((int64_t)b * S + s) * k, scratch[warp]);
    }
}

// ---- top-k over rows of a key matrix already in memory (feature-space kNN, knn_point)
template <int V>
__global__ void __launch_bounds__(256)
topk_rows_kernel(const float *__restrict__ keys, int64_t rows, int N, int64_t ld, int k, int largest,
                 int32_t *__restrict__ idx)
{
    __shared__ float scratch[8][TOPK_SCRATCH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * 8 + warp;
    if (row >= rows) return;
    const float *kr = keys + row * ld;
    float key[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int j = v * 32 + lane;
        const float x = j < N ? __ldcs(kr + j) : 0.0f;
        key[v] = j < N ? (largest ? x : -x) : -INFINITY;
    }
    warp_topk<V>(key, k, lane, idx + row * k, scratch[warp]);
}

// ---- exact re-rank of kNN candidates --------------------------------------------------------------
// The tcgen05 Gram keys only nominate 32 candidates per row; the final k neighbours are decided here on
// squared distances evaluated directly, sum_c (x_i[c] - x_j[c])^2, fp32 differences accumulated in float64,
// i.e. without the cancellation of the expanded form -|x_i|^2 + 2 x_i.x_j - |x_j|^2 the reference evaluates in
// fp32 (models/dgcnn.py:13-15).  Ties (coincident points) go to the lower point index.  One warp per row,
// lane = candidate; output sorted by (distance, index).
__global__ void __launch_bounds__(256)
knn_rerank_kernel(const float *__restrict__ x, int64_t ld, int C, const int32_t *__restrict__ cand, int64_t rows,
                  int N, int k, int32_t *__restrict__ idx)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const int64_t cloud0 = (row / N) * N;
    const int j = cand[row * 32 + lane];
    const float4 *xi = reinterpret_cast<const float4 *>(x + row * ld);
    const float4 *xj = reinterpret_cast<const float4 *>(x + (cloud0 + j) * ld);
    double acc = 0.0;
    for (int c = 0; c < (C >> 2); ++c) {
        const float4 a = __ldg(xi + c), b = __ldg(xj + c);
        const double d0 = (double)(a.x - b.x), d1 = (double)(a.y - b.y), d2 = (double)(a.z - b.z), d3 = (double)(a.w - b.w);
        acc = fma(d0, d0, acc); acc = fma(d1, d1, acc); acc = fma(d2, d2, acc); acc = fma(d3, d3, acc);
    }
    int rank = 0;
#pragma unroll 8
    for (int s = 0; s < 32; ++s) {
        const double o = __shfl_sync(FULL, acc, s);
        const int oj = __shfl_sync(FULL, j, s);
        rank += (o < acc || (o == acc && oj < j)) ? 1 : 0;
    }
    if (rank < k) idx[row * k + rank] = j;
}

__global__ void sqnorm_rows_kernel(const float *__restrict__ x, int64_t rows, int C, int64_t ld, float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    float s = 0.0f;
    for (int c = lane; c < C; c += 32) { const float v = x[row * ld + c]; s = fmaf(v, v, s); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if (lane == 0) out[row] = -s;
}

// ---- EdgeConv aggregation: out[i][c] = act(max_j P[idx[i][j]][c] + Q[i][c]); warp per point
template <int CPL>
__global__ void __launch_bounds__(256)
gather_max_kernel(const float *__restrict__ PQ, int64_t ldpq, const int32_t *__restrict__ idx, int64_t total, int N,
                  int k, int act, float *__restrict__ out, int64_t ldo, float *__restrict__ neg_sqnorm,
                  float *__restrict__ out_hi, float *__restrict__ out_lo, __half *__restrict__ h_hi,
                  __half *__restrict__ h_lo, int64_t ldh, float hscale)
{
    constexpr int Cout = CPL * 32;
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= total) return;
    const int64_t cloud0 = (i / N) * N;
    float mx[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) mx[c] = -INFINITY;
    for (int j0 = 0; j0 < k; j0 += 32) {
        const int nj = min(32, k - j0);
        const int my = lane < nj ? idx[i * k + j0 + lane] : 0;
#pragma unroll 4
        for (int j = 0; j < nj; ++j) {
            const int64_t src = cloud0 + __shfl_sync(FULL, my, j);
            const float *pr = PQ + src * ldpq + lane * CPL;
            if (CPL == 2) {
                const float2 v = *reinterpret_cast<const float2 *>(pr);
                mx[0] = fmaxf(mx[0], v.x); mx[1] = fmaxf(mx[1], v.y);
            } else {
#pragma unroll
                for (int q = 0; q < CPL / 4; ++q) {
                    const float4 v = *reinterpret_cast<const float4 *>(pr + 4 * q);
                    mx[4 * q] = fmaxf(mx[4 * q], v.x); mx[4 * q + 1] = fmaxf(mx[4 * q + 1], v.y);
                    mx[4 * q + 2] = fmaxf(mx[4 * q + 2], v.z); mx[4 * q + 3] = fmaxf(mx[4 * q + 3], v.w);
                }
            }
        }
    }
    const float *qr = PQ + i * ldpq + Cout + lane * CPL;
    float ss = 0.0f;
    float res[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        res[c] = apply_act(mx[c] + qr[c], act);
        ss = fmaf(res[c], res[c], ss);
    }
    if (out) {
        float *o = out + i * ldo + lane * CPL;
        if (CPL == 2) {
            *reinterpret_cast<float2 *>(o) = make_float2(res[0], res[1]);
        } else {
#pragma unroll
            for (int q = 0; q < CPL / 4; ++q)
                *reinterpret_cast<float4 *>(o + 4 * q) = make_float4(res[4 * q], res[4 * q + 1], res[4 * q + 2], res[4 * q + 3]);
        }
    }
    if (h_hi) {                                                     // two-term fp16 split for the kind::f16 consumers
        uint32_t wh[CPL / 2], wl[CPL / 2];
#pragma unroll
        for (int c = 0; c < CPL; c += 2) {
            __half a, b, la, lb;
            split_f16(res[c], hscale, a, la);
            split_f16(res[c + 1], hscale, b, lb);
            wh[c / 2] = pack_h2(a, b);
            wl[c / 2] = pack_h2(la, lb);
        }
        uint32_t *dh = reinterpret_cast<uint32_t *>(h_hi + i * ldh + lane * CPL);
        uint32_t *dl = reinterpret_cast<uint32_t *>(h_lo + i * ldh + lane * CPL);
#pragma unroll
        for (int c = 0; c < CPL / 2; ++c) { dh[c] = wh[c]; dl[c] = wl[c]; }
    }
    if (out_hi) {                                                   // tf32 hi/lo split for the tensor-core consumers
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const float h = __uint_as_float((__float_as_uint(res[c]) + 0x1000u) & 0xffffe000u);
            const float l = res[c] - h;
            out_hi[i * ldo + lane * CPL + c] = h;
            out_lo[i * ldo + lane * CPL + c] = __uint_as_float((__float_as_uint(l) + 0x1000u) & 0xffffe000u);
        }
    }
    if (neg_sqnorm) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(FULL, ss, s);
        if (lane == 0) neg_sqnorm[i] = -ss;
    }
}

// ---- EdgeConv aggregation with the neighbour table staged in shared memory ----------------------------------------
// The gather above reads k * Cout * 4 bytes per point out of L2 (168 MB per 32 clouds at Cout = 64) and is bound by
// the L2 -> SM path.  Here a CTA owns (cloud, 32-channel slice, range of points): the slice of P for ALL points of the
// cloud (N x 128 B) is copied once into shared memory with cp.async and the k gathers per point become LDS.128 -- 8
// lanes per point read one 128-byte row, so every quarter-warp access is conflict free.
// GMS_THREADS = 1024 for clouds whose table fills most of an SM's shared memory (one CTA per SM: all the warps it can hold);
// 512 for small (collapsed) clouds, so that three or four CTAs share an SM and one CTA's table fill overlaps the others' gathers.
constexpr int GM_SMALL_DEFAULT = 768;
template <int GMS_THREADS>
__global__ void __launch_bounds__(GMS_THREADS)
gather_max_smem_kernel(const float *__restrict__ PQ, int64_t ldpq, const int32_t *__restrict__ idx, int N, int k, int Cout,
                       int psplit, int act, float *__restrict__ out, int64_t ldo, float *__restrict__ out_hi,
                       float *__restrict__ out_lo, float *__restrict__ sq_part, __half *__restrict__ h_hi,
                       __half *__restrict__ h_lo, int64_t ldh, float hscale)
{
    extern __shared__ float4 ptab[];                               // N rows x 8 float4
    const int slices = Cout >> 5;
    const int unit = blockIdx.x;
    const int ps = unit % psplit, sl = (unit / psplit) % slices, cloud = unit / (psplit * slices);
    const int64_t cloud0 = (int64_t)cloud * N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 3, q = lane & 7;                          // 4 points per warp instruction, 8 lanes each
    const int per = N / psplit;
    const int p_end = (ps + 1) * per;
    // k = 20: the neighbour list of a point is five 16-byte words.  Each of the first five lanes of the point's group loads
    // ONE of them -- for the NEXT point the group will work on, a whole iteration ahead of its use (ncu: a third of all stall
    // samples sat on the first use of these indices, long scoreboard) -- and the group shares them by shuffle.
    auto load_nb = [&](int p0) -> int4 {
        const int32_t *row = idx + (cloud0 + p0 + g) * k;
        return q < 5 ? __ldg(reinterpret_cast<const int4 *>(row) + q) : make_int4(0, 0, 0, 0);
    };
    int4 nb_next = make_int4(0, 0, 0, 0);
    if (k == 20 && ps * per + warp * 4 < p_end) nb_next = load_nb(ps * per + warp * 4);   // in flight during the table fill
    {
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(ptab);
        for (int t = tid; t < N * 8; t += GMS_THREADS) {
            const int j = t >> 3, q = t & 7;
            const float *src = PQ + (cloud0 + j) * ldpq + sl * 32 + q * 4;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + (uint32_t)t * 16u), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    for (int p0 = ps * per + warp * 4; p0 < p_end; p0 += GMS_THREADS / 8) {
        const int i = p0 + g;
        const int32_t *row = idx + (cloud0 + i) * k;
        const float4 qv = *reinterpret_cast<const float4 *>(PQ + (cloud0 + i) * ldpq + Cout + sl * 32 + q * 4);   // in flight early
        float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (k == 20) {
            const int4 nb = nb_next;
            if (p0 + GMS_THREADS / 8 < p_end) nb_next = load_nb(p0 + GMS_THREADS / 8);
            const int lead = lane & ~7;
#pragma unroll
            for (int t = 0; t < 5; ++t) {
                const int js[4] = {__shfl_sync(FULL, nb.x, lead + t), __shfl_sync(FULL, nb.y, lead + t),
                                   __shfl_sync(FULL, nb.z, lead + t), __shfl_sync(FULL, nb.w, lead + t)};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 v = ptab[js[e] * 8 + q];
                    mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y); mx.z = fmaxf(mx.z, v.z); mx.w = fmaxf(mx.w, v.w);
                }
            }
        } else {
#pragma unroll 4
            for (int t = 0; t < k; ++t) {
                const float4 v = ptab[__ldg(row + t) * 8 + q];
                mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y); mx.z = fmaxf(mx.z, v.z); mx.w = fmaxf(mx.w, v.w);
            }
        }
        float4 r;
        r.x = apply_act(mx.x + qv.x, act); r.y = apply_act(mx.y + qv.y, act);
        r.z = apply_act(mx.z + qv.z, act); r.w = apply_act(mx.w + qv.w, act);
        const int64_t o = (cloud0 + i) * ldo + sl * 32 + q * 4;
        if (out) *reinterpret_cast<float4 *>(out + o) = r;
        if (h_hi) {                                                 // two-term fp16 split for the kind::f16 consumers
            __half h0, h1, h2, h3, l0, l1, l2, l3;
            split_f16(r.x, hscale, h0, l0); split_f16(r.y, hscale, h1, l1);
            split_f16(r.z, hscale, h2, l2); split_f16(r.w, hscale, h3, l3);
            const int64_t oh = (cloud0 + i) * ldh + sl * 32 + q * 4;
            *reinterpret_cast<uint2 *>(h_hi + oh) = make_uint2(pack_h2(h0, h1), pack_h2(h2, h3));
            *reinterpret_cast<uint2 *>(h_lo + oh) = make_uint2(pack_h2(l0, l1), pack_h2(l2, l3));
        }
        if (out_hi) {
            float4 h, l;
            h.x = __uint_as_float((__float_as_uint(r.x) + 0x1000u) & 0xffffe000u);
            h.y = __uint_as_float((__float_as_uint(r.y) + 0x1000u) & 0xffffe000u);
            h.z = __uint_as_float((__float_as_uint(r.z) + 0x1000u) & 0xffffe000u);
            h.w = __uint_as_float((__float_as_uint(r.w) + 0x1000u) & 0xffffe000u);
            l.x = __uint_as_float((__float_as_uint(r.x - h.x) + 0x1000u) & 0xffffe000u);
            l.y = __uint_as_float((__float_as_uint(r.y - h.y) + 0x1000u) & 0xffffe000u);
            l.z = __uint_as_float((__float_as_uint(r.z - h.z) + 0x1000u) & 0xffffe000u);
            l.w = __uint_as_float((__float_as_uint(r.w - h.w) + 0x1000u) & 0xffffe000u);
            *reinterpret_cast<float4 *>(out_hi + o) = h;
            *reinterpret_cast<float4 *>(out_lo + o) = l;
        }
        if (sq_part) {                                              // this slice's share of |x_i|^2 (fixed summation order)
            float ss = r.x * r.x;
            ss = fmaf(r.y, r.y, ss); ss = fmaf(r.z, r.z, ss); ss = fmaf(r.w, r.w, ss);
            ss += __shfl_xor_sync(FULL, ss, 4);
            ss += __shfl_xor_sync(FULL, ss, 2);
            ss += __shfl_xor_sync(FULL, ss, 1);
            if (q == 0) sq_part[(cloud0 + i) * slices + sl] = ss;
        }
    }
}

__global__ void xyz_to_point_major_kernel(const float *__restrict__ cf, int N, float *__restrict__ pm)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float *s = cf + (int64_t)b * 3 * N;
    float *d = pm + ((int64_t)b * N + i) * 3;
    d[0] = s[i]; d[1] = s[N + i]; d[2] = s[2 * N + i];
}

template <typename F>
int dispatch_v(int64_t N, F &&f)
{
    if (N <= 64) return f(std::integral_constant<int, 2>());
    if (N <= 128) return f(std::integral_constant<int, 4>());
    if (N <= 256) return f(std::integral_constant<int, 8>());
    if (N <= 512) return f(std::integral_constant<int, 16>());
    if (N <= 1024) return f(std::integral_constant<int, 32>());
    return f(std::integral_constant<int, 64>());
}

}  // namespace

int launch_knn_xyz(const float *xyz, int point_major, int64_t B, int64_t N, int k, int32_t *idx, cudaStream_t st)
{
    ProfileScope _ps("knn_xyz", st);
    IQ_CHECK(N >= 1 && N <= 2048, "knn: num_points must be in [1,2048]");
    IQ_CHECK(k >= 1 && k <= N, "knn: k must be in [1,num_points]");
    IQ_CHECK(B <= 65535, "knn: batch too large for one launch");
    if (B == 0) return 0;
    const int rows_per_cta = 64;
    dim3 grid((unsigned)ceil_div(N, rows_per_cta), (unsigned)B);
    const size_t smem = sizeof(float4) * (size_t)N;
    return dispatch_v(N, [&](auto v) {
        constexpr int V = decltype(v)::value;
        knn_xyz_kernel<V><<<grid, 256, smem, st>>>(xyz, point_major, (int)N, k, rows_per_cta, idx);
        IQ_COUNT_LAUNCH();
        IQ_LAUNCH_CHECK();
        return 0;
    });
}

int launch_knn_point(const float *xyz, const float *new_xyz, int64_t B, int64_t N, int64_t S, int k, int32_t *idx,
                     cudaStream_t st)
{
    ProfileScope _ps("knn_point", st);
    IQ_CHECK(N >= 1 && N <= 2048, "knn_point: num_points must be in [1,2048]");
    IQ_CHECK(k >= 1 && k <= N && B <= 65535, "knn_point: bad k or batch");
    if (B * S == 0) return 0;
    dim3 grid((unsigned)ceil_div(S, 64), (unsigned)B);
    const size_t smem = sizeof(float4) * (size_t)N;
    return dispatch_v(N, [&](auto v) {
        constexpr int V = decltype(v)::value;
        knn_point_kernel<V><<<grid, 256, smem, st>>>(xyz, new_xyz, (int)N, (int)S, k, idx);
        IQ_COUNT_LAUNCH();
        IQ_LAUNCH_CHECK();
        return 0;
    });
}

int launch_topk_rows(const float *keys, int64_t rows, int64_t N, int64_t ld, int k, int largest, int32_t *idx,
                     cudaStream_t st)
{
    ProfileScope _ps("topk_rows", st);
    IQ_CHECK(N >= 1 && N <= 2048, "topk: row length must be in [1,2048]");
    IQ_CHECK(k >= 1 && k <= N, "topk: k must be in [1,row length]");
    if (rows == 0) return 0;
    const unsigned grid = (unsigned)ceil_div(rows, 8);
    return dispatch_v(N, [&](auto v) {
        constexpr int V = decltype(v)::value;
        topk_rows_kernel<V><<<grid, 256, 0, st>>>(keys, rows, (int)N, ld, k, largest, idx);
        IQ_COUNT_LAUNCH();
        IQ_LAUNCH_CHECK();
        return 0;
    });
}

int launch_knn_rerank(const float *x, int64_t ld, int C, const int32_t *cand, int64_t rows, int64_t N, int k,
                      int32_t *idx, cudaStream_t st)
{
    ProfileScope _ps("knn_rerank", st);
    IQ_CHECK(C % 4 == 0 && ld % 4 == 0 && k <= 32 && N >= 32, "knn_rerank: bad shape");
    if (rows == 0) return 0;
    knn_rerank_kernel<<<(unsigned)ceil_div(rows * 32, 256), 256, 0, st>>>(x, ld, C, cand, rows, (int)N, k, idx);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

int launch_sqnorm_rows(const float *x, int64_t rows, int C, int64_t ld, float *out, cudaStream_t st)
{
    ProfileScope _ps("sqnorm_rows", st);
    if (rows == 0) return 0;
    sqnorm_rows_kernel<<<(unsigned)ceil_div(rows * 32, 256), 256, 0, st>>>(x, rows, C, ld, out);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

int launch_gather_max(const float *PQ, int64_t ldpq, const int32_t *idx, int64_t B, int64_t N, int k, int Cout,
                      int act, float *out, int64_t ldo, float *neg_sqnorm, int *sq_parts, float *out_hi, float *out_lo,
                      const GatherOut16 *h16, cudaStream_t st)
{
    if (sq_parts) *sq_parts = 1;
    __half *h_hi = h16 ? h16->hi : nullptr, *h_lo = h16 ? h16->lo : nullptr;
    const int64_t ldh = h16 ? h16->ld : 0;
    const float hscale = h16 ? h16->scale : 1.0f;
    IQ_CHECK(out || h_hi || out_hi, "gather_max: no output");
    IQ_CHECK(!out_hi || out_lo, "gather_max: the tf32 pair needs both halves");
    IQ_CHECK(!h_hi || (h_lo && ldh % 8 == 0), "gather_max: fp16 outputs need both halves and a leading dimension that is a multiple of 8");
    IQ_CHECK(out || !neg_sqnorm || sq_parts, "gather_max: the separate squared-norm pass reads the fp32 output");
    const int64_t total = B * N;
    if (total == 0) return 0;
    IQ_CHECK(Cout == 64 || Cout == 128 || Cout == 256, "gather_max: Cout must be 64, 128 or 256");
    IQ_CHECK(ldpq % 4 == 0 && ldo % 4 == 0, "gather_max: leading dimensions must be multiples of 4");
    const size_t smem = (size_t)N * 128;
    if (smem <= 200 * 1024 && N % 128 == 0 && B * (Cout / 32) * 8 < ((int64_t)1 << 31)) {
        {
            ProfileScope _ps("gather_max", st);
            if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(&gather_max_smem_kernel<1024>), 200 * 1024)) return rc;
            if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(&gather_max_smem_kernel<512>), 200 * 1024)) return rc;
            const int64_t base_units = B * (Cout / 32);
            int psplit = 1;
            while (psplit < 4 && base_units * psplit < 120) psplit *= 2;   // fill the SMs, but keep the units fat
            // the caller can take per-slice partial squared norms (sq_parts != null: |x_i|^2 = sum of Cout/32 parts)
            float *parts = (neg_sqnorm && sq_parts) ? neg_sqnorm : nullptr;
            // clouds of <= GM_SMALL_DEFAULT points (IQ_GM_SMALL overrides) take the 512-thread form: two to four CTAs then share an
            // SM.  Same box, headline step: threshold 512 / 768 / 1024 -> 123.8k / 126.3k / 123.9k forwards/s (gather_max 5.53 ms
            // with one form only, 5.17 / 5.05 / 4.97 ms; at 1024 the step loses elsewhere what the kernel gains)
            if (N <= env_int("IQ_GM_SMALL", GM_SMALL_DEFAULT))
                gather_max_smem_kernel<512><<<(unsigned)(base_units * psplit), 512, smem, st>>>(PQ, ldpq, idx, (int)N, k, Cout, psplit,
                                                                                      act, out, ldo, out_hi, out_lo, parts, h_hi, h_lo, ldh, hscale);
            else
                gather_max_smem_kernel<1024><<<(unsigned)(base_units * psplit), 1024, smem, st>>>(PQ, ldpq, idx, (int)N, k, Cout, psplit,
                                                                                        act, out, ldo, out_hi, out_lo, parts, h_hi, h_lo, ldh, hscale);
            IQ_COUNT_LAUNCH();
            IQ_LAUNCH_CHECK();
            if (parts) *sq_parts = Cout / 32;
        }
        if (neg_sqnorm && !sq_parts) return launch_sqnorm_rows(out, total, Cout, ldo, neg_sqnorm, st);
        return 0;
    }
    ProfileScope _ps("gather_max", st);
    const unsigned grid = (unsigned)ceil_div(total * 32, 256);
    if (Cout == 64) gather_max_kernel<2><<<grid, 256, 0, st>>>(PQ, ldpq, idx, total, (int)N, k, act, out, ldo, neg_sqnorm, out_hi, out_lo, h_hi, h_lo, ldh, hscale);
    else if (Cout == 128) gather_max_kernel<4><<<grid, 256, 0, st>>>(PQ, ldpq, idx, total, (int)N, k, act, out, ldo, neg_sqnorm, out_hi, out_lo, h_hi, h_lo, ldh, hscale);
    else gather_max_kernel<8><<<grid, 256, 0, st>>>(PQ, ldpq, idx, total, (int)N, k, act, out, ldo, neg_sqnorm, out_hi, out_lo, h_hi, h_lo, ldh, hscale);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

int launch_xyz_to_point_major(const float *x_cf, int64_t B, int64_t N, float *x_pm, cudaStream_t st)
{
    ProfileScope _ps("xyz_to_point_major", st);
    if (B * N == 0) return 0;
    IQ_CHECK(B <= 65535, "xyz_to_point_major: batch too large");
    dim3 grid((unsigned)ceil_div(N, 256), (unsigned)B);
    xyz_to_point_major_kernel<<<grid, 256, 0, st>>>(x_cf, (int)N, x_pm);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

}  // namespace iq
