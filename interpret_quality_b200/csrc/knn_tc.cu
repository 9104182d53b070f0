// Fused feature-space kNN for the dynamic graph of DGCNN (sm_100a only).
//
// Reference behaviour restated (never copied): knn(), models/dgcnn.py:12-18 -- per cloud the (N, N) matrix
// -|x_i|^2 + 2 x_i.x_j - |x_j|^2 is formed with two batched matmuls and torch.topk picks k columns per row.
//
// B200-first design: that N x N key matrix (134 MB per 32 clouds and layer) never leaves the SM.
//
//   gram_knn_kernel   persistent, one CTA per SM, unit = (cloud, 128-row tile of points i).
//       warp 0      TMA producer.  The unit's A tile (128 x K, tf32 hi and lo) is loaded ONCE and stays resident
//                   in shared memory; the cloud's points j stream through a ring of (BN x 32) hi/lo stages.
//       warp 1      one lane issues tcgen05.mma.kind::tf32, 3xTF32 (Alo*Bhi + Ahi*Blo + Ahi*Bhi), accumulators
//                   double buffered in TMEM.
//       warps 2-5   epilogue, thread = row i.  Every unit sweeps the cloud's columns TWICE (recomputing the MMAs
//                   is cheaper than holding 128 x N accumulators, which do not fit TMEM):
//                     sweep 1  keys 2*G - |x_j|^2 folded into 64 running maxima (column j -> block j mod 64);
//                              T0 = NOM-th largest block maximum (register bitonic network), so at least NOM
//                              columns of the row have key >= T0;
//                     sweep 2  every column with key > T0, and the first NOM with key == T0 (masked clouds are
//                              full of coincident points, i.e. exact ties), is appended to the row's candidate
//                              list: ~30 of 1024 columns.
//   knn_rerank_list_kernel   the k neighbours are decided among the candidates on squared distances evaluated
//                   directly, sum_c (x_i[c] - x_j[c])^2 with float64 accumulation -- without the cancellation of
//                   the expanded form -- ties to the lower index; half-warp per candidate, coalesced row reads.
//   knn_exact_rows_kernel    rows whose list overflowed (pathological column orders) or came up short are redone
//                   exhaustively; normally no row takes this path.
#include "tc_ptx.cuh"
#include "kernels.cuh"

namespace iq {

using namespace tc;

namespace {

constexpr int KNN_THREADS = 192;
constexpr unsigned FULL = 0xffffffffu;

struct KnnParams {
    int K;                 // feature width (multiple of 4; TMA zero-fills up to the next multiple of 32)
    int points;            // N, multiple of 128
    int m_tiles;           // N / 128
    int num_units;         // clouds * m_tiles
    const float *nxx;      // (rows) -|x_j|^2
    uint16_t *cand;        // (rows, KNN_CAND_CAP)
    int32_t *cnt;          // (rows) number of candidates found (may exceed the capacity: overflow)
};

template <int BN, int STAGES, int KMAX>
struct KnnSmem {
    static constexpr int A_TILE = TBM * TBK * 4;                   // one 32-wide k-block of the A tile, hi or lo
    static constexpr int A_BYTES = 2 * (KMAX / TBK) * A_TILE;
    static constexpr int B_TILE = BN * TBK * 4;
    static constexpr int STAGE_BYTES = 2 * B_TILE;
    static constexpr int NB_BYTES = 2048 * 4;
    static constexpr int BAR_BYTES = 256;
    static constexpr int TOTAL = A_BYTES + STAGES * STAGE_BYTES + NB_BYTES + BAR_BYTES + 1024;
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// descending bitonic network over 64 registers (672 compare-exchanges, fully unrolled)
__device__ __forceinline__ void sort64_desc(float (&a)[64])
{
#pragma unroll
    for (int kk = 2; kk <= 64; kk <<= 1) {
#pragma unroll
        for (int jj = kk >> 1; jj > 0; jj >>= 1) {
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                const int l = i ^ jj;
                if (l > i) {
                    const bool desc = (i & kk) == 0;
                    const float hi = fmaxf(a[i], a[l]), lo = fminf(a[i], a[l]);
                    a[i] = desc ? hi : lo;
                    a[l] = desc ? lo : hi;
                }
            }
        }
    }
}

template <int BN, int STAGES, int KMAX, int NOM>
__global__ void __launch_bounds__(KNN_THREADS, 1)
gram_knn_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
                const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
                const KnnParams p)
{
    using S = KnnSmem<BN, STAGES, KMAX>;
    static_assert(BN % 64 == 0 && BN <= 128, "column tile must be 64 or 128");
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *a_smem = smem;
    uint8_t *b_smem = smem + S::A_BYTES;
    float *nb = reinterpret_cast<float *>(b_smem + STAGES * S::STAGE_BYTES);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(nb) + S::NB_BYTES);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *a_full = empty_bar + STAGES;
    uint64_t *a_empty = a_full + 1;
    uint64_t *tmem_full = a_empty + 1;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = (p.K + TBK - 1) / TBK;
    const int T = p.points / BN;                                     // column tiles per sweep
    constexpr uint32_t TMEM_COLS = 2 * BN <= 128 ? 128 : 256;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_ahi); prefetch_tmap(&map_alo); prefetch_tmap(&map_bhi); prefetch_tmap(&map_blo);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(a_full, 1); mbar_init(a_empty, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, a_phase = 0;
            for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
                const int cloud = unit / p.m_tiles, mt = unit - cloud * p.m_tiles;
                const int cloud_row0 = cloud * p.points;
                mbar_wait(a_empty, a_phase ^ 1);                     // MMAs of the previous unit have retired
                mbar_arrive_expect_tx(a_full, (uint32_t)(2 * kblocks * S::A_TILE));
                for (int kb = 0; kb < kblocks; ++kb) {
                    tma_load_2d(a_smem + (2 * kb) * S::A_TILE, &map_ahi, a_full, kb * TBK, cloud_row0 + mt * TBM);
                    tma_load_2d(a_smem + (2 * kb + 1) * S::A_TILE, &map_alo, a_full, kb * TBK, cloud_row0 + mt * TBM);
                }
                a_phase ^= 1;
                for (int t = 0; t < 2 * T; ++t) {
                    const int b_row0 = cloud_row0 + (t >= T ? t - T : t) * BN;
                    for (int kb = 0; kb < kblocks; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t *st = b_smem + stage * S::STAGE_BYTES;
                        mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES);
                        tma_load_2d(st, &map_bhi, &full_bar[stage], kb * TBK, b_row0);
                        tma_load_2d(st + S::B_TILE, &map_blo, &full_bar[stage], kb * TBK, b_row0);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0, a_phase = 0;
            const uint32_t abase = smem_u32(a_smem);
            for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
                mbar_wait(a_full, a_phase);
                a_phase ^= 1;
                tc_fence_after();
                for (int t = 0; t < 2 * T; ++t) {
                    mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                    for (int kb = 0; kb < kblocks; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t sbase = smem_u32(b_smem + stage * S::STAGE_BYTES);
                        const uint64_t ahi = make_smem_desc(abase + (2 * kb) * S::A_TILE);
                        const uint64_t alo = make_smem_desc(abase + (2 * kb + 1) * S::A_TILE);
                        const uint64_t bhi = make_smem_desc(sbase), blo = make_smem_desc(sbase + S::B_TILE);
#pragma unroll
                        for (int term = 0; term < 3; ++term) {        // small terms first
                            const uint64_t ad = term == 0 ? alo : ahi;
                            const uint64_t bd = term == 1 ? blo : bhi;
#pragma unroll
                            for (int ks = 0; ks < TBK / UMMA_K; ++ks) {
                                const uint64_t koff = (uint64_t)((ks * UMMA_K * 4) >> 4);
                                umma_tf32(d_tmem, ad + koff, bd + koff, idesc, (kb | term | ks) != 0 ? 1u : 0u);
                            }
                        }
                        umma_commit(&empty_bar[stage]);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&tmem_full[acc]);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                umma_commit(a_empty);                                // the resident A tile may be overwritten
            }
        }
    } else {
        const int quad = warp & 3;                                   // TMEM lane quadrant this warp may read
        const int row_in_tile = quad * 32 + lane;
        const int etid = threadIdx.x - 64;                           // 0..127 among the epilogue threads
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
            const int cloud = unit / p.m_tiles, mt = unit - cloud * p.m_tiles;
            const int64_t cloud_row0 = (int64_t)cloud * p.points;
            const int64_t row = cloud_row0 + mt * TBM + row_in_tile;
            epi_bar_sync();                                          // everyone is done with the previous cloud's norms
            for (int i = etid; i < p.points; i += 128) nb[i] = __ldg(p.nxx + cloud_row0 + i);
            epi_bar_sync();

            // ---- sweep 1: running maxima of the 64 strided column blocks
            float bm[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) bm[i] = -INFINITY;
            for (int t = 0; t < T; ++t) {
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 64) {
                    float v[32];
                    const float4 *nb4 = reinterpret_cast<const float4 *>(nb + t * BN + c0);
                    tmem_ld32(taddr + c0, v);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b = nb4[q];
                        bm[4 * q] = fmaxf(bm[4 * q], fmaf(2.0f, v[4 * q], b.x));
                        bm[4 * q + 1] = fmaxf(bm[4 * q + 1], fmaf(2.0f, v[4 * q + 1], b.y));
                        bm[4 * q + 2] = fmaxf(bm[4 * q + 2], fmaf(2.0f, v[4 * q + 2], b.z));
                        bm[4 * q + 3] = fmaxf(bm[4 * q + 3], fmaf(2.0f, v[4 * q + 3], b.w));
                    }
                    tmem_ld32(taddr + c0 + 32, v);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b = nb4[8 + q];
                        bm[32 + 4 * q] = fmaxf(bm[32 + 4 * q], fmaf(2.0f, v[4 * q], b.x));
                        bm[32 + 4 * q + 1] = fmaxf(bm[32 + 4 * q + 1], fmaf(2.0f, v[4 * q + 1], b.y));
                        bm[32 + 4 * q + 2] = fmaxf(bm[32 + 4 * q + 2], fmaf(2.0f, v[4 * q + 2], b.z));
                        bm[32 + 4 * q + 3] = fmaxf(bm[32 + 4 * q + 3], fmaf(2.0f, v[4 * q + 3], b.w));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            sort64_desc(bm);
            const float T0 = bm[NOM - 1];                            // >= NOM columns of this row have key >= T0

            // ---- sweep 2: collect the columns at or above the threshold
            int n = 0, neq = 0;
            uint16_t *list = p.cand + row * KNN_CAND_CAP;
            for (int t = 0; t < T; ++t) {
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    float v[32];
                    const float4 *nb4 = reinterpret_cast<const float4 *>(nb + t * BN + c0);
                    tmem_ld32(taddr + c0, v);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b = nb4[q];
                        const float kq[4] = {fmaf(2.0f, v[4 * q], b.x), fmaf(2.0f, v[4 * q + 1], b.y),
                                             fmaf(2.0f, v[4 * q + 2], b.z), fmaf(2.0f, v[4 * q + 3], b.w)};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const bool eq = kq[e] == T0;
                            if (kq[e] > T0 || (eq && neq < NOM)) {
                                if (n < KNN_CAND_CAP) list[n] = (uint16_t)(t * BN + c0 + 4 * q + e);
                                ++n;
                            }
                            neq += eq ? 1 : 0;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            p.cnt[row] = n;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---- exact decision among the candidates ------------------------------------------------------------------
// One warp per row.  Half-warp h handles candidate slots c with (c & 16) == 16 h, so the 16 lanes of a half read one
// feature row as consecutive float4 (coalesced) and the lane that keeps the result, c & 31, lies in the same half.
template <int CV>                                                    // CV = C / 64: float4 per lane and row
__global__ void __launch_bounds__(256)
knn_rerank_list_kernel(const float *__restrict__ x, int64_t ld, const uint16_t *__restrict__ cand,
                       const int32_t *__restrict__ cnt, int64_t rows, int N, int k, int32_t *__restrict__ idx)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const int n = cnt[row];
    if (n > KNN_CAND_CAP || n < k) return;                           // redone by knn_exact_rows_kernel
    const int64_t cloud0 = (row / N) * N;
    const int sub = lane & 15, half = lane >> 4;
    int myj[2];
    myj[0] = lane < n ? (int)cand[row * KNN_CAND_CAP + lane] : 0;
    myj[1] = lane + 32 < n ? (int)cand[row * KNN_CAND_CAP + 32 + lane] : 0;
    float4 xi[CV];
#pragma unroll
    for (int q = 0; q < CV; ++q) xi[q] = __ldg(reinterpret_cast<const float4 *>(x + row * ld) + q * 16 + sub);
    double myd[2] = {INFINITY, INFINITY};
    const int rounds = n > 32 ? 2 : 1;
    for (int r = 0; r < rounds; ++r) {
#pragma unroll 4
        for (int tt = 0; tt < 16; ++tt) {
            const int slot = tt + 16 * half;                         // candidate slot (within this round) of my half
            const int j = __shfl_sync(FULL, myj[r], slot);
            const bool valid = r * 32 + slot < n;
            double acc = 0.0;
            if (valid) {
                const float4 *xj = reinterpret_cast<const float4 *>(x + (cloud0 + j) * ld);
#pragma unroll
                for (int q = 0; q < CV; ++q) {
                    const float4 b = __ldg(xj + q * 16 + sub);
                    const double d0 = (double)(xi[q].x - b.x), d1 = (double)(xi[q].y - b.y);
                    const double d2 = (double)(xi[q].z - b.z), d3 = (double)(xi[q].w - b.w);
                    acc = fma(d0, d0, acc); acc = fma(d1, d1, acc); acc = fma(d2, d2, acc); acc = fma(d3, d3, acc);
                }
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
            if (valid && lane == slot) myd[r] = acc;
        }
    }
    // rank of every candidate under (distance, index); the reduction order above is fixed, so equal points give
    // bit-equal distances and fall back to the index
    int rank0 = 0, rank1 = 0;
    for (int s = 0; s < n; ++s) {
        const double o = s < 32 ? __shfl_sync(FULL, myd[0], s) : __shfl_sync(FULL, myd[1], s - 32);
        const int oj = s < 32 ? __shfl_sync(FULL, myj[0], s) : __shfl_sync(FULL, myj[1], s - 32);
        rank0 += (o < myd[0] || (o == myd[0] && oj < myj[0])) ? 1 : 0;
        rank1 += (o < myd[1] || (o == myd[1] && oj < myj[1])) ? 1 : 0;
    }
    if (lane < n && rank0 < k) idx[row * k + rank0] = myj[0];
    if (lane + 32 < n && rank1 < k) idx[row * k + rank1] = myj[1];
}

// ---- exhaustive redo of a row (overflowed or short candidate list): distances to every point of the cloud, the 32
// nearest by the float32-rounded distance, then the same float64 (distance, index) ranking.  Slow and rarely taken.
__global__ void __launch_bounds__(256)
knn_exact_rows_kernel(const float *__restrict__ x, int64_t ld, int C, const int32_t *__restrict__ cnt, int64_t rows,
                      int N, int k, int32_t *__restrict__ idx)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const int n = cnt[row];
    if (!(n > KNN_CAND_CAP || n < k)) return;
    const int64_t cloud0 = (row / N) * N;
    const float4 *xi = reinterpret_cast<const float4 *>(x + row * ld);
    auto dist = [&](int j) {
        const float4 *xj = reinterpret_cast<const float4 *>(x + (cloud0 + j) * ld);
        double acc = 0.0;
        for (int c = 0; c < (C >> 2); ++c) {
            const float4 a = __ldg(xi + c), b = __ldg(xj + c);
            // same grouping as knn_rerank_list_kernel is not needed here: this kernel decides the whole row itself
            const double d0 = (double)(a.x - b.x), d1 = (double)(a.y - b.y), d2 = (double)(a.z - b.z), d3 = (double)(a.w - b.w);
            acc = fma(d0, d0, acc); acc = fma(d1, d1, acc); acc = fma(d2, d2, acc); acc = fma(d3, d3, acc);
        }
        return acc;
    };
    // k rounds of "smallest (distance, index) larger than the previous pick": O(k N C) per row, no scratch
    double last_d = -1.0;
    int last_j = -1;
    for (int r = 0; r < k; ++r) {
        double best_d = INFINITY;
        int best_j = 0x7fffffff;
        for (int j = lane; j < N; j += 32) {
            const double d = dist(j);
            const bool after = d > last_d || (d == last_d && j > last_j);
            if (after && (d < best_d || (d == best_d && j < best_j))) { best_d = d; best_j = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(FULL, best_d, o);
            const int oj = __shfl_xor_sync(FULL, best_j, o);
            if (od < best_d || (od == best_d && oj < best_j)) { best_d = od; best_j = oj; }
        }
        if (lane == 0) idx[row * k + r] = best_j;
        last_d = best_d;
        last_j = best_j;
    }
}

template <int BN, int STAGES, int KMAX, int NOM>
int launch_variant(const float *x_hi, const float *x_lo, int64_t ld, int64_t rows, const KnnParams &p, cudaStream_t st)
{
    using S = KnnSmem<BN, STAGES, KMAX>;
    static_assert(S::TOTAL <= 232448, "shared memory budget exceeded");
    CUtensorMap mahi, malo, mbhi, mblo;
    if (int rc = make_map(&mahi, x_hi, rows, p.K, ld, TBM)) return rc;
    if (int rc = make_map(&malo, x_lo, rows, p.K, ld, TBM)) return rc;
    if (int rc = make_map(&mbhi, x_hi, rows, p.K, ld, BN)) return rc;
    if (int rc = make_map(&mblo, x_lo, rows, p.K, ld, BN)) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        IQ_CUDA(cudaFuncSetAttribute(gram_knn_kernel<BN, STAGES, KMAX, NOM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     S::TOTAL));
        attr_set = true;
    }
    const int grid = std::min(p.num_units, sm_count());
    gram_knn_kernel<BN, STAGES, KMAX, NOM><<<grid, KNN_THREADS, S::TOTAL, st>>>(mahi, malo, mbhi, mblo, p);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

}  // namespace

bool knn_features_tc_supported(int64_t N, int C, int k)
{
    return N % 128 == 0 && N >= 128 && N <= 2048 && (C == 64 || C == 128) && k >= 1 && k <= 20;
}

int launch_knn_features_tc(const float *x, const float *x_hi, const float *x_lo, int64_t ld, int C, const float *nxx,
                           int64_t clouds, int64_t N, int k, uint16_t *cand, int32_t *cnt, int32_t *idx,
                           cudaStream_t st)
{
    IQ_CHECK(knn_features_tc_supported(N, C, k), "knn_features_tc: unsupported shape");
    IQ_CHECK(ld % 4 == 0, "knn_features_tc: leading dimension must be a multiple of 4");
    const int64_t rows = clouds * N;
    if (rows == 0) return 0;
    IQ_CHECK(rows < (int64_t)1 << 31, "knn_features_tc: too many rows");
    KnnParams p;
    p.K = C; p.points = (int)N; p.m_tiles = (int)(N / TBM); p.num_units = (int)(clouds * p.m_tiles);
    p.nxx = nxx; p.cand = cand; p.cnt = cnt;
    {
        ProfileScope _ps("tc_gram_knn", st);
        int rc = C <= 64 ? launch_variant<128, 4, 64, 24>(x_hi, x_lo, ld, rows, p, st)
                         : launch_variant<64, 5, 128, 24>(x_hi, x_lo, ld, rows, p, st);
        if (rc) return rc;
    }
    {
        ProfileScope _ps("knn_rerank", st);
        const unsigned grid = (unsigned)ceil_div(rows * 32, 256);
        if (C == 64) knn_rerank_list_kernel<1><<<grid, 256, 0, st>>>(x, ld, cand, cnt, rows, (int)N, k, idx);
        else knn_rerank_list_kernel<2><<<grid, 256, 0, st>>>(x, ld, cand, cnt, rows, (int)N, k, idx);
        IQ_COUNT_LAUNCH();
        IQ_LAUNCH_CHECK();
        knn_exact_rows_kernel<<<grid, 256, 0, st>>>(x, ld, C, cnt, rows, (int)N, k, idx);
        IQ_COUNT_LAUNCH();
        IQ_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace iq
