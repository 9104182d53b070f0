// Fused feature-space kNN for the dynamic graph of DGCNN (sm_100a only).
//
// Reference behaviour restated (never copied): knn(), models/dgcnn.py:12-18 -- per cloud the (N, N) matrix
// -|x_i|^2 + 2 x_i.x_j - |x_j|^2 is formed with two batched matmuls and torch.topk picks k columns per row.
//
// B200-first design: that N x N key matrix (134 MB per 32 clouds and layer) never leaves the SM.
//
//   gram_knn_kernel   persistent, one CTA per SM, unit = (cloud, 128-row tile of points i), 320 threads.
//       warp 0      TMA producer: the cloud's points j stream through a ring of (128 x 32) hi/lo stages.
//       warp 1      one lane issues tcgen05.mma.kind::tf32, 3xTF32 (Alo*Bhi + Ahi*Blo + Ahi*Bhi).  The unit's A tile
//                   (128 rows x K, tf32 hi and lo) is parked ONCE per unit in TMEM by the epilogue warps
//                   (tcgen05.st, lane = row, column = k) and every MMA reads it from there: an SS-mode 128x128 tf32
//                   MMA needs the whole 128 B/clk shared-memory read port for A and B, so with the TMA writes on
//                   top the pipe starved (measured 45 % active); A-from-TMEM halves that traffic.  The remaining
//                   TMEM columns hold 3 (K=64) or 2 (K=128) accumulator buffers.
//       warps 2-9   epilogue.  Thread = (row i, column half): two warps share each TMEM lane quadrant and split the
//                   columns of every tile (one warp per scheduler is latency bound: ncu showed ~0.2 IPC).
//                   Every unit sweeps the cloud's columns TWICE (recomputing the MMAs is cheaper than holding
//                   128 x N accumulators, which do not fit TMEM):
//                     sweep 1  keys 2*G - |x_j|^2 folded into running maxima of strided column blocks
//                              (column j -> block j mod 64); the two halves exchange their NOM largest block
//                              maxima through shared memory and T0 = NOM-th largest of the union, so at least
//                              NOM columns of the row have key >= T0;
//                     sweep 2  branch-free: one bit per column for key > T0 and one for key == T0 (masked clouds
//                              are full of coincident points, i.e. exact ties), written as (rows, 2, N/32) words.
//   knn_rerank_mask_kernel   warp per row: decodes the masks (all columns above T0, then the lowest-index ties up to
//                   NOM candidates: ~30 of 1024 columns) and decides the k neighbours on squared distances evaluated
//                   directly, sum_c (x_i[c] - x_j[c])^2 with float64 accumulation -- without the cancellation of
//                   the expanded form -- ties to the lower index; half-warp per candidate, coalesced row reads.
//   knn_exact_rows_kernel    rows whose candidate set overflowed (pathological column orders) or came up short are
//                   redone exhaustively; normally no row takes this path.
#include <stdlib.h>

#include "tc_ptx.cuh"
#include "kernels.cuh"

namespace iq {

using namespace tc;

namespace {

constexpr int KNN_THREADS = 320;                // producer warp, MMA warp, 8 epilogue warps
constexpr int KNN_EPI_THREADS = 256;
constexpr int KNN_NOM = 24;                     // columns nominated per row (>= k + slack for 3xTF32 key noise)
constexpr int XCH_STRIDE = KNN_NOM + 1;         // padded row of the threshold exchange (bank-conflict free)
constexpr unsigned FULL = 0xffffffffu;
constexpr int RERANK_F2_DEFAULT = 0;            // packed fp32 distance evaluation in the re-rank's fast path (IQ_RERANK_F2)

struct KnnParams {
    int K;                 // feature width (multiple of 4; TMA zero-fills up to the next multiple of 32)
    int points;            // N, multiple of 128
    int m_tiles;           // N / 128
    int num_units;         // clouds * m_tiles
    const void *x_hi, *x_lo;    // tf32 (or two-term fp16, template flag H) split of the features; the A rows are read directly
    int64_t ld_bytes;           // bytes between rows
    float key_scale;            // key = key_scale * G - |x_j|^2: 2, or 2 / s^2 when the fp16 operands carry s * x
    const float *nxx;      // (rows, nxx_parts): |x_j|^2 = |sum of the parts|
    int nxx_parts;
    uint32_t *masks;       // (rows, 2, N/32): bit j of [row][0] <=> key > T0, of [row][1] <=> key == T0
    int dbg;               // IQ_KNN_DBG (scripts/knn_probe.py): 1 = epilogue skips its math, 2 = no MMAs, 4 = no TMA loads
};

template <int BN, int STAGES, int KMAX>
struct KnnSmem {
    static constexpr int B_TILE = BN * TBK * 4;
    static constexpr int STAGE_BYTES = 2 * B_TILE;
    static constexpr int NB_BYTES = 2048 * 4;
    static constexpr int XCH_BYTES = 2 * TBM * XCH_STRIDE * 4;
    static constexpr int XPOSE_BYTES = 8 * 4096;                   // per epilogue warp: 32 rows x 128 B transpose tile
    static constexpr int BAR_BYTES = 256;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + XPOSE_BYTES + NB_BYTES + XCH_BYTES + BAR_BYTES + 1024;
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// descending bitonic network over W registers (fully unrolled)
template <int W>
__device__ __forceinline__ void sort_desc(float (&a)[W])
{
#pragma unroll
    for (int kk = 2; kk <= W; kk <<= 1) {
#pragma unroll
        for (int jj = kk >> 1; jj > 0; jj >>= 1) {
#pragma unroll
            for (int i = 0; i < W; ++i) {
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

// H: the operands are two-term fp16 splits (kind::f16 MMAs, 64 elements per 128-byte row): half the MMAs, half the
// operand bytes, the same three products.  In TMEM the A tile then holds two fp16 per column, so every column offset below
// (32 per ring stage, 8 per k-step) is the same in both formats; A takes KMAX / 2 columns per half instead of KMAX.
template <int BN, int STAGES, int KMAX, bool H = false>
__global__ void __launch_bounds__(KNN_THREADS, 1)
gram_knn_kernel(const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
                const KnnParams p)
{
    using S = KnnSmem<BN, STAGES, KMAX>;
    static_assert(BN == 64 || BN == 128, "column tile must be 64 or 128");
    constexpr int HC = BN / 2;                                       // columns of a tile handled by one epilogue warp
    constexpr int HG = HC / 32;                                      // 32-column TMEM loads per warp and tile
    static_assert(HC >= KNN_NOM, "each half must track at least NOM blocks");
    static_assert(KMAX == 64 || KMAX == 128, "feature width is 64 or 128");
    constexpr int KE = H ? 2 * TBK : TBK;                            // K elements per 128-byte row = per ring stage
    constexpr int AW = H ? KMAX / 2 : KMAX;                          // TMEM columns of one half (hi or lo) of the A tile
    constexpr int NACC = (512 - 2 * AW) / BN;                        // accumulator buffers: TMEM columns [0, NACC*BN)
    constexpr uint32_t A_COL = NACC * BN;                            // A hi at [A_COL, A_COL+AW), lo right after
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *b_smem = smem;
    uint8_t *a_xpose = b_smem + STAGES * S::STAGE_BYTES;            // 1 KB aligned: stages are multiples of 1 KB
    float *nb = reinterpret_cast<float *>(a_xpose + S::XPOSE_BYTES);
    float *xch = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(nb) + S::NB_BYTES);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(xch) + S::XCH_BYTES);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *a_full = empty_bar + STAGES;
    uint64_t *a_empty = a_full + 1;
    uint64_t *tmem_full = a_empty + 1;
    uint64_t *tmem_empty = tmem_full + NACC;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + NACC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = (p.K + KE - 1) / KE;
    const int T = p.points / BN;                                     // column tiles per sweep
    constexpr uint32_t TMEM_COLS = 512;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_bhi); prefetch_tmap(&map_blo);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(a_full, 8); mbar_init(a_empty, 1);
        for (int a = 0; a < NACC; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }
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
        int stage = 0;
        uint32_t phase = 0;
        for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
            const int cloud_row0 = (unit / p.m_tiles) * p.points;
            for (int t = 0; t < 2 * T; ++t) {
                const int b_row0 = cloud_row0 + (t >= T ? t - T : t) * BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t *st = b_smem + stage * S::STAGE_BYTES;
                    if (elect_one_sync()) {
                        if (p.dbg & 4) {                               // probe: no loads, the MMAs chew on stale smem
                            mbar_arrive(&full_bar[stage]);
                        } else {
                            mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES);
                            tma_load_2d(st, &map_bhi, &full_bar[stage], kb * KE, b_row0);
                            tma_load_2d(st + S::B_TILE, &map_blo, &full_bar[stage], kb * KE, b_row0);
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // the whole warp walks the schedule (uniform control flow); one elected lane issues the MMAs and commits
        constexpr uint32_t idesc = H ? make_idesc_f16(BN) : make_idesc(BN);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0, a_phase = 0;
        long long c_a = 0, c_acc = 0, c_full = 0, c_issue = 0, c0 = 0, c1 = 0;   // IQ_KNN_DBG & 16: where the issuer waits
        const bool prof = (p.dbg & 16) != 0;
        const long long c_begin = clock64();
        for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
            if (prof) c0 = clock64();
            mbar_wait(a_full, a_phase);
            if (prof) c_a += clock64() - c0;
            a_phase ^= 1;
            tc_fence_after();
            for (int t = 0; t < 2 * T; ++t) {
                if (prof) c0 = clock64();
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                if (prof) c_acc += clock64() - c0;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < kblocks; ++kb) {
                    if (prof) c0 = clock64();
                    mbar_wait(&full_bar[stage], phase);
                    if (prof) { c1 = clock64(); c_full += c1 - c0; }
                    tc_fence_after();
                    const uint32_t sbase = smem_u32(b_smem + stage * S::STAGE_BYTES);
                    const uint32_t ahi = tmem_base + A_COL + (uint32_t)(kb * TBK), alo = ahi + AW;
                    const uint64_t bhi = make_smem_desc(sbase), blo = make_smem_desc(sbase + S::B_TILE);
                    if (elect_one_sync()) {
                        if (!(p.dbg & 2)) {
#pragma unroll
                            for (int term = 0; term < 3; ++term) {    // small terms first
                                const uint32_t ad = term == 0 ? alo : ahi;
                                const uint64_t bd = term == 1 ? blo : bhi;
#pragma unroll
                                for (int ks = 0; ks < TBK / UMMA_K; ++ks) {
                                    const uint64_t koff = (uint64_t)((ks * UMMA_K * 4) >> 4);
                                    if (H)
                                        umma_f16_ts(d_tmem, ad + (uint32_t)(ks * UMMA_K), bd + koff, idesc,
                                                    (kb | term | ks) != 0 ? 1u : 0u);
                                    else
                                        umma_tf32_ts(d_tmem, ad + (uint32_t)(ks * UMMA_K), bd + koff, idesc,
                                                     (kb | term | ks) != 0 ? 1u : 0u);
                                }
                            }
                        }
                        umma_commit(&empty_bar[stage]);
                        if (kb == kblocks - 1) umma_commit(&tmem_full[acc]);
                    }
                    __syncwarp();
                    if (prof) c_issue += clock64() - c1;
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
            if (elect_one_sync()) umma_commit(a_empty);              // the A tile in TMEM may be overwritten
            __syncwarp();
        }
        if (prof && blockIdx.x == 0 && lane == 0)
            printf("gram_knn issuer (block 0): total %lld clk, wait A %lld, wait accumulator %lld, wait B stage %lld, issue %lld\n",
                   clock64() - c_begin, c_a, c_acc, c_full, c_issue);
    } else {
        const int quad = warp & 3;                                   // TMEM lane quadrant this warp may read
        const int half = (warp - 2) >> 2;                            // which half of every tile's columns
        const int row_in_tile = quad * 32 + lane;
        const int etid = threadIdx.x - 64;                           // 0..255 among the epilogue threads
        const uint32_t nb_s = smem_u32(nb);
        const int words = p.points >> 5;
        const float ksc = p.key_scale;
        int acc = 0;
        uint32_t acc_phase = 0, a_phase = 0;
        for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
            const int cloud = unit / p.m_tiles, mt = unit - cloud * p.m_tiles;
            const int64_t cloud_row0 = (int64_t)cloud * p.points;
            const int64_t row = cloud_row0 + mt * TBM + row_in_tile;
            // ---- park this unit's A rows in TMEM: half 0 writes the hi part of its rows, half 1 the lo part.  A thread
            // reading its own row touches 32 lines per load instruction (measured: ~8 k clk per unit with the tensor
            // pipe idle), so each warp loads its 32 rows coalesced (4 rows x 128 B per instruction), transposes them
            // through a 128B-swizzled 4 KB tile of shared memory and only then takes one row per lane.
            {
                // (byte arithmetic: a pass moves 128 bytes of each row = 32 TMEM columns in either operand format)
                const uint8_t *src = reinterpret_cast<const uint8_t *>(half ? p.x_lo : p.x_hi) +
                                     (cloud_row0 + mt * TBM + quad * 32) * p.ld_bytes;
                const uint32_t a_t = tmem_base + ((uint32_t)(quad * 32) << 16) + A_COL + (uint32_t)(half * AW);
                const uint32_t stg = smem_u32(a_xpose) + (uint32_t)(warp - 2) * 4096u;
                const int lr = lane >> 3, lq = lane & 7;
                float4 g[8];
#pragma unroll
                for (int it = 0; it < 8; ++it)
                    g[it] = __ldg(reinterpret_cast<const float4 *>(src + (int64_t)(it * 4 + lr) * p.ld_bytes) + lq);
                mbar_wait(a_empty, a_phase ^ 1);                     // MMAs of the previous unit have retired
                a_phase ^= 1;
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < AW / 32; ++c) {
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int r = it * 4 + lr;
                        sts128(stg + (uint32_t)r * 128u + (uint32_t)((lq ^ (r & 7)) << 4), g[it].x, g[it].y, g[it].z, g[it].w);
                    }
                    __syncwarp();
                    if (c + 1 < AW / 32) {
#pragma unroll
                        for (int it = 0; it < 8; ++it)
                            g[it] = __ldg(reinterpret_cast<const float4 *>(src + (int64_t)(it * 4 + lr) * p.ld_bytes) + (c + 1) * 8 + lq);
                    }
                    uint32_t r[32];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 f = lds128(stg + (uint32_t)lane * 128u + (uint32_t)((q ^ (lane & 7)) << 4));
                        r[4 * q] = __float_as_uint(f.x); r[4 * q + 1] = __float_as_uint(f.y);
                        r[4 * q + 2] = __float_as_uint(f.z); r[4 * q + 3] = __float_as_uint(f.w);
                    }
                    __syncwarp();
                    tmem_st32(a_t + 32 * c, r);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_full);
            }
            epi_bar_sync();                                          // everyone is done with the previous unit's smem
            for (int i = etid; i < p.points; i += KNN_EPI_THREADS) {
                const float *src = p.nxx + (cloud_row0 + i) * p.nxx_parts;
                float sum = __ldg(src);
                for (int q = 1; q < p.nxx_parts; ++q) sum += __ldg(src + q);
                nb[i] = -fabsf(sum);
            }
            epi_bar_sync();

            // ---- sweep 1: running maxima of this half's strided column blocks (column j -> block j mod HC)
            float bm[HC];
#pragma unroll
            for (int i = 0; i < HC; ++i) bm[i] = -INFINITY;
            for (int t = 0; t < T; ++t) {
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + half * HC);
                const uint32_t nba = nb_s + (uint32_t)(t * BN + half * HC) * 4u;
                uint32_t v[HG][32];
                if (!(p.dbg & 1)) {
#pragma unroll
                for (int g = 0; g < HG; ++g) tmem_ld32_issue(taddr + 32 * g, v[g]);
#pragma unroll
                for (int g = 0; g < HG; ++g) {
                    tmem_ld_wait_regs(v[g]);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b = lds128(nba + 128u * g + 16u * q);
                        const int o = 32 * g + 4 * q;
                        bm[o] = fmaxf(bm[o], fmaf(ksc, __uint_as_float(v[g][4 * q]), b.x));
                        bm[o + 1] = fmaxf(bm[o + 1], fmaf(ksc, __uint_as_float(v[g][4 * q + 1]), b.y));
                        bm[o + 2] = fmaxf(bm[o + 2], fmaf(ksc, __uint_as_float(v[g][4 * q + 2]), b.z));
                        bm[o + 3] = fmaxf(bm[o + 3], fmaf(ksc, __uint_as_float(v[g][4 * q + 3]), b.w));
                    }
                }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
            // T0 = NOM-th largest block maximum over both halves: the NOM largest of the union of two descending
            // lists a, b are { max(a[i], b[NOM-1-i]) }, and T0 is the smallest of them
            sort_desc<HC>(bm);
            {
                float *mine = xch + (half * TBM + row_in_tile) * XCH_STRIDE;
#pragma unroll
                for (int i = 0; i < KNN_NOM; ++i) mine[i] = bm[i];
            }
            epi_bar_sync();
            float T0 = INFINITY;
            {
                const float *other = xch + ((half ^ 1) * TBM + row_in_tile) * XCH_STRIDE;
#pragma unroll
                for (int i = 0; i < KNN_NOM; ++i) T0 = fminf(T0, fmaxf(bm[i], other[KNN_NOM - 1 - i]));
            }

            // ---- sweep 2: one bit per column for key > T0 and key == T0
            uint32_t *mrow = p.masks + row * (int64_t)(2 * words);
            for (int t = 0; t < T; ++t) {
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + half * HC);
                const uint32_t nba = nb_s + (uint32_t)(t * BN + half * HC) * 4u;
                uint32_t v[HG][32];
                uint32_t gtw[HG] = {}, eqw[HG] = {};
                if (!(p.dbg & 1)) {
#pragma unroll
                for (int g = 0; g < HG; ++g) tmem_ld32_issue(taddr + 32 * g, v[g]);
#pragma unroll
                for (int g = 0; g < HG; ++g) {
                    tmem_ld_wait_regs(v[g]);
                    uint32_t mgt[4] = {0u, 0u, 0u, 0u}, meq[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b = lds128(nba + 128u * g + 16u * q);
                        const float kq[4] = {fmaf(ksc, __uint_as_float(v[g][4 * q]), b.x),
                                             fmaf(ksc, __uint_as_float(v[g][4 * q + 1]), b.y),
                                             fmaf(ksc, __uint_as_float(v[g][4 * q + 2]), b.z),
                                             fmaf(ksc, __uint_as_float(v[g][4 * q + 3]), b.w)};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (kq[e] > T0) mgt[e] |= 1u << (4 * q + e);
                            if (kq[e] == T0) meq[e] |= 1u << (4 * q + e);
                        }
                    }
                    gtw[g] = mgt[0] | mgt[1] | mgt[2] | mgt[3];
                    eqw[g] = meq[0] | meq[1] | meq[2] | meq[3];
                }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
                const int w0 = (t * BN + half * HC) >> 5;
                if (HG == 2) {
                    *reinterpret_cast<uint2 *>(mrow + w0) = make_uint2(gtw[0], gtw[HG - 1]);
                    *reinterpret_cast<uint2 *>(mrow + words + w0) = make_uint2(eqw[0], eqw[HG - 1]);
                } else {
                    mrow[w0] = gtw[0];
                    mrow[words + w0] = eqw[0];
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}


// position of the r-th (0-based) set bit of m, by popcount bisection
__device__ __forceinline__ int nth_set_bit(uint32_t m, int r)
{
    int pos = 0, t;
    t = __popc(m & 0xffffu); if (r >= t) { r -= t; pos += 16; m >>= 16; }
    t = __popc(m & 0xffu);   if (r >= t) { r -= t; pos += 8;  m >>= 8; }
    t = __popc(m & 0xfu);    if (r >= t) { r -= t; pos += 4;  m >>= 4; }
    t = __popc(m & 0x3u);    if (r >= t) { r -= t; pos += 2;  m >>= 2; }
    if (r >= (int)(m & 1u)) pos += 1;
    return pos;
}

// Every lane holds one mask word m (word w0 + lane) and the inclusive prefix count incl of set bits over the lanes.
// Returns the column of the s-th set bit in word order (meaningless when s is out of range; all lanes must call).
__device__ __forceinline__ int column_of_slot(uint32_t m, int incl, int s, int w0)
{
    int w = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
        const int v = __shfl_sync(FULL, incl, w + step - 1);
        if (v <= s) w += step;
    }
    w = min(w, 31);
    const uint32_t mw = __shfl_sync(FULL, m, w);
    const int before = __shfl_sync(FULL, incl - __popc(m), w);
    return (w0 + w) * 32 + nth_set_bit(mw, s - before);
}

__device__ __forceinline__ int warp_inclusive_scan(int x, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(FULL, x, o);
        if (lane >= o) x += y;
    }
    return x;
}

// ---- exact decision among the candidates ------------------------------------------------------------------
// One warp per row.  The masks are decoded slot-parallel (lane s finds the s-th set bit by bisection over the prefix
// counts: no per-bit loops, which the masked rows with their hundreds of ties would run 24 times).  LPC = C/16 lanes
// share a candidate (each lane 16 channels as 4 float4, interleaved so a group reads 64 contiguous bytes per
// instruction), so a warp evaluates 32/LPC candidates at a time with all loads of four such steps in flight; the usual
// case of <= 32 candidates is then ordered by a 15-step bitonic network over the lanes, keyed by (float64 distance,
// index).
template <int C, bool F2>
__global__ void __launch_bounds__(256, 4)
knn_rerank_mask_kernel(const float *__restrict__ x, int64_t ld, const uint32_t *__restrict__ masks, int64_t rows, int N,
                       int k, int32_t *__restrict__ idx, int32_t *__restrict__ cnt)
{
    constexpr int LPC = C / 16;                                      // lanes per candidate: 4 (C = 64) or 8 (C = 128)
    constexpr int CPI = 32 / LPC;                                    // candidates per step
    constexpr int ITERS = 32 / CPI;                                  // steps per batch of 32 candidates
    __shared__ uint16_t cand_s[8][KNN_CAND_CAP];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const int words = N >> 5;
    const uint32_t *mg = masks + row * (int64_t)(2 * words), *me = mg + words;
    uint16_t *cl = cand_s[wib];
    const int grp = lane / LPC, sub = lane % LPC;
    float4 xi[4];                                                    // my 16 channels of the row itself, in flight early
#pragma unroll
    for (int q = 0; q < 4; ++q) xi[q] = __ldg(reinterpret_cast<const float4 *>(x + row * ld) + q * LPC + sub);

    // the first 32 words of the tie mask travel together with the first words of the "above" mask: 62 % of the rows need
    // them, and loaded on demand they cost a second L2 round trip on the critical path (ncu: long-scoreboard stalls)
    const uint32_t me_first = lane < words ? __ldg(me + lane) : 0u;
    int n = 0;
    for (int w0 = 0; w0 < words; w0 += 32) {                          // every column above the threshold
        const int w = w0 + lane;
        uint32_t m = w < words ? __ldg(mg + w) : 0u;
        const int incl = warp_inclusive_scan(__popc(m), lane);
        int pos = n + incl - __popc(m);
        while (m) {                                                  // ~1 bit per word
            const int b = __ffs(m) - 1;
            m &= m - 1;
            if (pos < KNN_CAND_CAP) cl[pos] = (uint16_t)(w * 32 + b);
            ++pos;
        }
        n += __shfl_sync(FULL, incl, 31);
    }
    int need = KNN_NOM - n;                                            // then the lowest-index ties, up to NOM in all
    for (int w0 = 0; w0 < words && need > 0; w0 += 32) {
        const uint32_t m = w0 == 0 ? me_first : (w0 + lane < words ? __ldg(me + w0 + lane) : 0u);
        const int incl = warp_inclusive_scan(__popc(m), lane);
        const int took = min(__shfl_sync(FULL, incl, 31), need);
        const int col = column_of_slot(m, incl, lane, w0);
        if (lane < took) cl[n + lane] = (uint16_t)col;               // n + lane < NOM <= capacity
        n += took;
        need -= took;
    }
    if (lane == 0) cnt[row] = n;
    if (n > KNN_CAND_CAP || n < k) return;                           // redone by knn_exact_rows_kernel
    __syncwarp();
    const int64_t cloud0 = (row / N) * N;
    const int self = (int)(row - cloud0);
    if (n <= 32) {
        // Fast path: the same direct distances accumulated in fp32.  Every term is non-negative, so the fp32 sum is
        // within gamma = 20 * 2^-24 (16-term FMA chain + lane tree) of the float64 sum, relatively.  The SET of the k
        // nearest is therefore already decided unless the k-th and (k+1)-th distances lie within 8e-6 of each other;
        // only those rows (about one in 10^4) fall through to the float64 evaluation below.  Equal fp32 distances come
        // from coincident points (identical rows, identical operation order) and are ordered by index.
        float d32[ITERS];
#pragma unroll
        for (int i0 = 0; i0 < ITERS; i0 += 2) {
            // two steps' worth of loads in flight; a step whose CPI slots lie beyond n is skipped (n is warp-uniform; about
            // 24.5 of the 32 slots are filled on average, so the last step is empty for half of the rows)
            float4 b[2][4];
            bool live[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                live[u] = (i0 + u) * CPI < n;
                if (!live[u]) continue;
                const int slot = (i0 + u) * CPI + grp;
                const int j = slot < n ? (int)cl[slot] : self;
                const float4 *xj = reinterpret_cast<const float4 *>(x + (cloud0 + j) * ld);
#pragma unroll
                for (int q = 0; q < 4; ++q) b[u][q] = __ldg(xj + q * LPC + sub);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                float acc = 0.0f;
                if (live[u]) {
                    if (F2) {
                        // packed fp32 (FADD2 / FFMA2): a scalar fp32 op issues every second cycle per scheduler on sm_100, and
                        // this loop is the kernel's fp32 work.  Two chains (even / odd channels) summed at the end: the same
                        // order for every candidate, so coincident points still give bit-equal distances.
                        unsigned long long a2 = 0ull;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const unsigned long long d01 = sub_f2(pack_f2(xi[q].x, xi[q].y), pack_f2(b[u][q].x, b[u][q].y));
                            const unsigned long long d23 = sub_f2(pack_f2(xi[q].z, xi[q].w), pack_f2(b[u][q].z, b[u][q].w));
                            a2 = fma_f2(d01, d01, a2);
                            a2 = fma_f2(d23, d23, a2);
                        }
                        float lo, hi;
                        unpack_f2(a2, lo, hi);
                        acc = lo + hi;
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float d0 = xi[q].x - b[u][q].x, d1 = xi[q].y - b[u][q].y;
                            const float d2 = xi[q].z - b[u][q].z, d3 = xi[q].w - b[u][q].w;
                            acc = fmaf(d0, d0, acc); acc = fmaf(d1, d1, acc); acc = fmaf(d2, d2, acc); acc = fmaf(d3, d3, acc);
                        }
                    }
#pragma unroll
                    for (int o = LPC / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
                }
                d32[i0 + u] = acc;
            }
        }
        float mine = 0.0f;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const float v = __shfl_sync(FULL, d32[it], (lane % CPI) * LPC);
            if (lane / CPI == it) mine = v;
        }
        if (n - k <= 12) {
            // Only n - k (4.5 on average) candidates have to go: drop the largest (distance, index) that many times -- two REDUX
            // per removal -- instead of sorting all 32 (a 15-stage network on 64-bit keys).  Distances are >= 0, so their bit
            // patterns order like unsigned integers.  The neighbours are written in candidate (= column) order: a SET, as the
            // max over the neighbourhood needs.
            const uint32_t dbits = __float_as_uint(mine);
            bool alive = lane < n;
            const uint32_t my_col = lane < n ? (uint32_t)cl[lane] + 1u : 0u;
            uint32_t last = 0;
            for (int r = n - k; r > 0; --r) {
                last = __reduce_max_sync(FULL, alive ? dbits : 0u);
                const bool top = alive && dbits == last;
                const uint32_t pick = __reduce_max_sync(FULL, top ? my_col : 0u);   // ties: the higher index goes first
                if (top && my_col == pick) alive = false;
            }
            const float dk1 = __uint_as_float(__reduce_max_sync(FULL, alive ? dbits : 0u));   // k-th smallest
            const float dk = __uint_as_float(last);                                           // (k+1)-th smallest, if n > k
            const bool ambiguous = n > k && dk != dk1 && dk - dk1 <= 8e-6f * dk;
            if (!ambiguous) {
                const unsigned keep = __ballot_sync(FULL, alive);
                if (alive) idx[row * k + __popc(keep & ((1u << lane) - 1u))] = (int)my_col - 1;
                return;
            }
        } else {
        unsigned long long key = lane < n ? ((unsigned long long)__float_as_uint(mine) << 32) | (unsigned)cl[lane] : ~0ull;
#pragma unroll
        for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
            for (int jj = kk >> 1; jj > 0; jj >>= 1) {
                const unsigned long long other = __shfl_xor_sync(FULL, key, jj);
                const bool keep_min = ((lane & kk) == 0) == ((lane & jj) == 0);
                key = keep_min ? min(key, other) : max(key, other);
            }
        }
        const float dk1 = __uint_as_float((unsigned)(__shfl_sync(FULL, key, k - 1) >> 32));
        const float dk = __uint_as_float((unsigned)(__shfl_sync(FULL, key, min(k, 31)) >> 32));
        const bool ambiguous = n > k && dk != dk1 && dk - dk1 <= 8e-6f * dk;
        if (!ambiguous) {
            if (lane < k) idx[row * k + lane] = (int)(key & 0xffffffffu);
            return;
        }
        }
    }
    const int batches = n > 32 ? 2 : 1;
    double myd[2] = {INFINITY, INFINITY};
    int myj[2] = {0x7fffffff, 0x7fffffff};
#pragma unroll
    for (int bt = 0; bt < 2; ++bt) {
        if (bt >= batches) break;
        double dacc[ITERS];
#pragma unroll
        for (int i0 = 0; i0 < ITERS; i0 += 2) {                      // two steps' worth of loads in flight
            float4 b[2][4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int slot = bt * 32 + (i0 + u) * CPI + grp;
                const int j = slot < n ? (int)cl[slot] : self;       // unconditional loads: out-of-range slots read the row itself
                const float4 *xj = reinterpret_cast<const float4 *>(x + (cloud0 + j) * ld);
#pragma unroll
                for (int q = 0; q < 4; ++q) b[u][q] = __ldg(xj + q * LPC + sub);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double d0 = (double)(xi[q].x - b[u][q].x), d1 = (double)(xi[q].y - b[u][q].y);
                    const double d2 = (double)(xi[q].z - b[u][q].z), d3 = (double)(xi[q].w - b[u][q].w);
                    acc = fma(d0, d0, acc); acc = fma(d1, d1, acc); acc = fma(d2, d2, acc); acc = fma(d3, d3, acc);
                }
#pragma unroll
                for (int o = LPC / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
                dacc[i0 + u] = acc;
            }
        }
        // candidate slot bt*32 + L moves to lane L: it was evaluated in step L / CPI by group L % CPI
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const double v = __shfl_sync(FULL, dacc[it], (lane % CPI) * LPC);
            if (lane / CPI == it) myd[bt] = v;
        }
        if (bt * 32 + lane < n) myj[bt] = cl[bt * 32 + lane];
        else myd[bt] = INFINITY;
    }
    // The reduction order above is fixed, so coincident points give bit-equal distances and fall back to the index.
    // Sort key: the bit pattern of a non-negative double orders like an integer; its 11 lowest mantissa bits (4.5e-13
    // relative, far below the 6e-8 rounding of the fp32 differences) make room for the point index as the tie-break.
    if (batches == 1) {
        unsigned long long key = ((unsigned long long)__double_as_longlong(myd[0]) & ~0x7ffull) | (unsigned)(myj[0] & 0x7ff);
        if (lane >= n) key = ~0ull;
#pragma unroll
        for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
            for (int jj = kk >> 1; jj > 0; jj >>= 1) {
                const unsigned long long other = __shfl_xor_sync(FULL, key, jj);
                const bool keep_min = ((lane & kk) == 0) == ((lane & jj) == 0);
                key = keep_min ? min(key, other) : max(key, other);
            }
        }
        if (lane < k) idx[row * k + lane] = (int)(key & 0x7ff);
        return;
    }
    int rank0 = 0, rank1 = 0;
    for (int s = 0; s < n; ++s) {
        const double o = s < 32 ? __shfl_sync(FULL, myd[0], s) : __shfl_sync(FULL, myd[1], s - 32);
        const int oj = cl[s];
        rank0 += (o < myd[0] || (o == myd[0] && oj < myj[0])) ? 1 : 0;
        rank1 += (o < myd[1] || (o == myd[1] && oj < myj[1])) ? 1 : 0;
    }
    if (lane < n && rank0 < k) idx[row * k + rank0] = myj[0];
    if (lane + 32 < n && rank1 < k) idx[row * k + rank1] = myj[1];
}

// ---- exhaustive redo of a row (overflowed or short candidate set): k rounds of "nearest point after the previous
// pick" under (float64 distance, index).  O(k N C) per row, slow and normally never taken: a small grid scans the
// candidate counts 32 rows per warp step and only stops at flagged rows.
__global__ void __launch_bounds__(256)
knn_exact_rows_kernel(const float *__restrict__ x, int64_t ld, int C, const int32_t *__restrict__ cnt, int64_t rows,
                      int N, int k, int32_t *__restrict__ idx)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r0 = warp * 32; r0 < rows; r0 += nwarps * 32) {
        const int n = r0 + lane < rows ? __ldg(cnt + r0 + lane) : k;
        unsigned todo = __ballot_sync(FULL, n > KNN_CAND_CAP || n < k);
        while (todo) {
            const int64_t row = r0 + __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t cloud0 = (row / N) * N;
            const float4 *xi = reinterpret_cast<const float4 *>(x + row * ld);
            double last_d = -1.0;
            int last_j = -1;
            for (int r = 0; r < k; ++r) {
                double best_d = INFINITY;
                int best_j = 0x7fffffff;
                for (int j = lane; j < N; j += 32) {
                    const float4 *xj = reinterpret_cast<const float4 *>(x + (cloud0 + j) * ld);
                    double d = 0.0;
                    for (int c = 0; c < (C >> 2); ++c) {
                        const float4 a = __ldg(xi + c), b = __ldg(xj + c);
                        const double d0 = (double)(a.x - b.x), d1 = (double)(a.y - b.y), d2 = (double)(a.z - b.z), d3 = (double)(a.w - b.w);
                        d = fma(d0, d0, d); d = fma(d1, d1, d); d = fma(d2, d2, d); d = fma(d3, d3, d);
                    }
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
    }
}

template <int BN, int STAGES, int KMAX, bool H>
int launch_variant(const void *x_hi, const void *x_lo, int64_t ld, int64_t rows, const KnnParams &p, cudaStream_t st)
{
    using S = KnnSmem<BN, STAGES, KMAX>;
    static_assert(S::TOTAL <= 232448, "shared memory budget exceeded");
    IQ_CHECK(p.K == KMAX, "knn_features_tc: feature width does not match the kernel variant");
    CUtensorMap mbhi, mblo;
    if (int rc = make_map_any(&mbhi, x_hi, rows, p.K, ld, BN, H ? 2 : 4)) return rc;
    if (int rc = make_map_any(&mblo, x_lo, rows, p.K, ld, BN, H ? 2 : 4)) return rc;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(&gram_knn_kernel<BN, STAGES, KMAX, H>), S::TOTAL)) return rc;
    const int grid = std::min(p.num_units, sm_count());
    gram_knn_kernel<BN, STAGES, KMAX, H><<<grid, KNN_THREADS, S::TOTAL, st>>>(mbhi, mblo, p);
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
                           int nxx_parts, int64_t clouds, int64_t N, int k, uint32_t *masks, int32_t *cnt, int32_t *idx,
                           cudaStream_t st, const KnnOperands16 *h16)
{
    IQ_CHECK(knn_features_tc_supported(N, C, k), "knn_features_tc: unsupported shape");
    IQ_CHECK(ld % 4 == 0, "knn_features_tc: leading dimension must be a multiple of 4");
    IQ_CHECK(!h16 || (h16->hi && h16->lo && h16->ld % 8 == 0 && h16->scale > 0.0f), "knn_features_tc: bad fp16 operands");
    const int64_t rows = clouds * N;
    if (rows == 0) return 0;
    IQ_CHECK(rows < (int64_t)1 << 31, "knn_features_tc: too many rows");
    KnnParams p;
    p.K = C; p.points = (int)N; p.m_tiles = (int)(N / TBM); p.num_units = (int)(clouds * p.m_tiles);
    p.nxx = nxx; p.nxx_parts = nxx_parts; p.masks = masks;
    p.dbg = env_int("IQ_KNN_DBG", 0);
    {
        ProfileScope _ps(C <= 64 ? "tc_gram_knn_c64" : "tc_gram_knn_c128", st);
        int rc;
        if (h16) {                                                   // nominate on kind::f16 MMAs over the fp16 pair of s * x
            p.x_hi = h16->hi; p.x_lo = h16->lo; p.ld_bytes = h16->ld * 2; p.key_scale = 2.0f / (h16->scale * h16->scale);
            rc = C == 64 ? launch_variant<128, 5, 64, true>(h16->hi, h16->lo, h16->ld, rows, p, st)
                         : launch_variant<128, 5, 128, true>(h16->hi, h16->lo, h16->ld, rows, p, st);
        } else {
            p.x_hi = x_hi; p.x_lo = x_lo; p.ld_bytes = ld * 4; p.key_scale = 2.0f;
            rc = C == 64 ? launch_variant<128, 5, 64, false>(x_hi, x_lo, ld, rows, p, st)
                         : launch_variant<128, 5, 128, false>(x_hi, x_lo, ld, rows, p, st);
        }
        if (rc) return rc;
    }
    {
        ProfileScope _ps("knn_rerank", st);
        const unsigned grid = (unsigned)ceil_div(rows * 32, 256);
        const bool f2 = env_int("IQ_RERANK_F2", RERANK_F2_DEFAULT) != 0;
        if (C == 64 && f2) knn_rerank_mask_kernel<64, true><<<grid, 256, 0, st>>>(x, ld, masks, rows, (int)N, k, idx, cnt);
        else if (C == 64) knn_rerank_mask_kernel<64, false><<<grid, 256, 0, st>>>(x, ld, masks, rows, (int)N, k, idx, cnt);
        else if (f2) knn_rerank_mask_kernel<128, true><<<grid, 256, 0, st>>>(x, ld, masks, rows, (int)N, k, idx, cnt);
        else knn_rerank_mask_kernel<128, false><<<grid, 256, 0, st>>>(x, ld, masks, rows, (int)N, k, idx, cnt);
        IQ_COUNT_LAUNCH();
        IQ_LAUNCH_CHECK();
        const unsigned scan_grid = (unsigned)std::min<int64_t>(ceil_div(rows, 32 * 8), 4 * sm_count());
        knn_exact_rows_kernel<<<scan_grid, 256, 0, st>>>(x, ld, C, cnt, rows, (int)N, k, idx);
        IQ_COUNT_LAUNCH();
        IQ_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace iq
