// tcgen05 / TMEM / TMA GEMM with 3xTF32 error compensation (sm_100a only).
//
//   D[m][n] = sum_k A[m][k] * B[n][k]          A (M,K), B (N,K) both K-major fp32
//
// Operand formats: tf32 hi/lo pairs (kind::tf32, 32 elements per 128-byte swizzle row) or, template flag H, two-term
// fp16 splits of the scaled operands (kind::f16, 64 elements per row: same bytes per stage, same descriptors, same
// three products, twice the K per MMA at the same MMA time -- half the stages and half the operand bytes per product).
//
// Single-pass TF32 (and BF16) fails the 1e-3 parity bar for every model of the
// reference, DGCNN worst because a rounded kNN key flips neighbours
// (SURVEY.md section 7.2), so each fp32 operand is pre-split by its producer into
// hi = tf32(x) and lo = tf32(x - hi) and the tile accumulates
//   Alo*Bhi + Ahi*Blo + Ahi*Bhi      (fp32 accumulation in TMEM).
//
// Kernel anatomy (persistent, one CTA per SM, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 128B-swizzled boxes of Ahi/Alo/Bhi/Blo
//               into a multi-stage shared-memory ring, mbarrier complete_tx.
//   warp 1      TMEM allocation; one elected lane issues tcgen05.mma.kind::tf32
//               (M=128, N=BN, K=8) x 4 k-steps x 3 split terms per stage, tcgen05.commit
//               releases the stage / publishes the accumulator.
//   warps 2-5   epilogue: tcgen05.ld 32x32b (thread = accumulator row), fused
//               scale / bias / activation, then either
//                 STORE  row-major fp32 store (EdgeConv P|Q rows, kNN keys 2*G - |x_j|^2), or
//                 POOL   rows are output channels, columns are the points of one cloud:
//                        running max / argmax / sum over all column tiles of the cloud, i.e.
//                        conv + BN + activation + global max/avg pooling without ever
//                        writing the (B, C, N) tensor (models/dgcnn.py:108-111,
//                        models/pointnet.py:35,83 of the reference).
//   Accumulators are double buffered in TMEM (2 x BN columns) so the epilogue of tile i
//   overlaps the main loop of tile i+1.
#include <stdlib.h>

#include "tc_ptx.cuh"
#include "kernels.cuh"

namespace iq {

using namespace tc;

namespace {

constexpr int TC_THREADS = 192;
constexpr int GA_THREADS = 384;          // gathered-A variant: three groups of four gather warps taking ring stages in turn

// Kernel parameters.  Together with the seven 128-byte tensor maps they must stay within 1024 bytes: one 64-byte field more
// and every variant of the kernel ran 25-50 % slower on B200 (measured by padding this struct: PointNet conv+pool 10.5 ->
// 15.6 ms), so the POOL-only and the gather-only fields share storage and the STORE outputs are flags (they leave via TMA).
struct TcParams {
    int K;                       // multiple of 4 (TMA zero-fills up to the next multiple of 32)
    union {
        struct {
            // STORE: units = (M/128) * (Ncols/BN), one accumulator tile each
            int n_tiles;             // Ncols / BN
            int rows_per_batch;      // > 0: B rows live in the same batch as the A rows (Gram), else B is shared
        };
        const float *pool_extra;     // POOL_RUN: weight of each group's last column beyond 1 (collapse.cu), or null
    };
    int num_units, tiles_per_unit;
    int act;
    int out_c, out_hilo;         // STORE outputs present: fp32 and / or its tf32 hi/lo split
    float alpha;
    int dbg;                     // IQ_TC_DBG (scripts/tc_probe.py): 1 = STORE epilogue skips its math and stores
    int four_terms;              // also accumulate Alo*Blo (the 2^-22 term 3xTF32 drops): fp32-FMA-chain accuracy
    const float *bias;           // STORE: per column (index b_row0 + c), POOL: per row (channel)
    union {
        // POOL: columns are grouped in runs of `points` consecutive B rows (points of a cloud / neighbours of a centroid).
        //   points >= BN: units = groups * m_tiles, tiles_per_unit = points / BN (running reduction across tiles)
        //   points <  BN: units = column tiles * m_tiles, BN / points groups reduced inside one tile
        struct {
            float *out_max, *out_mean;   // (clouds, ld_out)
            int64_t *out_arg;            // (clouds, Cout)
            int64_t ld_out;
            int cout;
            int m_tiles;                 // ceil(Cout / 128)
            int points;                  // columns per group
            int a_rows;                  // ATM variants: the A tile is parked in TMEM straight from global memory
            const float *a_hi, *a_lo;
            int64_t lda;
        };
        // gathered-A STORE: A[row][c] = act(U[cloud*nsrc + idx[row]][c] - V[row / gK][c] + gbias[c]), cloud = row / (gK * gS)
        struct {
            const float *gU, *gV, *gbias;
            const int32_t *gidx;
            int64_t gldu, gldv;
            int gK, gS, gnsrc, gact;
        };
    };
};
static_assert(sizeof(TcParams) + 7 * 128 <= 1024, "kernel parameters must stay within 1 KB (see above)");

template <int BN, int STAGES, bool ATM = false>
struct TcSmem {
    static constexpr int A_BYTES = ATM ? 0 : TBM * TBK * 4;                 // ATM: A lives in TMEM, the ring holds B only
    static constexpr int B_BYTES = BN * TBK * 4;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int BAR_BYTES = 256;
    static constexpr int OUT_BYTES = 4 * 2 * 4096;                          // per epilogue warp: two 32 x 32 store boxes
    static constexpr int TOTAL = STAGES * STAGE_BYTES + OUT_BYTES + BAR_BYTES + 1024;   // +1024: alignment slack
};

constexpr int tmem_cols_for(int bn) { return 2 * bn <= 32 ? 32 : 2 * bn <= 64 ? 64 : 2 * bn <= 128 ? 128 : 2 * bn <= 256 ? 256 : 512; }

// One warp's 32 x 32 block -> 128B-swizzled box in shared memory -> TMA store.  Boxes alternate between two buffers;
// before reusing one, the issuing lane waits until at most one earlier store is still reading shared memory.
__device__ __forceinline__ void store_box(const CUtensorMap *map, const float (&val)[32], uint32_t box0, int &nbox, int lane,
                                          int sw, int col, int row)
{
    if (lane == 0) bulk_wait_read<1>();
    __syncwarp();
    const uint32_t boxb = box0 + (uint32_t)(nbox & 1) * 4096u;
    const uint32_t mine = boxb + (uint32_t)lane * 128u;
#pragma unroll
    for (int c = 0; c < 8; ++c)
        sts128(mine + (uint32_t)((c ^ sw) << 4), val[4 * c], val[4 * c + 1], val[4 * c + 2], val[4 * c + 3]);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(map), "r"(boxb), "r"(col), "r"(row)
                     : "memory");
        bulk_commit();
    }
    ++nbox;
}

// GA ("gathered A", STORE mode): the first grouped-MLP layer of PointNet++ / PointConv, relu(U[idx] - V + b), used to be
// written to HBM as a tf32 hi/lo pair by one kernel and read back by this one -- 45 % of the HBM traffic of a grouped
// MLP chain that is HBM bound.  Eight extra warps (thread = row of the tile, two groups on alternate stages) now gather U, apply the layer and write the
// hi and lo operand tiles straight into the 128B-swizzled ring slots; TMA only brings the weights.
// The epilogue variant is a template parameter (VAR_*): with run-time mode switches inside one kernel the compiler's
// unswitching / unrolling choices for one variant moved whenever another was touched (POOL 85 -> 130 ms on PointNet++
// after an unrelated edit).
enum : int { VAR_STORE = 0, VAR_POOL_RUN = 1, VAR_POOL_TILE = 2, VAR_STORE_GATHER = 3, VAR_POOL_RUN_ATM = 4, VAR_POOL_TILE_ATM = 5 };
template <int BN, int STAGES, int VAR, bool H = false>
__global__ void __launch_bounds__(TC_THREADS + (VAR == VAR_STORE_GATHER ? GA_THREADS : 0), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
               const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
               const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_chi,
               const __grid_constant__ CUtensorMap map_clo, const TcParams p)
{
    // ATM ("A in tensor memory", POOL with K <= 128): a CTA keeps ONE 128-row tile of A -- the output channels it owns,
    // hi and lo -- parked in TMEM for its whole life (the grid is a multiple of the number of row tiles, so all its units
    // share it) and the MMAs read it from there; the ring carries B only, half the L2 operand stream of the SS form.
    constexpr bool ATM = VAR == VAR_POOL_RUN_ATM || VAR == VAR_POOL_TILE_ATM;
    using S = TcSmem<BN, STAGES, ATM>;
    constexpr bool GA = VAR == VAR_STORE_GATHER;
    constexpr bool STORE = VAR == VAR_STORE || VAR == VAR_STORE_GATHER;
    constexpr bool POOL_RUN = VAR == VAR_POOL_RUN || VAR == VAR_POOL_RUN_ATM;   // groups of >= BN columns: running reduction
    static_assert(!ATM || BN == 128, "the TMEM-resident A variants use 128-column tiles");
    static_assert(!H || (!ATM && !GA), "fp16 operands: plain STORE / POOL only");
    constexpr int KE = H ? 2 * TBK : TBK;                                  // K elements per 128-byte row = per ring stage
    constexpr uint32_t TMEM_COLS = ATM ? 512u : (uint32_t)tmem_cols_for(BN);
    constexpr uint32_t A_COL = 256;                                         // ATM: A hi at columns [256, 384), lo at [384, 512)
    constexpr uint32_t FULL_ARRIVALS = GA ? 5 : 1;                          // TMA producer (+ the four gather warps)
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *out_stage = smem + STAGES * S::STAGE_BYTES;                    // 1024-byte aligned (stages are multiples of 1 KB)
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(out_stage + S::OUT_BYTES);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full = empty_bar + STAGES;
    uint64_t *tmem_empty = tmem_full + 2;
    uint64_t *a_full = tmem_empty + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(a_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = (p.K + KE - 1) / KE;            // TMA zero-fills the K tail

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_ahi); prefetch_tmap(&map_alo); prefetch_tmap(&map_bhi); prefetch_tmap(&map_blo);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], FULL_ARRIVALS); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
        mbar_init(a_full, 4);
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

    auto tile_rows = [&](int unit, int j, int &a_row0, int &b_row0) {
        if (STORE) {
            const int mt = unit / p.n_tiles, nt = unit - mt * p.n_tiles;
            a_row0 = mt * TBM;
            b_row0 = nt * BN + (p.rows_per_batch > 0 ? (a_row0 / p.rows_per_batch) * p.rows_per_batch : 0);
        } else {
            const int grp = unit / p.m_tiles, mt = unit - grp * p.m_tiles;
            a_row0 = mt * TBM;
            b_row0 = POOL_RUN ? grp * p.points + j * BN : grp * BN;
        }
    };

    // Issuing warps: all 32 lanes walk the schedule in uniform control flow and one elected lane issues (tc_ptx.cuh,
    // elect_one_sync) -- tcgen05.mma issue is paced by the tensor pipe, so every clock the issuer spends elsewhere is lost.
    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x)
            for (int j = 0; j < p.tiles_per_unit; ++j) {
                int a_row0, b_row0;
                tile_rows(unit, j, a_row0, b_row0);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t *st = smem + stage * S::STAGE_BYTES;
                    if (elect_one_sync()) {
                        mbar_arrive_expect_tx(&full_bar[stage], GA ? 2 * S::B_BYTES : S::STAGE_BYTES);
                        if (!GA && !ATM) {
                            tma_load_2d(st, &map_ahi, &full_bar[stage], kb * KE, a_row0);
                            tma_load_2d(st + S::A_BYTES, &map_alo, &full_bar[stage], kb * KE, a_row0);
                        }
                        tma_load_2d(st + 2 * S::A_BYTES, &map_bhi, &full_bar[stage], kb * KE, b_row0);
                        tma_load_2d(st + 2 * S::A_BYTES + S::B_BYTES, &map_blo, &full_bar[stage], kb * KE, b_row0);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
    } else if (warp == 1) {
        constexpr uint32_t idesc = H ? make_idesc_f16(BN) : make_idesc(BN);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        if (ATM) {                                                          // the epilogue warps have parked this CTA's A tile
            mbar_wait(a_full, 0);
            tc_fence_after();
        }
        for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x)
            for (int j = 0; j < p.tiles_per_unit; ++j) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sbase = smem_u32(smem + stage * S::STAGE_BYTES);
                    const uint64_t ahi = make_smem_desc(sbase), alo = make_smem_desc(sbase + S::A_BYTES);
                    const uint64_t bhi = make_smem_desc(sbase + 2 * S::A_BYTES);
                    const uint64_t blo = make_smem_desc(sbase + 2 * S::A_BYTES + S::B_BYTES);
                    if (elect_one_sync()) {
                        if (!ATM && !H && p.four_terms) {
#pragma unroll
                            for (int ks = 0; ks < TBK / UMMA_K; ++ks) {
                                const uint64_t koff = (uint64_t)((ks * UMMA_K * 4) >> 4);
                                umma_tf32(d_tmem, alo + koff, blo + koff, idesc, (kb | ks) != 0 ? 1u : 0u);
                            }
                        }
#pragma unroll
                        for (int term = 0; term < 3; ++term) {            // small terms first
                            const uint64_t ad = term == 0 ? alo : ahi;
                            const uint64_t bd = term == 1 ? blo : bhi;
#pragma unroll
                            for (int ks = 0; ks < TBK / UMMA_K; ++ks) {
                                const uint64_t koff = (uint64_t)((ks * UMMA_K * 4) >> 4);
                                if (ATM)
                                    umma_tf32_ts(d_tmem, tmem_base + A_COL + (term == 0 ? 128u : 0u) + (uint32_t)(kb * TBK + ks * UMMA_K),
                                                 bd + koff, idesc, (kb | term | ks) != 0 ? 1u : 0u);
                                else if (H)                              // one k-step = the same 32 bytes = 16 fp16
                                    umma_f16(d_tmem, ad + koff, bd + koff, idesc, (kb | term | ks) != 0 ? 1u : 0u);
                                else
                                    umma_tf32(d_tmem, ad + koff, bd + koff, idesc, (kb | term | ks | p.four_terms) != 0 ? 1u : 0u);
                            }
                        }
                        umma_commit(&empty_bar[stage]);                  // frees the stage when the MMAs retire
                        if (kb == kblocks - 1) umma_commit(&tmem_full[acc]);   // accumulator complete
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
    } else if (GA && warp >= 6) {
        // ---- gather warps: thread = row of the tile, one 32-channel block per ring stage; the two groups of four warps
        // take alternate stages so that two stages' worth of gathers are in flight
        const int r = ((int)threadIdx.x - 192) & 127, grp = ((int)threadIdx.x - 192) >> 7;
        const int sw = r & 7;
        int stage = 0, turn = 0;
        uint32_t phase = 0;
        for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
            int a_row0, b_row0;
            tile_rows(unit, 0, a_row0, b_row0);
            const int64_t row = (int64_t)a_row0 + r;
            const int64_t cen = row / p.gK, cloud = cen / p.gS;
            const float *u = p.gU + (cloud * p.gnsrc + __ldg(p.gidx + row)) * p.gldu;
            const float *v = p.gV + cen * p.gldv;
            for (int kb = 0; kb < kblocks; ++kb, turn = turn + 1 == GA_THREADS / 128 ? 0 : turn + 1) {
                if (turn != grp) {                                    // the other group's stage
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    continue;
                }
                float4 uu[8];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    uu[q] = kb * TBK + 4 * q < p.K ? __ldg(reinterpret_cast<const float4 *>(u + kb * TBK) + q)
                                                   : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                mbar_wait(&empty_bar[stage], phase ^ 1);
                const uint32_t hi_row = smem_u32(smem + stage * S::STAGE_BYTES) + (uint32_t)r * 128u;
                const uint32_t lo_row = hi_row + S::A_BYTES;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float h[4] = {0.0f, 0.0f, 0.0f, 0.0f}, l[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                    if (kb * TBK + 4 * q < p.K) {
                        const float4 vv = __ldg(reinterpret_cast<const float4 *>(v + kb * TBK) + q);
                        const float4 bb = __ldg(reinterpret_cast<const float4 *>(p.gbias + kb * TBK) + q);
                        const float x[4] = {apply_act((uu[q].x - vv.x) + bb.x, p.gact), apply_act((uu[q].y - vv.y) + bb.y, p.gact),
                                            apply_act((uu[q].z - vv.z) + bb.z, p.gact), apply_act((uu[q].w - vv.w) + bb.w, p.gact)};
#pragma unroll
                        for (int e = 0; e < 4; ++e) { h[e] = tf32_round(x[e]); l[e] = tf32_round(x[e] - h[e]); }
                    }
                    const uint32_t off = (uint32_t)((q ^ sw) << 4);
                    sts128(hi_row + off, h[0], h[1], h[2], h[3]);
                    sts128(lo_row + off, l[0], l[1], l[2], l[3]);
                }
                fence_proxy_async();                                 // generic-proxy writes -> visible to the MMA's async reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        const int quad = warp & 3;                                       // TMEM lane quadrant this warp may read
        const int row_in_tile = quad * 32 + lane;
        int acc = 0, nbox = 0;
        uint32_t acc_phase = 0;
        if (ATM) {
            // thread = row of the tile = output channel: K hi values to columns [A_COL, A_COL+K), K lo values 128 further
            const int ch = (int)(blockIdx.x % p.m_tiles) * TBM + row_in_tile;
            const uint32_t a_t = tmem_base + ((uint32_t)(quad * 32) << 16) + A_COL;
            for (int half = 0; half < 2; ++half) {
                const float4 *src = reinterpret_cast<const float4 *>((half ? p.a_lo : p.a_hi) + (int64_t)ch * p.lda);
                for (int c = 0; c < kblocks; ++c) {
                    uint32_t r[32];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float4 f = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        if (ch < p.a_rows && c * TBK + 4 * q < p.K) f = __ldg(src + c * 8 + q);
                        r[4 * q] = __float_as_uint(f.x); r[4 * q + 1] = __float_as_uint(f.y);
                        r[4 * q + 2] = __float_as_uint(f.z); r[4 * q + 3] = __float_as_uint(f.w);
                    }
                    tmem_st32(a_t + (uint32_t)(half * 128 + c * TBK), r);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        }
        for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
            float run_max = -INFINITY, run_sum = 0.0f, y_last = 0.0f;
            int run_arg = 0;
            for (int j = 0; j < p.tiles_per_unit; ++j) {
                int a_row0, b_row0;
                tile_rows(unit, j, a_row0, b_row0);
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
                if (STORE) {
                    // The accumulator arrives one ROW per thread; stored that way a warp touches 32 rows per instruction
                    // and the epilogue, not the MMAs, paced the kernel (measured 76 us vs 25 us without it).  Each warp
                    // now writes its 32 x 32 block into a 128B-swizzled shared-memory box (conflict-free 16-byte
                    // stores) and one lane hands it to TMA: the stores leave asynchronously, fully coalesced.
                    const int nt = unit % p.n_tiles;
                    const uint32_t box0 = smem_u32(out_stage) + (uint32_t)(warp - 2) * 8192u;
                    const int out_row = a_row0 + quad * 32, sw = lane & 7;
#pragma unroll 1
                    for (int c0 = 0; c0 < BN; c0 += 32) {
                        if (p.dbg & 1) break;
                        float bv[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) bv[i] = p.bias ? __ldg(p.bias + b_row0 + c0 + i) : 0.0f;
                        float v[32];
                        tmem_ld32(taddr + c0, v);
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = apply_act(fmaf(p.alpha, v[i], bv[i]), p.act);
                        const int col = nt * BN + c0;
                        if (p.out_c) store_box(&map_c, v, box0, nbox, lane, sw, col, out_row);
                        if (p.out_hilo) {
                            float h[32];
#pragma unroll
                            for (int i = 0; i < 32; ++i) { h[i] = tf32_round(v[i]); v[i] = tf32_round(v[i] - h[i]); }
                            store_box(&map_chi, h, box0, nbox, lane, sw, col, out_row);
                            store_box(&map_clo, v, box0, nbox, lane, sw, col, out_row);
                        }
                    }
                } else {
                    const int ch = a_row0 + row_in_tile;
                    const float b = (p.bias && ch < p.cout) ? __ldg(p.bias + ch) : 0.0f;
                    const int grp0 = unit / p.m_tiles;
#pragma unroll 1
                    for (int c0 = 0; c0 < BN; c0 += 32) {
                        float v[32];
                        tmem_ld32(taddr + c0, v);
                        if (POOL_RUN) {
                            // the sum is taken per block of 32 columns first: the rounding error of a cloud's mean grows
                            // with sqrt(32) + sqrt(blocks) instead of sqrt(points), so the collapsed and the plain
                            // evaluation of a coalition cloud (different numbers of columns) agree to ~1e-6
                            float blk = 0.0f;
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const float y = apply_act(fmaf(p.alpha, v[i], b), p.act);
                                if (y > run_max) { run_max = y; run_arg = j * BN + c0 + i; }
                                blk += y;
                                if (i == 31) y_last = y;
                            }
                            run_sum += blk;
                        } else {                                            // several groups inside this tile
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const float y = apply_act(fmaf(p.alpha, v[i], b), p.act);
                                run_max = fmaxf(run_max, y);
                                if (((c0 + i + 1) & (p.points - 1)) == 0) {              // points is a power of two here
                                    const int64_t g = (int64_t)grp0 * (BN / p.points) + (c0 + i) / p.points;
                                    if (ch < p.cout) p.out_max[g * p.ld_out + ch] = run_max;
                                    run_max = -INFINITY;
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (POOL_RUN) {
                const int grp = unit / p.m_tiles, mt = unit - grp * p.m_tiles;
                const int ch = mt * TBM + row_in_tile;
                if (ch < p.cout) {
                    p.out_max[(int64_t)grp * p.ld_out + ch] = run_max;
                    if (p.out_mean) {
                        // the last column of a collapsed cloud stands for 1 + extra coincident points
                        const float extra = p.pool_extra ? __ldg(p.pool_extra + grp) : 0.0f;
                        p.out_mean[(int64_t)grp * p.ld_out + ch] = fmaf(extra, y_last, run_sum) / ((float)p.points + extra);
                    }
                    if (p.out_arg) p.out_arg[(int64_t)grp * p.cout + ch] = run_arg;
                }
            }
        }
    }

    if (warp >= 2 && warp < 6 && lane == 0) bulk_wait_all();                     // every TMA store of this warp has landed
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------ hi/lo split

__global__ void split_tf32_kernel(const float *__restrict__ x, int64_t rows, int cols, int64_t ldx,
                                  float *__restrict__ hi, float *__restrict__ lo, int64_t ldo)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int cv = cols >> 2;
    if (t >= rows * cv) return;
    const int64_t r = t / cv;
    const int c = (int)(t - r * cv) << 2;
    const float4 v = *reinterpret_cast<const float4 *>(x + r * ldx + c);
    float4 h, l;
    h.x = tf32_round(v.x); l.x = tf32_round(v.x - h.x);
    h.y = tf32_round(v.y); l.y = tf32_round(v.y - h.y);
    h.z = tf32_round(v.z); l.z = tf32_round(v.z - h.z);
    h.w = tf32_round(v.w); l.w = tf32_round(v.w - h.w);
    *reinterpret_cast<float4 *>(hi + r * ldo + c) = h;
    *reinterpret_cast<float4 *>(lo + r * ldo + c) = l;
}
}  // namespace

int launch_split_tf32(const float *x, int64_t rows, int cols, int64_t ldx, float *hi, float *lo, int64_t ldo,
                      cudaStream_t st)
{
    ProfileScope _ps("split_tf32", st);
    IQ_CHECK(cols % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "split_tf32: widths must be multiples of 4");
    if (rows == 0) return 0;
    const int64_t n = rows * (cols / 4);
    split_tf32_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(x, rows, cols, ldx, hi, lo, ldo);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

namespace {
__global__ void split_f16_kernel(const float *__restrict__ x, int64_t rows, int cols, int64_t ldx, float scale,
                                 __half *__restrict__ hi, __half *__restrict__ lo, int64_t ldo)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int cv = cols >> 2;
    if (t >= rows * cv) return;
    const int64_t r = t / cv;
    const int c = (int)(t - r * cv) << 2;
    const float4 v = *reinterpret_cast<const float4 *>(x + r * ldx + c);
    __half h0, h1, h2, h3, l0, l1, l2, l3;
    split_f16(v.x, scale, h0, l0); split_f16(v.y, scale, h1, l1);
    split_f16(v.z, scale, h2, l2); split_f16(v.w, scale, h3, l3);
    *reinterpret_cast<uint2 *>(hi + r * ldo + c) = make_uint2(pack_h2(h0, h1), pack_h2(h2, h3));
    *reinterpret_cast<uint2 *>(lo + r * ldo + c) = make_uint2(pack_h2(l0, l1), pack_h2(l2, l3));
}
}  // namespace

int launch_split_f16(const float *x, int64_t rows, int cols, int64_t ldx, float scale, __half *hi, __half *lo, int64_t ldo,
                     cudaStream_t st)
{
    ProfileScope _ps("split_f16", st);
    IQ_CHECK(cols % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "split_f16: widths must be multiples of 4");
    if (rows == 0) return 0;
    const int64_t n = rows * (cols / 4);
    split_f16_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(x, rows, cols, ldx, scale, hi, lo, ldo);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

static int pick_bn(int n)
{
    for (int bn : {128, 96, 64, 32})
        if (n % bn == 0) return bn;
    return 0;
}

bool tc_gemm_supported(const TcGemm &g)
{
    if (g.K < 4 || g.K % 4 != 0) return false;
    // STORE: a ragged last row tile is fine for plain operands (TMA zero-fills the A rows beyond M and clips the stores);
    // the gathered-A and the batched-Gram forms index per row and need whole tiles
    if (g.mode == 0) return g.M >= 1 && pick_bn(g.N) != 0 && (g.M % TBM == 0 || (!g.gather.U && g.rows_per_batch == 0));
    if (g.cout < 1 || g.points < 1 || g.clouds < 1) return false;
    if (g.points >= 128) return g.points % 128 == 0;
    return 128 % g.points == 0 && ((int64_t)g.clouds * g.points) % 128 == 0;
}

template <int BN, int STAGES, int VAR, bool H = false>
static int launch_tc_variant(const TcGemm &g, TcParams p, int64_t a_rows, int64_t b_rows, cudaStream_t st)
{
    constexpr bool ATM = VAR == VAR_POOL_RUN_ATM || VAR == VAR_POOL_TILE_ATM;
    using S = TcSmem<BN, STAGES, ATM>;
    CUtensorMap mahi, malo, mbhi, mblo;
    constexpr bool GA = VAR == VAR_STORE_GATHER;
    if (H) {
        if (int rc = make_map_any(&mbhi, g.Bh_hi, b_rows, g.K, g.ldb, BN, 2)) return rc;
        if (int rc = make_map_any(&mblo, g.Bh_lo, b_rows, g.K, g.ldb, BN, 2)) return rc;
        if (int rc = make_map_any(&mahi, g.Ah_hi, a_rows, g.K, g.lda, TBM, 2)) return rc;
        if (int rc = make_map_any(&malo, g.Ah_lo, a_rows, g.K, g.lda, TBM, 2)) return rc;
    } else {
        if (int rc = make_map(&mbhi, g.B_hi, b_rows, g.K, g.ldb, BN)) return rc;
        if (int rc = make_map(&mblo, g.B_lo, b_rows, g.K, g.ldb, BN)) return rc;
        if (GA) {
            mahi = mbhi; malo = mblo;                           // unused: the A tiles are produced in the kernel
        } else {
            if (int rc = make_map(&mahi, g.A_hi, a_rows, g.K, g.lda, TBM)) return rc;
            if (int rc = make_map(&malo, g.A_lo, a_rows, g.K, g.lda, TBM)) return rc;
        }
    }
    // STORE outputs leave through TMA: 32 x 32 boxes (128 bytes wide) over each fp32 output matrix
    CUtensorMap mc = mahi, mchi = mahi, mclo = mahi;
    if (g.mode == 0) {
        if (g.C) { if (int rc = make_map(&mc, g.C, g.M, g.N, g.ldc, 32)) return rc; }
        if (g.C_hi) {
            if (int rc = make_map(&mchi, g.C_hi, g.M, g.N, g.ldc, 32)) return rc;
            if (int rc = make_map(&mclo, g.C_lo, g.M, g.N, g.ldc, 32)) return rc;
        }
    }
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(&gemm_tc_kernel<BN, STAGES, VAR, H>), S::TOTAL)) return rc;
    int grid = std::min(p.num_units, sm_count());
    if (ATM) {                                                  // a CTA keeps one row tile of A: grid = multiple of m_tiles
        grid = std::max(p.m_tiles, grid / p.m_tiles * p.m_tiles);
        p.a_hi = g.A_hi; p.a_lo = g.A_lo; p.lda = g.lda; p.a_rows = (int)a_rows;
    }
    gemm_tc_kernel<BN, STAGES, VAR, H><<<grid, TC_THREADS + (GA ? GA_THREADS : 0), S::TOTAL, st>>>(mahi, malo, mbhi, mblo, mc, mchi, mclo, p);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

int launch_gemm_tc(const TcGemm &g, cudaStream_t st)
{
    ProfileScope _ps(g.tag, st);
    IQ_CHECK(tc_gemm_supported(g), "gemm_tc: unsupported shape");
    TcParams p = {};
    p.K = g.K; p.alpha = g.alpha; p.bias = g.bias; p.act = g.act; p.four_terms = g.four_terms;
    p.dbg = env_int("IQ_TC_DBG", 0);
    int64_t a_rows, b_rows;
    int bn;
    if (g.mode == 0) {
        bn = pick_bn(g.N);
        p.n_tiles = g.N / bn; p.rows_per_batch = g.rows_per_batch; p.out_c = g.C != nullptr; p.out_hilo = g.C_hi != nullptr;
        p.num_units = (int)ceil_div(g.M, TBM) * p.n_tiles; p.tiles_per_unit = 1;
        a_rows = g.M;
        b_rows = g.rows_per_batch > 0 ? g.M : g.N;
        IQ_CHECK((g.C || (g.C_hi && g.C_lo)) && g.ldc % 4 == 0, "gemm_tc: no output or misaligned leading dimension");
    } else {
        bn = 128;
        p.m_tiles = (g.cout + TBM - 1) / TBM; p.points = g.points; p.cout = g.cout;
        b_rows = (int64_t)g.clouds * g.points;
        if (g.points >= bn) { p.num_units = g.clouds * p.m_tiles; p.tiles_per_unit = g.points / bn; }
        else { p.num_units = (int)(b_rows / bn) * p.m_tiles; p.tiles_per_unit = 1; }
        p.out_max = g.out_max; p.out_mean = g.out_mean; p.out_arg = g.out_arg; p.ld_out = g.ld_out;
        p.pool_extra = g.pool_extra;
        IQ_CHECK(!g.pool_extra || g.points >= bn, "gemm_tc: pool_extra needs groups of >= 128 columns");
        a_rows = g.cout;
        IQ_CHECK(g.out_max, "gemm_tc: pooling output missing");
        IQ_CHECK(g.points >= bn || (!g.out_mean && !g.out_arg), "gemm_tc: mean / argmax need groups of >= 128 columns");
    }
    if (p.num_units == 0) return 0;
    if (g.Ah_hi) {                                              // two-term fp16 operands (kind::f16)
        IQ_CHECK(g.Ah_lo && g.Bh_hi && g.Bh_lo && !g.gather.U && g.rows_per_batch == 0 && !g.four_terms,
                 "gemm_tc: fp16 operands take the plain STORE / POOL forms");
        IQ_CHECK(g.K % 8 == 0 && g.lda % 8 == 0 && g.ldb % 8 == 0, "gemm_tc: fp16 operands need K and leading dimensions in multiples of 8");
        if (g.mode == 1) {
            IQ_CHECK(g.points >= bn, "gemm_tc: fp16 POOL needs groups of >= 128 columns");
            return launch_tc_variant<128, 3, VAR_POOL_RUN, true>(g, p, a_rows, b_rows, st);
        }
        IQ_CHECK(bn == 128, "gemm_tc: fp16 STORE needs N % 128 == 0");
        return launch_tc_variant<128, 3, VAR_STORE, true>(g, p, a_rows, b_rows, st);
    }
    if (g.gather.U) {
        IQ_CHECK(g.mode == 0 && g.N == bn, "gemm_tc: the gathered-A variant takes one column tile (N in {32, 64, 96, 128})");
        IQ_CHECK(g.gather.ldu % 4 == 0 && g.gather.ldv % 4 == 0 && g.K % 4 == 0, "gemm_tc: gathered-A widths must be multiples of 4");
        p.gU = g.gather.U; p.gV = g.gather.V; p.gbias = g.gather.bias; p.gidx = g.gather.idx; p.gldu = g.gather.ldu;
        p.gldv = g.gather.ldv; p.gK = g.gather.K; p.gS = g.gather.S; p.gnsrc = g.gather.nsrc; p.gact = g.gather.act;
        switch (bn) {
        case 128: return launch_tc_variant<128, 3, VAR_STORE_GATHER>(g, p, a_rows, b_rows, st);
        case 96: return launch_tc_variant<96, 3, VAR_STORE_GATHER>(g, p, a_rows, b_rows, st);
        case 64: return launch_tc_variant<64, 4, VAR_STORE_GATHER>(g, p, a_rows, b_rows, st);
        default: return launch_tc_variant<32, 4, VAR_STORE_GATHER>(g, p, a_rows, b_rows, st);
        }
    }
    if (g.mode == 1 && g.K <= 128 && g.lda % 4 == 0 && p.m_tiles <= sm_count() && !env_int("IQ_TC_NO_ATM", 0))
        return g.points >= bn ? launch_tc_variant<128, 6, VAR_POOL_RUN_ATM>(g, p, a_rows, b_rows, st)
                              : launch_tc_variant<128, 6, VAR_POOL_TILE_ATM>(g, p, a_rows, b_rows, st);
    if (g.mode == 1)
        return g.points >= bn ? launch_tc_variant<128, 3, VAR_POOL_RUN>(g, p, a_rows, b_rows, st)
                              : launch_tc_variant<128, 3, VAR_POOL_TILE>(g, p, a_rows, b_rows, st);
    switch (bn) {
    case 128: return launch_tc_variant<128, 3, VAR_STORE>(g, p, a_rows, b_rows, st);
    case 96: return launch_tc_variant<96, 3, VAR_STORE>(g, p, a_rows, b_rows, st);
    case 64: return launch_tc_variant<64, 4, VAR_STORE>(g, p, a_rows, b_rows, st);
    default: return launch_tc_variant<32, 4, VAR_STORE>(g, p, a_rows, b_rows, st);
    }
}

}  // namespace iq
