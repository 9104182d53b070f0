// Grouped shared MLP of a PointNet++ set-abstraction scale as ONE tcgen05 kernel (sm_100a):
//
//   H1 = relu(U[idx] - V + b1)          gathered per 128-row tile straight into the A ring (never in HBM)
//   H2 = relu(H1 W2^T + b2)             tcgen05.mma, accumulator in TMEM
//   H3 = relu(H2 W3^T + b3)             tcgen05.mma with the A operand READ FROM TMEM: the epilogue warps turn the
//                                       layer-2 accumulator into its tf32 hi / lo pair in place (tcgen05.ld -> bias,
//                                       ReLU, split -> tcgen05.st), so H2 never leaves the SM
//   out = max over the K rows of a group (the K neighbours of one centroid)      REDUX per column
//
// Reference behaviour restated (never copied): models/pointnet2.py:215-232 (per scale: three 1x1 Conv2d + BN + ReLU on
// the grouped tensor, torch.max over the neighbour axis).  Round 1 ran layers 1+2 (gemm_tc.cu, gathered-A STORE) and
// layer 3 + pool (POOL) as two kernels with H2 written to HBM as a hi/lo pair and read back: 16 bytes per element of
// the largest activation of the network, 88 MB + 88 MB per cloud, the bound of both kernels.
//
// Layout of a CTA (576 threads, one CTA per SM, persistent over 128-row tiles):
//   warp 0        TMA producer: W2 k-blocks (layer-2 stages), then W3 k-blocks (layer-3 stages) through one ring
//   warp 1        MMA issuer: layer 2 in SS mode (A = gathered tile in the ring), layer 3 in TS mode (A = TMEM)
//   warps 2-9     epilogue: thread = row of the tile (TMEM lane), two warps per lane quadrant sharing the column blocks
//                 (with one warp per quadrant the conversion + pooling of a tile, not its MMAs, paced the kernel:
//                 ncu 43 % tensor-pipe active)
//   warps 10-17   gather warps: two groups of four, thread = row, taking layer-2 ring stages in turn
// TMEM columns: layer-2 accumulator, overwritten in place by H2 hi; H2 lo; layer-3 accumulator -- two layer-2 / H2 buffers
// when 4*C2 + C3 <= 512 (the layer-2 MMAs of the next tile then run while this tile is converted), one at C2 = 128,
// C3 = 256.  3xTF32 split products as in gemm_tc.cu (Alo*Bhi + Ahi*Blo + Ahi*Bhi, small terms first).
#include "tc_ptx.cuh"
#include "kernels.cuh"

namespace iq {

using namespace tc;

namespace {

constexpr int CH_EPI_WARPS = 8;                        // two per TMEM lane quadrant, each taking every other 32-column block
constexpr int CH_GROUPS = 2;                           // gather groups of four warps (thread = row of the tile)
constexpr int CH_GATHER0 = 64 + 32 * CH_EPI_WARPS;     // first gather thread
constexpr int CH_THREADS = CH_GATHER0 + 128 * CH_GROUPS;
constexpr int CH_STAGES = 3;
constexpr int CH_A_BYTES = TBM * TBK * 4;              // one gathered A tile (hi or lo) per k-block: 16 KB

struct ChainParams {
    int C1;                      // layer-1 width = K of layer 2 (multiple of 4; zero-filled up to a multiple of 32)
    int num_units;               // 128-row tiles
    int gK, gS, gnsrc;           // neighbours per centroid, centroids per cloud, source points per cloud
    int act1;
    const float *gU, *gV, *b1;   // layer 1: relu(U[cloud*nsrc + idx[row]] - V[row / gK] + b1)
    const int32_t *gidx;
    int64_t gldu, gldv;
    const float *b2, *b3;        // folded biases of layers 2 and 3
    float *out;                  // (rows / gK, ld_out): max over each group of gK rows
    int64_t ld_out;
    int dbg;                     // IQ_CHAIN_DBG (bisecting): 1 no layer-3 MMAs, 2 no H2 conversion, 4 no gather loads, 8 no layer-2 MMAs
};

template <int C2, int C3>
struct ChainSmem {
    static constexpr int L2_STAGE = 2 * CH_A_BYTES + 2 * C2 * 128;          // A hi | A lo | W2 hi | W2 lo
    static constexpr int L3_STAGE = 2 * C3 * 128;                           // W3 hi | W3 lo
    static constexpr int STAGE_BYTES = ((L2_STAGE > L3_STAGE ? L2_STAGE : L3_STAGE) + 1023) / 1024 * 1024;
    static constexpr int POOL_BYTES = 4 * C3 * 4;                           // per-warp column maxima (groups of > 32 rows)
    static constexpr int BAR_BYTES = 256;
    static constexpr int TOTAL = CH_STAGES * STAGE_BYTES + POOL_BYTES + BAR_BYTES + 1024;
};

__device__ __forceinline__ void mbar_arrive_n(uint64_t *bar, uint32_t n)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
// the epilogue threads only (named barrier 1), leaving barrier 0 to __syncthreads
__device__ __forceinline__ void epilogue_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * CH_EPI_WARPS) : "memory"); }

// The four roles of a CTA walk the SAME sequence of ring segments.  With two layer-2 buffers in TMEM (NBUF = 2: 4*C2 + C3
// <= 512 columns) the layer-2 MMAs of tile i+1 are issued BEFORE the layer-3 MMAs of tile i, so the tensor pipe works on
// them while the epilogue warps convert / pool tile i:   L2(0) | L2(1) L3(0) | L2(2) L3(1) | ...   With one buffer
// (C2 = 128, C3 = 256 fills TMEM) the sequence is L2(0) L3(0) L2(1) L3(1) ...  l2(tile, buffer, use) / l3(tile, buffer, use,
// ordinal): `use` counts the uses of that buffer (barrier parity), `ordinal` the tiles of this CTA.
template <int NBUF, typename F2, typename F3>
__device__ __forceinline__ void chain_schedule(int first, int step, int n, F2 l2, F3 l3)
{
    if (first >= n) return;
    if (NBUF == 2) {
        l2(first, 0, 0);
        int i = 0;
        for (int u = first; u < n; u += step, ++i) {
            if (u + step < n) l2(u + step, (i + 1) & 1, (i + 1) >> 1);
            l3(u, i & 1, i >> 1, i);
        }
    } else {
        int i = 0;
        for (int u = first; u < n; u += step, ++i) {
            l2(u, 0, i);
            l3(u, 0, i, i);
        }
    }
}

template <int C2, int C3>
__global__ void __launch_bounds__(CH_THREADS, 1)
sa_chain_kernel(const __grid_constant__ CUtensorMap map_w2hi, const __grid_constant__ CUtensorMap map_w2lo,
                const __grid_constant__ CUtensorMap map_w3hi, const __grid_constant__ CUtensorMap map_w3lo,
                const ChainParams p)
{
    using S = ChainSmem<C2, C3>;
    static_assert(C2 % 32 == 0 && C2 <= 128 && C3 % 32 == 0 && C3 <= 256, "chain: widths out of range");
    constexpr int KB3 = C2 / TBK;                                           // layer-3 k-blocks
    constexpr int NBUF = 4 * C2 + C3 <= 512 ? 2 : 1;                        // layer-2 accumulator / H2 buffers in TMEM
    // TMEM columns: H2 hi (over the layer-2 accumulator) of buffer b at b*C2, H2 lo at LO + b*C2, layer-3 accumulator at ACC3
    constexpr uint32_t LO = NBUF == 2 ? 2 * C2 : 128, ACC3 = NBUF == 2 ? 4 * C2 : 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float *pool_s = reinterpret_cast<float *>(smem + CH_STAGES * S::STAGE_BYTES);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + CH_STAGES * S::STAGE_BYTES + S::POOL_BYTES);
    uint64_t *empty_bar = full_bar + CH_STAGES;
    uint64_t *acc2_full = empty_bar + CH_STAGES;                            // [2]
    uint64_t *a3_ready = acc2_full + 2;                                     // [2]
    uint64_t *acc3_full = a3_ready + 2;
    uint64_t *acc3_empty = acc3_full + 1;
    uint64_t *go = acc3_empty + 1;                                          // [groups][2]: producer -> gather group "your slot is free"
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(go + 2 * CH_GROUPS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb2n = (p.C1 + TBK - 1) / TBK;                                 // layer-2 k-blocks (TMA zero-fills the K tail)
    const int first = blockIdx.x, step = gridDim.x, n = p.num_units;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_w2hi); prefetch_tmap(&map_w2lo); prefetch_tmap(&map_w3hi); prefetch_tmap(&map_w3lo);
        for (int s = 0; s < CH_STAGES; ++s) { mbar_init(&full_bar[s], 5); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc2_full[b], 1); mbar_init(&a3_ready[b], CH_EPI_WARPS); }
        mbar_init(acc3_full, 1);
        mbar_init(acc3_empty, CH_EPI_WARPS);
        for (int g = 0; g < 2 * CH_GROUPS; ++g) mbar_init(&go[g], 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // ring position shared by every role: each thread advances it over all segments, whether it takes part or not
    int stage = 0;
    uint32_t phase = 0;
    auto advance = [&]() { if (++stage == CH_STAGES) { stage = 0; phase ^= 1; } };

    if (warp == 0) {
        // ---- TMA producer
        // The producer visits every ring position, so its parity waits on the empty barriers are always exactly one phase
        // back.  The gather groups only take part in every third layer-2 position: a parity wait of theirs on a slot could
        // be two phases ahead (reads "free" too early) or two behind (waits for a phase that needs their own data), so the
        // producer hands each layer-2 slot to its owner group through that group's own barriers instead.  The ring lets the
        // producer run at most three positions -- one more slot of the same group -- ahead of the slot a group is working on,
        // so a group alternates between two barriers (its j-th slot uses barrier j & 1, phase j >> 1): never more than one
        // arrival outstanding per barrier.
        int turn = 0;
        uint32_t owned[CH_GROUPS] = {};
        chain_schedule<NBUF>(first, step, n,
            [&](int, int, int) {
                for (int kb = 0; kb < kb2n; ++kb, turn = turn + 1 == CH_GROUPS ? 0 : turn + 1) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t *st = smem + stage * S::STAGE_BYTES;
                    if (elect_one_sync()) {
                        mbar_arrive(&go[2 * turn + (int)(owned[turn] & 1u)]);
                        mbar_arrive_expect_tx(&full_bar[stage], 2 * C2 * 128);
                        tma_load_2d(st + 2 * CH_A_BYTES, &map_w2hi, &full_bar[stage], kb * TBK, 0);
                        tma_load_2d(st + 2 * CH_A_BYTES + C2 * 128, &map_w2lo, &full_bar[stage], kb * TBK, 0);
                    }
                    __syncwarp();
                    ++owned[turn];
                    advance();
                }
            },
            [&](int, int, int, int) {
                for (int kb = 0; kb < KB3; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t *st = smem + stage * S::STAGE_BYTES;
                    if (elect_one_sync()) {
                        mbar_arrive_n(&full_bar[stage], 4);                  // the gather warps have no part in these stages
                        mbar_arrive_expect_tx(&full_bar[stage], 2 * C3 * 128);
                        tma_load_2d(st, &map_w3hi, &full_bar[stage], kb * TBK, 0);
                        tma_load_2d(st + C3 * 128, &map_w3lo, &full_bar[stage], kb * TBK, 0);
                    }
                    __syncwarp();
                    advance();
                }
            });
    } else if (warp == 1) {
        // ---- MMA issuer
        constexpr uint32_t idesc2 = make_idesc(C2), idesc3 = make_idesc(C3);
        chain_schedule<NBUF>(first, step, n,
            [&](int, int buf, int) {
                const uint32_t d2 = tmem_base + (uint32_t)(buf * C2);
                for (int kb = 0; kb < kb2n; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sbase = smem_u32(smem + stage * S::STAGE_BYTES);
                    const uint64_t ahi = make_smem_desc(sbase), alo = make_smem_desc(sbase + CH_A_BYTES);
                    const uint64_t bhi = make_smem_desc(sbase + 2 * CH_A_BYTES);
                    const uint64_t blo = make_smem_desc(sbase + 2 * CH_A_BYTES + C2 * 128);
                    if (elect_one_sync()) {
#pragma unroll
                        for (int term = 0; term < 3; ++term) {
                            if (p.dbg & 8) break;
                            const uint64_t ad = term == 0 ? alo : ahi;
                            const uint64_t bd = term == 1 ? blo : bhi;
#pragma unroll
                            for (int ks = 0; ks < TBK / UMMA_K; ++ks) {
                                const uint64_t koff = (uint64_t)((ks * UMMA_K * 4) >> 4);
                                umma_tf32(d2, ad + koff, bd + koff, idesc2, (kb | term | ks) != 0 ? 1u : 0u);
                            }
                        }
                        umma_commit(&empty_bar[stage]);
                        if (kb == kb2n - 1) umma_commit(&acc2_full[buf]);
                    }
                    __syncwarp();
                    advance();
                }
            },
            [&](int, int buf, int use, int ord) {
                mbar_wait(&a3_ready[buf], (uint32_t)(use & 1));              // H2 hi / lo parked in TMEM by the epilogue warps
                mbar_wait(acc3_empty, (uint32_t)((ord & 1) ^ 1));            // the previous tile's pooling has read its accumulator
                tc_fence_after();
                const uint32_t a_hi = tmem_base + (uint32_t)(buf * C2), a_lo = tmem_base + LO + (uint32_t)(buf * C2);
                for (int kb = 0; kb < KB3; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sbase = smem_u32(smem + stage * S::STAGE_BYTES);
                    const uint64_t bhi = make_smem_desc(sbase), blo = make_smem_desc(sbase + C3 * 128);
                    if (elect_one_sync()) {
#pragma unroll
                        for (int term = 0; term < 3; ++term) {
                            if (p.dbg & 1) break;
                            const uint64_t bd = term == 1 ? blo : bhi;
#pragma unroll
                            for (int ks = 0; ks < TBK / UMMA_K; ++ks) {
                                const uint64_t koff = (uint64_t)((ks * UMMA_K * 4) >> 4);
                                umma_tf32_ts(tmem_base + ACC3, (term == 0 ? a_lo : a_hi) + (uint32_t)(kb * TBK + ks * UMMA_K),
                                             bd + koff, idesc3, (kb | term | ks) != 0 ? 1u : 0u);
                            }
                        }
                        umma_commit(&empty_bar[stage]);
                        if (kb == KB3 - 1) umma_commit(acc3_full);
                    }
                    __syncwarp();
                    advance();
                }
            });
    } else if (warp >= 2 + CH_EPI_WARPS) {
        // ---- gather warps: thread = row of the tile; the three groups of four warps take layer-2 stages in turn and are
        // told by the producer (go[group], one phase per owned slot) when their slot is free
        const int r = ((int)threadIdx.x - CH_GATHER0) & 127, grp = ((int)threadIdx.x - CH_GATHER0) >> 7;
        const int sw = r & 7;
        int turn = 0;
        uint32_t mine_n = 0;                                                 // layer-2 slots this group has taken
        chain_schedule<NBUF>(first, step, n,
            [&](int unit, int, int) {
                const int64_t row = (int64_t)unit * TBM + r;
                const int64_t cen = row / p.gK, cloud = cen / p.gS;
                const float *u = p.gU + (cloud * p.gnsrc + __ldg(p.gidx + row)) * p.gldu;
                const float *v = p.gV + cen * p.gldv;
                for (int kb = 0; kb < kb2n; ++kb, turn = turn + 1 == CH_GROUPS ? 0 : turn + 1) {
                    if (turn != grp) {
                        advance();
                        continue;
                    }
                    float4 uu[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        uu[q] = (kb * TBK + 4 * q < p.C1 && !(p.dbg & 4)) ? __ldg(reinterpret_cast<const float4 *>(u + kb * TBK) + q)
                                                                          : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    mbar_wait(&go[2 * grp + (int)(mine_n & 1u)], (mine_n >> 1) & 1u);
                    ++mine_n;
                    const uint32_t hi_row = smem_u32(smem + stage * S::STAGE_BYTES) + (uint32_t)r * 128u;
                    const uint32_t lo_row = hi_row + CH_A_BYTES;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float h[4] = {0.0f, 0.0f, 0.0f, 0.0f}, l[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                        if (kb * TBK + 4 * q < p.C1) {
                            const float4 vv = __ldg(reinterpret_cast<const float4 *>(v + kb * TBK) + q);
                            const float4 bb = __ldg(reinterpret_cast<const float4 *>(p.b1 + kb * TBK) + q);
                            const float x[4] = {apply_act((uu[q].x - vv.x) + bb.x, p.act1), apply_act((uu[q].y - vv.y) + bb.y, p.act1),
                                                apply_act((uu[q].z - vv.z) + bb.z, p.act1), apply_act((uu[q].w - vv.w) + bb.w, p.act1)};
#pragma unroll
                            for (int e = 0; e < 4; ++e) { h[e] = tf32_round(x[e]); l[e] = tf32_round(x[e] - h[e]); }
                        }
                        const uint32_t off = (uint32_t)((q ^ sw) << 4);
                        sts128(hi_row + off, h[0], h[1], h[2], h[3]);
                        sts128(lo_row + off, l[0], l[1], l[2], l[3]);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_bar[stage]);
                    advance();
                }
            },
            [&](int, int, int, int) {
                for (int kb = 0; kb < KB3; ++kb) advance();                  // layer-3 slots belong to the producer alone
            });
    } else {
        // ---- epilogue warps: thread = row of the tile = TMEM lane
        const int quad = warp & 3;                                          // TMEM lane quadrant = rows [32 * quad, 32 * quad + 32)
        const int half = (warp - 2) >> 2;                                   // which of the quadrant's warps: 32-column blocks half, half + 2, ...
        constexpr int CSTEP = 32 * (CH_EPI_WARPS / 4);
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        chain_schedule<NBUF>(first, step, n,
            [&](int, int buf, int use) {
                // layer-2 accumulator -> H2 = relu(acc + b2) -> tf32 hi over the accumulator, lo in the lo buffer
                mbar_wait(&acc2_full[buf], (uint32_t)(use & 1));
                tc_fence_after();
#pragma unroll 1
                for (int c0 = 32 * half; c0 < C2; c0 += CSTEP) {
                    if (p.dbg & 2) break;
                    float v[32];
                    tmem_ld32(lane_base + (uint32_t)(buf * C2 + c0), v);
                    uint32_t hi[32], lo[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float y = fmaxf(v[i] + __ldg(p.b2 + c0 + i), 0.0f);
                        const float h = tf32_round(y);
                        hi[i] = __float_as_uint(h);
                        lo[i] = __float_as_uint(tf32_round(y - h));
                    }
                    tmem_st32(lane_base + (uint32_t)(buf * C2 + c0), hi);
                    tmem_st32(lane_base + LO + (uint32_t)(buf * C2 + c0), lo);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a3_ready[buf]);
            },
            [&](int unit, int, int, int ord) {
                // layer-3 accumulator -> relu(acc + b3) -> max over each group of gK rows -> out
                mbar_wait(acc3_full, (uint32_t)(ord & 1));
                tc_fence_after();
                const int64_t grp0 = (int64_t)unit * (TBM / p.gK);           // first group of this tile (gK divides 128)
#pragma unroll 1
                for (int c0 = 32 * half; c0 < C3; c0 += CSTEP) {
                    float v[32];
                    tmem_ld32(lane_base + ACC3 + c0, v);
                    float mine = 0.0f;                                      // lane l keeps column c0 + l of its group
                    if (p.gK >= 32) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float y = fmaxf(v[i] + __ldg(p.b3 + c0 + i), 0.0f);      // >= 0: uint order == float order
                            const uint32_t m = __reduce_max_sync(0xffffffffu, __float_as_uint(y));
                            if (lane == i) mine = __uint_as_float(m);
                        }
                        if (p.gK == 32) p.out[(grp0 + quad) * p.ld_out + c0 + lane] = mine;
                        else pool_s[quad * C3 + c0 + lane] = mine;
                    } else {                                                // gK == 16: two groups per warp
                        // REDUX returns one warp-uniform value, so the two half-warp groups are reduced one after the other
                        // with the other half's lanes contributing the neutral 0 (ReLU outputs are >= 0)
                        float other = 0.0f;                                 // lanes hold columns l and l + 16 of their half's group
                        const bool upper = lane >= 16;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const uint32_t y = __float_as_uint(fmaxf(v[i] + __ldg(p.b3 + c0 + i), 0.0f));
                            const uint32_t m0 = __reduce_max_sync(0xffffffffu, upper ? 0u : y);
                            const uint32_t m1 = __reduce_max_sync(0xffffffffu, upper ? y : 0u);
                            const float m = __uint_as_float(upper ? m1 : m0);
                            if ((lane & 15) == (i & 15)) { if (i < 16) mine = m; else other = m; }
                        }
                        float *o = p.out + (grp0 + quad * 2 + (lane >> 4)) * p.ld_out + c0 + (lane & 15);
                        o[0] = mine;
                        o[16] = other;
                    }
                }
                if (p.gK > 32) {                                            // groups of 64 / 128 rows: combine the warps' maxima
                    epilogue_sync();
                    const int per = p.gK / 32;                              // warps per group
                    const int groups = 4 / per;
                    for (int t = (int)threadIdx.x - 64; t < groups * C3; t += 32 * CH_EPI_WARPS) {
                        const int g = t / C3, c = t - g * C3;
                        float m = pool_s[(g * per) * C3 + c];
                        for (int w = 1; w < per; ++w) m = fmaxf(m, pool_s[(g * per + w) * C3 + c]);
                        p.out[(grp0 + g) * p.ld_out + c] = m;
                    }
                    epilogue_sync();                                        // pool_s is reused by the next tile
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc3_empty);
            });
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int C2, int C3>
int launch_chain_variant(const SaChain &g, const ChainParams &p, cudaStream_t st)
{
    using S = ChainSmem<C2, C3>;
    static_assert(S::TOTAL <= 232448, "chain: shared memory budget exceeded");
    CUtensorMap w2hi, w2lo, w3hi, w3lo;
    if (int rc = make_map(&w2hi, g.W2_hi, C2, g.C1, g.ldw2, C2)) return rc;
    if (int rc = make_map(&w2lo, g.W2_lo, C2, g.C1, g.ldw2, C2)) return rc;
    if (int rc = make_map(&w3hi, g.W3_hi, C3, C2, g.ldw3, C3)) return rc;
    if (int rc = make_map(&w3lo, g.W3_lo, C3, C2, g.ldw3, C3)) return rc;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(&sa_chain_kernel<C2, C3>), S::TOTAL)) return rc;
    const int grid = std::min(p.num_units, sm_count());
    sa_chain_kernel<C2, C3><<<grid, CH_THREADS, S::TOTAL, st>>>(w2hi, w2lo, w3hi, w3lo, p);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

}  // namespace

bool sa_chain_supported(const SaChain &g)
{
    const bool widths = (g.C2 == 32 && g.C3 == 64) || (g.C2 == 64 && g.C3 == 128) || (g.C2 == 96 && g.C3 == 128) ||
                        (g.C2 == 128 && g.C3 == 256);
    return widths && g.C1 >= 4 && g.C1 % 4 == 0 && g.C1 <= 128 && g.rows % TBM == 0 &&
           (g.K == 16 || g.K == 32 || g.K == 64 || g.K == 128) && g.ldu % 4 == 0 && g.ldv % 4 == 0 && g.ldw2 % 4 == 0 &&
           g.ldw3 % 4 == 0 && g.rows / TBM < ((int64_t)1 << 31);
}

int launch_sa_chain(const SaChain &g, cudaStream_t st)
{
    ProfileScope _ps(g.tag, st);
    IQ_CHECK(sa_chain_supported(g), "sa_chain: unsupported shape");
    if (g.rows == 0) return 0;
    ChainParams p = {};
    p.C1 = g.C1; p.num_units = (int)(g.rows / TBM); p.gK = g.K; p.gS = g.S; p.gnsrc = g.nsrc; p.act1 = ACT_RELU;
    p.gU = g.U; p.gV = g.V; p.b1 = g.b1; p.gidx = g.idx; p.gldu = g.ldu; p.gldv = g.ldv; p.b2 = g.b2; p.b3 = g.b3;
    p.out = g.out; p.ld_out = g.ld_out;
    p.dbg = env_int("IQ_CHAIN_DBG", 0);
    if (g.C2 == 32) return launch_chain_variant<32, 64>(g, p, st);
    if (g.C2 == 64) return launch_chain_variant<64, 128>(g, p, st);
    if (g.C2 == 96) return launch_chain_variant<96, 128>(g, p, st);
    return launch_chain_variant<128, 256>(g, p, st);
}

}  // namespace iq
