/*
 * ORACLE (test infrastructure, not product code).
 *
 * Plain-C restatement of the integer-valued geometry steps of the reference's
 * coalition path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (interpret_quality_b200/) never does.
 *
 * Compile with -ffp-contract=off: every rounding below is spelled out, and
 * where the reference's CPU BLAS fuses a multiply-add it is written as fmaf().
 *
 * Pinned against the reference itself (imported from /root/reference in the
 * build container) by tests/golden/make_golden.py -> tests/golden/geometry.npz.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* Farthest point sampling.
 * Follows final_save_fps.py:10-31 (same loop in models/pointnet2.py:45-68 and
 * models/pointconv.py:54-77): start at index 0, running distance initialised
 * to 1e10, d = sum((xyz-c)**2,-1) evaluated as ((dx*dx + dy*dy) + dz*dz) in
 * fp32, update where d < running, next = argmax with the lowest index on ties.
 */
void oracle_fps(const float *xyz, int64_t B, int64_t N, int64_t npoint, int64_t *out)
{
    float *dist = (float *)malloc(sizeof(float) * (size_t)N);
    for (int64_t b = 0; b < B; ++b) {
        const float *p = xyz + b * N * 3;
        for (int64_t i = 0; i < N; ++i) dist[i] = 1e10f;
        int64_t far = 0;
        for (int64_t s = 0; s < npoint; ++s) {
            out[b * npoint + s] = far;
            const float cx = p[far * 3 + 0], cy = p[far * 3 + 1], cz = p[far * 3 + 2];
            float best = -1.0f;
            int64_t arg = 0;
            for (int64_t i = 0; i < N; ++i) {
                const float dx = p[i * 3 + 0] - cx, dy = p[i * 3 + 1] - cy, dz = p[i * 3 + 2] - cz;
                const float xx = dx * dx, yy = dy * dy, zz = dz * dz;
                const float d = (xx + yy) + zz;
                if (d < dist[i]) dist[i] = d;
                if (dist[i] > best) { best = dist[i]; arg = i; }
            }
            far = arg;
        }
    }
    free(dist);
}

static inline float sqnorm3(const float *p)
{
    const float xx = p[0] * p[0], yy = p[1] * p[1], zz = p[2] * p[2];
    return (xx + yy) + zz;
}

/* square_distance for C == 3.
 * Follows tools/final_util.py:134-147 (copies: models/pointnet2.py:12-25,
 * models/pointconv.py:13-31): dist = -2 * (src @ dst^T); dist += |src|^2;
 * dist += |dst|^2.  The K=3 product on the CPU BLAS is an fma chain over k.
 * src (B,N,3), dst (B,M,3) -> out (B,N,M).
 */
void oracle_square_distance3(const float *src, const float *dst, int64_t B, int64_t N, int64_t M, float *out)
{
    for (int64_t b = 0; b < B; ++b)
        for (int64_t i = 0; i < N; ++i) {
            const float *s = src + (b * N + i) * 3;
            const float ss = sqnorm3(s);
            for (int64_t j = 0; j < M; ++j) {
                const float *d = dst + (b * M + j) * 3;
                float dot = s[0] * d[0];
                dot = fmaf(s[1], d[1], dot);
                dot = fmaf(s[2], d[2], dot);
                float t = -2.0f * dot;
                t = t + ss;
                t = t + sqnorm3(d);
                out[(b * N + i) * M + j] = t;
            }
        }
}

/* Nearest-centre region assignment.
 * Follows cal_region_id, final_shapley_value.py:20-35: argmin over the 32
 * centres of square_distance(data, data[fps_index]); first minimum wins.
 */
void oracle_region_id(const float *xyz, const int64_t *fps_index, int64_t N, int64_t R, int64_t *region_id)
{
    float *c = (float *)malloc(sizeof(float) * (size_t)R * 3);
    float *d = (float *)malloc(sizeof(float) * (size_t)N * (size_t)R);
    for (int64_t r = 0; r < R; ++r)
        for (int k = 0; k < 3; ++k) c[r * 3 + k] = xyz[fps_index[r] * 3 + k];
    oracle_square_distance3(xyz, c, 1, N, R, d);
    for (int64_t i = 0; i < N; ++i) {
        int64_t arg = 0;
        float best = d[i * R];
        for (int64_t r = 1; r < R; ++r)
            if (d[i * R + r] < best) { best = d[i * R + r]; arg = r; }
        region_id[i] = arg;
    }
    free(c);
    free(d);
}

/* Ball query.
 * Follows query_ball_point, models/pointnet2.py:70-91: candidates whose squared
 * distance is NOT greater than radius**2 are kept in index order, the first K
 * are taken and the tail is padded with the first hit.  radius**2 is evaluated
 * in double like the Python scalar and compared after promotion of the fp32
 * distance, which is what `sqrdists > radius**2` does for a 0-dim double.
 * xyz (B,N,3), new_xyz (B,S,3) -> idx (B,S,K).
 */
void oracle_ball_query(const float *xyz, const float *new_xyz, int64_t B, int64_t N, int64_t S, int64_t K,
                       double radius, int64_t *idx)
{
    /* torch compares a float tensor with a python scalar in the tensor's dtype */
    const float r2 = (float)(radius * radius);
    float *d = (float *)malloc(sizeof(float) * (size_t)N);
    for (int64_t b = 0; b < B; ++b)
        for (int64_t s = 0; s < S; ++s) {
            oracle_square_distance3(new_xyz + (b * S + s) * 3, xyz + b * N * 3, 1, 1, N, d);
            int64_t cnt = 0;
            int64_t *o = idx + (b * S + s) * K;
            for (int64_t i = 0; i < N && cnt < K; ++i)
                if (!(d[i] > r2)) o[cnt++] = i;
            /* a centroid is one of the points, so cnt >= 1 unless the distance to itself rounds above r2 */
            const int64_t first = cnt > 0 ? o[0] : N;
            for (; cnt < K; ++cnt) o[cnt] = first;
        }
    free(d);
}

/* Shapley coalition expansion.
 * Follows mask_data_batch, tools/final_common.py:46-61 (single-permutation
 * form mask_data, final_shapley_value.py:74-88): row r of permutation p keeps
 * the regions order[0..r-1] and has the points of order[r..R-1] set to center.
 * data (N,3), orders (bs,R), region_id (N) -> out ((R+1)*bs, N, 3).
 */
void oracle_mask_shapley(const float *data, const float *center, const int64_t *orders, const int64_t *region_id,
                         int64_t bs, int64_t R, int64_t N, float *out)
{
    int64_t *rank = (int64_t *)malloc(sizeof(int64_t) * (size_t)R);
    for (int64_t p = 0; p < bs; ++p) {
        for (int64_t r = 0; r < R; ++r) rank[r] = -1;
        for (int64_t r = 0; r < R; ++r) rank[orders[p * R + r]] = r;
        for (int64_t row = 0; row <= R; ++row) {
            float *o = out + ((p * (R + 1) + row) * N) * 3;
            for (int64_t i = 0; i < N; ++i) {
                /* region with position q in the order is masked in rows 0..q */
                const int64_t q = rank[region_id[i]];
                const int masked = (q >= 0) && (row <= q);
                for (int k = 0; k < 3; ++k) o[i * 3 + k] = masked ? center[k] : data[i * 3 + k];
            }
        }
    }
    free(rank);
}

/* Interaction coalition expansion.
 * Follows final_point_binary_interaction_logits.py:42-56: for context k the
 * four clouds 4k..4k+3 keep S+{i,j}, S+{i}, S+{j}, S; everything else is
 * center.  The reference evaluates data*mask + center*(1-mask)-style
 * arithmetic: kept entries are data*1 + 0, masked entries are data*0 + center.
 * data (N,3) point-major, contexts (ctx,m) -> out (4*ctx, 3, N) channel-first.
 */
void oracle_mask_interaction(const float *data, const float *center, const int64_t *contexts, int64_t ctx, int64_t m,
                             int64_t region_i, int64_t region_j, const int64_t *region_id, int64_t R, int64_t N,
                             float *out)
{
    unsigned char *inS = (unsigned char *)malloc((size_t)R);
    for (int64_t k = 0; k < ctx; ++k) {
        for (int64_t r = 0; r < R; ++r) inS[r] = 0;
        for (int64_t t = 0; t < m; ++t) {
            const int64_t r = contexts[k * m + t];
            if (r >= 0 && r < R) inS[r] = 1;
        }
        for (int v = 0; v < 4; ++v)
            for (int64_t i = 0; i < N; ++i) {
                const int64_t r = region_id[i];
                int keep = inS[r];
                if (r == region_i && (v == 0 || v == 1)) keep = 1;
                if (r == region_j && (v == 0 || v == 2)) keep = 1;
                for (int c = 0; c < 3; ++c) {
                    const float x = data[i * 3 + c];
                    float y = keep ? x * 1.0f : x * 0.0f;
                    y = y + (keep ? 0.0f : center[c]);
                    out[((k * 4 + v) * 3 + c) * N + i] = y;
                }
            }
    }
    free(inS);
}
