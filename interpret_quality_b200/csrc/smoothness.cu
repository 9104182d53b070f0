// Region geometry ascent / descent: one epoch of final_smoothness_center_enum_all.py (update_region :184-243 for
// every region that is still being updated, :305-321) in ONE launch.
//
// The reference runs, per region and per gradient step, ~40 tiny torch ops plus two .item() round trips
// (cal_variance :48-62, apply_var_bound :65-73, sort_var :85-100, the mode's ratio and its autograd backward
// :205-221, gradient_descent :123-139, apply_distance_bound :103-120, check_stop_condition :167-181): ~10^5
// launches per epoch.  A region only reads and writes its own points, so the regions of an epoch are independent:
// here a CTA owns a region, keeps its points in shared memory and walks the whole while-loop (up to
// max_iteration+1 steps) with the gradient of the variance ratio in closed form, in fp32 like the reference.
//
// Work per step is O(S) flops on S ~ N/R points; the kernel is latency-bound by its block reductions (4 per step),
// not by HBM or the ALUs: it exists to remove the launch and synchronisation overhead, not to hit a roofline.
#include "common.cuh"
#include "kernels.cuh"

namespace iq {

namespace {

constexpr int SM_THREADS = 128;
constexpr int SM_WARPS = SM_THREADS / 32;

struct Sum4 {
    float a, b, c, d;
};

// fixed-order block sum of four floats; every thread receives the totals
__device__ __forceinline__ Sum4 block_sum4(Sum4 v, float (*scratch)[4])
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.a += __shfl_xor_sync(0xffffffffu, v.a, o);
        v.b += __shfl_xor_sync(0xffffffffu, v.b, o);
        v.c += __shfl_xor_sync(0xffffffffu, v.c, o);
        v.d += __shfl_xor_sync(0xffffffffu, v.d, o);
    }
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        scratch[warp][0] = v.a;
        scratch[warp][1] = v.b;
        scratch[warp][2] = v.c;
        scratch[warp][3] = v.d;
    }
    __syncthreads();
    Sum4 r = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int w = 0; w < SM_WARPS; ++w) {
        r.a += scratch[w][0];
        r.b += scratch[w][1];
        r.c += scratch[w][2];
        r.d += scratch[w][3];
    }
    __syncthreads();
    return r;
}

// modes as final_smoothness_center_enum_all.py:205-221
enum { MODE_LINEARITY = 0, MODE_PLANARITY = 1, MODE_SCATTERING = 2 };

__global__ void __launch_bounds__(SM_THREADS)
region_smoothness_epoch_kernel(float *__restrict__ data, const float *__restrict__ data_orig,
                               const int32_t *__restrict__ offsets, const int32_t *__restrict__ members,
                               const float *__restrict__ orient, const float *__restrict__ var_ub,
                               const float *__restrict__ var_lb, double *__restrict__ smooth, int32_t *__restrict__ alive,
                               int32_t *__restrict__ iters, float *__restrict__ last_var, int32_t *__restrict__ stop_flags,
                               int mode, int rising, float step, double enum_step, float dist_thr, double stop_ratio,
                               int max_iteration, int clamp)
{
    extern __shared__ float sm_pts[];                 // cur (S,3) | orig (S,3)
    __shared__ float scratch[SM_WARPS][4];
    const int r = blockIdx.x;
    const int tid = threadIdx.x;
    if (!alive[r]) {
        if (tid == 0) iters[r] = 0;
        return;
    }
    const int beg = offsets[r], S = offsets[r + 1] - beg;
    float *cur = sm_pts, *org = sm_pts + 3 * S;
    for (int i = tid; i < 3 * S; i += SM_THREADS) {
        const int p = members[beg + i / 3], c = i % 3;
        cur[i] = data[3 * (size_t)p + c];
        org[i] = data_orig[3 * (size_t)p + c];
    }
    float o[3][3], ub[3], lb[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int c = 0; c < 3; ++c) o[k][c] = orient[r * 9 + k * 3 + c];
        ub[k] = var_ub[r * 3 + k];
        lb[k] = var_lb[r * 3 + k];
    }
    __syncthreads();

    const double smooth0 = smooth[r];
    const double target = rising ? smooth0 + enum_step : smooth0 - enum_step;       // :198, Python floats
    double sm_now = smooth0;
    int it = 0, keep = 1, flags = 0;
    float v[3] = {0.f, 0.f, 0.f};
    const float inv_sm1 = 2.0f / (float)(S - 1);                                      // torch.var backward: 2/(S-1)

    while (rising ? (sm_now < target) : (sm_now > target)) {
        // unbiased variance of the three projections (cal_variance :48-62), two passes
        Sum4 acc = {0.f, 0.f, 0.f, 0.f};
        for (int s = tid; s < S; s += SM_THREADS) {
            const float x = cur[3 * s], y = cur[3 * s + 1], z = cur[3 * s + 2];
            acc.a += fmaf(z, o[0][2], fmaf(y, o[0][1], x * o[0][0]));
            acc.b += fmaf(z, o[1][2], fmaf(y, o[1][1], x * o[1][0]));
            acc.c += fmaf(z, o[2][2], fmaf(y, o[2][1], x * o[2][0]));
        }
        acc = block_sum4(acc, scratch);
        const float mean[3] = {acc.a / (float)S, acc.b / (float)S, acc.c / (float)S};
        Sum4 dev = {0.f, 0.f, 0.f, 0.f};
        for (int s = tid; s < S; s += SM_THREADS) {
            const float x = cur[3 * s], y = cur[3 * s + 1], z = cur[3 * s + 2];
            const float d0 = fmaf(z, o[0][2], fmaf(y, o[0][1], x * o[0][0])) - mean[0];
            const float d1 = fmaf(z, o[1][2], fmaf(y, o[1][1], x * o[1][0])) - mean[1];
            const float d2 = fmaf(z, o[2][2], fmaf(y, o[2][1], x * o[2][0])) - mean[2];
            dev.a = fmaf(d0, d0, dev.a);
            dev.b = fmaf(d1, d1, dev.b);
            dev.c = fmaf(d2, d2, dev.c);
        }
        dev = block_sum4(dev, scratch);
        v[0] = dev.a / (float)(S - 1);
        v[1] = dev.b / (float)(S - 1);
        v[2] = dev.c / (float)(S - 1);
        // apply_var_bound :65-73: a variance outside its bound is a constant for the gradient
        bool live[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) live[k] = !(v[k] > ub[k] || v[k] < lb[k]);
        // sort_var :85-100 (np.argsort: ascending, first index wins a tie)
        int i_min = 0, i_mid = 1, i_max = 2;
        if (v[i_mid] < v[i_min]) { int t = i_min; i_min = i_mid; i_mid = t; }
        if (v[i_max] < v[i_mid]) { int t = i_mid; i_mid = i_max; i_max = t; }
        if (v[i_mid] < v[i_min]) { int t = i_min; i_min = i_mid; i_mid = t; }
        const float s_min = v[i_min], s_mid = v[i_mid], s_max = v[i_max];
        // the mode's ratio and d ratio / d variance as autograd forms them (a / b: 1/b and -(a/b)/b)
        float coef[3] = {0.f, 0.f, 0.f}, f;
        bool any_grad;
        if (mode == MODE_LINEARITY) {
            const float a = s_max - s_mid;
            f = a / s_max;
            any_grad = live[i_max] || live[i_mid];
            coef[i_max] = 1.0f / s_max - (a / s_max) / s_max;
            coef[i_mid] = -(1.0f / s_max);
        } else if (mode == MODE_PLANARITY) {
            const float a = s_mid - s_min;
            f = a / s_max;
            any_grad = live[i_max] || live[i_mid] || live[i_min];
            coef[i_max] = -((a / s_max) / s_max);
            coef[i_mid] = 1.0f / s_max;
            coef[i_min] = -(1.0f / s_max);
        } else {
            f = s_min / s_max;
            any_grad = live[i_max] || live[i_min];
            coef[i_max] = -((s_min / s_max) / s_max);
            coef[i_min] = 1.0f / s_max;
        }
        sm_now = (double)f;
#pragma unroll
        for (int k = 0; k < 3; ++k) coef[k] = live[k] ? coef[k] * inv_sm1 : 0.0f;
        const bool grad_none = !any_grad;                                             // :131, :137
        // gradient_descent :123-139: x +- step * g / |g|
        if (!grad_none) {
            Sum4 nrm = {0.f, 0.f, 0.f, 0.f};
            for (int s = tid; s < S; s += SM_THREADS) {
                const float x = cur[3 * s], y = cur[3 * s + 1], z = cur[3 * s + 2];
                const float w0 = coef[0] * (fmaf(z, o[0][2], fmaf(y, o[0][1], x * o[0][0])) - mean[0]);
                const float w1 = coef[1] * (fmaf(z, o[1][2], fmaf(y, o[1][1], x * o[1][0])) - mean[1]);
                const float w2 = coef[2] * (fmaf(z, o[2][2], fmaf(y, o[2][1], x * o[2][0])) - mean[2]);
                const float gx = w0 * o[0][0] + w1 * o[1][0] + w2 * o[2][0];
                const float gy = w0 * o[0][1] + w1 * o[1][1] + w2 * o[2][1];
                const float gz = w0 * o[0][2] + w1 * o[1][2] + w2 * o[2][2];
                nrm.a += gx * gx + gy * gy + gz * gz;
            }
            nrm = block_sum4(nrm, scratch);
            const float norm = sqrtf(nrm.a);
            for (int s = tid; s < S; s += SM_THREADS) {
                const float x = cur[3 * s], y = cur[3 * s + 1], z = cur[3 * s + 2];
                const float w0 = coef[0] * (fmaf(z, o[0][2], fmaf(y, o[0][1], x * o[0][0])) - mean[0]);
                const float w1 = coef[1] * (fmaf(z, o[1][2], fmaf(y, o[1][1], x * o[1][0])) - mean[1]);
                const float w2 = coef[2] * (fmaf(z, o[2][2], fmaf(y, o[2][1], x * o[2][0])) - mean[2]);
                const float gx = w0 * o[0][0] + w1 * o[1][0] + w2 * o[2][0];
                const float gy = w0 * o[0][1] + w1 * o[1][1] + w2 * o[2][1];
                const float gz = w0 * o[0][2] + w1 * o[1][2] + w2 * o[2][2];
                float dx = 1e-8f, dy = 1e-8f, dz = 1e-8f;                             // :135, |g| == 0
                if (norm != 0.0f) {
                    dx = step * gx / norm;
                    dy = step * gy / norm;
                    dz = step * gz / norm;
                }
                cur[3 * s] = rising ? x + dx : x - dx;
                cur[3 * s + 1] = rising ? y + dy : y - dy;
                cur[3 * s + 2] = rising ? z + dz : z - dz;
            }
        }
        // apply_distance_bound :103-120.  The reference counts the points beyond dist_threshold; its pull-back
        // assigns to a temporary row view and has no effect, so clamp == 0 is the reference's behaviour.
        Sum4 cnt = {0.f, 0.f, 0.f, 0.f};
        for (int s = tid; s < S; s += SM_THREADS) {
            const float dx = cur[3 * s] - org[3 * s], dy = cur[3 * s + 1] - org[3 * s + 1], dz = cur[3 * s + 2] - org[3 * s + 2];
            const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
            if (dist > dist_thr) {
                cnt.a += 1.0f;
                if (clamp) {
                    cur[3 * s] = org[3 * s] + dist_thr * dx / dist;
                    cur[3 * s + 1] = org[3 * s + 1] + dist_thr * dy / dist;
                    cur[3 * s + 2] = org[3 * s + 2] + dist_thr * dz / dist;
                }
            }
        }
        cnt = block_sum4(cnt, scratch);
        ++it;
        // check_stop_condition :167-181
        const bool too_far = (double)cnt.a / (double)S > stop_ratio;
        const bool too_long = it > max_iteration;
        if (too_far || grad_none || too_long) {
            flags = (too_far ? 1 : 0) | (grad_none ? 2 : 0) | (too_long ? 4 : 0);
            keep = 0;
            break;
        }
    }
    __syncthreads();
    for (int i = tid; i < 3 * S; i += SM_THREADS) data[3 * (size_t)members[beg + i / 3] + i % 3] = cur[i];
    if (tid == 0) {
        smooth[r] = sm_now;
        alive[r] = keep;
        iters[r] = it;
        stop_flags[r] = flags;
        last_var[r * 3] = v[0];
        last_var[r * 3 + 1] = v[1];
        last_var[r * 3 + 2] = v[2];
    }
}

}  // namespace

int launch_region_smoothness_epoch(float *data, const float *data_orig, const int32_t *offsets, const int32_t *members,
                                   const float *orient, const float *var_ub, const float *var_lb, double *smooth,
                                   int32_t *alive, int32_t *iters, float *last_var, int32_t *stop_flags, int64_t R,
                                   int64_t max_region, int mode, int rising, double step, double enum_step, double dist_thr,
                                   double stop_ratio, int max_iteration, int clamp, cudaStream_t st)
{
    ProfileScope _ps("region_smoothness_epoch", st);
    IQ_CHECK(mode >= 0 && mode <= 2, "region_smoothness_epoch: mode must be 0 (linearity), 1 (planarity) or 2 (scattering)");
    IQ_CHECK(max_region >= 2, "region_smoothness_epoch: a region needs at least two points for an unbiased variance");
    if (R == 0) return 0;
    const size_t smem = sizeof(float) * 6 * (size_t)max_region;
    IQ_CHECK(smem <= 200 * 1024, "region_smoothness_epoch: region larger than 8533 points");
    if (smem > 48 * 1024)
        IQ_CUDA(cudaFuncSetAttribute(region_smoothness_epoch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    region_smoothness_epoch_kernel<<<(unsigned)R, SM_THREADS, smem, st>>>(
        data, data_orig, offsets, members, orient, var_ub, var_lb, smooth, alive, iters, last_var, stop_flags, mode, rising,
        (float)step, enum_step, (float)dist_thr, stop_ratio, max_iteration, clamp);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

}  // namespace iq
