// Farthest point sampling, nearest-centre region assignment and K=3 squared distances.
//
// Reference behaviour restated (never copied):
//   farthest_point_sample   final_save_fps.py:10-31, models/pointnet2.py:45-68, models/pointconv.py:54-77
//   square_distance         tools/final_util.py:134-147 (models/pointnet2.py:12-25, models/pointconv.py:13-31)
//   cal_region_id           final_shapley_value.py:20-35
//
// Integer outputs must match the reference bit for bit, so every rounding is
// pinned with __fmul_rn/__fadd_rn/__fmaf_rn (never contracted by the compiler):
//   FPS      d = ((dx*dx + dy*dy) + dz*dz), running min, argmax -> lowest index on ties
//   sqdist   dot = fma(z,z', fma(y,y', x*x')); t = -2*dot; t += |src|^2; t += |dst|^2
#include <limits.h>

#include "common.cuh"
#include "kernels.cuh"

namespace iq {

// ------------------------------------------------------------------ FPS
// One CTA per cloud.  Each of the 256 threads keeps PPT points and their running
// distances in registers; a round is: distance update, per-thread argmax,
// two redux.sync per warp, one __syncthreads, 8-way combine.  The chain of
// npoint dependent rounds bounds the kernel, so many
// clouds are kept in flight per SM instead.
template <int PPT>
__global__ void __launch_bounds__(256)
fps_kernel(const float *__restrict__ xyz, int N, int npoint, int64_t *__restrict__ idx64,
           int32_t *__restrict__ idx32, float *__restrict__ new_xyz)
{
    extern __shared__ float sxyz[];                               // 3N floats, point-major copy of the cloud
    __shared__ int wbest[2][8];
    __shared__ int warg[2][8];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *p = xyz + (int64_t)b * N * 3;
    for (int i = tid; i < 3 * N; i += 256) sxyz[i] = p[i];
    __syncthreads();
    float px[PPT], py[PPT], pz[PPT], dist[PPT];
#pragma unroll
    for (int t = 0; t < PPT; ++t) {
        const int i = t * 256 + tid;
        const bool ok = i < N;
        px[t] = ok ? sxyz[3 * i] : 0.0f;
        py[t] = ok ? sxyz[3 * i + 1] : 0.0f;
        pz[t] = ok ? sxyz[3 * i + 2] : 0.0f;
        dist[t] = 1e10f;
    }
    int far = 0;
    for (int s = 0; s < npoint; ++s) {
        const float cx = sxyz[3 * far], cy = sxyz[3 * far + 1], cz = sxyz[3 * far + 2];
        if (tid == 0) {
            if (idx64) idx64[(int64_t)b * npoint + s] = far;
            if (idx32) idx32[(int64_t)b * npoint + s] = far;
            if (new_xyz) {
                float *o = new_xyz + ((int64_t)b * npoint + s) * 3;
                o[0] = cx; o[1] = cy; o[2] = cz;
            }
        }
        int best = -1, arg = INT_MAX;                              // distances are >= 0: int order == float order
#pragma unroll
        for (int t = 0; t < PPT; ++t) {
            const int i = t * 256 + tid;
            const float dx = __fsub_rn(px[t], cx), dy = __fsub_rn(py[t], cy), dz = __fsub_rn(pz[t], cz);
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (d < dist[t]) dist[t] = d;
            const int bits = __float_as_int(dist[t]);
            if (i < N && bits > best) { best = bits; arg = i; }
        }
        const int wmax = __reduce_max_sync(0xffffffffu, best);
        const int wmin = __reduce_min_sync(0xffffffffu, best == wmax ? arg : INT_MAX);
        const int buf = s & 1;
        if (lane == 0) { wbest[buf][warp] = wmax; warg[buf][warp] = wmin; }
        __syncthreads();
        // combine the 8 warp results: lanes 0-7 take one each, two more REDUX (largest distance, lowest index on ties)
        const int vb = lane < 8 ? wbest[buf][lane] : -1;
        const int va = lane < 8 ? warg[buf][lane] : INT_MAX;
        const int gb = __reduce_max_sync(0xffffffffu, vb);
        far = __reduce_min_sync(0xffffffffu, vb == gb ? va : INT_MAX);
    }
}

int launch_fps(const float *xyz, int64_t B, int64_t N, int64_t npoint, int64_t *idx64, int32_t *idx32, float *new_xyz,
               cudaStream_t st)
{
    ProfileScope _ps("fps", st);
    IQ_CHECK(N >= 1 && N <= 4096, "fps: num_points must be in [1,4096]");
    IQ_CHECK(npoint >= 1, "fps: npoint must be positive");
    if (B == 0) return 0;
    const size_t smem = sizeof(float) * 3 * (size_t)N;
    const unsigned grid = (unsigned)B;
    if (N <= 512) fps_kernel<2><<<grid, 256, smem, st>>>(xyz, (int)N, (int)npoint, idx64, idx32, new_xyz);
    else if (N <= 1024) fps_kernel<4><<<grid, 256, smem, st>>>(xyz, (int)N, (int)npoint, idx64, idx32, new_xyz);
    else if (N <= 2048) fps_kernel<8><<<grid, 256, smem, st>>>(xyz, (int)N, (int)npoint, idx64, idx32, new_xyz);
    else {
        IQ_CUDA(cudaFuncSetAttribute(fps_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fps_kernel<16><<<grid, 256, smem, st>>>(xyz, (int)N, (int)npoint, idx64, idx32, new_xyz);
    }
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------ squared distances, K = 3
__device__ __forceinline__ float sqnorm3(float x, float y, float z)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}
__device__ __forceinline__ float sqdist3(float sx, float sy, float sz, float ss, float dx, float dy, float dz, float dd)
{
    float dot = __fmul_rn(sx, dx);
    dot = __fmaf_rn(sy, dy, dot);
    dot = __fmaf_rn(sz, dz, dot);
    float t = __fmul_rn(-2.0f, dot);
    t = __fadd_rn(t, ss);
    return __fadd_rn(t, dd);
}

__global__ void square_distance3_kernel(const float *__restrict__ src, const float *__restrict__ dst, int N, int M,
                                        float *__restrict__ out)
{
    const int b = blockIdx.z;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= M) return;
    const float *s = src + ((int64_t)b * N + i) * 3;
    const float *d = dst + ((int64_t)b * M + j) * 3;
    const float ss = sqnorm3(s[0], s[1], s[2]), dd = sqnorm3(d[0], d[1], d[2]);
    out[((int64_t)b * N + i) * M + j] = sqdist3(s[0], s[1], s[2], ss, d[0], d[1], d[2], dd);
}

int launch_square_distance3(const float *src, const float *dst, int64_t B, int64_t N, int64_t M, float *out,
                            cudaStream_t st)
{
    ProfileScope _ps("square_distance3", st);
    if (B * N * M == 0) return 0;
    IQ_CHECK(N <= 65535 && B <= 65535, "square_distance3: N and B must be <= 65535");
    dim3 grid((unsigned)ceil_div(M, 128), (unsigned)N, (unsigned)B);
    square_distance3_kernel<<<grid, 128, 0, st>>>(src, dst, (int)N, (int)M, out);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------ nearest-centre region ids
__global__ void region_id_kernel(const float *__restrict__ xyz, const int64_t *__restrict__ fps_index, int N, int R,
                                 int64_t *__restrict__ region_id)
{
    extern __shared__ float cen[];                                // R * 4: x,y,z,|c|^2
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const int64_t src = fps_index[r];
        const float x = xyz[3 * src], y = xyz[3 * src + 1], z = xyz[3 * src + 2];
        cen[4 * r] = x; cen[4 * r + 1] = y; cen[4 * r + 2] = z; cen[4 * r + 3] = sqnorm3(x, y, z);
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    const float ss = sqnorm3(x, y, z);
    float best = sqdist3(x, y, z, ss, cen[0], cen[1], cen[2], cen[3]);
    int arg = 0;
    for (int r = 1; r < R; ++r) {
        const float d = sqdist3(x, y, z, ss, cen[4 * r], cen[4 * r + 1], cen[4 * r + 2], cen[4 * r + 3]);
        if (d < best) { best = d; arg = r; }                      // argmin: first minimum wins
    }
    region_id[i] = arg;
}

int launch_region_id(const float *xyz, const int64_t *fps_index, int64_t N, int64_t R, int64_t *region_id,
                     cudaStream_t st)
{
    ProfileScope _ps("region_id", st);
    IQ_CHECK(R >= 1 && R <= 2048, "region_id: num_regions out of range");
    if (N == 0) return 0;
    region_id_kernel<<<(unsigned)ceil_div(N, 128), 128, sizeof(float) * 4 * (size_t)R, st>>>(xyz, fps_index, (int)N,
                                                                                           (int)R, region_id);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------ centre of the cloud
// mean over the points, accumulated in float64 and rounded once (the reference's
// torch.mean(data, dim=1), tools/final_common.py:80, to within its own rounding).
__global__ void __launch_bounds__(256) center_kernel(const float *__restrict__ xyz, int N, float *__restrict__ center)
{
    __shared__ double part[3][256];
    double a[3] = {0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < N; i += 256)
        for (int c = 0; c < 3; ++c) a[c] += (double)xyz[3 * i + c];
    for (int c = 0; c < 3; ++c) part[c][threadIdx.x] = a[c];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s)
            for (int c = 0; c < 3; ++c) part[c][threadIdx.x] += part[c][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x < 3) center[threadIdx.x] = (float)(part[threadIdx.x][0] / (double)N);
}

int launch_center(const float *xyz, int64_t N, float *center, cudaStream_t st)
{
    ProfileScope _ps("center", st);
    IQ_CHECK(N >= 1, "center: empty cloud");
    center_kernel<<<1, 256, 0, st>>>(xyz, (int)N, center);
    IQ_COUNT_LAUNCH();
    IQ_LAUNCH_CHECK();
    return 0;
}

}  // namespace iq
