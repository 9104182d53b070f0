// Measures the issue-to-retire rate of tcgen05.mma on this GPU: one CTA per SM issues REPS MMAs of shape
// 128 x N x (32 bytes of K) back to back on stale shared memory, commits, waits, and reports clk per MMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../interpret_quality_b200/csrc umma_rate.cu -o umma_rate
#include <cstdio>
#include <cstdlib>
#include "tc_ptx.cuh"
using namespace iq::tc;

__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int N, int KIND, int TS>   // KIND 0 = tf32, 1 = bf16
__global__ void __launch_bounds__(128, 1) rate_kernel(int reps, long long *out)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (128 + N) * 128 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = tptr;
    if (warp == 0) {
        // idesc: c F32 (1<<4); tf32 a/b = 2, bf16 a/b = 1 at bits 7 and 10
        const uint32_t fmt = KIND == 0 ? 2u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t ad = make_smem_desc(smem_u32(smem)), bd = make_smem_desc(smem_u32(smem + 128 * 128));
        long long t0 = clock64();
        for (int r = 0; r < reps; r += 8) {
            if (elect_one_sync()) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint64_t koff = (uint64_t)((u & 3) * 2);
                    if (TS) umma_tf32_ts(tb + (u & 1) * 256, tb + 480 + (u & 3) * 8, bd + koff, idesc, 1u);
                    else if (KIND == 0) umma_tf32(tb + (u & 1) * 256, ad + koff, bd + koff, idesc, 1u);
                    else umma_f16(tb + (u & 1) * 256, ad + koff, bd + koff, idesc, 1u);
                }
            }
            __syncwarp();
        }
        if (elect_one_sync()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int N, int KIND, int TS>
void run(const char *name, int reps)
{
    long long *d; cudaMalloc(&d, 148 * 8);
    const int smem = (128 + N) * 128 + 2048;
    cudaFuncSetAttribute(rate_kernel<N, KIND, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate_kernel<N, KIND, TS><<<148, 128, smem>>>(reps, d);
    cudaEventRecord(e0);
    rate_kernel<N, KIND, TS><<<148, 128, smem>>>(reps, d);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const double kelems = KIND == 0 ? 8 : 16;
    const double flop = 2.0 * 128 * N * kelems * reps * 148;
    printf("%-28s %s  clk/MMA %.1f  event %.3f ms  -> %.0f TFLOP/s  (SM clock ~%.0f MHz)\n", name, cudaGetErrorString(err),
           (double)h[0] / reps, ms, flop / (ms * 1e-3) / 1e12, (double)h[0] / (ms * 1e3));
    cudaFree(d);
}

int main()
{
    const int reps = 20000;
    run<128, 0, 0>("tf32 128x128x8  SS", reps);
    run<256, 0, 0>("tf32 128x256x8  SS", reps);
    run<128, 0, 1>("tf32 128x128x8  TS", reps);
    run<256, 0, 1>("tf32 128x256x8  TS", reps);
    run<64, 0, 0>("tf32 128x64x8   SS", reps);
    run<128, 1, 0>("bf16 128x128x16 SS", reps);
    run<256, 1, 0>("bf16 128x256x16 SS", reps);
    return 0;
}
