// Shared helpers for the iq_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

namespace iq {

// ---- error plumbing: every extern "C" entry returns 0 or a negative code and
// leaves a message retrievable through iq_last_error().
void set_error(const std::string &msg);
const char *last_error();

#define IQ_CHECK(cond, msg)                                                                        \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            ::iq::set_error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + (msg)); \
            return -1;                                                                             \
        }                                                                                          \
    } while (0)

#define IQ_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ::iq::set_error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " +        \
                            cudaGetErrorString(_e));                                               \
            return -2;                                                                             \
        }                                                                                          \
    } while (0)

#define IQ_LAUNCH_CHECK() IQ_CUDA(cudaGetLastError())

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE setting: cached per (kernel, current device), so a
// process that drives several GPUs opts every one of them in.  Thread-safe.  Returns 0 or -2 (error recorded).
int ensure_dynamic_smem(const void *func, int bytes);
// environment switches of the diagnostics scripts, read once per process
int env_int(const char *name, int fallback);
void env_forget();                 // the next env_int() reads the environment again (iq_debug_reload_env)
// Which DGCNN / GCNN products run on kind::f16 MMAs over two-term fp16 operands instead of 3xTF32 (edgeconv_model.cu):
// bit 0 conv5 + pooling, bit 1 the tcgen05 EdgeConv products, bit 2 the Gram kNN nomination (bits 1-2 need bit 0, whose
// fp16 activation buffers they read).  Defaults below, overridden by IQ_F16_CONV5 / IQ_F16_STORE / IQ_F16_GRAM = 0 | 1.
enum : int { F16_CONV5 = 1, F16_STORE = 2, F16_GRAM = 4 };
constexpr int F16_DEFAULT_PATHS = F16_CONV5 | F16_STORE | F16_GRAM;
int f16_paths();

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// number of kernels this library has launched (bench.py reports it as gpu_launches)
extern std::atomic<unsigned long long> g_launch_count;
#define IQ_COUNT_LAUNCH() (::iq::g_launch_count.fetch_add(1, std::memory_order_relaxed))

// ---- optional per-kernel timing (bench.py's roofline leg): CUDA events around every launch
struct ProfileScope {
    ProfileScope(const char *name, cudaStream_t st);
    ~ProfileScope();
    int slot;
    cudaStream_t st;
};
void profile_enable(bool on);
// synchronises, then fills up to `cap` entries; returns the number of distinct kernel names
int profile_report(const char **names, double *ms, long long *counts, int cap);

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2 };

__device__ __forceinline__ float apply_act(float v, int act)
{
    if (act == ACT_RELU) return fmaxf(v, 0.0f);
    if (act == ACT_LRELU) return v > 0.0f ? v : 0.2f * v;
    return v;
}

// Two-term fp16 split of x * scale for the kind::f16 tensor-core products (gemm_tc.cu): hi = fp16(xs), lo = fp16(xs - hi).
// 22 significand bits (the tf32 pair has 21) while lo is a normal fp16 number, |xs| >= 2^-3; below that the absolute error
// is <= 2^-25 (the fp16 subnormal spacing), i.e. 2^-25 / scale on x.  Conversions saturate at +-65504 instead of
// producing infinities: hi + lo follows x up to |xs| = 131008 (with 2^-12 relative accuracy beyond 65504).
__host__ __device__ __forceinline__ void split_f16(float x, float scale, __half &hi, __half &lo)
{
    const float xs = x * scale;
    hi = __float2half_rn(fminf(fmaxf(xs, -65504.0f), 65504.0f));
    lo = __float2half_rn(fminf(fmaxf(xs - __half2float(hi), -65504.0f), 65504.0f));
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b)
{
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// Packed fp32 arithmetic (sm_100: FADD2 / FFMA2, two IEEE fp32 operations per instruction and lane).  A scalar FFMA / FADD
// issues every second cycle per scheduler on this part, so fp32-bound loops run up to twice as fast packed; each half
// rounds exactly like the scalar instruction.
__device__ __forceinline__ unsigned long long pack_f2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f2(unsigned long long v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma_f2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long sub_f2(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// monotone float -> uint key (larger float <=> larger uint); -0.0 sorts just below +0.0
__device__ __forceinline__ uint32_t ordered_u32(float f)
{
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_u32(uint32_t u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

}  // namespace iq
