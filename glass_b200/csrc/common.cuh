// Shared helpers for the glass_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/glass_b200.h"

namespace glass {

void set_error(const char* fmt, ...);

#define GLASS_CHECK_ARG(cond, ...)                 \
    do {                                           \
        if (!(cond)) {                             \
            ::glass::set_error(__VA_ARGS__);       \
            return GLASS_ERR_BAD_ARG;              \
        }                                          \
    } while (0)

#define GLASS_CUDA(call)                                                                      \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            ::glass::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return GLASS_ERR_CUDA;                                                            \
        }                                                                                     \
    } while (0)

#define GLASS_LAUNCH_CHECK() GLASS_CUDA(cudaPeekAtLastError())

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();  // cached per process

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// expm1 for v <= 0 without the slow libm path: degree-5 Taylor below 1/16 (error < 1e-9 relative),
// MUFU-based exp elsewhere (2 ulp of exp(v), i.e. < 3e-6 relative for |v| >= 1/16).
// The exponential is one MUFU.EX2 (ex2.approx.ftz, relative error 2^-22) issued through inline PTX: __expf adds a
// denormal-range fix-up whose control flow made nvcc emit a divergent branch region PER ELEMENT in the GEMM
// epilogues (ncu source view: 16 BSSY/BSYNC pairs per 8 columns, ~800 instructions per epilogue iteration).
// For v <= 0 the result of ex2 is in (0, 1]: nothing to fix up; positive v only produces a value the caller discards.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float expm1_neg(float v) {
    const float t = v * (1.f + v * (0.5f + v * (0.16666667f + v * (0.041666668f + v * 0.0083333338f))));
    const float e = ex2_approx(v * 1.4426950408889634f) - 1.f;
    return v > -0.0625f ? t : e;
}

// ELU without a branch: nvcc turns `v > 0 ? v : expm1(v)` into a divergent branch around the exponential (one
// BSSY/BSYNC region per element); max(v, 0) + expm1(min(v, 0)) is the same value (expm1(0) == 0 exactly) in
// straight-line code.  The asm is volatile so that it cannot be sunk back into a conditional region.
__device__ __forceinline__ float elu_fwd(float v) {
    const float m = fminf(v, 0.f);
    float e;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(m * 1.4426950408889634f));
    const float t = m * (1.f + m * (0.5f + m * (0.16666667f + m * (0.041666668f + m * 0.0083333338f))));
    return fmaxf(v, 0.f) + (m > -0.0625f ? t : e - 1.f);
}

__device__ __forceinline__ float act_fwd(float v, int act) {
    if (act == GLASS_ACT_RELU) return fmaxf(v, 0.f);
    if (act == GLASS_ACT_ELU) return elu_fwd(v);
    return v;
}
// derivative expressed through the POST-activation value (ELU alpha = 1: d/dx = y + 1 for x <= 0)
__device__ __forceinline__ float act_grad_from_out(float y, int act) {
    if (act == GLASS_ACT_RELU) return y > 0.f ? 1.f : 0.f;
    if (act == GLASS_ACT_ELU) return y > 0.f ? 1.f : y + 1.f;
    return 1.f;
}

// the same derivative from the PRE-activation value: exp(x) has no cancellation near 0, so no expm1 blend
__device__ __forceinline__ float act_grad_from_pre(float v, int act) {
    if (act == GLASS_ACT_RELU) return v > 0.f ? 1.f : 0.f;
    if (act == GLASS_ACT_ELU) return v > 0.f ? 1.f : ex2_approx(v * 1.4426950408889634f);
    return 1.f;
}

// rows of the GraphNorm statistics table stats[6, c] (graphnorm.cu writes it; pool.cu / gemm_tc.cu consume it)
enum { ST_SCALE = 0, ST_AM = 1, ST_MU = 2, ST_RSTD = 3, ST_BIAS = 4, ST_RNG = 5 };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace glass
