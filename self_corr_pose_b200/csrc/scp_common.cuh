// Shared helpers for the sm_100a kernels of the self-corr-pose hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace scp {

void set_last_error(const char *fmt, ...);

// Launch-status check used at the end of every C-ABI entry point (the reference only printf's
// launch failures, soft_rasterize_cuda_kernel.cu:710-712; we return them).
inline int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

// 2^x on the MUFU pipe (ex2.approx.ftz: relative error 2^-22, results below 2^-126 flush to 0) -- one instruction
// instead of exp2f's range-handling sequence; used where x <= 0 up to rounding (softmax numerators)
__device__ __forceinline__ float ex2_approx(float x)
{
#ifdef SCP_HOST_EMU      // host emulation of the plain-CUDA kernels (tools/emu, tests only)
    return exp2f(x);
#else
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}

#ifndef SCP_HOST_EMU
// Packed fp32 pairs (sm_100 FFMA2 / FADD2: two independent fp32 operations per issue slot, each rounded like the scalar
// instruction) for issue-bound per-element loops.  A pair lives in a 64-bit register: {lo = first, hi = second}.
__device__ __forceinline__ uint64_t f2_pack(float a, float b)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float &a, float &b)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
#endif

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace scp
