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
