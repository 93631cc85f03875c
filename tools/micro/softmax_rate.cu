// Microbenchmark: the fa3 softmax inner loop (64 scores per thread: scale, 2^x, row sum, hi/lo bf16 split) in isolation,
// WPS warps per scheduler, registers only.  Variants: 0 = as in the kernel; 1 = MUFU replaced by an FMUL (floor without the
// MUFU pipe); 2 = half of the exponentials by a degree-5 polynomial on the FMA pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I self_corr_pose_b200/csrc -o tools/micro/softmax_rate tools/micro/softmax_rate.cu
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "scp_common.cuh"
using namespace scp;

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t *>(&v); }
__device__ __forceinline__ uint32_t split_bf16x2(float a, float b, uint32_t &lo_pair)
{
    const uint32_t h = pack_bf16x2(a, b);
    uint16_t h0, h1;
    asm("mov.b32 {%0, %1}, %2;" : "=h"(h0), "=h"(h1) : "r"(h));
    const uint16_t m1 = 0xBF80;
    float la, lb;
    asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(la) : "h"(h0), "h"(m1), "f"(a));
    asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(lb) : "h"(h1), "h"(m1), "f"(b));
    lo_pair = pack_bf16x2(la, lb);
    return h;
}
// 2^x for x <= 0 (down to ~-126) on the FMA / ALU pipes: x = n + f, n = round(x), f in [-0.5, 0.5], degree-5 minimax of 2^f,
// exponent added into the bits.  ~2e-7 relative.
__device__ __forceinline__ float ex2_poly(float x)
{
    x = fmaxf(x, -125.f);
    const float t = x + 12582912.f;            // 1.5 * 2^23: round to nearest integer in the low mantissa bits
    const float n = t - 12582912.f;
    const float f = x - n;
    float p = 1.3333558146e-3f;
    p = fmaf(p, f, 9.6181291076e-3f);
    p = fmaf(p, f, 5.5504108665e-2f);
    p = fmaf(p, f, 2.4022650696e-1f);
    p = fmaf(p, f, 6.9314718056e-1f);
    p = fmaf(p, f, 1.f);
    return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}

template <int VAR>
__global__ void __launch_bounds__(512, 1) k(int iters, float scale, float seed, long long *out, uint32_t *sink)
{
    float v[64];
#pragma unroll
    for (int i = 0; i < 64; i++) v[i] = -seed * (float)((threadIdx.x * 7 + i * 13) % 97);
    uint32_t ph[32], pl[32];
    float l_run = 0.f;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const uint64_t sc2 = f2_pack(scale, scale), nm2 = f2_pack(-l_run * 1e-30f, -l_run * 1e-30f);
        uint64_t ra = f2_pack(0.f, 0.f), rb = ra;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            float x0, x1, x2, x3;
            f2_unpack(f2_fma(f2_pack(v[2 * i], v[2 * i + 1]), sc2, nm2), x0, x1);
            f2_unpack(f2_fma(f2_pack(v[2 * i + 2], v[2 * i + 3]), sc2, nm2), x2, x3);
            float p0, p1, p2, p3;
            if (VAR == 0) { p0 = ex2(x0); p1 = ex2(x1); p2 = ex2(x2); p3 = ex2(x3); }
            if (VAR == 1) { p0 = x0 * 0.99f; p1 = x1 * 0.98f; p2 = x2 * 0.97f; p3 = x3 * 0.96f; }
            if (VAR == 2) { p0 = ex2(x0); p1 = ex2_poly(x1); p2 = ex2(x2); p3 = ex2_poly(x3); }
            ra = f2_add(ra, f2_pack(p0, p1));
            rb = f2_add(rb, f2_pack(p2, p3));
            ph[i] = split_bf16x2(p0, p1, pl[i]);
            ph[i + 1] = split_bf16x2(p2, p3, pl[i + 1]);
        }
        float r0, r1, r2, r3;
        f2_unpack(ra, r0, r1);
        f2_unpack(rb, r2, r3);
        l_run += (r0 + r1) + (r2 + r3);
#pragma unroll
        for (int i = 0; i < 32; i++) acc ^= ph[i] + pl[i];      // consume (the kernel stores them to TMEM): 64 extra ALU ops
    }
    const long long t1 = clock64();
    if (acc == 0x12345u && l_run == 1.f) sink[0] = acc;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int VAR>
void run(const char *name, int threads, long long *d, uint32_t *sink)
{
    const int iters = 500;
    for (int rep = 0; rep < 2; rep++) { k<VAR><<<148, threads>>>(iters, 0.18f, 0.37f, d, sink); cudaDeviceSynchronize(); }
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < 148; i++) s += h[i];
    const double cyc = s / 148 / iters;
    printf("%-44s %2d warps/SM: %7.1f cycles per 64-score tile per warp, %6.1f scores / clk / SM\n", name, threads / 32, cyc,
           threads * 64.0 / cyc);
}

int main()
{
    long long *d;
    uint32_t *sink;
    cudaMalloc(&d, 148 * sizeof(long long));
    cudaMalloc(&sink, 4);
    for (int threads : { 128, 256, 512 }) {
        run<0>("kernel form (MUFU.EX2 per score)", threads, d, sink);
        run<1>("MUFU replaced by FMUL", threads, d, sink);
        run<2>("half MUFU, half degree-5 polynomial", threads, d, sink);
    }
    return 0;
}
