// Microbenchmark: per-SM throughput (thread-operations per clock) of the instructions the softmax / split epilogues lean on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/micro/inst_rate tools/micro/inst_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(int iters, float seed, long long *out, float *sink)
{
    float x[8];
    uint32_t u[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = seed + 0.001f * (threadIdx.x + i); u[i] = __float_as_uint(x[i]); }
    unsigned long long w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(w[i]) : "f"(x[i]), "f"(x[i] + 1.f));
    const unsigned short m1 = 0xBF80;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
            if (OP == 2) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(x[i]), "f"(__uint_as_float(u[i]))); }
            if (OP == 3) { unsigned short h = (unsigned short)u[i]; asm volatile("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(x[i]) : "h"(h), "h"(m1)); }
            if (OP == 4) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(w[i]));
            if (OP == 5) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 6) asm volatile("add.rn.f32 %0, %0, %0;" : "+f"(x[i]));
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) { float a, b; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(w[i])); s += x[i] + __uint_as_float(u[i]) + a + b; }
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, long long *d, float *sink, int per_inst)
{
    const int iters = 4000;
    for (int rep = 0; rep < 2; rep++) { k<OP><<<148, 1024>>>(iters, 0.5f, d, sink); cudaDeviceSynchronize(); }
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < 148; i++) s += h[i];
    const double cyc = s / 148;
    printf("%-28s %6.1f thread-instructions / clk / SM   (%d result%s each)\n", name, 1024.0 * iters * 8 / cyc, per_inst, per_inst > 1 ? "s" : "");
}

int main()
{
    long long *d;
    float *sink;
    cudaMalloc(&d, 148 * sizeof(long long));
    cudaMalloc(&sink, 4);
    run<0>("MUFU.EX2 (ex2.approx.ftz)", d, sink, 1);
    run<5>("MUFU.RCP", d, sink, 1);
    run<1>("FFMA", d, sink, 1);
    run<6>("FADD", d, sink, 1);
    run<4>("FFMA2 (fma.rn.f32x2)", d, sink, 2);
    run<2>("F2FP.BF16.F32.PACK_AB", d, sink, 2);
    run<3>("FHFMA.BF16 (fma.f32.bf16)", d, sink, 1);
    return 0;
}
