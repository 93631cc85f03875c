// Microbenchmark: does tcgen05.ld / tcgen05.st traffic slow the tensor pipe down?  One CTA per SM: an elected thread issues
// back-to-back tcgen05.mma (M = 128, N, K = 16, A from TMEM) while NLD warps loop over tcgen05.ld.32x32b.x64 (+ optional st).
//   nvcc -gencode arch=compute_100a,code=sm_100a -I self_corr_pose_b200/csrc -o tools/micro/umma_ldtm tools/micro/umma_ldtm.cu -lcuda
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "scp_tc5.cuh"
using namespace scp;

__global__ void __launch_bounds__(192, 1) k(int N, int iters, int nld, int do_st, int mma_on, long long *out, float *sink)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { tc5::mbar_init(&bar, 1); tc5::mbar_fence_init(); stop = 0; }
    if (warp == 0) tc5::tmem_alloc(&slot, 512);
    tc5::tc_fence_before();
    __syncthreads();
    tc5::tc_fence_after();
    const uint32_t tm = slot;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        if (tc5::elect_one()) {
            const uint32_t idesc = tc5::umma_idesc_bf16(128, N);
            const uint32_t b = tc5::smem_u32(smem);
            const long long t0 = clock64();
            if (mma_on) {
                for (int it = 0; it < iters; it++) {
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) tc5::umma_bf16_ts(tm, tm + 384 + kk * 8, tc5::umma_desc_sw128(b + kk * 32), idesc, 1);
                }
                tc5::umma_commit(&bar);
                tc5::mbar_wait(&bar, 0);
            } else {
                while (clock64() - t0 < 400000) {}
            }
            const long long t1 = clock64();
            out[blockIdx.x * 2] = t1 - t0;
            *reinterpret_cast<volatile int *>(&stop) = 1;
        }
    } else if (warp >= 2 && warp - 2 < nld) {
        const uint32_t t_lane = (uint32_t)((warp & 3) * 32) << 16;
        float acc = 0.f;
        long long n = 0;
        const long long t0 = clock64();
        while (*reinterpret_cast<volatile int *>(&stop) == 0) {
            float v[64];
            tc5::tmem_ld64(tm + t_lane + 256, v);       // columns 256..319: not touched by the MMAs (D = 0..N-1, A = 384..)
#pragma unroll
            for (int i = 0; i < 64; i++) acc += v[i];
            if (do_st) {
                uint32_t w[32];
#pragma unroll
                for (int i = 0; i < 32; i++) w[i] = __float_as_uint(v[i]);
                tc5::tmem_st32(tm + t_lane + 320, w);
                tc5::tmem_st_wait();
            }
            n++;
        }
        const long long t1 = clock64();
        if (lane == 0 && warp == 2) { out[blockIdx.x * 2 + 1] = (t1 - t0) / (n > 0 ? n : 1); }
        if (acc == 123.456f) sink[0] = acc;
    }
    tc5::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc5::tmem_dealloc(tm, 512);
}

int main()
{
    long long *d;
    float *sink;
    cudaMalloc(&d, 148 * 2 * sizeof(long long));
    cudaMalloc(&sink, 4);
    const int smem = 32768 + 2048;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;
    for (int mma_on = 1; mma_on >= 0; mma_on--)
        for (int N : { 64, 128 })
            for (int do_st = 0; do_st < 2; do_st++)
                for (int nld : { 0, 1, 2, 4 }) {
                    if (!mma_on && (nld == 0 || N != 64)) continue;
                    cudaMemset(d, 0, 148 * 2 * sizeof(long long));
                    k<<<148, 192, smem>>>(N, iters, nld, do_st, mma_on, d, sink);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    long long h[296];
                    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                    double s = 0, l = 0;
                    for (int i = 0; i < 148; i++) { s += h[2 * i]; l += h[2 * i + 1]; }
                    printf("mma %s N=%3d  ld warps %d%s: %6.1f cycles per MMA (floor %3d)   %7.1f cycles per 8 KB tcgen05.ld.x64%s per warp\n",
                           mma_on ? "on " : "off", N, nld, do_st ? " (+st x32)" : "          ", mma_on ? s / 148 / (iters * 4.0) : 0.0, N / 2,
                           l / 148, do_st ? "+st" : "");
                }
    return 0;
}
