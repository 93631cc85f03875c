// Microbenchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16, cta_group::1) as a function of N and of where the
// A operand lives (shared memory descriptor vs tensor memory), one CTA per SM, back-to-back issue from one elected thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I self_corr_pose_b200/csrc -o /tmp/umma_rate tools/micro/umma_rate.cu -lcuda
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "scp_tc5.cuh"
using namespace scp;

template <bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int ctas_per_sm_dummy, long long *out)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { tc5::mbar_init(&bar, 1); tc5::mbar_fence_init(); }
    if (warp == 0) tc5::tmem_alloc(&slot, 512);
    tc5::tc_fence_before();
    __syncthreads();
    tc5::tc_fence_after();
    const uint32_t tm = slot;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    long long t0 = 0, t1 = 0;
    if (warp == 1) {
        if (tc5::elect_one()) {
            const uint32_t idesc = tc5::umma_idesc_bf16(128, N);
            const uint32_t a = tc5::smem_u32(smem), b = a + 16384;
            t0 = clock64();
            for (int it = 0; it < iters; it++) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (TS) tc5::umma_bf16_ts(tm, tm + 256 + k * 8, tc5::umma_desc_sw128(b + k * 32), idesc, 1);
                    else tc5::umma_bf16(tm, tc5::umma_desc_sw128(a + k * 32), tc5::umma_desc_sw128(b + k * 32), idesc, 1);
                }
            }
            tc5::umma_commit(&bar);
            tc5::mbar_wait(&bar, 0);
            t1 = clock64();
            out[blockIdx.x] = t1 - t0;
        }
    }
    tc5::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc5::tmem_dealloc(tm, 512);
}

int main()
{
    long long *d;
    cudaMalloc(&d, 148 * sizeof(long long));
    const int smem = 16384 + 32768 + 2048;
    cudaFuncSetAttribute(rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;
    for (int ts = 0; ts < 2; ts++)
        for (int N : { 16, 32, 64, 96, 128, 192, 256 }) {
            for (int rep = 0; rep < 2; rep++) {
                if (ts) rate_kernel<true><<<148, 128, smem>>>(N, iters, 0, d);
                else rate_kernel<false><<<148, 128, smem>>>(N, iters, 0, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            long long h[148];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            double s = 0;
            for (int i = 0; i < 148; i++) s += h[i];
            const double cyc = s / 148 / (iters * 4.0);
            printf("%s M=128 N=%3d K=16: %6.1f cycles per MMA  (arithmetic floor %5.1f = N/2; %.0f%% of it)\n", ts ? "A=TMEM" : "A=SMEM", N, cyc,
                   N / 2.0, 100.0 * (N / 2.0) / cyc);
        }
    return 0;
}
