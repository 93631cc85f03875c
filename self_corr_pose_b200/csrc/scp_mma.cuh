// Warp-level tensor-core helpers for the correspondence kernels: m16n8k8 TF32 MMA with the
// 3-term split (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo) that recovers ~fp32 accuracy, and cp.async.
// (The ViT kernels use tcgen05/TMEM instead -- see scp_vit.cu; these small K=64 similarity
// products live behind HBM-bound softmax/reduction epilogues and are issued per warp.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace scp {

__device__ __forceinline__ uint32_t f2tf32(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo)
{
    hi = f2tf32(x);
    lo = f2tf32(x - __uint_as_float(hi));
}

// D(16x8) += A(16x8, row) * B(8x8, col); fragments per the PTX ISA m16n8k8 .tf32 layout:
//   a0:(g,t) a1:(g+8,t) a2:(g,t+4) a3:(g+8,t+4);  b0:(k=t,n=g) b1:(k=t+4,n=g);
//   d0:(g,2t) d1:(g,2t+1) d2:(g+8,2t) d3:(g+8,2t+1)      with g = lane/4, t = lane%4
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                 "{%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

}  // namespace scp
