// Warp-level tensor-core helpers for the correspondence kernels: m16n8k8 TF32 MMA with the
// 3-term split (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo) that recovers ~fp32 accuracy, and cp.async.
// (The ViT kernels use tcgen05/TMEM instead -- see scp_vit.cu; these small K=64 similarity
// products live behind HBM-bound softmax/reduction epilogues and are issued per warp.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace scp {

#ifdef SCP_HOST_EMU
// Host emulation (tools/emu, tests only): the same six helpers in plain C++.  mma_tf32 is a warp collective there too:
// every lane publishes its fragments, then computes its four outputs of D += A B with the layout documented below.
__device__ __forceinline__ uint32_t f2tf32(float x)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    return (u + 0x1000u) & ~0x1fffu;      // cvt.rna: nearest, ties away from zero, 10 mantissa bits kept
}
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo)
{
    hi = f2tf32(x);
    lo = f2tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    uint32_t (*frag)[8] = scp_emu_warp_fragments();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int i = 0; i < 4; i++) frag[lane][i] = a[i];
    frag[lane][4] = b[0];
    frag[lane][5] = b[1];
    __syncwarp();
    for (int i = 0; i < 4; i++) {
        const int row = g + (i >= 2 ? 8 : 0), col = 2 * t + (i & 1);
        float acc = d[i];
        for (int k = 0; k < 8; k++) {
            const float av = __uint_as_float(frag[(row & 7) * 4 + (k & 3)][(row >= 8 ? 1 : 0) + (k >= 4 ? 2 : 0)]);
            const float bv = __uint_as_float(frag[col * 4 + (k & 3)][4 + (k >= 4 ? 1 : 0)]);
            acc += av * bv;
        }
        d[i] = acc;
    }
    __syncwarp();
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) { memcpy(smem, gmem, 16); }
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) { memcpy(smem, gmem, 8); }
__device__ __forceinline__ void cp_async_commit() {}
template <int N>
__device__ __forceinline__ void cp_async_wait() {}
#else
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

#endif  // SCP_HOST_EMU

}  // namespace scp
