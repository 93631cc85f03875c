// Flash attention for the ViT on tcgen05 (sm_100a): S = Q K^T and O += P V on the 5th-gen tensor cores with
// accumulators in TMEM, operands staged by TMA, online softmax in registers (one query row per thread).
//
//   one CTA = 128 queries of one (image, head); key/value tiles of 128, double-buffered
//   warp 0      : TMA producer (Q once; K tile [128 keys][64] and V^T tile [64 d][128 keys] per step)
//   warp 1      : TMEM allocator + MMA issuer:  S(128x128) = Q K^T (4 x K=16);  O_tile(128x64) = P V (8 x K=16)
//   warps 2..5  : softmax: tcgen05.ld S (thread = query row), running max / sum with exp2, P -> bf16 into a
//                 128-byte-swizzled smem tile (the A operand of the second MMA), then tcgen05.ld O_tile and
//                 accumulate o = o * corr + O_tile in registers
// Replaces the (b,6,1025,1025) attention materialisation of vision_transformer_flexible.py:90-94.
#pragma once
#include <cuda_bf16.h>

#include "scp_common.cuh"
#include "scp_tc5.cuh"

namespace scp {
namespace fa {

constexpr int BQ = 128, BKV = 128, HD = 64, HEADS = 6;
constexpr int NTHREADS = 192;
constexpr int Q_BYTES = BQ * HD * 2;                 // 16 KiB
constexpr int K_BYTES = BKV * HD * 2;                // 16 KiB  [128 keys][64 d]
constexpr int V_BYTES = HD * BKV * 2;                // 16 KiB  2 atoms of [64 d][64 keys]
constexpr int P_BYTES = BQ * BKV * 2;                // 32 KiB  2 atoms of [128 rows][64 keys]
constexpr int SMEM_BYTES = Q_BYTES + 2 * K_BYTES + 2 * V_BYTES + P_BYTES + 256 + 1024;
constexpr int TMEM_COLS = 256;                       // S: cols [0,128), O_tile: cols [128,192)

__device__ __forceinline__ float ex2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(NTHREADS, 2)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
              const __grid_constant__ CUtensorMap tmap_vt, __nv_bfloat16 *__restrict__ o, int T, float scale_log2e)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem, *sK = sQ + Q_BYTES, *sV = sK + 2 * K_BYTES, *sP = sV + 2 * V_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sP + P_BYTES);
    uint64_t *q_full = bars, *kv_full = bars + 1, *kv_empty = bars + 3, *s_full = bars + 5, *p_full = bars + 6,
             *o_full = bars + 7;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.y, q0 = blockIdx.x * BQ;
    const int ntiles = (T + BKV - 1) / BKV;

    if (warp == 0 && lane == 0) {
        tc5::tma_prefetch_desc(&tmap_q);
        tc5::tma_prefetch_desc(&tmap_k);
        tc5::tma_prefetch_desc(&tmap_vt);
        tc5::mbar_init(q_full, 1);
        for (int i = 0; i < 2; i++) { tc5::mbar_init(kv_full + i, 1); tc5::mbar_init(kv_empty + i, 1); }
        tc5::mbar_init(s_full, 1);
        tc5::mbar_init(p_full, 4);
        tc5::mbar_init(o_full, 1);
        tc5::mbar_fence_init();
    }
    if (warp == 1) tc5::tmem_alloc(tmem_slot, TMEM_COLS);
    tc5::tc_fence_before();
    __syncthreads();
    tc5::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (tc5::elect_one()) {   // ===== TMA producer =====
            tc5::mbar_expect_tx(q_full, Q_BYTES);
            tc5::tma_load_2d(sQ, &tmap_q, q_full, 0, bh * T + q0);
            for (int j = 0; j < ntiles; j++) {
                const int st = j & 1, ph = (j >> 1) & 1;
                tc5::mbar_wait(kv_empty + st, ph ^ 1);
                tc5::mbar_expect_tx(kv_full + st, K_BYTES + V_BYTES);
                tc5::tma_load_2d(sK + st * K_BYTES, &tmap_k, kv_full + st, 0, bh * T + j * BKV);
                tc5::tma_load_2d(sV + st * V_BYTES, &tmap_vt, kv_full + st, j * BKV, bh * HD);
                tc5::tma_load_2d(sV + st * V_BYTES + V_BYTES / 2, &tmap_vt, kv_full + st, j * BKV + 64, bh * HD);
            }
        }
    } else if (warp == 1) {
        if (tc5::elect_one()) {   // ===== MMA issuer =====
            constexpr uint32_t idesc_qk = tc5::umma_idesc_bf16(BQ, BKV), idesc_pv = tc5::umma_idesc_bf16(BQ, HD);
            const uint32_t aQ = tc5::smem_u32(sQ), aP = tc5::smem_u32(sP);
            tc5::mbar_wait(q_full, 0);
            for (int j = 0; j < ntiles; j++) {
                const int st = j & 1, ph = (j >> 1) & 1;
                tc5::mbar_wait(kv_full + st, ph);
                tc5::tc_fence_after();
                const uint32_t aK = tc5::smem_u32(sK + st * K_BYTES), aV = tc5::smem_u32(sV + st * V_BYTES);
                // keys of this tile that exist, rounded up to the MMA granularity (N multiple of 16, K = 16)
                const int nk = min(BKV, (T - j * BKV + 15) & ~15);
                const uint32_t idesc_s = nk == BKV ? idesc_qk : tc5::umma_idesc_bf16(BQ, nk);
#pragma unroll
                for (int k = 0; k < HD / 16; k++)
                    tc5::umma_bf16(tmem_base, tc5::umma_desc_sw128(aQ + k * 32), tc5::umma_desc_sw128(aK + k * 32), idesc_s, k != 0);
                tc5::umma_commit(s_full);
                tc5::mbar_wait(p_full, j & 1);                    // P tile written (and O_tile of step j-1 consumed)
                tc5::tc_fence_after();
                for (int k = 0; k < nk / 16; k++) {
                    const uint32_t atom = k >> 2, off = (k & 3) * 32;
                    tc5::umma_bf16(tmem_base + 128, tc5::umma_desc_sw128(aP + atom * (P_BYTES / 2) + off),
                                   tc5::umma_desc_sw128(aV + atom * (V_BYTES / 2) + off), idesc_pv, k != 0);
                }
                tc5::umma_commit(o_full);
                tc5::umma_commit(kv_empty + st);
            }
        }
    } else {
        // ===== softmax warps: thread = query row (TMEM lane) =====
        const int quarter = warp & 3, row = quarter * 32 + lane;
        const uint32_t t_lane = (uint32_t)(quarter * 32) << 16;
        float o_acc[HD];
#pragma unroll
        for (int i = 0; i < HD; i++) o_acc[i] = 0.f;
        float m_run = -1e30f, l_run = 0.f;
        for (int j = 0; j < ntiles; j++) {
            tc5::mbar_wait(s_full, j & 1);
            tc5::tc_fence_after();
            const int nvalid = T - j * BKV;                       // keys of this tile that exist
            const bool full_tile = nvalid >= BKV;                 // only the last tile needs masking
            const int ncols = min(BKV, (nvalid + 15) & ~15);     // columns the MMAs of this tile produce / consume
            float mraw = -3.0e38f;
#pragma unroll 1
            for (int c = 0; c < ncols; c += 32) {
                float v[32];
                tc5::tmem_ld32(tmem_base + t_lane + c, v);
                if (full_tile) {
#pragma unroll
                    for (int i = 0; i < 32; i++) mraw = fmaxf(mraw, v[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; i++) mraw = fmaxf(mraw, (c + i < nvalid) ? v[i] : -3.0e38f);
                }
            }
            const float mx = fmaxf(m_run, mraw * scale_log2e);   // scale > 0: max commutes with the scaling
            const float corr = ex2(m_run - mx);
            m_run = mx;
            float rs = 0.f;
#pragma unroll 1
            for (int c = 0; c < ncols; c += 32) {
                float v[32];
                tc5::tmem_ld32(tmem_base + t_lane + c, v);
                uint8_t *prow = sP + (c >> 6) * (P_BYTES / 2) + row * 128;
#pragma unroll
                for (int q8 = 0; q8 < 4; q8++) {                  // four 16-byte chunks of 8 bf16
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int i = q8 * 8 + 2 * e;
                        float p0 = ex2(fmaf(v[i], scale_log2e, -mx)), p1 = ex2(fmaf(v[i + 1], scale_log2e, -mx));
                        if (!full_tile) {
                            p0 = (c + i < nvalid) ? p0 : 0.f;
                            p1 = (c + i + 1 < nvalid) ? p1 : 0.f;
                        }
                        rs += p0 + p1;
                        __nv_bfloat162 pk = __floats2bfloat162_rn(p0, p1);
                        w[e] = *reinterpret_cast<uint32_t *>(&pk);
                    }
                    const int chunk = ((c & 63) >> 3) + q8;       // 16-byte chunk index inside the 128-byte row
                    *reinterpret_cast<uint4 *>(prow + ((chunk ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            l_run = l_run * corr + rs;
            tc5::fence_proxy_async();          // P (generic-proxy stores) -> visible to the tensor core's async proxy
            tc5::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc5::mbar_arrive(p_full);
            tc5::mbar_wait(o_full, j & 1);
            tc5::tc_fence_after();
#pragma unroll
            for (int c = 0; c < HD; c += 32) {
                float v[32];
                tc5::tmem_ld32(tmem_base + t_lane + 128 + c, v);
#pragma unroll
                for (int i = 0; i < 32; i++) o_acc[c + i] = o_acc[c + i] * corr + v[i];
            }
        }
        const int t = q0 + row;
        if (t < T) {
            const float inv = 1.f / l_run;
            const int b = bh / HEADS, h = bh - b * HEADS;
            __nv_bfloat16 *dst = o + ((long)b * T + t) * (HEADS * HD) + h * HD;
#pragma unroll
            for (int c = 0; c < HD; c += 8) {
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    __nv_bfloat162 pk = __floats2bfloat162_rn(o_acc[c + 2 * e] * inv, o_acc[c + 2 * e + 1] * inv);
                    w[e] = *reinterpret_cast<uint32_t *>(&pk);
                }
                *reinterpret_cast<uint4 *>(dst + c) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        tc5::tc_fence_before();
    }
    tc5::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc5::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace fa
}  // namespace scp
