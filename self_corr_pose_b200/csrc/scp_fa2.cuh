// Flash attention v2 for the ViT on tcgen05 (sm_100a).  Differences from scp_fa.cuh (kept as variant "1"):
//   * key/value tiles of 64 with TWO S accumulators in TMEM: the tensor core computes S(j+1) = Q K(j+1)^T while the
//     softmax warps work on S(j);
//   * P never touches shared memory: the softmax warps write bf16 P back into TENSOR MEMORY (tcgen05.st) and the
//     second product O += P V reads its A operand from there (tcgen05.mma with A in TMEM);
//   * O accumulates in TMEM across the whole key loop (no per-tile read-back); the running maximum is LAZY: the
//     accumulator is rescaled (tcgen05.ld / st by the row's own thread) only when a row maximum grows by more
//     than 2^8, so after the first tiles the softmax warps do one TMEM read of S and one exp2 per element;
//   * separate K and V^T rings (4 stages each) so the next K tile lands long before its S product is issued;
//   * query tiles with fewer than 128 valid rows (T = 1025: the last tile holds ONE token) only run the softmax
//     warps that own a valid row.
//
// Measured and rejected variants (B = 64, per layer; the kernel stays at ~1000 cycles per 128 x 64 score tile per SM):
//   * two softmax threads per row (8 softmax warps, partial maxima exchanged through smem): 0.211 -> 0.232 ms;
//   * exp2 emulation on the FMA pipe for every 4th element: 0.211 -> 0.224 ms (below).
//   * THREE S accumulators with P(j) written over the head of S(j) (S(j+3) issued right after PV(j)): 0.211 -> 0.202 ms
//     but WRONG results once a buffer is reused (T = 1025 fails, T <= 128 passes) -- not kept; to be debugged with GPU
//     time (suspect: the product that overwrites the aliased columns starts before PV(j) has read P(j)).
// What the source page shows (profiles/r1_fa2_source_page.csv.gz): 68 % of the softmax warps' waits for the next S tile
// spin (11.7 % of all stall samples): S(j+2) is only issued after PV(j), so the tensor pipe's latency is exposed once per
// tile; a deeper S ring is the lever, more softmax warps or fewer MUFU ops are not.
//
//   TMEM columns (256 per CTA, two CTAs per SM):  S0 [0,64)  S1 [64,128)  P0 [128,160)  P1 [160,192)  O [192,256)
//   warp 0 : TMA producer     warp 1 : TMEM allocator + MMA issuer     warps 2..5 : softmax (thread = query row)
// Replaces the (b,6,1025,1025) attention materialisation of vision_transformer_flexible.py:90-94.
#pragma once
#include <cuda_bf16.h>

#include "scp_common.cuh"
#include "scp_tc5.cuh"

namespace scp {
namespace fa2 {

constexpr int BQ = 128, BKV = 64, HD = 64, HEADS = 6;
constexpr int NTHREADS = 192;
constexpr int NK = 4, NV = 4;                        // K / V^T ring depths
constexpr int Q_BYTES = BQ * HD * 2;                 // 16 KiB  [128 queries][64 d]
constexpr int KT_BYTES = BKV * HD * 2;               //  8 KiB  [64 keys][64 d]
constexpr int VT_BYTES = HD * BKV * 2;               //  8 KiB  [64 d][64 keys]
constexpr int SMEM_BYTES = Q_BYTES + NK * KT_BYTES + NV * VT_BYTES + 256 + 1024;
constexpr int TMEM_COLS = 256;
constexpr uint32_t COL_S = 0, COL_P = 128, COL_O = 192;
constexpr float LAZY = 8.f;                          // log2 units: P stays below 2^8

__device__ __forceinline__ float ex2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 2^x for x <= ~8 on the FMA / ALU pipes (no MUFU): Cody-Waite split x = n + f, f in [-0.5, 0.5], degree-3 minimax
// polynomial for 2^f (relative error 7.7e-5, far below the bf16 rounding of P) and n added into the exponent field.
// Measured and rejected for this kernel (SCP_FA2_POLY_EVERY = 4: 0.211 -> 0.224 ms per layer at B = 64): with two
// softmax warps per scheduler the kernel is issue / latency bound, not bound by the 16 ex2 / clk / SM of the MUFU pipe,
// so trading one MUFU op for nine ALU ops loses.  Kept behind the macro for wider softmax configurations.
__device__ __forceinline__ float ex2_poly(float x)
{
    x = fmaxf(x, -126.f);
    const float r = x + 12582912.f;                 // 1.5 * 2^23: n = round(x) lands in the low mantissa bits
    const float f = x - (r - 12582912.f);
    const float p = fmaf(fmaf(fmaf(0.05508868396282196f, f, 0.24260404706001282f), f, 0.6932762265205383f), f,
                         0.9999289512634277f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}

// Next-round candidate, NOT validated on a GPU yet (default off): a separate "S consumed" barrier, arrived by the softmax
// warps right after their tcgen05.ld of S(j), lets the MMA warp issue Q K(j+2)^T during the softmax of tile j instead of
// after it (P and S do not alias, so only the read of S(j) has to be over) -- addresses the s_full starvation above.
#ifndef SCP_FA2_EARLY_QK
#define SCP_FA2_EARLY_QK 0
#endif

#ifndef SCP_FA2_POLY_EVERY
#define SCP_FA2_POLY_EVERY 0      // N > 0: every N-th element of a row takes the polynomial path
#endif

__global__ void __launch_bounds__(NTHREADS, 2)
fa2_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
               const __grid_constant__ CUtensorMap tmap_vt, __nv_bfloat16 *__restrict__ o, int T, float scale_log2e,
               int token_major)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem, *sK = sQ + Q_BYTES, *sV = sK + NK * KT_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sV + NV * VT_BYTES);
    uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = k_full + NK, *v_full = k_empty + NK, *v_empty = v_full + NV,
             *s_full = v_empty + NV, *p_full = s_full + 2, *pv_done = p_full + 2;
#if SCP_FA2_EARLY_QK
    uint64_t *s_free = pv_done + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(s_free + 2);
#else
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(pv_done + 1);
#endif

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.y, q0 = blockIdx.x * BQ;
    // q / k operand layouts: head-major [B*6][T][64] (tmap rows = (b*6+h)*T + t, column 0), or token-major as the QKV
    // GEMM writes them, one [B*T][768] matrix (tmap rows = b*T + t, columns h*64 for q and 384 + h*64 for k)
    const int qk_row0 = token_major ? (bh / HEADS) * T : bh * T;
    const int q_col = token_major ? (bh % HEADS) * HD : 0, k_col = token_major ? HEADS * HD + q_col : 0;
    const int nt = (T + BKV - 1) / BKV;
    const int rows_valid = T - q0;                                  // > 0 by the grid size
    const int n_active = min(4, (rows_valid + 31) >> 5);            // softmax warps that own a valid query row

    if (warp == 0 && lane == 0) {
        tc5::tma_prefetch_desc(&tmap_q);
        tc5::tma_prefetch_desc(&tmap_k);
        tc5::tma_prefetch_desc(&tmap_vt);
        tc5::mbar_init(q_full, 1);
        for (int i = 0; i < NK; i++) { tc5::mbar_init(k_full + i, 1); tc5::mbar_init(k_empty + i, 1); }
        for (int i = 0; i < NV; i++) { tc5::mbar_init(v_full + i, 1); tc5::mbar_init(v_empty + i, 1); }
        for (int i = 0; i < 2; i++) { tc5::mbar_init(s_full + i, 1); tc5::mbar_init(p_full + i, n_active); }
        tc5::mbar_init(pv_done, 1);
#if SCP_FA2_EARLY_QK
        for (int i = 0; i < 2; i++) tc5::mbar_init(s_free + i, n_active);
#endif
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
            tc5::tma_load_2d(sQ, &tmap_q, q_full, q_col, qk_row0 + q0);
            for (int j = 0; j < nt; j++) {
                const int ks = j % NK, vs = j % NV;
                tc5::mbar_wait(k_empty + ks, ((j / NK) & 1) ^ 1);
                tc5::mbar_expect_tx(k_full + ks, KT_BYTES);
                tc5::tma_load_2d(sK + ks * KT_BYTES, &tmap_k, k_full + ks, k_col, qk_row0 + j * BKV);
                tc5::mbar_wait(v_empty + vs, ((j / NV) & 1) ^ 1);
                tc5::mbar_expect_tx(v_full + vs, VT_BYTES);
                tc5::tma_load_2d(sV + vs * VT_BYTES, &tmap_vt, v_full + vs, j * BKV, bh * HD);
            }
        }
    } else if (warp == 1) {
        if (tc5::elect_one()) {   // ===== MMA issuer =====
            constexpr uint32_t idesc_pv = tc5::umma_idesc_bf16(BQ, HD);
            const uint32_t aQ = tc5::smem_u32(sQ);
            // S(jj) = Q K(jj)^T into S buffer jj & 1 (4 x K = 16); frees the K stage and publishes S when done
            auto issue_qk = [&](int jj) {
                const int ks = jj % NK;
                tc5::mbar_wait(k_full + ks, (jj / NK) & 1);
                tc5::tc_fence_after();
                const int ncols = min(BKV, (T - jj * BKV + 15) & ~15);   // keys that exist, MMA N granularity 16
                const uint32_t idesc = tc5::umma_idesc_bf16(BQ, ncols);
                const uint32_t aK = tc5::smem_u32(sK + ks * KT_BYTES);
#pragma unroll
                for (int k = 0; k < HD / 16; k++)
                    tc5::umma_bf16(tmem_base + COL_S + (jj & 1) * BKV, tc5::umma_desc_sw128(aQ + k * 32),
                                   tc5::umma_desc_sw128(aK + k * 32), idesc, k != 0);
                tc5::umma_commit(k_empty + ks);
                tc5::umma_commit(s_full + (jj & 1));
            };
            tc5::mbar_wait(q_full, 0);
            issue_qk(0);
            if (nt > 1) issue_qk(1);
            for (int j = 0; j < nt; j++) {
                const int buf = j & 1, vs = j % NV;
#if SCP_FA2_EARLY_QK
                if (j + 2 < nt) {                                // S(j) has been read: its buffer can take S(j+2) now
                    tc5::mbar_wait(s_free + buf, (j >> 1) & 1);
                    tc5::tc_fence_after();
                    issue_qk(j + 2);
                }
#endif
                tc5::mbar_wait(p_full + buf, (j >> 1) & 1);      // P(j) is in TMEM, S buffer `buf` is free again
                tc5::mbar_wait(v_full + vs, (j / NV) & 1);
                tc5::tc_fence_after();
                const int ncols = min(BKV, (T - j * BKV + 15) & ~15);
                const uint32_t aV = tc5::smem_u32(sV + vs * VT_BYTES);
                for (int k = 0; k < ncols / 16; k++)             // O += P(:, 16k..16k+15) V(16k..16k+15, :)
                    tc5::umma_bf16_ts(tmem_base + COL_O, tmem_base + COL_P + buf * (BKV / 2) + k * 8,
                                      tc5::umma_desc_sw128(aV + k * 32), idesc_pv, (j | k) != 0);
                tc5::umma_commit(v_empty + vs);
                tc5::umma_commit(pv_done);
#if !SCP_FA2_EARLY_QK
                if (j + 2 < nt) issue_qk(j + 2);
#endif
            }
        }
    } else if ((warp & 3) < n_active) {
        // ===== softmax warps: thread = query row (TMEM lane); warp w may only touch lanes 32 (w % 4) .. + 31 =====
        const int quarter = warp & 3, row = quarter * 32 + lane;
        const uint32_t t_lane = (uint32_t)(quarter * 32) << 16;
        const uint32_t tO = tmem_base + t_lane + COL_O;
        float m_run = -1e30f, l_run = 0.f;
        for (int j = 0; j < nt; j++) {
            const int buf = j & 1;
            tc5::mbar_wait(s_full + buf, (j >> 1) & 1);
            tc5::tc_fence_after();
            float v[BKV];
            tc5::tmem_ld64(tmem_base + t_lane + COL_S + buf * BKV, v);
#if SCP_FA2_EARLY_QK
            tc5::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc5::mbar_arrive(s_free + buf);
#endif
            const int nvalid = T - j * BKV;
            if (nvalid < BKV) {                                  // last tile: keys past the sequence
#pragma unroll
                for (int i = 0; i < BKV; i++) v[i] = i < nvalid ? v[i] : -3.0e38f;
            }
            float mx0 = fmaxf(v[0], v[1]), mx1 = fmaxf(v[2], v[3]), mx2 = fmaxf(v[4], v[5]), mx3 = fmaxf(v[6], v[7]);
#pragma unroll
            for (int i = 8; i < BKV; i += 4) {
                mx0 = fmaxf(mx0, v[i]); mx1 = fmaxf(mx1, v[i + 1]); mx2 = fmaxf(mx2, v[i + 2]); mx3 = fmaxf(mx3, v[i + 3]);
            }
            const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2e;   // scale > 0
            const bool grow = mx > m_run + LAZY;
            float f = 1.f;
            if (grow) {                                          // (first tile: m_run = -1e30 -> f = 0, l_run = 0)
                f = ex2(m_run - mx);
                m_run = mx;
                l_run *= f;
            }
            if (j > 0 && __any_sync(0xffffffffu, grow)) {        // rescale this warp's rows of the O accumulator
                tc5::mbar_wait(pv_done, (j - 1) & 1);            // every product issued so far has landed
                tc5::tc_fence_after();
#pragma unroll
                for (int c = 0; c < HD; c += 32) {
                    float ov[32];
                    uint32_t w[32];
                    tc5::tmem_ld32(tO + c, ov);
#pragma unroll
                    for (int i = 0; i < 32; i++) w[i] = __float_as_uint(ov[i] * f);
                    tc5::tmem_st32(tO + c, w);
                }
            }
            const float nm = -m_run;
            uint32_t pw[BKV / 2];
            float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
            for (int i = 0; i < BKV / 2; i++) {
                const float x0 = fmaf(v[2 * i], scale_log2e, nm), x1 = fmaf(v[2 * i + 1], scale_log2e, nm);
                const float p0 = ex2(x0);
                constexpr int every = SCP_FA2_POLY_EVERY > 0 ? SCP_FA2_POLY_EVERY : 1;
                const float p1 = (SCP_FA2_POLY_EVERY > 0 && ((2 * i + 1) % every) == every - 1)
                                     ? ex2_poly(x1) : ex2(x1);
                rs0 += p0;
                rs1 += p1;
                __nv_bfloat162 pk = __floats2bfloat162_rn(p0, p1);   // low half = even key
                pw[i] = *reinterpret_cast<uint32_t *>(&pk);
            }
            l_run += rs0 + rs1;
            tc5::tmem_st32(tmem_base + t_lane + COL_P + buf * (BKV / 2), pw);
            tc5::tmem_st_wait();
            tc5::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc5::mbar_arrive(p_full + buf);
        }
        tc5::mbar_wait(pv_done, (nt - 1) & 1);
        tc5::tc_fence_after();
        const int t = q0 + row;
        const float inv = 1.f / l_run;
        const int b = bh / HEADS, h = bh - b * HEADS;
        __nv_bfloat16 *dst = o + ((long)b * T + t) * (HEADS * HD) + h * HD;
#pragma unroll
        for (int c = 0; c < HD; c += 32) {
            float ov[32];
            tc5::tmem_ld32(tO + c, ov);
            if (t < T) {
#pragma unroll
                for (int c8 = 0; c8 < 32; c8 += 8) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        __nv_bfloat162 pk = __floats2bfloat162_rn(ov[c8 + 2 * e] * inv, ov[c8 + 2 * e + 1] * inv);
                        w[e] = *reinterpret_cast<uint32_t *>(&pk);
                    }
                    *reinterpret_cast<uint4 *>(dst + c + c8) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    }
    tc5::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc5::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace fa2
}  // namespace scp
