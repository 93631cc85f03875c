// Flash attention for the ViT on tcgen05 (sm_100a) at fp32-class precision ("x3" mode, see scp_gemm.cuh): the same
// pipeline as scp_fa2.cuh (S double-buffered in TMEM, P written back into TENSOR MEMORY, O accumulated in TMEM with a
// lazy running maximum) with every tensor-core product carried as three bf16 products of split operands:
//     S  = Qhi Khi^T + Qhi Klo^T + Qlo Khi^T            (q, k, v arrive split as hi + lo bf16 pairs)
//     O += Phi Vhi   + Phi Vlo   + Plo Vhi              (P = exp2(..) is split by the softmax warps: 16 mantissa bits)
// so the attention output carries ~2^-16 relative error instead of the 2^-9 of bf16 P / bf16 operands -- this is what the
// 1e-3 / 99.9 % arg-max parity contract against the reference's fp32 attention
// (third-party/zsp/zsp/method/vision_transformer_flexible.py:85-101) needs.
//
//   q | k : token-major split matrix [B*T][1536] as the QKV GEMM writes it (i32 layout: groups of 32 logical columns stored
//           as [32 hi | 32 lo]); head h of q = physical columns [128 h, 128 h + 128), of k = 768 + the same
//   v     : transposed per head, two planes: vt[2][B*6*64][Tp] (plane 0 = hi, plane 1 = lo), keys contiguous
//   o     : split i32 layout [B*T][768] (the A operand of the projection GEMM)
//
// Round 2, second version -- Q lives in TENSOR MEMORY.  The ncu capture of the first version (Q in shared memory, S double
// buffered) shows the tensor pipe and the tensor cores' shared-memory read path (l1tex__data_pipe_tc_wavefronts_mem_shared)
// saturating together at ~58 %: an M = 128, N = 64, K = 16 product with both operands in shared memory reads 4 KB of Q and
// 2 KB of K for 32 cycles of tensor work = 192 B/clk against the SM's 128 B/clk, so every score product ran at 2/3 rate and
// Q was re-read 12 times per key tile.  Now the softmax threads load their own query row (256 B, split) from global memory
// once and store it into TMEM as the A operand of the score products (the form the P V products already use); the score
// products read only K (64 B/clk), shared memory holds only the K / V^T rings (3 stages each), and the S accumulator is
// single-buffered: S(j+1) is issued as soon as the softmax warps have READ S(j) into registers, so it runs while they
// exponentiate, ahead of P V(j) in the in-order tensor queue.
//   TMEM columns (256 per CTA, two CTAs per SM):  Qhi [0,32)  Qlo [32,64)  S [64,128)  Phi [128,160)  Plo [160,192)  O [192,256)
//   shared memory: K ring 3 x 16 KiB ([64 keys][32 d: hi | lo] x 2), V^T ring 3 x 16 KiB (hi tile + lo tile) = 96 KiB per CTA
//   warp 0 : TMA producer     warp 1 : TMEM allocator + MMA issuer     warps 2..5 : softmax (thread = query row)
#pragma once
#include <cuda_bf16.h>

#include "scp_common.cuh"
#include "scp_gemm.cuh"
#include "scp_tc5.cuh"

namespace scp {
namespace fa3 {

constexpr int BQ = 128, BKV = 64, HD = 64, HEADS = 6;
constexpr int NTHREADS = 192;
constexpr int NK = 3, NV = 3;                         // K / V^T ring depths
constexpr int K_TILE = BKV * 128;                     // [64 keys][32 d: hi | lo]
constexpr int KT_BYTES = 2 * K_TILE;                  // 16 KiB
constexpr int V_TILE = HD * BKV * 2;                  // [64 d][64 keys] bf16, one plane
constexpr int VT_BYTES = 2 * V_TILE;                  // 16 KiB: hi plane, lo plane
constexpr int SMEM_BYTES = NK * KT_BYTES + NV * VT_BYTES + 256 + 1024;
constexpr int TMEM_COLS = 256;
constexpr uint32_t COL_QH = 0, COL_QL = 32, COL_S = 64, COL_PH = 128, COL_PL = 160, COL_O = 192;
constexpr float LAZY_SUM = 1099511627776.f;           // 2^40: a tile's row sum of P above this triggers the re-referencing

__device__ __forceinline__ float ex2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(NTHREADS, 2)
fa3_fwd_kernel(const __nv_bfloat16 *__restrict__ qk, const __grid_constant__ CUtensorMap tmap_k,
               const __grid_constant__ CUtensorMap tmap_vt, __nv_bfloat16 *__restrict__ o, int T, float scale_log2e,
               int vt_plane_rows)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sK = smem, *sV = sK + NK * KT_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sV + NV * VT_BYTES);
    uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = k_full + NK, *v_full = k_empty + NK, *v_empty = v_full + NV,
             *s_full = v_empty + NV, *p_full = s_full + 1, *pv_done = p_full + 1, *s_free = pv_done + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(s_free + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.y, q0 = blockIdx.x * BQ;
    const int b = bh / HEADS, h = bh - b * HEADS;
    const int row0 = b * T;                                          // first token row of this image
    const int q_col = h * 2 * HD, k_col = HEADS * 2 * HD + q_col;    // physical columns (hi | lo interleaved by 32)
    const int nt = (T + BKV - 1) / BKV;
    const int rows_valid = T - q0;                                  // > 0 by the grid size
    const int n_active = min(4, (rows_valid + 31) >> 5);            // softmax warps that own a valid query row

    if (warp == 0 && lane == 0) {
        tc5::tma_prefetch_desc(&tmap_k);
        tc5::tma_prefetch_desc(&tmap_vt);
        tc5::mbar_init(q_full, 4);                                   // all four row quarters are stored (zeros past T)
        for (int i = 0; i < NK; i++) { tc5::mbar_init(k_full + i, 1); tc5::mbar_init(k_empty + i, 1); }
        for (int i = 0; i < NV; i++) { tc5::mbar_init(v_full + i, 1); tc5::mbar_init(v_empty + i, 1); }
        tc5::mbar_init(s_full, 1);
        tc5::mbar_init(s_free, n_active);
        tc5::mbar_init(p_full, n_active);
        tc5::mbar_init(pv_done, 1);
        tc5::mbar_fence_init();
    }
    if (warp == 1) tc5::tmem_alloc(tmem_slot, TMEM_COLS);
    tc5::tc_fence_before();
    __syncthreads();
    tc5::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (tc5::elect_one()) {   // ===== TMA producer =====
            for (int j = 0; j < nt; j++) {
                const int ks = j % NK, vs = j % NV;
                tc5::mbar_wait(k_empty + ks, ((j / NK) & 1) ^ 1);
                tc5::mbar_expect_tx(k_full + ks, KT_BYTES);
                tc5::tma_load_2d(sK + ks * KT_BYTES, &tmap_k, k_full + ks, k_col, row0 + j * BKV);
                tc5::tma_load_2d(sK + ks * KT_BYTES + K_TILE, &tmap_k, k_full + ks, k_col + 64, row0 + j * BKV);
                tc5::mbar_wait(v_empty + vs, ((j / NV) & 1) ^ 1);
                tc5::mbar_expect_tx(v_full + vs, VT_BYTES);
                tc5::tma_load_2d(sV + vs * VT_BYTES, &tmap_vt, v_full + vs, j * BKV, bh * HD);
                tc5::tma_load_2d(sV + vs * VT_BYTES + V_TILE, &tmap_vt, v_full + vs, j * BKV, vt_plane_rows + bh * HD);
            }
        }
    } else if (warp == 1) {
        // elect.sync, not lane == 0: ptxas then knows a single thread runs the loop and emits the tcgen05.mma sequence
        // without a per-instruction uniformisation loop (ELECT / PLOP3 / BRA.U.ANY, ~10 dependent instructions per MMA)
        if (tc5::elect_one()) {   // ===== MMA issuer =====
            constexpr uint32_t idesc_pv = tc5::umma_idesc_bf16(BQ, HD);
            const uint32_t tQh = tmem_base + COL_QH, tQl = tmem_base + COL_QL, tS = tmem_base + COL_S;
            // S(jj) = Q K(jj)^T: four K=16 steps over d, three split products each; Q from TMEM (8 columns per step)
            auto issue_qk = [&](int jj) {
                const int ks = jj % NK;
                tc5::mbar_wait(k_full + ks, (jj / NK) & 1);
                tc5::tc_fence_after();
                const int ncols = min(BKV, (T - jj * BKV + 15) & ~15);   // keys that exist, MMA N granularity 16
                const uint32_t idesc = tc5::umma_idesc_bf16(BQ, ncols);
                const uint32_t aK = tc5::smem_u32(sK + ks * KT_BYTES);
#pragma unroll
                for (int s = 0; s < HD / 16; s++) {
                    const uint32_t kh = aK + (s >> 1) * K_TILE + (s & 1) * 32;
                    tc5::umma_bf16_ts(tS, tQh + s * 8, tc5::umma_desc_sw128(kh), idesc, s != 0);
                    tc5::umma_bf16_ts(tS, tQh + s * 8, tc5::umma_desc_sw128(kh + 64), idesc, 1);
                    tc5::umma_bf16_ts(tS, tQl + s * 8, tc5::umma_desc_sw128(kh), idesc, 1);
                }
                tc5::umma_commit(k_empty + ks);
                tc5::umma_commit(s_full);
            };
            tc5::mbar_wait(q_full, 0);
            tc5::tc_fence_after();
            issue_qk(0);
            for (int j = 0; j < nt; j++) {
                const int vs = j % NV;
                if (j + 1 < nt) {          // S(j) is in the softmax warps' registers: its buffer takes S(j+1) now
                    tc5::mbar_wait(s_free, j & 1);
                    tc5::tc_fence_after();
                    issue_qk(j + 1);
                }
                tc5::mbar_wait(p_full, j & 1);                    // P(j) is in TMEM
                tc5::mbar_wait(v_full + vs, (j / NV) & 1);
                tc5::tc_fence_after();
                const int ncols = min(BKV, (T - j * BKV + 15) & ~15);
                const uint32_t aV = tc5::smem_u32(sV + vs * VT_BYTES);
                for (int k = 0; k < ncols / 16; k++) {            // O += P(:, 16k..16k+15) V(16k..16k+15, :)
                    const uint32_t ph = tmem_base + COL_PH + k * 8, pl = tmem_base + COL_PL + k * 8;
                    tc5::umma_bf16_ts(tmem_base + COL_O, ph, tc5::umma_desc_sw128(aV + k * 32), idesc_pv, (j | k) != 0);
                    tc5::umma_bf16_ts(tmem_base + COL_O, ph, tc5::umma_desc_sw128(aV + V_TILE + k * 32), idesc_pv, 1);
                    tc5::umma_bf16_ts(tmem_base + COL_O, pl, tc5::umma_desc_sw128(aV + k * 32), idesc_pv, 1);
                }
                tc5::umma_commit(v_empty + vs);
                tc5::umma_commit(pv_done);
            }
        }
    } else {
        // ===== Q into tensor memory: thread = query row; its 256 bytes are [hi d0..31 | lo d0..31 | hi d32..63 | lo d32..63],
        // a 32-bit word holds two consecutive d (low half = even d) -- exactly the packed A-operand word of a K = 16 step
        const int quarter = warp & 3, row = quarter * 32 + lane;
        const uint32_t t_lane = (uint32_t)(quarter * 32) << 16;
        uint32_t qh[32], ql[32];
        if (q0 + row < T) {
            const uint4 *src = reinterpret_cast<const uint4 *>(qk + ((long)row0 + q0 + row) * (4 * HEADS * HD) + q_col);
#pragma unroll
            for (int g = 0; g < 2; g++) {                         // d 0..31, d 32..63
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint4 a = __ldg(src + g * 8 + i), c = __ldg(src + g * 8 + 4 + i);
                    qh[g * 16 + 4 * i] = a.x; qh[g * 16 + 4 * i + 1] = a.y; qh[g * 16 + 4 * i + 2] = a.z; qh[g * 16 + 4 * i + 3] = a.w;
                    ql[g * 16 + 4 * i] = c.x; ql[g * 16 + 4 * i + 1] = c.y; ql[g * 16 + 4 * i + 2] = c.z; ql[g * 16 + 4 * i + 3] = c.w;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; i++) qh[i] = ql[i] = 0u;
        }
        tc5::tmem_st32(tmem_base + t_lane + COL_QH, qh);
        tc5::tmem_st32(tmem_base + t_lane + COL_QL, ql);
        tc5::tmem_st_wait();
        tc5::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc5::mbar_arrive(q_full);
    }
    if (warp >= 2 && (warp & 3) < n_active) {
        // ===== softmax warps: thread = query row (TMEM lane); warp w may only touch lanes 32 (w % 4) .. + 31 =====
        const int quarter = warp & 3, row = quarter * 32 + lane;
        const uint32_t t_lane = (uint32_t)(quarter * 32) << 16;
        const uint32_t tO = tmem_base + t_lane + COL_O;
        float m_run = -1e30f, l_run = 0.f;
        for (int j = 0; j < nt; j++) {
            tc5::mbar_wait(s_full, j & 1);
            tc5::tc_fence_after();
            float v[BKV];
            tc5::tmem_ld64(tmem_base + t_lane + COL_S, v);
            tc5::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc5::mbar_arrive(s_free);              // S(j) consumed: the buffer may take S(j+1)
            const int nvalid = T - j * BKV;
            if (nvalid < BKV) {                                  // last tile: keys past the sequence
#pragma unroll
                for (int i = 0; i < BKV; i++) v[i] = i < nvalid ? v[i] : -3.0e38f;
            }
            // Row maximum only where it is needed: on the first tile (it becomes the reference point m_run) and on the slow
            // path below.  Later tiles exponentiate against m_run directly; a row whose scores outgrow it by more than
            // 2^LAZY is detected AFTER the fact from the row sum it has to form anyway (sum <= 2^LAZY bounds every term;
            // inf / NaN fail the test too) and redone against its own maximum, with the O rescale of the lazy scheme.
            // fp32 P, the hi/lo split and the fp32 accumulators are all relative-precision, so a reference point up to
            // 2^LAZY below the true maximum costs nothing.  Per element: 1/2 FFMA2 + MUFU + 1/2 FADD2 + split (2) = 4
            // issue slots (was 7.5: FMNMX, FFMA, MUFU, FADD, split 3) -- the softmax warps' issue rate bounds this kernel.
            uint32_t ph[BKV / 2], pl[BKV / 2];
            const uint64_t sc2 = f2_pack(scale_log2e, scale_log2e);
            auto row_max = [&]() {
                float mx0 = fmaxf(v[0], v[1]), mx1 = fmaxf(v[2], v[3]), mx2 = fmaxf(v[4], v[5]), mx3 = fmaxf(v[6], v[7]);
#pragma unroll
                for (int i = 8; i < BKV; i += 4) {
                    mx0 = fmaxf(mx0, v[i]); mx1 = fmaxf(mx1, v[i + 1]); mx2 = fmaxf(mx2, v[i + 2]); mx3 = fmaxf(mx3, v[i + 3]);
                }
                return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2e;   // scale > 0
            };
            auto exps = [&](float nm) {   // P = 2^(s * scale - m_run), split into TMEM words; returns the row sum
                const uint64_t nm2 = f2_pack(nm, nm);
                uint64_t ra = f2_pack(0.f, 0.f), rb = ra;
#pragma unroll
                for (int i = 0; i < BKV / 2; i += 2) {
                    float x0, x1, x2, x3;
                    f2_unpack(f2_fma(f2_pack(v[2 * i], v[2 * i + 1]), sc2, nm2), x0, x1);
                    f2_unpack(f2_fma(f2_pack(v[2 * i + 2], v[2 * i + 3]), sc2, nm2), x2, x3);
                    const float p0 = ex2(x0), p1 = ex2(x1), p2 = ex2(x2), p3 = ex2(x3);
                    ra = f2_add(ra, f2_pack(p0, p1));
                    rb = f2_add(rb, f2_pack(p2, p3));
                    ph[i] = gemm::split_bf16x2(p0, p1, pl[i]);        // low half = even key
                    ph[i + 1] = gemm::split_bf16x2(p2, p3, pl[i + 1]);
                }
                float r0, r1, r2, r3;
                f2_unpack(ra, r0, r1);
                f2_unpack(rb, r2, r3);
                return (r0 + r1) + (r2 + r3);
            };
            bool grow = false;
            float f = 1.f, rs;
            if (j == 0) {
                m_run = row_max();                               // l_run = 0, O not written yet: nothing to rescale
                rs = exps(-m_run);
            } else {
                rs = exps(-m_run);
                if (!(rs <= LAZY_SUM)) {                         // rare: scores far above the reference point
                    const float mx = row_max();
                    grow = true;
                    f = ex2(m_run - mx);
                    m_run = mx;
                    l_run *= f;
                    rs = exps(-m_run);
                }
            }
            l_run += rs;
            if (j > 0) {                                         // PV(j-1) has read the P buffer (and landed in O)
                tc5::mbar_wait(pv_done, (j - 1) & 1);
                tc5::tc_fence_after();
                if (__any_sync(0xffffffffu, grow)) {             // rescale this warp's rows of the O accumulator
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
            }
            tc5::tmem_st32(tmem_base + t_lane + COL_PH, ph);
            tc5::tmem_st32(tmem_base + t_lane + COL_PL, pl);
            tc5::tmem_st_wait();
            tc5::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc5::mbar_arrive(p_full);
        }
        tc5::mbar_wait(pv_done, (nt - 1) & 1);
        tc5::tc_fence_after();
        const int t = q0 + row;
        const float inv = 1.f / l_run;
        // split output row: physical columns [128 h, 128 h + 128) = [hi d0..31 | lo d0..31 | hi d32..63 | lo d32..63]
        __nv_bfloat16 *dst = o + ((long)row0 + t) * (2 * HEADS * HD) + h * 2 * HD;
#pragma unroll
        for (int c = 0; c < HD; c += 32) {
            float ov[32];
            tc5::tmem_ld32(tO + c, ov);
            if (t < T) {
#pragma unroll
                for (int c8 = 0; c8 < 32; c8 += 8) {
                    uint32_t wh[4], wl[4];
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        wh[e] = gemm::split_bf16x2(ov[c8 + 2 * e] * inv, ov[c8 + 2 * e + 1] * inv, wl[e]);
                    *reinterpret_cast<uint4 *>(dst + 2 * c + c8) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
                    *reinterpret_cast<uint4 *>(dst + 2 * c + 32 + c8) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
                }
            }
        }
    }
    tc5::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc5::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace fa3
}  // namespace scp
