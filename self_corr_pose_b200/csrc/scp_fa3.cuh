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
//   warp 0 : TMA producer     warp 1 : TMEM allocator + MMA issuer     warps 2.. : softmax
//
// Two forms.  Default: FOUR softmax warps, thread = query row, per-tile re-referencing (fa3_fwd_kernel<4, false>).
// SCP_FA3_WARPS=8 selects EIGHT softmax warps, two per TMEM lane quarter: a query row is shared by two threads, each owning
// 32 of a key tile's 64 columns (and 32 of the 64 output columns); they agree on the reference point once (maximum of the
// first key tile, exchanged through shared memory) and never synchronise again.  A row whose later scores outgrow the
// reference by more than 2^100 (sum test, also catches inf / NaN) would need a rescale of O that both threads would have to
// coordinate -- the CTA raises a retry flag instead and fa3_fwd_kernel<4, true>, launched right behind it, recomputes
// exactly the flagged query tiles.  Measured (B = 64, per layer): 4 warps 0.305 ms, 8 warps + retry launch 0.336 ms --
// the softmax warps are not what bounds the kernel any more: the MMA-issuing thread stalls on the tensor pipe's queue
// (stall_mio on every UTCHMMA), i.e. the 48 M=128 x N=64 x K=16 products per pair of key tiles take ~63 cycles each
// against 32 cycles of arithmetic (operand fetch is not hidden behind products this small; N = 64 is the head dimension).
// Both forms stay parity-tested (tests/test_vit_gpu.py).
#pragma once
#include <cuda_bf16.h>

#include "scp_common.cuh"
#include "scp_gemm.cuh"
#include "scp_tc5.cuh"

namespace scp {
namespace fa3 {

constexpr int BQ = 128, BKV = 64, HD = 64, HEADS = 6;
constexpr int NFLAGS = 1 << 12;                        // retry flags, indexed by (query tile) mod NFLAGS
constexpr int NK = 3, NV = 3;                         // K / V^T ring depths
constexpr int K_TILE = BKV * 128;                     // [64 keys][32 d: hi | lo]
constexpr int KT_BYTES = 2 * K_TILE;                  // 16 KiB
constexpr int V_TILE = HD * BKV * 2;                  // [64 d][64 keys] bf16, one plane
constexpr int VT_BYTES = 2 * V_TILE;                  // 16 KiB: hi plane, lo plane
constexpr int SMEM_BYTES = NK * KT_BYTES + NV * VT_BYTES + 256 + 4 * BQ * 4 + 1024;   // rings, barriers, exchange, alignment slack
constexpr int TMEM_COLS = 256;
constexpr uint32_t COL_QH = 0, COL_QL = 32, COL_S = 64, COL_PH = 128, COL_PL = 160, COL_O = 192;
constexpr float LAZY_SUM = 1099511627776.f;           // 2^40: robust form, a tile's row sum of P above this re-references
constexpr float FAST_SUM = 1.2676506e30f;             // 2^100: fast form, above this the query tile is flagged for the robust form

__device__ __forceinline__ float ex2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ int g_retry[NFLAGS];   // zero-filled by the launcher before the fast kernel (one attention call at a time per device)

// P = 2^(s * scale + nm) for N scores, split into packed hi / lo TMEM words; returns the row sum.
// Per element: 1/2 FFMA2 + MUFU + 1/2 FADD2 + split (F2FP + 2 FHFMA + F2FP per pair) = 4 issue slots.
template <int N>
__device__ __forceinline__ float exps(const float (&v)[N], float scale_log2e, float nm, uint32_t (&ph)[N / 2], uint32_t (&pl)[N / 2])
{
    const uint64_t sc2 = f2_pack(scale_log2e, scale_log2e), nm2 = f2_pack(nm, nm);
    uint64_t ra = f2_pack(0.f, 0.f), rb = ra;
#pragma unroll
    for (int i = 0; i < N / 2; i += 2) {
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
}
template <int N>
__device__ __forceinline__ float row_max(const float (&v)[N])
{
    float mx0 = fmaxf(v[0], v[1]), mx1 = fmaxf(v[2], v[3]), mx2 = fmaxf(v[4], v[5]), mx3 = fmaxf(v[6], v[7]);
#pragma unroll
    for (int i = 8; i < N; i += 4) {
        mx0 = fmaxf(mx0, v[i]); mx1 = fmaxf(mx1, v[i + 1]); mx2 = fmaxf(mx2, v[i + 2]); mx3 = fmaxf(mx3, v[i + 3]);
    }
    return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
}

// NSW = softmax warps: 4 = thread per query row with per-tile re-referencing (robust; the default), 8 = two threads per
// query row with a fixed reference point (flags what it cannot handle).  RETRY: run only on flagged query tiles.
template <int NSW, bool RETRY>
__global__ void __launch_bounds__(64 + 32 * NSW, 2)
fa3_fwd_kernel(const __nv_bfloat16 *__restrict__ qk, const __grid_constant__ CUtensorMap tmap_k,
               const __grid_constant__ CUtensorMap tmap_vt, __nv_bfloat16 *__restrict__ o, int T, float scale_log2e,
               int vt_plane_rows)
{
    constexpr bool FAST = NSW == 8;
    constexpr int SPLIT = NSW / 4;                                   // threads per query row
    constexpr int CW = BKV / SPLIT;                                  // score columns per thread and tile
    int *retry = g_retry + ((blockIdx.y * gridDim.x + blockIdx.x) & (NFLAGS - 1));
    if (RETRY && *reinterpret_cast<volatile int *>(retry) == 0) return;   // uniform over the CTA

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sK = smem, *sV = sK + NK * KT_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sV + NV * VT_BYTES);
    uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = k_full + NK, *v_full = k_empty + NK, *v_empty = v_full + NV,
             *s_full = v_empty + NV, *p_full = s_full + 1, *pv_done = p_full + 1, *s_free = pv_done + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(s_free + 1);
    float *s_xch = reinterpret_cast<float *>(bars + 32);              // [2 halves][128 rows] reference points, then the same for row sums

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.y, q0 = blockIdx.x * BQ;
    const int b = bh / HEADS, h = bh - b * HEADS;
    const int row0 = b * T;                                          // first token row of this image
    const int q_col = h * 2 * HD, k_col = HEADS * 2 * HD + q_col;    // physical columns (hi | lo interleaved by 32)
    const int nt = (T + BKV - 1) / BKV;
    const int rows_valid = T - q0;                                  // > 0 by the grid size
    const int n_active = min(4, (rows_valid + 31) >> 5);            // row quarters that own a valid query row

    if (warp == 0 && lane == 0) {
        tc5::tma_prefetch_desc(&tmap_k);
        tc5::tma_prefetch_desc(&tmap_vt);
        tc5::mbar_init(q_full, NSW);                                 // all row quarters are stored (zeros past T)
        for (int i = 0; i < NK; i++) { tc5::mbar_init(k_full + i, 1); tc5::mbar_init(k_empty + i, 1); }
        for (int i = 0; i < NV; i++) { tc5::mbar_init(v_full + i, 1); tc5::mbar_init(v_empty + i, 1); }
        tc5::mbar_init(s_full, 1);
        tc5::mbar_init(s_free, SPLIT * n_active);
        tc5::mbar_init(p_full, SPLIT * n_active);
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
        // ===== softmax warps: thread = query row (TMEM lane; warp w may only touch lanes 32 (w % 4) .. + 31) x column half =====
        const int quarter = warp & 3, half = FAST ? (warp - 2) >> 2 : 0, row = quarter * 32 + lane;
        const uint32_t t_lane = (uint32_t)(quarter * 32) << 16;
        const int t = q0 + row;
        {
            // Q into tensor memory.  A row's 256 bytes are [hi d0..31 | lo d0..31 | hi d32..63 | lo d32..63]; a 32-bit word
            // holds two consecutive d (low half = even d) = the packed A-operand word of a K = 16 step.  In the fast form
            // thread `half` moves d 32 half .. + 31.
            const uint4 *src = reinterpret_cast<const uint4 *>(qk + ((long)row0 + t) * (4 * HEADS * HD) + q_col);
#pragma unroll
            for (int g = FAST ? half : 0; g < (FAST ? half + 1 : 2); g++) {
                uint32_t qh[16], ql[16];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    uint4 a = make_uint4(0u, 0u, 0u, 0u), c = a;
                    if (t < T) { a = __ldg(src + g * 8 + i); c = __ldg(src + g * 8 + 4 + i); }
                    qh[4 * i] = a.x; qh[4 * i + 1] = a.y; qh[4 * i + 2] = a.z; qh[4 * i + 3] = a.w;
                    ql[4 * i] = c.x; ql[4 * i + 1] = c.y; ql[4 * i + 2] = c.z; ql[4 * i + 3] = c.w;
                }
                tc5::tmem_st16(tmem_base + t_lane + COL_QH + g * 16, qh);
                tc5::tmem_st16(tmem_base + t_lane + COL_QL + g * 16, ql);
            }
            tc5::tmem_st_wait();
            tc5::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc5::mbar_arrive(q_full);
        }
        if (quarter < n_active) {
            const uint32_t tS = tmem_base + t_lane + COL_S + half * CW, tO = tmem_base + t_lane + COL_O;
            const uint32_t tPh = tmem_base + t_lane + COL_PH + half * (CW / 2), tPl = tmem_base + t_lane + COL_PL + half * (CW / 2);
            float m_run = -1e30f, l_run = 0.f;
            bool flagged = false;
            for (int j = 0; j < nt; j++) {
                tc5::mbar_wait(s_full, j & 1);
                tc5::tc_fence_after();
                float v[CW];
                if constexpr (FAST) tc5::tmem_ld32(tS, v);
                else tc5::tmem_ld64(tS, v);
                tc5::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc5::mbar_arrive(s_free);              // S(j) consumed: the buffer may take S(j+1)
                const int nvalid = T - j * BKV - half * CW;
                if (nvalid < CW) {                                   // last tile: keys past the sequence
#pragma unroll
                    for (int i = 0; i < CW; i++) v[i] = i < nvalid ? v[i] : -3.0e38f;
                }
                // Reference point = maximum of the FIRST key tile; later tiles exponentiate against it directly (fp32 P, the
                // hi/lo split and the fp32 accumulators are all relative-precision, so a reference point below the true
                // maximum costs nothing until the sums leave the fp32 range).  The row sum the loop forms anyway bounds every
                // term: above the limit (inf / NaN fail the test too) the robust form re-references against the tile's own
                // maximum and rescales O; the fast form flags the query tile for the robust form.
                uint32_t ph[CW / 2], pl[CW / 2];
                bool grow = false;
                float f = 1.f, rs;
                if (j == 0) {
                    m_run = row_max(v) * scale_log2e;                // scale > 0; l_run = 0, O not written yet
                    if constexpr (FAST) {                            // the row's two threads agree on the reference
                        s_xch[half * BQ + row] = m_run;
                        asm volatile("bar.sync 1, %0;" ::"r"(64 * n_active) : "memory");
                        m_run = fmaxf(m_run, s_xch[(half ^ 1) * BQ + row]);
                    }
                    rs = exps(v, scale_log2e, -m_run, ph, pl);
                } else {
                    rs = exps(v, scale_log2e, -m_run, ph, pl);
                    if constexpr (FAST) {
                        flagged |= !(rs <= FAST_SUM);
                    } else if (!(rs <= LAZY_SUM)) {                  // rare: scores far above the reference point
                        const float mx = row_max(v) * scale_log2e;
                        grow = true;
                        f = ex2(m_run - mx);
                        m_run = mx;
                        l_run *= f;
                        rs = exps(v, scale_log2e, -m_run, ph, pl);
                    }
                }
                l_run += rs;
                if (j > 0) {                                         // PV(j-1) has read the P buffer (and landed in O)
                    tc5::mbar_wait(pv_done, (j - 1) & 1);
                    tc5::tc_fence_after();
                    if constexpr (!FAST) {
                        if (__any_sync(0xffffffffu, grow)) {         // rescale this warp's rows of the O accumulator
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
                }
                if constexpr (FAST) {
                    tc5::tmem_st16(tPh, ph);
                    tc5::tmem_st16(tPl, pl);
                } else {
                    tc5::tmem_st32(tPh, ph);
                    tc5::tmem_st32(tPl, pl);
                }
                tc5::tmem_st_wait();
                tc5::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc5::mbar_arrive(p_full);
            }
            if constexpr (FAST) {                                    // row sum = both halves; retry flag
                s_xch[(2 + half) * BQ + row] = l_run;
                asm volatile("bar.sync 1, %0;" ::"r"(64 * n_active) : "memory");
                l_run += s_xch[(2 + (half ^ 1)) * BQ + row];
                if (__any_sync(0xffffffffu, flagged) && lane == 0) atomicOr(retry, 1);
            }
            tc5::mbar_wait(pv_done, (nt - 1) & 1);
            tc5::tc_fence_after();
            const float inv = 1.f / l_run;
            // split output row: physical columns [128 h, 128 h + 128) = [hi d0..31 | lo d0..31 | hi d32..63 | lo d32..63]
            __nv_bfloat16 *dst = o + ((long)row0 + t) * (2 * HEADS * HD) + h * 2 * HD;
#pragma unroll
            for (int c = FAST ? 32 * half : 0; c < (FAST ? 32 * half + 32 : HD); c += 32) {
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
    }
    tc5::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc5::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace fa3
}  // namespace scp
