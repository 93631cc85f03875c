// Row-statistics GEMM for sm_100a: S = A W^T with K = 64 on the tcgen05 tensor cores, consumed in place -- the
// accumulator never leaves the SM.  Sibling of gemm_bf16_tn_kernel (scp_gemm.cuh: same warp roles, barriers, tile shape)
// for the correspondence soft-maxes (scp_corr_tc.cu), different in three ways:
//
//   * operands are fp32 matrices split into TF32 (hi, lo) pairs, groups of 16 logical columns stored as
//     [16 hi | 16 lo] fp32 = 128 bytes (one swizzle-atom row); per group three kind::tf32 products per K = 8 chunk,
//     hi*hi + hi*lo + lo*hi: the accuracy of the 3-term mma.sync split this replaces (~2^-21 per product);
//   * the A tile is STATIONARY: K = 64 is exactly four 128-byte k-blocks = the four stages of the shared-memory ring, so
//     the ring's A halves (4 x 32 KiB) hold the whole 256-row A tile and only the W halves are refilled while a CTA walks
//     the n tiles of a "unit" = (m tile, range of NU n tiles).  Operand traffic from the L2 per 256 x 128 tile drops from
//     192 KiB to 64 KiB + 128 KiB / NU (the first version, which reloaded A per tile, ran at the L2 -> SM limit:
//     profiles/r2_corr_tcgen05.md);
//   * the epilogue is a functor that folds 32-column chunks of the accumulator into eight registers per row, and units /
//     n tiles without live rows or columns are skipped by all three roles.  An epilogue warp owns one TMEM lane quarter
//     and one 64-column half of the tile for BOTH M halves, i.e. a thread holds two rows (r, r + 128): whatever the
//     functor loads per column (warp-uniform addresses: four LSU write-back cycles per 16-byte load, the unit that bounded
//     the first version) serves two accumulator elements.
//       __device__ int  live_n_tiles(int m_blk) const;            n tiles to visit for this m tile (0: skip the m tile)
//       __device__ bool chunk_live(int row0, int col0) const;     warp-uniform; false skips the chunk (no TMEM load);
//                                                                 row0 = first row of the warp in M half 0
//       __device__ void accum2(int row, int col0, const float (&v0)[32], const float (&v1)[32], float (&a0)[8],
//                              float (&a1)[8]) const;             lane = row (v0) and row + 128 (v1)
//       __device__ void finish(int row, int n_blk, int chalf, const float (&a)[8]) const;   per row, tile, column half
//   * units are enumerated position-major (n range, then position of the m tile inside its problem, then problem), so
//     that neighbouring units -- which the static round-robin hands to different CTAs -- have the same liveness.
//
// Batched: the M axis is NB stacked problems of rows_a rows (a multiple of 256); problem b multiplies its rows of A with
// rows [b * rows_w, +rows_w) of W (rows_w a multiple of 128); `col` in the functor calls is the column inside the problem.
#pragma once
#include "scp_gemm.cuh"

namespace scp {
namespace gemm_rs {

using gemm::BM;
using gemm::BN;
using gemm::BK;
constexpr int KB = 4;                                 // k-blocks of 128 bytes = ring stages (K = 64 logical columns)
constexpr int STAGES = KB;
constexpr int ACC_STAGES = 2;
#ifndef SCP_RS_EPI_WARPS
#define SCP_RS_EPI_WARPS 8          // 8: a warp owns 64 columns (two chunks) of a lane quarter; 16: 32 columns (one chunk)
#endif
constexpr int EPI_WARPS = SCP_RS_EPI_WARPS;
constexpr int CSPLIT = EPI_WARPS / 4;                 // column ranges per tile = partial slots per row and tile
constexpr int NTHREADS = 64 + 32 * EPI_WARPS;
constexpr int STAGE_BYTES = (BM + BN) * BK * 2;     // 48 KiB: [A 256 rows | W 128 rows] x 128 B
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 /*barriers*/ + 1024 /*align slack*/;
constexpr int ACC_COLS = 2 * BN;                     // TMEM columns of one accumulator stage: [M half][BN]
constexpr int TMEM_COLS = ACC_STAGES * ACC_COLS;    // 512

struct Shape {
    int tiles_m, tiles_n;      // m tiles over all problems, n tiles per problem
    int nu, units_n;           // n tiles per unit, units per m tile
    int tpb;                   // m tiles per problem
    int rows_w;                // W rows per problem
};

// unit u -> m tile and first n tile: n range slowest, then the m tile's position inside its problem, then the problem
__device__ __forceinline__ void decode_unit(const Shape &s, int u, int &m_blk, int &n_beg)
{
    const int nr = u / s.tiles_m, r = u - nr * s.tiles_m, nb = s.tiles_m / s.tpb;
    const int pos = r / nb, b = r - pos * nb;
    m_blk = b * s.tpb + pos;
    n_beg = nr * s.nu;
}

template <class Epi>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_rowstats_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, Shape s, Epi epi)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *full = bars, *empty = bars + STAGES, *tfull = bars + 2 * STAGES, *tempty = bars + 2 * STAGES + ACC_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 2 * ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nunits = s.tiles_m * s.units_n;

    if (warp == 0 && lane == 0) {
        tc5::tma_prefetch_desc(&tmap_a);
        tc5::tma_prefetch_desc(&tmap_w);
        for (int i = 0; i < STAGES; i++) { tc5::mbar_init(full + i, 1); tc5::mbar_init(empty + i, 1); }
        for (int i = 0; i < ACC_STAGES; i++) { tc5::mbar_init(tfull + i, 1); tc5::mbar_init(tempty + i, EPI_WARPS); }
        tc5::mbar_fence_init();
    }
    if (warp == 1) tc5::tmem_alloc(tmem_slot, TMEM_COLS);
    tc5::tc_fence_before();
    __syncthreads();
    tc5::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // every role walks the same sequence: unit u -> (m tile, n range clipped to the live n tiles of that m tile)
    if (warp == 0) {
        // ===== TMA producer: stage kb of the ring carries k-block kb; its A half is loaded once per unit =====
        if (tc5::elect_one()) {
            int phase = 0;
            for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
                int m_blk, n_beg;
                decode_unit(s, u, m_blk, n_beg);
                const int n_end = min(n_beg + s.nu, epi.live_n_tiles(m_blk));
                const int w_base = (m_blk / s.tpb) * s.rows_w;
                for (int n_blk = n_beg; n_blk < n_end; n_blk++) {
                    for (int kb = 0; kb < KB; kb++) {
                        tc5::mbar_wait(empty + kb, phase ^ 1);
                        uint8_t *sa = smem + kb * STAGE_BYTES, *sb = sa + BM * BK * 2;
                        if (n_blk == n_beg) {
                            tc5::mbar_expect_tx(full + kb, (BM + BN) * BK * 2);
                            tc5::tma_load_2d(sa, &tmap_a, full + kb, kb * BK, m_blk * BM);
                        } else {
                            tc5::mbar_expect_tx(full + kb, BN * BK * 2);
                        }
                        tc5::tma_load_2d(sb, &tmap_w, full + kb, kb * BK, w_base + n_blk * BN);
                    }
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (tc5::elect_one()) {
            constexpr uint32_t idesc = tc5::umma_idesc_tf32(128, BN);
            int phase = 0, acc = 0, acc_phase = 0;
            for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
                int m_blk, n_beg;
                decode_unit(s, u, m_blk, n_beg);
                const int n_end = min(n_beg + s.nu, epi.live_n_tiles(m_blk));
                for (int n_blk = n_beg; n_blk < n_end; n_blk++) {
                    tc5::mbar_wait(tempty + acc, acc_phase ^ 1);       // epilogue has drained this accumulator
                    tc5::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
                    for (int kb = 0; kb < KB; kb++) {
                        tc5::mbar_wait(full + kb, phase);              // TMA bytes have landed
                        tc5::tc_fence_after();
                        const uint32_t sa = tc5::smem_u32(smem + kb * STAGE_BYTES), sb = sa + BM * BK * 2;
#pragma unroll
                        for (int mh = 0; mh < 2; mh++) {
#pragma unroll
                            for (int c = 0; c < 2; c++) {              // K = 8 chunk c: hi at byte c*32, lo at 64 + c*32
                                const uint32_t ah = sa + mh * (128 * BK * 2) + c * 32, bh = sb + c * 32;
                                tc5::umma_tf32(d_tmem + mh * BN, tc5::umma_desc_sw128(ah), tc5::umma_desc_sw128(bh), idesc,
                                               (kb | c) != 0);
                                tc5::umma_tf32(d_tmem + mh * BN, tc5::umma_desc_sw128(ah), tc5::umma_desc_sw128(bh + 64), idesc, 1);
                                tc5::umma_tf32(d_tmem + mh * BN, tc5::umma_desc_sw128(ah + 64), tc5::umma_desc_sw128(bh), idesc, 1);
                            }
                        }
                        tc5::umma_commit(empty + kb);                  // frees the stage's W half (and A at a unit's end)
                    }
                    phase ^= 1;
                    tc5::umma_commit(tfull + acc);                     // accumulator complete -> epilogue
                    if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else {
        // ===== epilogue warps: TMEM lane quarter = warp % 4, column half of the tile = (warp - 2) / 4, both M halves =====
        const int quarter = warp & 3, chalf = (warp - 2) >> 2;     // chalf: column range 0 .. CSPLIT-1 of BN / CSPLIT columns
        const uint32_t t_lane = (uint32_t)(quarter * 32) << 16;
        int acc = 0, acc_phase = 0;
        for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
            int m_blk, n_beg;
            decode_unit(s, u, m_blk, n_beg);
            const int n_end = min(n_beg + s.nu, epi.live_n_tiles(m_blk));
            const int row0 = m_blk * BM + quarter * 32;
            for (int n_blk = n_beg; n_blk < n_end; n_blk++) {
                tc5::mbar_wait(tfull + acc, acc_phase);
                tc5::tc_fence_after();
                float a0[8], a1[8];
#pragma unroll
                for (int k = 0; k < 8; k++) a0[k] = a1[k] = 0.f;
#pragma unroll 1
                for (int cc = 0; cc < BN / 32 / CSPLIT; cc++) {
                    const int c0 = chalf * (BN / CSPLIT) + cc * 32;
                    if (!epi.chunk_live(row0, n_blk * BN + c0)) continue;      // warp-uniform (same address in every lane)
                    uint32_t r0[32], r1[32];
                    tc5::tmem_ld32_nowait(tmem_base + t_lane + acc * ACC_COLS + c0, r0);
                    tc5::tmem_ld32_nowait(tmem_base + t_lane + acc * ACC_COLS + BN + c0, r1);
                    tc5::tmem_ld_wait();
                    float v0[32], v1[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) { v0[j] = __uint_as_float(r0[j]); v1[j] = __uint_as_float(r1[j]); }
                    epi.accum2(row0 + lane, n_blk * BN + c0, v0, v1, a0, a1);
                }
                tc5::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc5::mbar_arrive(tempty + acc);                 // accumulator drained: the MMAs may go on
                epi.finish(row0 + lane, n_blk, chalf, a0);
                epi.finish(row0 + 128 + lane, n_blk, chalf, a1);
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
        }
    }
    tc5::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc5::tmem_dealloc(tmem_base, TMEM_COLS);
}

// A: [nb * rows_a][128 fp32] split rows, W: [nb * rows_w][128 fp32]; nu = n tiles per unit (A reloaded once per unit)
template <class Epi>
int launch(const void *A, const void *W, int nb, int rows_a, int rows_w, int nu, const Epi &epi, cudaStream_t st)
{
    if (nb <= 0 || rows_a % BM != 0 || rows_w % BN != 0 || nu <= 0) {
        set_last_error("tcgen05 row-statistics gemm: unsupported shape nb=%d rows_a=%d rows_w=%d", nb, rows_a, rows_w);
        return -1;
    }
    CUtensorMap ta, tw;
    const int phys = KB * BK;   // row length in 2-byte units (the tensor maps only move bytes)
    if (!gemm::make_tmap_bf16(&ta, A, phys, (uint64_t)nb * rows_a, phys, BM) ||
        !gemm::make_tmap_bf16(&tw, W, phys, (uint64_t)nb * rows_w, phys, BN)) {
        set_last_error("tcgen05 row-statistics gemm: cuTensorMapEncodeTiled failed");
        return -1;
    }
    static bool attr_done = false;  // per template instantiation
    if (!attr_done) {
        cudaFuncSetAttribute(gemm_rowstats_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        attr_done = true;
    }
    Shape s;
    s.tpb = rows_a / BM;
    s.tiles_m = nb * s.tpb;
    s.tiles_n = rows_w / BN;
    s.nu = nu < s.tiles_n ? nu : s.tiles_n;
    s.units_n = (s.tiles_n + s.nu - 1) / s.nu;
    s.rows_w = rows_w;
    const int nunits = s.tiles_m * s.units_n;
    const int grid = nunits < gemm::num_sms() ? nunits : gemm::num_sms();
    gemm_rowstats_kernel<Epi><<<grid, NTHREADS, SMEM_BYTES, st>>>(ta, tw, s, epi);
    return 0;
}

}  // namespace gemm_rs
}  // namespace scp
