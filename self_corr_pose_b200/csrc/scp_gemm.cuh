// Persistent warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  (bf16 in, fp32
// accumulate in TMEM), with a functor epilogue that receives 32 consecutive columns of one row.
//
//   warp 0      : TMA producer  (cp.async.bulk.tensor, 128-byte swizzled K-major tiles, mbarrier expect_tx)
//   warp 1      : TMEM allocator + MMA issuer (one elected lane: tcgen05.mma cta_group::1, M=128, N=128, K=16;
//                 a CTA tile is 256 x 128 = two M=128 accumulators sharing every B (weight) tile)
//   warps 2..9  : epilogue (tcgen05.ld 32x32b.x32 -> registers -> smem transpose -> functor -> global);
//                 warp w owns TMEM lanes 32*(w%4).. of the M half (w-2)/4 of the tile
// Tile shape: operand traffic per MAC is (BM + BN) / (BM * BN); at 128 x 128 the 148 SMs ask the L2 for ~120 B/clk
// each at full tensor rate against a measured chip-wide L2 cap of ~43 B/clk/SM, i.e. the GEMM is L2-bound at ~35 %
// of the tensor peak.  256 x 128 (A: 32 KiB, B: 16 KiB per 64-deep k block) lowers the demand by 25 % per MAC.
// Three pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty double buffer (MMA <-> epilogue),
// and a static round-robin tile schedule (tile = blockIdx.x + i * gridDim.x), one CTA per SM.
#pragma once
#include <cuda_bf16.h>

#include "scp_common.cuh"
#include "scp_tc5.cuh"

namespace scp {
namespace gemm {

constexpr int BM = 256, BN = 128, BK = 64;          // BK * sizeof(bf16) = 128 B = one swizzle atom row
constexpr int MH = BM / 128;                          // M = 128 accumulators per tile
constexpr int STAGES = 4;
constexpr int ACC_STAGES = 2;
constexpr int EPI_WARPS = 8;                          // two warps per TMEM lane quarter, 64 columns each
constexpr int NTHREADS = 64 + 32 * EPI_WARPS;
constexpr int STAGE_BYTES = (BM + BN) * BK * 2;     // 48 KiB
constexpr int STG_FLOATS = 32 * 33;                   // per-epilogue-warp transpose buffer (padded) ...
constexpr int TMA_TILE_BYTES = 32 * 32 * 4;           // ... or a 32x32 fp32 box in TMA SWIZZLE_128B layout (same region)
constexpr int EPI_BYTES = EPI_WARPS * STG_FLOATS * 4;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 256 /*barriers*/ + 1024 /*align slack*/;
constexpr int ACC_COLS = MH * BN;                    // TMEM columns of one accumulator stage: [M half][BN]
constexpr int TMEM_COLS = ACC_STAGES * ACC_COLS;    // 512

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}

// ---- fp32-class precision on the bf16 tensor cores ("x3" mode, NT = 3) -----------------------------------------
// A value x is carried as TWO bf16 numbers, hi = bf16(x) and lo = bf16(x - hi) (16 mantissa bits together), and a
// product of two such operands as three tensor-core products:  a*w ~= a_hi*w_hi + a_hi*w_lo + a_lo*w_hi  (the dropped
// a_lo*w_lo term and the residual beyond lo are ~2^-16 relative), accumulated in fp32 in TMEM.  This is the mode that
// meets the 1e-3 / 99.9 % arg-max parity contract against the reference's fp32 ViT; plain bf16 (NT = 1) is 5e-3.
//
// Memory layout of a split operand ("i32" layout): row r of a logical [rows][K] matrix is 2K bf16 numbers, groups of
// 32 logical columns stored as [32 hi | 32 lo] (128 bytes).  A 64-element (128-byte) TMA box row therefore holds both
// halves of 32 logical k values: one smem stage carries everything its three products need, the GEMM main loop is
// the plain one over 2K physical columns with six K=16 instructions per stage and accumulator instead of four, and an
// epilogue's 32-column accumulator chunk maps to exactly one 128-byte box row of the split output.
// Split a pair of fp32 values: returns the packed hi pair, writes the packed lo pair.
// The residual x - float(hi) is ONE mixed-precision instruction per element (fma.rn.f32.bf16: hi * (-1) + x, the bf16 half
// selected inside the instruction, FHFMA.BF16 in SASS) instead of unpack + subtract: 4 instructions per pair, was 6; the
// difference is exact in fp32 either way, so the result is bit-identical to the unpack form.
#ifndef SCP_SPLIT_MIXED
#define SCP_SPLIT_MIXED 1
#endif
__device__ __forceinline__ uint32_t split_bf16x2(float a, float b, uint32_t &lo_pair)
{
    const uint32_t h = pack_bf16x2(a, b);
#if SCP_SPLIT_MIXED
    uint16_t h0, h1;
    asm("mov.b32 {%0, %1}, %2;" : "=h"(h0), "=h"(h1) : "r"(h));
    const uint16_t m1 = 0xBF80;   // bf16 -1.0
    float la, lb;
    asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(la) : "h"(h0), "h"(m1), "f"(a));
    asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(lb) : "h"(h1), "h"(m1), "f"(b));
    lo_pair = pack_bf16x2(la, lb);
#else
    lo_pair = pack_bf16x2(a - __uint_as_float(h << 16), b - __uint_as_float(h & 0xffff0000u));
#endif
    return h;
}

struct Shape {
    int M, N, K;
    int a_box_rows;   // rows of the A tensor-map box (BM; 128 for a matrix of at most 128 rows)
    // batched mode (a_idx != nullptr): the M axis is NB stacked problems of rows_per_batch rows; problem p multiplies
    // rows [a_idx[p] * rows_per_batch, +rows_per_batch) of A with rows [w_idx[p] * N, +N) of W (N = s.N columns)
    const long long *a_idx, *w_idx;
    int rows_per_batch;
};

// optional fourth epilogue mode (Epi::kTmaStoreBf16 == true): out[tile] = bf16(epi.apply(acc + bias[col])) computed in
// registers (thread = row, 32 independent columns in flight), packed into a 32 x 64 bf16 box in TMA SWIZZLE_128B
// layout and written with ONE cp.async.bulk.tensor store per warp and tile -- the SM issues no global stores
template <class E, class = void>
struct tma_store_mode { static constexpr bool value = false; };
template <class E>
struct tma_store_mode<E, decltype((void)E::kTmaStoreBf16)> { static constexpr bool value = E::kTmaStoreBf16; };

// Epi: struct with
//   static constexpr bool kStaged;
//   kStaged == true : __device__ void chunk(int row0, int nrows, int col, const float *stg, int lane) const;
//                     called per 32x32 accumulator chunk after it has been transposed through a padded
//                     shared-memory buffer: value of (row0 + r, col) is stg[r * 33 + lane], lane = column, so a
//                     warp touches 32 consecutive columns of ONE row per instruction (coalesced)
//   kStaged == false: __device__ void operator()(int row, int col0, const float (&acc)[32]) const;  lane = row
//                     (for outputs that are contiguous along the rows, e.g. the transposed feature store)
//   static constexpr bool kTmaReduceAdd (optional third mode, takes precedence): C[tile] += acc + bias[col] through
//                     cp.reduce.async.bulk.tensor (.add, fp32) from a swizzled smem box -- the read-modify-write of
//                     the residual stream happens at L2, the SM issues no loads; needs `const float *bias` member
template <class Epi, int NT = 1>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_c, Shape s, Epi epi)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *epi_buf = smem + STAGES * STAGE_BYTES;   // 1024-aligned
    uint64_t *bars = reinterpret_cast<uint64_t *>(epi_buf + EPI_BYTES);
    uint64_t *full = bars, *empty = bars + STAGES, *tfull = bars + 2 * STAGES, *tempty = bars + 2 * STAGES + ACC_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 2 * ACC_STAGES);
    float *stage_buf = reinterpret_cast<float *>(epi_buf);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (s.M + BM - 1) / BM, tiles_n = s.N / BN, ntiles = tiles_m * tiles_n, nkb = s.K / BK;

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

    if (warp == 0) {
        // ===== TMA producer =====
        if (tc5::elect_one()) {
            int stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;   // n fastest: A tile reused from L2
                int a_row = m_blk * BM, w_row = n_blk * BN;
                if (s.a_idx != nullptr) {
                    const int tpb = s.rows_per_batch / BM, p = m_blk / tpb;
                    a_row = (int)s.a_idx[p] * s.rows_per_batch + (m_blk - p * tpb) * BM;
                    w_row += (int)s.w_idx[p] * s.N;
                }
                for (int kb = 0; kb < nkb; kb++) {
                    tc5::mbar_wait(empty + stage, phase ^ 1);
                    uint8_t *sa = smem + stage * STAGE_BYTES, *sb = sa + BM * BK * 2;
                    tc5::mbar_expect_tx(full + stage, (s.a_box_rows + BN) * BK * 2);
                    tc5::tma_load_2d(sa, &tmap_a, full + stage, kb * BK, a_row);
                    tc5::tma_load_2d(sb, &tmap_w, full + stage, kb * BK, w_row);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====   (elect.sync, not lane == 0: ptxas then knows a single thread runs the loop and
        // emits the tcgen05.mma sequence without a per-instruction uniformisation loop -- see scp_tc5.cuh)
        if (tc5::elect_one()) {
            constexpr uint32_t idesc = tc5::umma_idesc_bf16(128, BN);
            int stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                tc5::mbar_wait(tempty + acc, acc_phase ^ 1);       // epilogue has drained this accumulator
                tc5::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
                for (int kb = 0; kb < nkb; kb++) {
                    tc5::mbar_wait(full + stage, phase);           // TMA bytes have landed
                    tc5::tc_fence_after();
                    const uint32_t sa = tc5::smem_u32(smem + stage * STAGE_BYTES), sb = sa + BM * BK * 2;
#pragma unroll
                    for (int mh = 0; mh < MH; mh++) {
                        if constexpr (NT == 3) {
                            // split operands: a 128-byte tile row = [32 hi | 32 lo] of 32 logical k values; per K=16
                            // chunk c: hi*hi, hi*lo, lo*hi (byte offsets c*32 for hi, 64 + c*32 for lo)
#pragma unroll
                            for (int c = 0; c < 2; c++) {
                                const uint32_t ah = sa + mh * (128 * BK * 2) + c * 32, bh = sb + c * 32;
                                tc5::umma_bf16(d_tmem + mh * BN, tc5::umma_desc_sw128(ah), tc5::umma_desc_sw128(bh), idesc,
                                               (kb | c) != 0);
                                tc5::umma_bf16(d_tmem + mh * BN, tc5::umma_desc_sw128(ah), tc5::umma_desc_sw128(bh + 64), idesc, 1);
                                tc5::umma_bf16(d_tmem + mh * BN, tc5::umma_desc_sw128(ah + 64), tc5::umma_desc_sw128(bh), idesc, 1);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < BK / 16; k++) {
                                // advance 16 elements (32 B) along K inside the 128-byte swizzle atom
                                tc5::umma_bf16(d_tmem + mh * BN, tc5::umma_desc_sw128(sa + mh * (128 * BK * 2) + k * 32),
                                               tc5::umma_desc_sw128(sb + k * 32), idesc, (kb | k) != 0);
                            }
                        }
                    }
                    tc5::umma_commit(empty + stage);               // frees the smem slot when the MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc5::umma_commit(tfull + acc);                     // accumulator complete -> epilogue
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue warps: TMEM lane quarter = warp % 4, M half of the tile = (warp - 2) / 4 =====
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const uint32_t t_acc = ((uint32_t)(quarter * 32) << 16) + half * BN;
        float *stg = stage_buf + (warp - 2) * STG_FLOATS;
        int acc = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;
            tc5::mbar_wait(tfull + acc, acc_phase);
            tc5::tc_fence_after();
            const int row0 = m_blk * BM + half * 128 + quarter * 32;
            if constexpr (tma_store_mode<Epi>::value && NT == 3) {
              // split output: one 32-column accumulator chunk -> one box row [32 hi | 32 lo] (128 B); physical column 2c
              uint8_t *box = epi_buf + (warp - 2) * TMA_TILE_BYTES;
#pragma unroll 1
              for (int c0 = 0; c0 < BN; c0 += 32) {
                float v[32];
                tc5::tmem_ld32(tmem_base + t_acc + acc * ACC_COLS + c0, v);
                if constexpr (Epi::kMixed) {
                    if (epi.direct(n_blk * BN + c0)) {
                        if (row0 + lane < s.M) epi(row0 + lane, n_blk * BN + c0, v);
                        continue;
                    }
                }
                if (lane == 0) tc5::tma_store_wait_read();
                __syncwarp();
                const float4 *b4 = reinterpret_cast<const float4 *>(epi.bias + n_blk * BN + c0);
#pragma unroll
                for (int j = 0; j < 4; j++) {                                   // 8 columns -> one 16-byte chunk hi + one lo
                    const float4 ba = __ldg(b4 + 2 * j), bb = __ldg(b4 + 2 * j + 1);
                    uint32_t h[4], l[4];
                    h[0] = split_bf16x2(epi.apply(v[8 * j + 0] + ba.x), epi.apply(v[8 * j + 1] + ba.y), l[0]);
                    h[1] = split_bf16x2(epi.apply(v[8 * j + 2] + ba.z), epi.apply(v[8 * j + 3] + ba.w), l[1]);
                    h[2] = split_bf16x2(epi.apply(v[8 * j + 4] + bb.x), epi.apply(v[8 * j + 5] + bb.y), l[2]);
                    h[3] = split_bf16x2(epi.apply(v[8 * j + 6] + bb.z), epi.apply(v[8 * j + 7] + bb.w), l[3]);
                    *reinterpret_cast<uint4 *>(box + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<uint4 *>(box + lane * 128 + (((4 + j) ^ (lane & 7)) << 4)) = make_uint4(l[0], l[1], l[2], l[3]);
                }
                tc5::fence_proxy_async();
                __syncwarp();
                if (lane == 0 && row0 < s.M) {   // rows past M are clipped by the tensor map
                    tc5::tma_store_2d(&tmap_c, box, 2 * (n_blk * BN + c0), row0);
                    tc5::tma_store_commit();
                }
              }
            } else if constexpr (tma_store_mode<Epi>::value) {
              uint8_t *box = epi_buf + (warp - 2) * TMA_TILE_BYTES;            // 32 rows x 128 B (64 bf16), SWIZZLE_128B
#pragma unroll 1
              for (int cg = 0; cg < BN; cg += 64) {
                if constexpr (Epi::kMixed) {   // column ranges that are contiguous along the rows go through operator()
                    if (epi.direct(n_blk * BN + cg)) {
#pragma unroll 1
                        for (int cc = 0; cc < 64; cc += 32) {
                            float v[32];
                            tc5::tmem_ld32(tmem_base + t_acc + acc * ACC_COLS + cg + cc, v);
                            if (row0 + lane < s.M) epi(row0 + lane, n_blk * BN + cg + cc, v);
                        }
                        continue;
                    }
                }
                if (lane == 0) tc5::tma_store_wait_read();                      // previous bulk store has read the box
                __syncwarp();
#pragma unroll
                for (int cc = 0; cc < 64; cc += 32) {
                    const int c0 = cg + cc;
                    float v[32];
                    tc5::tmem_ld32(tmem_base + t_acc + acc * ACC_COLS + c0, v);
                    const float4 *b4 = reinterpret_cast<const float4 *>(epi.bias + n_blk * BN + c0);
#pragma unroll
                    for (int j = 0; j < 4; j++) {                               // 8 columns -> one 16-byte chunk
                        const float4 ba = __ldg(b4 + 2 * j), bb = __ldg(b4 + 2 * j + 1);
                        uint32_t w[4];
                        w[0] = pack_bf16x2(epi.apply(v[8 * j + 0] + ba.x), epi.apply(v[8 * j + 1] + ba.y));
                        w[1] = pack_bf16x2(epi.apply(v[8 * j + 2] + ba.z), epi.apply(v[8 * j + 3] + ba.w));
                        w[2] = pack_bf16x2(epi.apply(v[8 * j + 4] + bb.x), epi.apply(v[8 * j + 5] + bb.y));
                        w[3] = pack_bf16x2(epi.apply(v[8 * j + 6] + bb.z), epi.apply(v[8 * j + 7] + bb.w));
                        const int chunk = (cc >> 3) + j;
                        *reinterpret_cast<uint4 *>(box + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
                tc5::fence_proxy_async();
                __syncwarp();
                if (lane == 0 && row0 < s.M) {   // rows past M are clipped by the tensor map
                    tc5::tma_store_2d(&tmap_c, box, n_blk * BN + cg, row0);
                    tc5::tma_store_commit();
                }
              }
            } else
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                float v[32];
                tc5::tmem_ld32(tmem_base + t_acc + acc * ACC_COLS + c0, v);
                if constexpr (Epi::kTmaReduceAdd) {
                    uint8_t *box = epi_buf + (warp - 2) * TMA_TILE_BYTES;      // 32 rows x 128 B, SWIZZLE_128B
                    if (lane == 0) tc5::tma_store_wait_read();                  // previous bulk store has read the box
                    __syncwarp();
                    const float4 *b4 = reinterpret_cast<const float4 *>(epi.bias + n_blk * BN + c0);
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float4 bb = __ldg(b4 + j);
                        const float4 o = make_float4(v[4 * j] + bb.x, v[4 * j + 1] + bb.y, v[4 * j + 2] + bb.z, v[4 * j + 3] + bb.w);
                        *reinterpret_cast<float4 *>(box + lane * 128 + ((j ^ (lane & 7)) << 4)) = o;
                    }
                    tc5::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0 && row0 < s.M) {   // rows past M are clipped by the tensor map
                        tc5::tma_reduce_add_2d(&tmap_c, box, n_blk * BN + c0, row0);
                        tc5::tma_store_commit();
                    }
                } else if constexpr (Epi::kStaged) {
                    if constexpr (Epi::kMixed) {   // some column ranges are contiguous along the rows instead
                        if (epi.direct(n_blk * BN + c0)) {
                            if (row0 + lane < s.M) epi(row0 + lane, n_blk * BN + c0, v);
                            continue;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 32; j++) stg[lane * 33 + j] = v[j];
                    __syncwarp();
                    const int nrows = min(32, s.M - row0);
                    if (nrows > 0) epi.chunk(row0, nrows, n_blk * BN + c0 + lane, stg, lane);
                    __syncwarp();
                } else {
                    if (row0 + lane < s.M) epi(row0 + lane, n_blk * BN + c0, v);
                }
            }
            tc5::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc5::mbar_arrive(tempty + acc);
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
        if constexpr (Epi::kTmaReduceAdd || tma_store_mode<Epi>::value) {
            if (lane == 0) tc5::tma_store_wait_all();   // all bulk stores / reduce-adds of this warp have completed
        }
    }
    tc5::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc5::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// row-major bf16 matrix [rows][inner] (inner contiguous, row pitch `pitch_elems`), box = 64 x 128, SWIZZLE_128B
inline bool make_tmap_bf16(CUtensorMap *m, const void *ptr, uint64_t inner, uint64_t rows, uint64_t pitch_elems,
                           uint32_t box_rows = BM)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = { inner, rows };
    const cuuint64_t strides[1] = { pitch_elems * 2 };
    const cuuint32_t box[2] = { (cuuint32_t)BK, (cuuint32_t)box_rows };
    const cuuint32_t estr[2] = { 1, 1 };
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// row-major fp32 matrix [rows][cols] (pitch in elements), box = 32 x 32, SWIZZLE_128B (reduce-add epilogue target)
inline bool make_tmap_f32_box32(CUtensorMap *m, const void *ptr, uint64_t cols, uint64_t rows, uint64_t pitch_elems)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = { cols, rows };
    const cuuint64_t strides[1] = { pitch_elems * 4 };
    const cuuint32_t box[2] = { 32, 32 };
    const cuuint32_t estr[2] = { 1, 1 };
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(ptr), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline int num_sms()
{
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

// A[M,K] (pitch lda), W[N,K] (pitch ldw); N % 128 == 0, K % 64 == 0
// c_out / ldc: fp32 output matrix of the kTmaReduceAdd epilogue / bf16 output of the kTmaStoreBf16 epilogue (ignored otherwise)
// NT = 3: A and W are split operands in the i32 layout (see above): K is the LOGICAL depth (K % 32 == 0), lda / ldw /
// ldc are PHYSICAL pitches in bf16 elements (>= 2K); a bf16 output (kTmaStoreBf16) is written split as well.
template <class Epi, int NT = 1>
int launch(const void *A, int lda, const void *W, int ldw, int M, int N, int K, const Epi &epi, cudaStream_t st,
           void *c_out = nullptr, int ldc = 0, const long long *a_idx = nullptr, const long long *w_idx = nullptr,
           int rows_per_batch = 0, long a_rows_total = 0, long w_rows_total = 0)
{
    // batched mode: M = NB * rows_per_batch virtual rows; A has a_rows_total rows, W has w_rows_total rows
    const bool batched = a_idx != nullptr;
    if (batched && (rows_per_batch % BM != 0 || M % rows_per_batch != 0 || !w_idx)) {
        set_last_error("tcgen05 gemm: batched mode needs rows_per_batch %% %d == 0", BM);
        return -1;
    }
    if (M <= 0 || N % BN != 0 || (NT == 3 ? K % 32 : K % BK) != 0) {
        set_last_error("tcgen05 gemm: unsupported shape M=%d N=%d K=%d", M, N, K);
        return -1;
    }
    if (NT == 3) K *= 2;   // physical columns of the split operands
    CUtensorMap ta, tw;
    int a_box = BM;
    if (!make_tmap_bf16(&ta, A, K, batched ? a_rows_total : M, lda, BM)) {
        a_box = 128;   // a matrix of at most 128 rows: only the first M = 128 accumulator holds stored rows
        if (M > 128 || !make_tmap_bf16(&ta, A, K, M, lda, 128)) {
            set_last_error("tcgen05 gemm: cuTensorMapEncodeTiled failed (A)");
            return -1;
        }
    }
    if (!make_tmap_bf16(&tw, W, K, batched ? w_rows_total : N, ldw, BN)) {
        set_last_error("tcgen05 gemm: cuTensorMapEncodeTiled failed (W)");
        return -1;
    }
    CUtensorMap tc = ta;
    if constexpr (Epi::kTmaReduceAdd) {
        if (!c_out || !make_tmap_f32_box32(&tc, c_out, N, M, ldc)) {
            set_last_error("tcgen05 gemm: output tensor map failed");
            return -1;
        }
    }
    if constexpr (tma_store_mode<Epi>::value) {   // bf16 output [M][N], box = 64 columns x 32 rows
        // (an epilogue with direct column ranges may store fewer than N columns through TMA: the matrix is ldc wide)
        const int out_cols = NT == 3 ? 2 * N : N;
        if (!c_out || !make_tmap_bf16(&tc, c_out, out_cols < ldc ? out_cols : ldc, M, ldc, 32)) {
            set_last_error("tcgen05 gemm: bf16 output tensor map failed");
            return -1;
        }
    }
    static bool attr_done = false;  // per template instantiation
    if (!attr_done) {
        cudaFuncSetAttribute(gemm_bf16_tn_kernel<Epi, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        attr_done = true;
    }
    const int ntiles = ((M + BM - 1) / BM) * (N / BN);
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    Shape s{ M, N, K, a_box, a_idx, w_idx, rows_per_batch };
    gemm_bf16_tn_kernel<Epi, NT><<<grid, NTHREADS, SMEM_BYTES, st>>>(ta, tw, tc, s, epi);
    return 0;
}

}  // namespace gemm
}  // namespace scp
