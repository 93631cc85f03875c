// Correspondence.match forward on the 5th-generation tensor cores (model/module/correspondence.py:42-53 of the
// reference; training path, i.e. no full-resolution pointcorr output).  Same results as corr_fwd_kernel of scp_corr.cu
// (the mma.sync version, which stays the evaluation-mode / odd-shape path), different machine mapping:
//
//   * the similarity S = img_feat^T mesh_feat^T is formed by the persistent tcgen05 GEMM of scp_gemm.cuh in its NT = 4
//     mode: both operands are fp32 rows split into TF32 (hi, lo) pairs, three kind::tf32 products per logical product
//     (hi*hi + hi*lo + lo*hi: the accuracy of the 3-term mma.sync split it replaces), accumulator in tensor memory;
//   * the two soft-maxes are the GEMM's epilogue (kRowStats mode): the epilogue warps read the accumulator from TMEM
//     (thread = row), exponentiate once per element against the fixed reference point tau * 1 and fold
//     sum_col e * w_k[col] into per-row registers -- S is never written;
//   * a row soft-max (over vertices) and a column soft-max (over pixels) need reductions along different axes of S.  A
//     thread owns a ROW of the accumulator, so instead of folding columns across lanes (3 shuffles per element) the
//     product is formed twice, once as S (rows = pixels: rsum, match) and once as S^T (rows = vertices: csum, imatch,
//     and the 2x2-pooled similarity + its column soft-max for the pre-training cycle loss).  The tensor pipe is idle
//     otherwise; both passes are the same kernel with "weights" (1, v) per vertex or mask * (1, grid) per pixel;
//   * pixels are stored block-major (the four pixels of a 2x2 block are consecutive rows), so the 2x2 mean of S^T is a
//     sum over four consecutive accumulator columns inside one thread, and its store is coalesced along the vertices;
//   * background: per 32 consecutive (block-major) pixels one liveness word; an epilogue warp skips chunks without a
//     foreground pixel (pass S: its 32 rows; pass S^T: the 32 columns), corr_fill_kernel writes their constants.
#include <stdlib.h>

#include "../../include/scp_b200.h"
#include "scp_common.cuh"
#ifndef SCP_HOST_EMU
#include "scp_gemm_rs.cuh"
#else
// Host emulation (tools/emu, tests only): the tcgen05 GEMM is replaced by its plain statement -- the same unit / tile
// walk, live-tile clipping and epilogue calls (row0 / col0 / lane conventions of gemm_rowstats_kernel), the accumulator
// formed from the split operands exactly as the issue loop combines them (hi*hi + hi*lo + lo*hi per [16 hi | 16 lo]
// group).  Everything else of this file (operand preparation, epilogue functors, combining kernels, workspace layout) is
// compiled unchanged.
namespace scp {
namespace gemm_rs {
constexpr int BM = 256, BN = 128;
#ifndef SCP_RS_EPI_WARPS
#define SCP_RS_EPI_WARPS 8
#endif
constexpr int CSPLIT = SCP_RS_EPI_WARPS / 4;
template <class Epi>
int launch(const void *A, const void *W, int nb, int rows_a, int rows_w, int nu, const Epi &epi, cudaStream_t)
{
    const float *Af = (const float *)A, *Wf = (const float *)W;
    const int tpb = rows_a / BM, tiles_m = nb * tpb, tiles_n = rows_w / BN;
    if (nu > tiles_n) nu = tiles_n;
    const int units_n = (tiles_n + nu - 1) / nu;
    for (int u = 0; u < tiles_m * units_n; u++) {
        const int nr = u / tiles_m, r = u - nr * tiles_m, pos = r / nb, b = r - pos * nb;   // decode_unit
        const int m_blk = b * tpb + pos, n_beg = nr * nu;
        const int lim = epi.live_n_tiles(m_blk), n_end = n_beg + nu < lim ? n_beg + nu : lim;
        const int w_base = (m_blk / tpb) * rows_w;
        for (int n_blk = n_beg; n_blk < n_end; n_blk++)
            for (int sub = 0; sub < 4 * CSPLIT; sub++) {  // epilogue warp: column range sub / 4, TMEM lane quarter sub % 4
                const int row0 = m_blk * BM + (sub & 3) * 32, chalf = sub >> 2;
                for (int lane = 0; lane < 32; lane++) {
                    float a0[8] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f }, a1[8] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
                    for (int cc = 0; cc < BN / 32 / CSPLIT; cc++) {
                        const int c0 = chalf * (BN / CSPLIT) + cc * 32;
                        if (!epi.chunk_live(row0, n_blk * BN + c0)) continue;
                        float v[2][32];
                        for (int h = 0; h < 2; h++) {
                            const float *ar = Af + (size_t)(row0 + 128 * h + lane) * 128;
                            for (int j = 0; j < 32; j++) {
                                const float *wr = Wf + (size_t)(w_base + n_blk * BN + c0 + j) * 128;
                                float acc = 0.f;
                                for (int g = 0; g < 4; g++)
                                    for (int k = 0; k < 16; k++) {
                                        const float ah = ar[32 * g + k], al = ar[32 * g + 16 + k];
                                        const float wh = wr[32 * g + k], wl = wr[32 * g + 16 + k];
                                        acc += ah * wh + ah * wl + al * wh;
                                    }
                                v[h][j] = acc;
                            }
                        }
                        epi.accum2(row0 + lane, n_blk * BN + c0, v[0], v[1], a0, a1);
                    }
                    epi.finish(row0 + lane, n_blk, chalf, a0);
                    epi.finish(row0 + 128 + lane, n_blk, chalf, a1);
                }
            }
    }
    return 0;
}
}  // namespace gemm_rs
}  // namespace scp
#endif

namespace scp {
namespace corr_tc {

constexpr int C = 64;            // feature channels
constexpr int ROWF = 2 * C;      // floats per split operand row: 4 groups of [16 hi | 16 lo]
constexpr int NT = 256;
constexpr int TM = gemm_rs::BM, TN = gemm_rs::BN;   // GEMM tile: 256 rows x 128 columns
constexpr int CS = gemm_rs::CSPLIT;                  // partial slots per row and tile (column ranges of the epilogue warps)
constexpr float LOG2E = 1.4426950408889634f;
#ifndef SCP_CORR_SLIM
#define SCP_CORR_SLIM 1       // partial loads of the per-column weight records (0.466 -> 0.454 ms at B = 64, profiles/r2_corr_tcgen05.md)
#endif

__device__ __forceinline__ float tf32_rna(float x)
{
#ifdef SCP_HOST_EMU
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & ~0x1fffu);   // cvt.rna: nearest, ties away, 10 mantissa bits
#else
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
#endif
}

// Pixel rows are COMPACTED and block-major: row i of image b is sub-pixel i % 4 (= 2 * dy + dx) of the (i / 4)-th 2x2
// block of the image's foreground block list (corr_blocklist_kernel: blk[0] = count, blk[1..] = pooled-pixel ids,
// ascending); rows past 4 * count up to the next multiple of 256 are zero.  -1 past the list.
__device__ __forceinline__ int row_pixel(const int *__restrict__ blk, int i, int wf)
{
    if ((i >> 2) >= blk[0]) return -1;
    const int id = blk[1 + (i >> 2)], sub = i & 3, w2 = wf >> 1;
    const int by = id / w2, bx = id - by * w2;
    return (2 * by + (sub >> 1)) * wf + 2 * bx + (sub & 1);
}

// ---- operand preparation ----------------------------------------------------------------------------------------
// img_feat[b][c][p] -> a_img[b][i][128] (compacted block-major rows, split), wc[b][i] = mask * (gx, gy, 1, 0),
// wp[b][i / 4] = (pooled gx, pooled gy, pooled-pixel id as int bits, 1), live[b][i / 32] = bit mask of the foreground
// rows.  CTA = 64 rows; CTAs past the image's last (256-row) tile exit.
__global__ void __launch_bounds__(NT) prep_img_kernel(int P, int wf, const float *__restrict__ img_feat,
                                                      const float *__restrict__ mask_down,
                                                      const float *__restrict__ meshgrid, const int *__restrict__ blocks,
                                                      float *__restrict__ a_img, float4 *__restrict__ wc,
                                                      float4 *__restrict__ wp, uint32_t *__restrict__ live)
{
    __shared__ float s[C][65];
    __shared__ float s_g[2][64];
    __shared__ int s_pix[64];
    const int b = blockIdx.y, i0 = blockIdx.x * 64, tid = threadIdx.x;
    const int *blk = blocks + (size_t)b * ((P >> 2) + 1);
    const int rows = 4 * blk[0];
    if (i0 >= (rows + TM - 1) / TM * TM) return;
    const float *img_b = img_feat + (size_t)b * C * P;
    if (tid < 64) {
        const int p = row_pixel(blk, i0 + tid, wf);
        const bool fg = p >= 0 && mask_down[(size_t)b * P + p] != 0.f;
        const float gx = p >= 0 ? meshgrid[p] : 0.f, gy = p >= 0 ? meshgrid[P + p] : 0.f;
        s_pix[tid] = p;
        s_g[0][tid] = gx;
        s_g[1][tid] = gy;
        wc[(size_t)b * P + i0 + tid] = fg ? make_float4(gx, gy, 1.f, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t bal = __ballot_sync(0xffffffffu, fg);
        if ((tid & 31) == 0) live[((size_t)b * P + i0 + tid) >> 5] = bal;
    }
    __syncthreads();
    for (int idx = tid; idx < C * 64; idx += NT) {
        const int c = idx >> 6, r = idx & 63, p = s_pix[r];
        s[c][r] = p >= 0 ? img_b[(size_t)c * P + p] : 0.f;
    }
    if (tid < 16) {   // bilinear 1/2 of the meshgrid = 2x2 mean: (top + bottom) of each column, then the two columns
        const int k = 4 * tid, li = (i0 >> 2) + tid;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (li < blk[0]) {
            o.x = 0.25f * ((s_g[0][k] + s_g[0][k + 2]) + (s_g[0][k + 1] + s_g[0][k + 3]));
            o.y = 0.25f * ((s_g[1][k] + s_g[1][k + 2]) + (s_g[1][k + 1] + s_g[1][k + 3]));
            o.z = __int_as_float(blk[1 + li]);
            o.w = 1.f;
        }
        wp[(size_t)b * (P >> 2) + li] = o;
    }
    __syncthreads();
    float4 *out = reinterpret_cast<float4 *>(a_img + ((size_t)b * P + i0) * ROWF);
    for (int idx = tid; idx < 64 * (ROWF / 4); idx += NT) {
        const int r = idx >> 5, q = idx & 31;                       // row, float4 of the 128-float row
        const int g = q >> 3, lo = (q >> 2) & 1, c = 16 * g + 4 * (q & 3);
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float x = s[c + k][r], hi = tf32_rna(x);
            v[k] = lo ? tf32_rna(x - hi) : hi;
        }
        out[idx] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// mesh_feat[b][n][c] -> a_mesh[b][n][128] (split; rows N..Npad-1 zero), wr[b][n] = (v, 1) (zero past N)
__global__ void __launch_bounds__(NT) prep_mesh_kernel(int N, int Npad, const float *__restrict__ mesh_feat,
                                                       const float *__restrict__ pred_v, float *__restrict__ a_mesh,
                                                       float4 *__restrict__ wr)
{
    const int b = blockIdx.y, idx = blockIdx.x * NT + threadIdx.x;
    const int n = idx >> 5, q = idx & 31;
    if (n >= Npad) return;
    const int g = q >> 3, lo = (q >> 2) & 1, c = 16 * g + 4 * (q & 3);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) {
        const float4 x = *reinterpret_cast<const float4 *>(mesh_feat + ((size_t)b * N + n) * C + c);
        const float h0 = tf32_rna(x.x), h1 = tf32_rna(x.y), h2 = tf32_rna(x.z), h3 = tf32_rna(x.w);
        o = lo ? make_float4(tf32_rna(x.x - h0), tf32_rna(x.y - h1), tf32_rna(x.z - h2), tf32_rna(x.w - h3))
               : make_float4(h0, h1, h2, h3);
    }
    reinterpret_cast<float4 *>(a_mesh + ((size_t)b * Npad + n) * ROWF)[q] = o;
    if (q == 0) {
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < N) {
            const float *v = pred_v + ((size_t)b * N + n) * 3;
            w = make_float4(v[0], v[1], v[2], 1.f);
        }
        wr[(size_t)b * Npad + n] = w;
    }
}

// ---- epilogues ---------------------------------------------------------------------------------------------------
// pass S (rows = compacted pixels, columns = vertices): a[0..3] += e * (1, v.x, v.y, v.z); wr holds (v, valid)
struct EpiRows {
    const float4 *wr;        // [B][Npad]
    const uint32_t *live;    // [B * P / 32]
    const int *blocks;       // [B][P / 4 + 1]
    float4 *part;            // [CS * Npad / 128][B * P]: per n tile and column range
    int P, Npad, M, nreal;   // nreal = N: columns past it are padding
    float kexp;
    // an m tile is visited when its first row lies inside the image's compacted list
    __device__ int live_n_tiles(int m_blk) const
    {
        const int tpb = P / TM, b = m_blk / tpb;
        return (m_blk - b * tpb) * TM < 4 * blocks[(size_t)b * ((P >> 2) + 1)] ? Npad / TN : 0;
    }
    __device__ bool chunk_live(int row0, int) const { return (live[row0 >> 5] | live[(row0 + 128) >> 5]) != 0u; }
    __device__ void accum2(int row, int col0, const float (&v0)[32], const float (&v1)[32], float (&a0)[8],
                           float (&a1)[8]) const
    {
        const float4 *w = wr + (size_t)(row / P) * Npad + col0;
#if SCP_CORR_SLIM
        // chunks that hold only real vertices (all but the image's last): 12 of the record's 16 bytes are loaded (three
        // L1 write-back cycles per warp instead of four; the loads share the data port with the tensor core's operand reads)
        if (col0 + 32 <= nreal) {
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const float2 q = __ldg(reinterpret_cast<const float2 *>(w + j) + 0);
                const float qz = __ldg(reinterpret_cast<const float *>(w + j) + 2);
                const float e0 = ex2_approx(fmaf(v0[j], kexp, -kexp)), e1 = ex2_approx(fmaf(v1[j], kexp, -kexp));
                a0[0] += e0;
                a0[1] = fmaf(e0, q.x, a0[1]);
                a0[2] = fmaf(e0, q.y, a0[2]);
                a0[3] = fmaf(e0, qz, a0[3]);
                a1[0] += e1;
                a1[1] = fmaf(e1, q.x, a1[1]);
                a1[2] = fmaf(e1, q.y, a1[2]);
                a1[3] = fmaf(e1, qz, a1[3]);
            }
            return;
        }
#endif
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const float4 q = __ldg(w + j);
            const float e0 = ex2_approx(fmaf(v0[j], kexp, -kexp)), e1 = ex2_approx(fmaf(v1[j], kexp, -kexp));
            a0[0] = fmaf(e0, q.w, a0[0]);
            a0[1] = fmaf(e0, q.x, a0[1]);
            a0[2] = fmaf(e0, q.y, a0[2]);
            a0[3] = fmaf(e0, q.z, a0[3]);
            a1[0] = fmaf(e1, q.w, a1[0]);
            a1[1] = fmaf(e1, q.x, a1[1]);
            a1[2] = fmaf(e1, q.y, a1[2]);
            a1[3] = fmaf(e1, q.z, a1[3]);
        }
    }
    __device__ void finish(int row, int n_blk, int chalf, const float (&a)[8]) const
    {
        part[(size_t)(CS * n_blk + chalf) * M + row] = make_float4(a[0], a[1], a[2], a[3]);
    }
};

// pass S^T (rows = vertices, columns = compacted pixels): a[0..2] += e * mask * (1, gx, gy); with POOL the 2x2 mean of
// the masked similarity is written (coalesced along the vertices) and a[4..6] += e_pool * (1, pooled gx, pooled gy)
template <bool POOL>
struct EpiCols {
    const float4 *wc;        // [B][P]
    const float4 *wp;        // [B][P / 4]
    const uint32_t *live;    // [B][P / 32]
    const int *blocks;       // [B][P / 4 + 1]
    float4 *part, *part_pool;   // [CS * P / 128][B * Npad]: per n tile and column range
    float *pc_pool;          // [B][P / 4][N]
    int P, N, Npad, M;
    float kexp;
    // the n tiles that hold listed pixels of the image
    __device__ int live_n_tiles(int m_blk) const
    {
        const int b = m_blk / (Npad / TM);
        return (4 * blocks[(size_t)b * ((P >> 2) + 1)] + TN - 1) / TN;
    }
    __device__ bool chunk_live(int row0, int col0) const { return live[(size_t)(row0 / Npad) * (P >> 5) + (col0 >> 5)] != 0u; }
    __device__ void accum2(int row, int col0, const float (&v0)[32], const float (&v1)[32], float (&a0)[8],
                           float (&a1)[8]) const
    {
        const int b = row / Npad, n = row - b * Npad;     // rows n and n + 128 of the same image (Npad % 256 == 0)
        const float4 *w = wc + (size_t)b * P + col0;
#if SCP_CORR_SLIM
        const uint32_t bits = live[(size_t)b * (P >> 5) + (col0 >> 5)];   // foreground flags of the chunk's 32 pixels
#endif
#pragma unroll
        for (int k = 0; k < 8; k++) {
            float s0[4], s1[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const float x0 = v0[4 * k + t], x1 = v1[4 * k + t];
#if SCP_CORR_SLIM
                // 8 of the record's 16 bytes (grid x, y); the mask comes from the warp-uniform liveness word
                const float2 q = __ldg(reinterpret_cast<const float2 *>(w + 4 * k + t) + 0);
                const bool on = (bits >> (4 * k + t)) & 1u;
                const float e0 = on ? ex2_approx(fmaf(x0, kexp, -kexp)) : 0.f, e1 = on ? ex2_approx(fmaf(x1, kexp, -kexp)) : 0.f;
                a0[0] += e0;
                a0[1] = fmaf(e0, q.x, a0[1]);
                a0[2] = fmaf(e0, q.y, a0[2]);
                a1[0] += e1;
                a1[1] = fmaf(e1, q.x, a1[1]);
                a1[2] = fmaf(e1, q.y, a1[2]);
#else
                const float4 q = __ldg(w + 4 * k + t);
                const bool on = q.z != 0.f;
                const float e0 = ex2_approx(fmaf(x0, kexp, -kexp)), e1 = ex2_approx(fmaf(x1, kexp, -kexp));
                a0[0] = fmaf(e0, q.z, a0[0]);
                a0[1] = fmaf(e0, q.x, a0[1]);
                a0[2] = fmaf(e0, q.y, a0[2]);
                a1[0] = fmaf(e1, q.z, a1[0]);
                a1[1] = fmaf(e1, q.x, a1[1]);
                a1[2] = fmaf(e1, q.y, a1[2]);
#endif
                s0[t] = on ? x0 : -1e5f;
                s1[t] = on ? x1 : -1e5f;
            }
            if constexpr (POOL) {
                const float4 g = __ldg(wp + (size_t)b * (P >> 2) + (col0 >> 2) + k);
                // (top + bottom) per column, then the two columns: the summation order of corr_fwd_kernel
                const float p0 = 0.25f * ((s0[0] + s0[2]) + (s0[1] + s0[3])), p1 = 0.25f * ((s1[0] + s1[2]) + (s1[1] + s1[3]));
                if (g.w != 0.f) {
                    float *dst = pc_pool + ((size_t)b * (P >> 2) + __float_as_int(g.z)) * N;
                    if (n < N) dst[n] = p0;
                    if (n + 128 < N) dst[n + 128] = p1;
                }
                const float f0 = ex2_approx(fmaf(p0, kexp, -kexp)), f1 = ex2_approx(fmaf(p1, kexp, -kexp));
                a0[4] += f0;                                 // blocks with a background pixel underflow to 0
                a0[5] = fmaf(f0, g.x, a0[5]);
                a0[6] = fmaf(f0, g.y, a0[6]);
                a1[4] += f1;
                a1[5] = fmaf(f1, g.x, a1[5]);
                a1[6] = fmaf(f1, g.y, a1[6]);
            }
        }
    }
    __device__ void finish(int row, int n_blk, int chalf, const float (&a)[8]) const
    {
        part[(size_t)(CS * n_blk + chalf) * M + row] = make_float4(a[0], a[1], a[2], 0.f);
        if constexpr (POOL) part_pool[(size_t)(CS * n_blk + chalf) * M + row] = make_float4(a[4], a[5], a[6], 0.f);
    }
};

// ---- combining the per-tile partials -----------------------------------------------------------------------------
// rows: match[b][p] = sum_n Pi v, rsum for the listed pixels; background pixels inside listed blocks: uniform soft-max
// -> mean vertex, rsum = N (correspondence.py:44,48); pixels of unlisted blocks are written by corr_fill_kernel
__global__ void finish_rows_kernel(int P, int N, int wf, int tiles, size_t M, const float4 *__restrict__ part,
                                   const float *__restrict__ mask_down, const float *__restrict__ vmean,
                                   const int *__restrict__ blocks, float *__restrict__ match, float *__restrict__ rsum)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    const int p = row_pixel(blocks + (size_t)b * ((P >> 2) + 1), i, wf);
    if (p < 0) return;
    const size_t bp = (size_t)b * P + p;
    if (mask_down[bp] == 0.f) {
        match[bp * 3 + 0] = vmean[b * 4 + 0];
        match[bp * 3 + 1] = vmean[b * 4 + 1];
        match[bp * 3 + 2] = vmean[b * 4 + 2];
        rsum[bp] = (float)N;
        return;
    }
    float l = 0.f, x = 0.f, y = 0.f, z = 0.f;
    for (int t = 0; t < CS * tiles; t++) {
        const float4 q = part[(size_t)t * M + (size_t)b * P + i];
        l += q.x; x += q.y; y += q.z; z += q.w;
    }
    const float inv = 1.f / l;
    match[bp * 3 + 0] = x * inv;
    match[bp * 3 + 1] = y * inv;
    match[bp * 3 + 2] = z * inv;
    rsum[bp] = l;
}

// columns: csum[b][n], imatch[b][:, n] = sum_p Pm grid (also run on the pooled partials: the uniform fallback then
// averages the pooled grid, which has the same mean as the full one)
__global__ void finish_cols_kernel(int P, int N, int Npad, size_t M, const float4 *__restrict__ part,
                                   const float *__restrict__ meshgrid, const int *__restrict__ blocks,
                                   float *__restrict__ imatch, float *__restrict__ csum)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (n >= N) return;
    const int tiles = (4 * blocks[(size_t)b * ((P >> 2) + 1)] + TN - 1) / TN;
    float s = 0.f, gx = 0.f, gy = 0.f;
    for (int t = 0; t < CS * tiles; t++) {
        const float4 q = part[(size_t)t * M + (size_t)b * Npad + n];
        s += q.x; gx += q.y; gy += q.z;
    }
    float ix, iy;
    if (s > 0.f) {
        ix = gx / s; iy = gy / s;
    } else {  // every pixel is background: the reference's softmax is uniform over all P pixels
        float mx = 0.f, my = 0.f;
        for (int p = 0; p < P; p++) { mx += meshgrid[p]; my += meshgrid[P + p]; }
        ix = mx / P; iy = my / P;
    }
    imatch[((size_t)b * 2 + 0) * N + n] = ix;
    imatch[((size_t)b * 2 + 1) * N + n] = iy;
    csum[(size_t)b * N + n] = s;
}

static size_t al(size_t x) { return (x + 255) / 256 * 256; }

struct Layout {
    int Npad;
    size_t a_img, a_mesh, wc, wp, wr, live, part_r, part_c, part_cp, total;
};

static Layout make_layout(int B, int P, int N)
{
    Layout L;
    L.Npad = (N + TM - 1) / TM * TM;
    size_t o = 0;
    L.a_img = o;   o += al((size_t)B * P * ROWF * 4);
    L.a_mesh = o;  o += al((size_t)B * L.Npad * ROWF * 4);
    L.wc = o;      o += al((size_t)B * P * 16);
    L.wp = o;      o += al((size_t)B * (P / 4) * 16);
    L.wr = o;      o += al((size_t)B * L.Npad * 16);
    L.live = o;    o += al((size_t)B * (P / 32) * 4);
    L.part_r = o;  o += al((size_t)(CS * L.Npad / TN) * B * P * 16);
    L.part_c = o;  o += al((size_t)(CS * P / TN) * B * L.Npad * 16);
    L.part_cp = o; o += al((size_t)(CS * P / TN) * B * L.Npad * 16);
    L.total = o;
    return L;
}

// shapes the tensor-core path covers (the callers fall back to the mma.sync kernel otherwise)
bool eligible(int B, int hf, int wf, int N, int Cc)
{
    const long P = (long)hf * wf;
    if (Cc != C || B <= 0 || N <= 0 || (wf & 1) || (hf & 1)) return false;
    if (P % TM != 0) return false;                                 // rows per problem of pass S, columns of pass S^T
    const long Npad = ((long)N + TM - 1) / TM * TM;
    if ((long)B * P >= (1l << 30) || (long)B * Npad >= (1l << 30)) return false;   // int row arithmetic in the GEMM kernel
    return true;
}

size_t workspace_bytes(int B, int hf, int wf, int N) { return make_layout(B, hf * wf, N).total; }

// blocks / vmean: outputs of corr_blocklist_kernel (already launched by the caller on `st`); corr_fill_kernel has
// written the constants of the all-background 2x2 blocks.  ws: workspace_bytes(...) bytes, 256-byte aligned.
int forward(const float *img_feat, const float *mesh_feat, const float *mask_down, const float *pred_v,
            const float *meshgrid, float tau, int B, int hf, int wf, int N, float *pc_pool, float *match,
            float *imatch, float *rsum, float *csum, float *A_pool, float *csum_pool, const int *blocks,
            const float *vmean, void *ws, cudaStream_t st)
{
    const int P = hf * wf;
    const Layout L = make_layout(B, P, N);
    char *w = (char *)ws;
    float *a_img = (float *)(w + L.a_img), *a_mesh = (float *)(w + L.a_mesh);
    float4 *wc = (float4 *)(w + L.wc), *wp = (float4 *)(w + L.wp), *wr = (float4 *)(w + L.wr);
    uint32_t *live = (uint32_t *)(w + L.live);
    float4 *part_r = (float4 *)(w + L.part_r), *part_c = (float4 *)(w + L.part_c), *part_cp = (float4 *)(w + L.part_cp);
    const int Npad = L.Npad;
    const float kexp = tau * LOG2E;

    prep_img_kernel<<<dim3(P / 64, B), NT, 0, st>>>(P, wf, img_feat, mask_down, meshgrid, blocks, a_img, wc, wp, live);
    prep_mesh_kernel<<<dim3((Npad * 32 + NT - 1) / NT, B), NT, 0, st>>>(N, Npad, mesh_feat, pred_v, a_mesh, wr);

    // pass S: rows = pixels (B problems of P rows), columns = vertices; one unit = all n tiles of an m tile
    EpiRows er{ wr, live, blocks, part_r, P, Npad, B * P, N, kexp };
    int rc = gemm_rs::launch(a_img, a_mesh, B, P, Npad, 16, er, st);
    if (rc != 0) return rc;
    // pass S^T: rows = vertices (B problems of Npad rows), columns = pixels; units of 4 n tiles
    if (pc_pool != nullptr) {
        EpiCols<true> ec{ wc, wp, live, blocks, part_c, part_cp, pc_pool, P, N, Npad, B * Npad, kexp };
        rc = gemm_rs::launch(a_mesh, a_img, B, Npad, P, 4, ec, st);
    } else {
        EpiCols<false> ec{ wc, wp, live, blocks, part_c, part_cp, nullptr, P, N, Npad, B * Npad, kexp };
        rc = gemm_rs::launch(a_mesh, a_img, B, Npad, P, 4, ec, st);
    }
    if (rc != 0) return rc;
    finish_rows_kernel<<<dim3(P / 256, B), 256, 0, st>>>(P, N, wf, Npad / TN, (size_t)B * P, part_r, mask_down, vmean, blocks,
                                                        match, rsum);
    finish_cols_kernel<<<dim3((N + 127) / 128, B), 128, 0, st>>>(P, N, Npad, (size_t)B * Npad, part_c, meshgrid, blocks, imatch,
                                                                 csum);
    if (A_pool != nullptr)
        finish_cols_kernel<<<dim3((N + 127) / 128, B), 128, 0, st>>>(P, N, Npad, (size_t)B * Npad, part_cp, meshgrid, blocks,
                                                                     A_pool, csum_pool);
    return 0;
}

}  // namespace corr_tc
}  // namespace scp
