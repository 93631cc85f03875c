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
#include "scp_gemm.cuh"
#else
// Host emulation (tools/emu, tests only): the tcgen05 GEMM is replaced by its plain statement -- the same tile walk,
// batched row arithmetic and kRowStats epilogue calls (row0 / col0 / lane conventions of gemm_bf16_tn_kernel), the
// accumulator formed from the split operands exactly as the NT = 4 issue loop combines them (hi*hi + hi*lo + lo*hi per
// [16 hi | 16 lo] group).  Everything else of this file (operand preparation, epilogue functors, combining kernels,
// workspace layout) is compiled unchanged.
namespace scp {
namespace gemm {
constexpr int BM = 256, BN = 128;
template <class Epi, int NTm>
int launch(const void *A, int lda, const void *W, int ldw, int M, int N, int K, const Epi &epi, cudaStream_t,
           void * = nullptr, int = 0, const long long *a_idx = nullptr, const long long *w_idx = nullptr,
           int rows_per_batch = 0, long = 0, long = 0)
{
    static_assert(NTm == 4, "emulated for the TF32-split mode only");
    const float *Af = (const float *)A, *Wf = (const float *)W;
    const int pa = lda / 2, pw = ldw / 2;   // pitches arrive in 2-byte units
    const int tiles_m = (M + BM - 1) / BM, tiles_n = N / BN;
    for (int tile = 0; tile < tiles_m * tiles_n; tile++) {
        const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;
        int a_row = m_blk * BM, w_row = n_blk * BN;
        if (a_idx != nullptr) {
            const int tpb = rows_per_batch / BM, p = m_blk / tpb;
            a_row = (int)a_idx[p] * rows_per_batch + (m_blk - p * tpb) * BM;
            w_row += (int)w_idx[p] * N;
        }
        for (int sub = 0; sub < 8; sub++) {           // epilogue warp: M half sub / 4, TMEM lane quarter sub % 4
            const int row0 = m_blk * BM + (sub >> 2) * 128 + (sub & 3) * 32;
            for (int lane = 0; lane < 32; lane++) {
                float a8[8] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
                const float *ar = Af + (size_t)(a_row + (row0 - m_blk * BM) + lane) * pa;
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    if (!epi.chunk_live(row0, n_blk * BN + c0)) continue;
                    float v[32];
                    for (int j = 0; j < 32; j++) {
                        const float *wr = Wf + (size_t)(w_row + c0 + j) * pw;
                        float acc = 0.f;
                        for (int g = 0; g < K / 16; g++)
                            for (int k = 0; k < 16; k++) {
                                const float ah = ar[32 * g + k], al = ar[32 * g + 16 + k];
                                const float wh = wr[32 * g + k], wl = wr[32 * g + 16 + k];
                                acc += ah * wh + ah * wl + al * wh;
                            }
                        v[j] = acc;
                    }
                    epi.accum(row0 + lane, n_blk * BN + c0, v, a8);
                }
                epi.finish(row0 + lane, n_blk, a8);
            }
        }
    }
    return 0;
}
}  // namespace gemm
}  // namespace scp
#endif

namespace scp {
namespace corr_tc {

constexpr int C = 64;            // feature channels
constexpr int ROWF = 2 * C;      // floats per split operand row: 4 groups of [16 hi | 16 lo]
constexpr int NT = 256;
constexpr float LOG2E = 1.4426950408889634f;

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

// block-major pixel order: row i -> 2x2 block i / 4 (row-major over the pooled map), sub-pixel i % 4 = 2 * dy + dx
__host__ __device__ __forceinline__ int perm_pixel(int i, int wf)
{
    const int id = i >> 2, sub = i & 3, w2 = wf >> 1;
    const int by = id / w2, bx = id - by * w2;
    return (2 * by + (sub >> 1)) * wf + 2 * bx + (sub & 1);
}

// ---- operand preparation ----------------------------------------------------------------------------------------
// img_feat[b][c][p] -> a_img[b][i][128] (block-major rows, split), wc[b][i] = mask * (1, gx, gy, 0),
// wp[b][i / 4] = (pooled gx, pooled gy, 0, 0), live[b][i / 32] = bit mask of the foreground rows.  CTA = 64 rows.
__global__ void __launch_bounds__(NT) prep_img_kernel(int P, int wf, const float *__restrict__ img_feat,
                                                      const float *__restrict__ mask_down,
                                                      const float *__restrict__ meshgrid, float *__restrict__ a_img,
                                                      float4 *__restrict__ wc, float4 *__restrict__ wp,
                                                      uint32_t *__restrict__ live)
{
    __shared__ float s[C][65];
    __shared__ float s_g[2][64];
    const int b = blockIdx.y, i0 = blockIdx.x * 64, tid = threadIdx.x;
    const float *img_b = img_feat + (size_t)b * C * P;
    for (int idx = tid; idx < C * 64; idx += NT) {
        const int c = idx >> 6, r = idx & 63;
        s[c][r] = img_b[(size_t)c * P + perm_pixel(i0 + r, wf)];
    }
    if (tid < 64) {
        const int p = perm_pixel(i0 + tid, wf);
        const bool fg = mask_down[(size_t)b * P + p] != 0.f;
        const float gx = meshgrid[p], gy = meshgrid[P + p];
        s_g[0][tid] = gx;
        s_g[1][tid] = gy;
        wc[(size_t)b * P + i0 + tid] = fg ? make_float4(1.f, gx, gy, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t bal = __ballot_sync(0xffffffffu, fg);
        if ((tid & 31) == 0) live[((size_t)b * P + i0 + tid) >> 5] = bal;
    }
    __syncthreads();
    if (tid < 16) {   // bilinear 1/2 of the meshgrid = 2x2 mean: (top + bottom) of each column, then the two columns
        const int k = 4 * tid;
        const float px = 0.25f * ((s_g[0][k] + s_g[0][k + 2]) + (s_g[0][k + 1] + s_g[0][k + 3]));
        const float py = 0.25f * ((s_g[1][k] + s_g[1][k + 2]) + (s_g[1][k + 1] + s_g[1][k + 3]));
        wp[((size_t)b * P + i0) / 4 + tid] = make_float4(px, py, 0.f, 0.f);
    }
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

// mesh_feat[b][n][c] -> a_mesh[b][n][128] (split; rows N..Npad-1 zero), wr[b][n] = (1, v) (zero past N)
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
            w = make_float4(1.f, v[0], v[1], v[2]);
        }
        wr[(size_t)b * Npad + n] = w;
    }
}

// ---- epilogues ---------------------------------------------------------------------------------------------------
// pass S (rows = block-major pixels, columns = vertices): a[0..3] += e * (1, v.x, v.y, v.z)
struct EpiRows {
    static constexpr bool kRowStats = true, kStaged = false, kTmaReduceAdd = false, kMixed = false;
    const float4 *wr;        // [B][Npad]
    const uint32_t *live;    // [B * P / 32]
    float4 *part;            // [Npad / 128][B * P]
    int P, Npad, M;
    float kexp;
    __device__ bool chunk_live(int row0, int) const { return live[row0 >> 5] != 0u; }
    __device__ void accum(int row, int col0, const float (&v)[32], float (&a)[8]) const
    {
        const float4 *w = wr + (size_t)(row / P) * Npad + col0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const float4 q = __ldg(w + j);
            const float e = ex2_approx(fmaf(v[j], kexp, -kexp));
            a[0] = fmaf(e, q.x, a[0]);
            a[1] = fmaf(e, q.y, a[1]);
            a[2] = fmaf(e, q.z, a[2]);
            a[3] = fmaf(e, q.w, a[3]);
        }
    }
    __device__ void finish(int row, int n_blk, const float (&a)[8]) const
    {
        part[(size_t)n_blk * M + row] = make_float4(a[0], a[1], a[2], a[3]);
    }
};

// pass S^T (rows = vertices, columns = block-major pixels): a[0..2] += e * mask * (1, gx, gy); with POOL the 2x2 mean of
// the masked similarity is written (coalesced along the vertices) and a[4..6] += e_pool * (1, pooled gx, pooled gy)
template <bool POOL>
struct EpiCols {
    static constexpr bool kRowStats = true, kStaged = false, kTmaReduceAdd = false, kMixed = false;
    const float4 *wc;        // [B][P]
    const float4 *wp;        // [B][P / 4]
    const uint32_t *live;    // [B][P / 32]
    float4 *part, *part_pool;   // [P / 128][B * Npad]
    float *pc_pool;          // [B][P / 4][N]
    int P, N, Npad, M;
    float kexp;
    __device__ bool chunk_live(int row0, int col0) const { return live[(size_t)(row0 / Npad) * (P >> 5) + (col0 >> 5)] != 0u; }
    __device__ void accum(int row, int col0, const float (&v)[32], float (&a)[8]) const
    {
        const int b = row / Npad, n = row - b * Npad;
        const float4 *w = wc + (size_t)b * P + col0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            float sm[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const float4 q = __ldg(w + 4 * k + t);
                const float s = v[4 * k + t];
                const float e = ex2_approx(fmaf(s, kexp, -kexp));
                a[0] = fmaf(e, q.x, a[0]);
                a[1] = fmaf(e, q.y, a[1]);
                a[2] = fmaf(e, q.z, a[2]);
                sm[t] = q.x != 0.f ? s : -1e5f;
            }
            if constexpr (POOL) {
                const float pv = 0.25f * ((sm[0] + sm[2]) + (sm[1] + sm[3]));   // (top + bottom) per column, then the columns
                const size_t id = (size_t)b * (P >> 2) + (col0 >> 2) + k;
                if (n < N) pc_pool[id * N + n] = pv;
                const float4 g = __ldg(wp + id);
                const float ep = ex2_approx(fmaf(pv, kexp, -kexp));   // blocks with a background pixel underflow to 0
                a[4] += ep;
                a[5] = fmaf(ep, g.x, a[5]);
                a[6] = fmaf(ep, g.y, a[6]);
            }
        }
    }
    __device__ void finish(int row, int n_blk, const float (&a)[8]) const
    {
        part[(size_t)n_blk * M + row] = make_float4(a[0], a[1], a[2], 0.f);
        if constexpr (POOL) part_pool[(size_t)n_blk * M + row] = make_float4(a[4], a[5], a[6], 0.f);
    }
};

// ---- combining the per-tile partials -----------------------------------------------------------------------------
// rows: match[b][p] = sum_n Pi v, rsum; background pixels: uniform soft-max -> mean vertex, rsum = N (correspondence.py:44,48)
__global__ void finish_rows_kernel(int B, int P, int N, int wf, int tiles, const float4 *__restrict__ part,
                                   const float *__restrict__ mask_down, const float *__restrict__ vmean,
                                   float *__restrict__ match, float *__restrict__ rsum)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= B * P) return;
    const int b = row / P, p = perm_pixel(row - b * P, wf);
    const size_t bp = (size_t)b * P + p;
    if (mask_down[bp] == 0.f) {
        match[bp * 3 + 0] = vmean[b * 4 + 0];
        match[bp * 3 + 1] = vmean[b * 4 + 1];
        match[bp * 3 + 2] = vmean[b * 4 + 2];
        rsum[bp] = (float)N;
        return;
    }
    float l = 0.f, x = 0.f, y = 0.f, z = 0.f;
    for (int t = 0; t < tiles; t++) {
        const float4 q = part[(size_t)t * B * P + row];
        l += q.x; x += q.y; y += q.z; z += q.w;
    }
    const float inv = 1.f / l;
    match[bp * 3 + 0] = x * inv;
    match[bp * 3 + 1] = y * inv;
    match[bp * 3 + 2] = z * inv;
    rsum[bp] = l;
}

// columns: csum[b][n], imatch[b][:, n] = sum_p Pm grid (pooled = 1: the same over the pooled map)
__global__ void finish_cols_kernel(int B, int P, int N, int Npad, int tiles, const float4 *__restrict__ part,
                                   const float *__restrict__ meshgrid, float *__restrict__ imatch,
                                   float *__restrict__ csum)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (n >= N) return;
    float s = 0.f, gx = 0.f, gy = 0.f;
    for (int t = 0; t < tiles; t++) {
        const float4 q = part[(size_t)t * B * Npad + (size_t)b * Npad + n];
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

__global__ void iota_kernel(long long *idx, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = i;
}

static size_t al(size_t x) { return (x + 255) / 256 * 256; }

struct Layout {
    int Npad;
    size_t a_img, a_mesh, wc, wp, wr, live, part_r, part_c, part_cp, idx, total;
};

static Layout make_layout(int B, int P, int N)
{
    Layout L;
    L.Npad = (N + gemm::BM - 1) / gemm::BM * gemm::BM;
    size_t o = 0;
    L.a_img = o;   o += al((size_t)B * P * ROWF * 4);
    L.a_mesh = o;  o += al((size_t)B * L.Npad * ROWF * 4);
    L.wc = o;      o += al((size_t)B * P * 16);
    L.wp = o;      o += al((size_t)B * (P / 4) * 16);
    L.wr = o;      o += al((size_t)B * L.Npad * 16);
    L.live = o;    o += al((size_t)B * (P / 32) * 4);
    L.part_r = o;  o += al((size_t)(L.Npad / gemm::BN) * B * P * 16);
    L.part_c = o;  o += al((size_t)(P / gemm::BN) * B * L.Npad * 16);
    L.part_cp = o; o += al((size_t)(P / gemm::BN) * B * L.Npad * 16);
    L.idx = o;     o += al((size_t)B * 8);
    L.total = o;
    return L;
}

// shapes the tensor-core path covers (the callers fall back to the mma.sync kernel otherwise)
bool eligible(int B, int hf, int wf, int N, int Cc)
{
    const long P = (long)hf * wf;
    if (Cc != C || B <= 0 || N <= 0 || (wf & 1) || (hf & 1)) return false;
    if (P % gemm::BM != 0) return false;                          // rows per problem of pass S, columns of pass S^T
    const long Npad = ((long)N + gemm::BM - 1) / gemm::BM * gemm::BM;
    if ((long)B * P >= (1l << 30) || (long)B * Npad >= (1l << 30)) return false;   // int row arithmetic in the GEMM kernel
    return true;
}

size_t workspace_bytes(int B, int hf, int wf, int N) { return make_layout(B, hf * wf, N).total; }

// blocks / vmean: outputs of corr_blocklist_kernel (already launched by the caller on `st`); corr_fill_kernel has
// written the constants of the all-background 2x2 blocks.  ws: workspace_bytes(...) bytes, 256-byte aligned.
int forward(const float *img_feat, const float *mesh_feat, const float *mask_down, const float *pred_v,
            const float *meshgrid, float tau, int B, int hf, int wf, int N, float *pc_pool, float *match,
            float *imatch, float *rsum, float *csum, float *A_pool, float *csum_pool, const float *vmean, void *ws,
            cudaStream_t st)
{
    const int P = hf * wf;
    const Layout L = make_layout(B, P, N);
    char *w = (char *)ws;
    float *a_img = (float *)(w + L.a_img), *a_mesh = (float *)(w + L.a_mesh);
    float4 *wc = (float4 *)(w + L.wc), *wp = (float4 *)(w + L.wp), *wr = (float4 *)(w + L.wr);
    uint32_t *live = (uint32_t *)(w + L.live);
    float4 *part_r = (float4 *)(w + L.part_r), *part_c = (float4 *)(w + L.part_c), *part_cp = (float4 *)(w + L.part_cp);
    long long *idx = (long long *)(w + L.idx);
    const int Npad = L.Npad;
    const float kexp = tau * LOG2E;

    prep_img_kernel<<<dim3(P / 64, B), NT, 0, st>>>(P, wf, img_feat, mask_down, meshgrid, a_img, wc, wp, live);
    prep_mesh_kernel<<<dim3((Npad * 32 + NT - 1) / NT, B), NT, 0, st>>>(N, Npad, mesh_feat, pred_v, a_mesh, wr);
    iota_kernel<<<(B + 127) / 128, 128, 0, st>>>(idx, B);

    // pass S: rows = pixels (B problems of P rows), columns = vertices
    EpiRows er{ wr, live, part_r, P, Npad, B * P, kexp };
    int rc = gemm::launch<EpiRows, 4>(a_img, 2 * ROWF, a_mesh, 2 * ROWF, B * P, Npad, C, er, st, nullptr, 0, idx, idx, P,
                                      (long)B * P, (long)B * Npad);
    if (rc != 0) return rc;
    // pass S^T: rows = vertices (B problems of Npad rows), columns = pixels
    if (pc_pool != nullptr) {
        EpiCols<true> ec{ wc, wp, live, part_c, part_cp, pc_pool, P, N, Npad, B * Npad, kexp };
        rc = gemm::launch<EpiCols<true>, 4>(a_mesh, 2 * ROWF, a_img, 2 * ROWF, B * Npad, P, C, ec, st, nullptr, 0, idx, idx,
                                            Npad, (long)B * Npad, (long)B * P);
    } else {
        EpiCols<false> ec{ wc, wp, live, part_c, part_cp, nullptr, P, N, Npad, B * Npad, kexp };
        rc = gemm::launch<EpiCols<false>, 4>(a_mesh, 2 * ROWF, a_img, 2 * ROWF, B * Npad, P, C, ec, st, nullptr, 0, idx, idx,
                                             Npad, (long)B * Npad, (long)B * P);
    }
    if (rc != 0) return rc;
    finish_rows_kernel<<<(B * P + 255) / 256, 256, 0, st>>>(B, P, N, wf, Npad / gemm::BN, part_r, mask_down, vmean, match, rsum);
    finish_cols_kernel<<<dim3((N + 127) / 128, B), 128, 0, st>>>(B, P, N, Npad, P / gemm::BN, part_c, meshgrid, imatch, csum);
    if (A_pool != nullptr)
        finish_cols_kernel<<<dim3((N + 127) / 128, B), 128, 0, st>>>(B, P, N, Npad, P / gemm::BN, part_cp, meshgrid, A_pool,
                                                                     csum_pool);
    return 0;
}

}  // namespace corr_tc
}  // namespace scp
