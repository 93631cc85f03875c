// Dense 2D<->3D feature correspondence (Correspondence.match of the reference,
// model/module/correspondence.py:36-73) as fused sm_100a kernels, forward + backward.
//
//   S[p,n]   = <mesh_feat[n,:], img_feat[:,p]>, rows of background pixels := -1e5      (:42-44)
//   Pm       = softmax(tau*S, dim=p)   imatch[:,n] = sum_p meshgrid[:,p] Pm[p,n]        (:47,52)
//   Pi       = softmax(tau*S, dim=n)   match[p,:]  = sum_n Pi[p,n] pred_v[n,:]          (:48,53)
//
// The reference materialises S and both softmaxes ((B,P,N) fp32 each, 16-21 MB per image) plus
// a (B,P,N,3) broadcast product.  Here one CTA owns 128 pixels (two rows of the 64x64 map, ordered
// as 8x2 patches so that the 2x2 down-sampling used by the pre-training cycle loss completes
// inside a thread quad), streams the vertices in tiles of 64, forms the S tile on the tensor
// cores (m16n8k8 TF32 with the 3-term split, ~fp32 accuracy; K = 64 only) and reduces it in
// registers: ONE exp per element serves both softmaxes (reference point tau*1: features are
// L2-normalised, so S <= 1), row sums stay in the thread, column sums are folded across the warp
// and written as per-row-block partials that a small second kernel combines.  S itself is
// written at most once (full, eval mode) or only as its 2x2 mean (training).  The backward
// recomputes the S tile instead of reading saved softmaxes and runs the two gradient products on
// the tensor cores as well.
//
// Background skipping: a background pixel contributes exactly nothing to the column softmax and its own
// row outputs are constants (uniform row softmax -> match = mean vertex, pointcorr = -1e5, zero gradient).
// A per-image pre-pass lists the 2x2 pixel blocks that contain at least one foreground pixel; the row
// blocks of every kernel are built from that list (32 blocks = 128 rows per CTA), so the P x N work
// shrinks to the silhouette's share of the map (the object mask covers ~40 % of a crop), and a fill
// kernel writes the constants of the dropped blocks.
#include <stdlib.h>

#include "../../include/scp_b200.h"
#include "scp_common.cuh"
#include "scp_mma.cuh"

namespace scp {
namespace corr {

constexpr int C = 64;        // feature channels (n_corr_feat of every shipped config)
constexpr int BM = 128;      // pixels per row block
constexpr int BN = 64;       // vertices per tile
constexpr int NT = 256;
constexpr int AS = BM + 8;   // As[c][r]   (k-major A operand)
constexpr int BS = C + 4;    // Bs[n][c]
constexpr int DS = BN + 4;   // Ds[r][n]
constexpr float LOG2E = 1.4426950408889634f;

struct Geo {
    int B, P, N, hf, wf, npblk, ntile;
    float tau;
};

// Per-image list of kept 2x2 pixel blocks: blk[0] = count, blk[1..count] = pooled-pixel ids (ascending).
// MMA row r (0..127) of row block `pblk` -> pixel index (-1 past the end of the list).  A 16-row MMA tile holds
// four listed blocks: rows q and q^1 are horizontal neighbours, rows q and q+8 vertical neighbours of one block.
constexpr int BLK_PER_CTA = BM / 4;
__device__ __forceinline__ int row_pixel(const Geo &g, const int *__restrict__ blk, int pblk, int r, int &pool_idx)
{
    const int j = r >> 4, q = r & 15;
    const int li = pblk * BLK_PER_CTA + 4 * j + ((q & 7) >> 1);
    if (li >= blk[0]) { pool_idx = -1; return -1; }
    const int id = blk[1 + li], w2 = g.wf >> 1;
    const int by = id / w2, bx = id - by * w2;
    pool_idx = id;
    return (2 * by + (q >> 3)) * g.wf + 2 * bx + (q & 1);
}
__device__ __forceinline__ int active_row_blocks(const int *__restrict__ blk) { return (blk[0] + BLK_PER_CTA - 1) / BLK_PER_CTA; }

// ---- tile loaders --------------------------------------------------------------------------
// As[c][r] <- img_feat[b][c][pixel(r)]: rows 2i, 2i+1 are horizontally adjacent pixels (one 8-byte cp.async);
// rows past the block list are zero-filled.  s_pix must be visible (barrier) before the call.
__device__ __forceinline__ void load_A(const Geo &g, float *As, const float *__restrict__ img_b, const int *s_pix)
{
    for (int i = threadIdx.x; i < C * (BM / 2); i += NT) {
        const int c = i / (BM / 2), r2 = (i - c * (BM / 2)) * 2;
        const int p = s_pix[r2];
        if (p >= 0) cp_async8(As + c * AS + r2, img_b + (size_t)c * g.P + p);
        else *reinterpret_cast<float2 *>(As + c * AS + r2) = make_float2(0.f, 0.f);
    }
}

// Bs[n][c] <- mesh_feat[b][n0+n][c], rows past N zero-filled
__device__ __forceinline__ void load_B(const Geo &g, float *Bs, const float *__restrict__ mesh_b, int n0)
{
    for (int i = threadIdx.x; i < BN * (C / 4); i += NT) {
        const int n = i / (C / 4), c4 = (i - n * (C / 4)) * 4;
        if (n0 + n < g.N) cp_async16(Bs + n * BS + c4, mesh_b + (size_t)(n0 + n) * C + c4);
        else *reinterpret_cast<float4 *>(Bs + n * BS + c4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ---- S tile on the tensor cores: acc[mi][ni] (warp tile 32 rows x 32 cols, K = 64) ----------
__device__ __forceinline__ void mma_S(float (&acc)[2][4][4], const float *As, const float *Bs, int wm, int wn,
                                      int g, int t)
{
#pragma unroll
    for (int mi = 0; mi < 2; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++)
#pragma unroll
            for (int k = 0; k < 4; k++) acc[mi][ni][k] = 0.f;
#pragma unroll
    for (int k0 = 0; k0 < C; k0 += 8) {
        uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
        for (int mi = 0; mi < 2; mi++) {
            const float *a = As + (k0 + t) * AS + 32 * wm + 16 * mi + g;
            split_tf32(a[0], ah[mi][0], al[mi][0]);
            split_tf32(a[8], ah[mi][1], al[mi][1]);
            split_tf32(a[4 * AS], ah[mi][2], al[mi][2]);
            split_tf32(a[4 * AS + 8], ah[mi][3], al[mi][3]);
        }
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            const float *b = Bs + (32 * wn + 8 * ni + g) * BS + k0 + t;
            split_tf32(b[0], bh[ni][0], bl[ni][0]);
            split_tf32(b[4], bh[ni][1], bl[ni][1]);
        }
#pragma unroll
        for (int mi = 0; mi < 2; mi++)
#pragma unroll
            for (int ni = 0; ni < 4; ni++) {
                mma_tf32(acc[mi][ni], al[mi], bh[ni]);
                mma_tf32(acc[mi][ni], ah[mi], bl[ni]);
                mma_tf32(acc[mi][ni], ah[mi], bh[ni]);
            }
    }
}

// fold 24 per-lane column partials ([c = 2*ni + j][3]) across the 8 row-lanes (lane bits 2..4): 24 -> 12 -> 6 -> 3.
// Lane (g,t) ends with column c = g of its t-group, i.e. tile column 32*wn + 8*(g>>1) + 2*t + (g&1).
__device__ __forceinline__ void fold24(const float (&cv)[24], int lane, float (&a3)[3])
{
    float a12[12], a6[6];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const float send = b4 ? cv[i] : cv[12 + i], keep = b4 ? cv[12 + i] : cv[i];
        a12[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
        const float send = b3 ? a12[i] : a12[6 + i], keep = b3 ? a12[6 + i] : a12[i];
        a6[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float send = b2 ? a6[i] : a6[3 + i], keep = b2 ? a6[3 + i] : a6[i];
        a3[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
}

// ---- forward --------------------------------------------------------------------------------
// dynamic shared memory layout (floats)
constexpr int F_AS = 0;
constexpr int F_BS = F_AS + C * AS;              // 2 buffers
constexpr int F_VS = F_BS + 2 * BN * BS;         // 2 buffers of [BN][4]: x, y, z, valid
constexpr int F_COL = F_VS + 2 * BN * 4;         // [2: full-res / pooled][4 wm][BN][4]
constexpr int F_ROW = F_COL + 2 * 4 * BN * 4;    // [2 wn][BM][4]
constexpr int F_INFO = F_ROW + 2 * BM * 4;       // mask[BM], gx[BM], gy[BM], pixel[BM] (int), pooled pixel[BM] (int)
constexpr int F_TOTAL = F_INFO + 5 * BM;

__device__ __forceinline__ void load_V(const Geo &g, float *Vs, const float *__restrict__ v_b, int n0)
{
    if (threadIdx.x < BN) {
        const int n = n0 + threadIdx.x;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < g.N) q = make_float4(v_b[3 * n], v_b[3 * n + 1], v_b[3 * n + 2], 1.f);
        *reinterpret_cast<float4 *>(Vs + 4 * threadIdx.x) = q;
    }
}

__global__ void __launch_bounds__(NT, 2)
corr_fwd_kernel(Geo geo, const float *__restrict__ img_feat, const float *__restrict__ mesh_feat,
                const float *__restrict__ mask_down, const float *__restrict__ pred_v,
                const float *__restrict__ meshgrid, float *__restrict__ pc_full, float *__restrict__ pc_pool,
                float *__restrict__ match, float *__restrict__ rsum, float *__restrict__ colpart,
                float *__restrict__ colpart_pool, const int *__restrict__ blocks)
{
    extern __shared__ __align__(16) float sm[];
    float *As = sm + F_AS, *Bs = sm + F_BS, *Vs = sm + F_VS, *s_col = sm + F_COL, *s_row = sm + F_ROW;
    float *s_mask = sm + F_INFO, *s_gx = s_mask + BM, *s_gy = s_gx + BM;
    int *s_pix = reinterpret_cast<int *>(s_gy + BM), *s_pool = s_pix + BM;

    const int pblk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int *blk = blocks + (size_t)b * ((geo.P >> 2) + 1);
    if (pblk >= active_row_blocks(blk)) return;   // row blocks past the foreground list have no work
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3, wm = warp & 3, wn = warp >> 2;
    const float *img_b = img_feat + (size_t)b * C * geo.P;
    const float *mesh_b = mesh_feat + (size_t)b * geo.N * C;
    const float *v_b = pred_v + (size_t)b * geo.N * 3;
    const float kexp = geo.tau * LOG2E;

    if (tid < BM) {
        int pool;
        const int p = row_pixel(geo, blk, pblk, tid, pool);
        s_pix[tid] = p;
        s_pool[tid] = pool;
        s_mask[tid] = p >= 0 ? mask_down[(size_t)b * geo.P + p] : 0.f;
        s_gx[tid] = p >= 0 ? meshgrid[p] : 0.f;
        s_gy[tid] = p >= 0 ? meshgrid[geo.P + p] : 0.f;
    }
    __syncthreads();
    load_A(geo, As, img_b, s_pix);
    load_B(geo, Bs, mesh_b, 0);
    cp_async_commit();
    load_V(geo, Vs, v_b, 0);

    // this thread's four rows: ri = 2*mi + h  ->  r = 32*wm + 16*mi + 8*h + g
    float rmask[4], rgx[4], rgy[4];
    int rpix[4], rpool[2];
#pragma unroll
    for (int ri = 0; ri < 4; ri++) {
        const int r = 32 * wm + 16 * (ri >> 1) + 8 * (ri & 1) + g;
        rmask[ri] = s_mask[r]; rgx[ri] = s_gx[r]; rgy[ri] = s_gy[r]; rpix[ri] = s_pix[r];
    }
#pragma unroll
    for (int mi = 0; mi < 2; mi++) rpool[mi] = s_pool[32 * wm + 16 * mi + g];
    // grid coordinates of the 2x2-pooled pixel of each row pair (bilinear 1/2 of the meshgrid = 2x2 mean)
    float pgx[2], pgy[2];
#pragma unroll
    for (int mi = 0; mi < 2; mi++) {
        float sx = rgx[2 * mi] + rgx[2 * mi + 1], sy = rgy[2 * mi] + rgy[2 * mi + 1];
        sx += __shfl_xor_sync(0xffffffffu, sx, 4);
        sy += __shfl_xor_sync(0xffffffffu, sy, 4);
        pgx[mi] = 0.25f * sx; pgy[mi] = 0.25f * sy;
    }
    const bool want_pool_stats = colpart_pool != nullptr;
    float rl[4] = { 0.f, 0.f, 0.f, 0.f }, rax[4] = { 0.f, 0.f, 0.f, 0.f }, ray[4] = { 0.f, 0.f, 0.f, 0.f },
          raz[4] = { 0.f, 0.f, 0.f, 0.f };

    for (int it = 0; it < geo.ntile; it++) {
        const int n0 = it * BN;
        float *Bt = Bs + (it & 1) * BN * BS, *Vt = Vs + (it & 1) * BN * 4;
        cp_async_wait<0>();
        __syncthreads();  // tile `it` has landed; every warp is done with iteration it-1
        if (it + 1 < geo.ntile) {  // prefetch overlaps the MMAs below
            load_B(geo, Bs + ((it + 1) & 1) * BN * BS, mesh_b, n0 + BN);
            cp_async_commit();
            load_V(geo, Vs + ((it + 1) & 1) * BN * 4, v_b, n0 + BN);
        }

        float acc[2][4][4];
        mma_S(acc, As, Bt, wm, wn, g, t);

        float cv[24];   // column partials: [c = 2*ni + j][sum e, sum e*gx, sum e*gy]
        float cvp[24];  // the same for the 2x2-pooled similarity (column softmax of the pre-training cycle loss)
#pragma unroll
        for (int k = 0; k < 24; k++) cv[k] = cvp[k] = 0.f;
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int cl = 32 * wn + 8 * ni + 2 * t + j;
                const float4 vq = *reinterpret_cast<const float4 *>(Vt + 4 * cl);
                const bool valid = vq.w != 0.f;
                const int n = n0 + cl;
#pragma unroll
                for (int mi = 0; mi < 2; mi++) {
                    float sm2[2];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int ri = 2 * mi + h;
                        const bool masked = rmask[ri] == 0.f;
                        const float s = masked ? -1e5f : acc[mi][ni][2 * h + j];
                        sm2[h] = s;
                        const float e = (valid && !masked) ? ex2_approx((s - 1.f) * kexp) : 0.f;
                        const float er = valid ? (masked ? 1.f : e) : 0.f;  // background rows: uniform softmax
                        rl[ri] += er; rax[ri] += er * vq.x; ray[ri] += er * vq.y; raz[ri] += er * vq.z;
                        cv[(2 * ni + j) * 3 + 0] += e;
                        cv[(2 * ni + j) * 3 + 1] += e * rgx[ri];
                        cv[(2 * ni + j) * 3 + 2] += e * rgy[ri];
                        if (pc_full != nullptr && valid && rpix[ri] >= 0) pc_full[((size_t)b * geo.P + rpix[ri]) * geo.N + n] = s;
                    }
                    if (pc_pool != nullptr) {
                        float q = sm2[0] + sm2[1];
                        q += __shfl_xor_sync(0xffffffffu, q, 4);   // horizontal neighbour (g ^ 1)
                        const float pv = 0.25f * q;
                        if (valid && !(g & 1) && rpool[mi] >= 0) {
                            pc_pool[((size_t)b * (geo.P >> 2) + rpool[mi]) * geo.N + n] = pv;
                            if (want_pool_stats) {   // pooled rows containing background pixels underflow to 0
                                const float ep = ex2_approx((pv - 1.f) * kexp);
                                cvp[(2 * ni + j) * 3 + 0] += ep;
                                cvp[(2 * ni + j) * 3 + 1] += ep * pgx[mi];
                                cvp[(2 * ni + j) * 3 + 2] += ep * pgy[mi];
                            }
                        }
                    }
                }
            }
        }
        {
            float a3[3];
            const int cl = 32 * wn + 8 * (g >> 1) + 2 * t + (g & 1);
            fold24(cv, lane, a3);
            *reinterpret_cast<float4 *>(s_col + (wm * BN + cl) * 4) = make_float4(a3[0], a3[1], a3[2], 0.f);
            if (want_pool_stats) {
                fold24(cvp, lane, a3);
                *reinterpret_cast<float4 *>(s_col + ((4 + wm) * BN + cl) * 4) = make_float4(a3[0], a3[1], a3[2], 0.f);
            }
        }
        __syncthreads();
        if (tid < BN * 3) {
            const int cl = tid / 3, k = tid - cl * 3;
            if (n0 + cl < geo.N) {
                const size_t dst = (((size_t)b * geo.npblk + pblk) * geo.N + n0 + cl) * 4 + k;
                colpart[dst] = s_col[(0 * BN + cl) * 4 + k] + s_col[(1 * BN + cl) * 4 + k] +
                               s_col[(2 * BN + cl) * 4 + k] + s_col[(3 * BN + cl) * 4 + k];
                if (want_pool_stats)
                    colpart_pool[dst] = s_col[(4 * BN + cl) * 4 + k] + s_col[(5 * BN + cl) * 4 + k] +
                                        s_col[(6 * BN + cl) * 4 + k] + s_col[(7 * BN + cl) * 4 + k];
            }
        }
    }

    // rows: combine the 4 t-lanes, then the two column-halves (wn) through shared memory
#pragma unroll
    for (int ri = 0; ri < 4; ri++) {
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
            rl[ri] += __shfl_xor_sync(0xffffffffu, rl[ri], o);
            rax[ri] += __shfl_xor_sync(0xffffffffu, rax[ri], o);
            ray[ri] += __shfl_xor_sync(0xffffffffu, ray[ri], o);
            raz[ri] += __shfl_xor_sync(0xffffffffu, raz[ri], o);
        }
        if (t == 0) {
            const int r = 32 * wm + 16 * (ri >> 1) + 8 * (ri & 1) + g;
            *reinterpret_cast<float4 *>(s_row + (wn * BM + r) * 4) = make_float4(rl[ri], rax[ri], ray[ri], raz[ri]);
        }
    }
    __syncthreads();
    if (tid < BM && s_pix[tid] >= 0) {
        const float4 a = *reinterpret_cast<const float4 *>(s_row + tid * 4);
        const float4 c = *reinterpret_cast<const float4 *>(s_row + (BM + tid) * 4);
        const float l = a.x + c.x, inv = 1.f / l;
        const size_t p = (size_t)b * geo.P + s_pix[tid];
        match[p * 3 + 0] = (a.y + c.y) * inv;
        match[p * 3 + 1] = (a.z + c.z) * inv;
        match[p * 3 + 2] = (a.w + c.w) * inv;
        rsum[p] = l;
    }
}

// combine the per-row-block column partials: csum[n], imatch[:,n]
// (launched a second time on the pooled partials with pooled = 1: the uniform fallback then averages the
// pooled grid, which has the same mean as the full one)
__global__ void corr_colreduce_kernel(Geo geo, const float *__restrict__ colpart,
                                      const float *__restrict__ meshgrid, float *__restrict__ imatch,
                                      float *__restrict__ csum, const int *__restrict__ blocks)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (n >= geo.N) return;
    const int nact = active_row_blocks(blocks + (size_t)b * ((geo.P >> 2) + 1));
    float s = 0.f, gx = 0.f, gy = 0.f;
    for (int k = 0; k < nact; k++) {
        const float4 q = *reinterpret_cast<const float4 *>(colpart + (((size_t)b * geo.npblk + k) * geo.N + n) * 4);
        s += q.x; gx += q.y; gy += q.z;
    }
    float ix, iy;
    if (s > 0.f) {
        ix = gx / s; iy = gy / s;
    } else {  // every pixel is background: the reference's softmax is uniform over all P pixels
        float mx = 0.f, my = 0.f;
        for (int p = 0; p < geo.P; p++) { mx += meshgrid[p]; my += meshgrid[geo.P + p]; }
        ix = mx / geo.P; iy = my / geo.P;
    }
    imatch[((size_t)b * 2 + 0) * geo.N + n] = ix;
    imatch[((size_t)b * 2 + 1) * geo.N + n] = iy;
    csum[(size_t)b * geo.N + n] = s;
}

// ---- backward -------------------------------------------------------------------------------
// dS[p,n] = mask[p] * ( tau*Pi*(gm[p].v[n] - gm[p].match[p]) + tau*Pm*(gi[:,n].grid[:,p] - gi[:,n].imatch[:,n])
//                       + 0.25*g_pool[pool(p),n] + g_full[p,n] ),  Pi = e/rsum[p], Pm = e/csum[n]
// row parameters (8 floats): mask, gx, gy, tau/rsum, gm.x, gm.y, gm.z, gm.match
// col parameters (12 floats): v.x, v.y, v.z, tau/csum | gi.x, gi.y, gi.imatch, valid |
//                             tau/csum_pool, gA.x, gA.y, gA.A_pool   (pooled column softmax, zeros when unused)
struct BwdArgs {
    const float *img_feat, *mesh_feat, *mask_down, *pred_v, *meshgrid;
    const float *match, *imatch, *rsum, *csum;
    const float *g_match, *g_imatch, *g_pool, *g_full;
    const float *A_pool, *csum_pool, *g_A_pool;   // pooled column softmax (may be NULL)
    float *g_img_feat, *g_mesh_feat;
    const int *blocks;   // per-image foreground block lists (see row_pixel)
};
constexpr int CPS = 12;   // floats per column-parameter record

__device__ __forceinline__ void load_rowparams(const Geo &geo, const BwdArgs &a, int b, int pblk, float *Rp,
                                               int *s_pix, int *s_pool)
{
    if (threadIdx.x < BM) {
        const int r = threadIdx.x;
        int pool;
        const int p = row_pixel(geo, a.blocks + (size_t)b * ((geo.P >> 2) + 1), pblk, r, pool);
        float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;   // rows past the list: mask 0 -> dS = 0
        if (p >= 0) {
            const size_t bp = (size_t)b * geo.P + p;
            const float mk = a.mask_down[bp];
            const float gx = a.meshgrid[p], gy = a.meshgrid[geo.P + p];
            const float l = a.rsum[bp];
            const float g0 = a.g_match[bp * 3], g1 = a.g_match[bp * 3 + 1], g2 = a.g_match[bp * 3 + 2];
            const float dot = g0 * a.match[bp * 3] + g1 * a.match[bp * 3 + 1] + g2 * a.match[bp * 3 + 2];
            q0 = make_float4(mk, gx, gy, geo.tau / l);
            q1 = make_float4(g0, g1, g2, dot);
        }
        float4 *dst = reinterpret_cast<float4 *>(Rp + 8 * r);
        dst[0] = q0;
        dst[1] = q1;
        s_pix[r] = p;
        s_pool[r] = pool;
    }
}

__device__ __forceinline__ void load_colparams(const Geo &geo, const BwdArgs &a, int b, int n0, float *Cp)
{
    if (threadIdx.x < BN) {
        const int n = n0 + threadIdx.x;
        float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0, q2 = q0;
        if (n < geo.N) {
            const float *v = a.pred_v + ((size_t)b * geo.N + n) * 3;
            const size_t i0 = ((size_t)b * 2 + 0) * geo.N + n, i1 = ((size_t)b * 2 + 1) * geo.N + n;
            const float cs = a.csum[(size_t)b * geo.N + n];
            const float gi0 = a.g_imatch[i0], gi1 = a.g_imatch[i1];
            q0 = make_float4(v[0], v[1], v[2], cs > 0.f ? geo.tau / cs : 0.f);
            q1 = make_float4(gi0, gi1, gi0 * a.imatch[i0] + gi1 * a.imatch[i1], 1.f);
            if (a.g_A_pool != nullptr) {
                const float cp = a.csum_pool[(size_t)b * geo.N + n];
                const float ga0 = a.g_A_pool[i0], ga1 = a.g_A_pool[i1];
                q2 = make_float4(cp > 0.f ? geo.tau / cp : 0.f, ga0, ga1, ga0 * a.A_pool[i0] + ga1 * a.A_pool[i1]);
            }
        }
        float4 *dst = reinterpret_cast<float4 *>(Cp + CPS * threadIdx.x);
        dst[0] = q0;
        dst[1] = q1;
        dst[2] = q2;
    }
}

// turns the S tile in acc into dS and stores it to Ds[r][n_local]
__device__ __forceinline__ void make_dS(const Geo &geo, const BwdArgs &a, int b, int n0, float (&acc)[2][4][4],
                                        const float *Rp, const float *Cp, const int *s_pix, const int *s_pool,
                                        float *Ds, int wm, int wn, int g, int t)
{
    const float kexp = geo.tau * LOG2E;
    const bool pooled_softmax = a.g_A_pool != nullptr;
#pragma unroll
    for (int mi = 0; mi < 2; mi++) {
        const int r_lo = 32 * wm + 16 * mi + g;          // rows r_lo (h = 0) and r_lo + 8 (h = 1): a vertical pair
        float4 ra[2], rb[2];
        int pix[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            ra[h] = *reinterpret_cast<const float4 *>(Rp + 8 * (r_lo + 8 * h));
            rb[h] = *reinterpret_cast<const float4 *>(Rp + 8 * (r_lo + 8 * h) + 4);
            pix[h] = s_pix[r_lo + 8 * h];
        }
        const int pool = s_pool[r_lo];
        // pooled grid coordinates of this 2x2 block
        float pgx = ra[0].y + ra[1].y, pgy = ra[0].z + ra[1].z;
        pgx += __shfl_xor_sync(0xffffffffu, pgx, 4);
        pgy += __shfl_xor_sync(0xffffffffu, pgy, 4);
        pgx *= 0.25f; pgy *= 0.25f;
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            float d2[2][2];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int cl = 32 * wn + 8 * ni + 2 * t + j;
                const float4 c0 = *reinterpret_cast<const float4 *>(Cp + CPS * cl);
                const float4 c1 = *reinterpret_cast<const float4 *>(Cp + CPS * cl + 4);
                const bool valid = c1.w != 0.f;
                const int n = n0 + cl;
                // gradient through the pooled similarity: shared by the four pixels of the block
                float dpool = 0.f;
                if (a.g_pool != nullptr && valid && pool >= 0) dpool = 0.25f * a.g_pool[((size_t)b * (geo.P >> 2) + pool) * geo.N + n];
                if (pooled_softmax) {
                    const float4 c2 = *reinterpret_cast<const float4 *>(Cp + CPS * cl + 8);
                    float q = (ra[0].x == 0.f ? -1e5f : acc[mi][ni][j]) + (ra[1].x == 0.f ? -1e5f : acc[mi][ni][2 + j]);
                    q += __shfl_xor_sync(0xffffffffu, q, 4);
                    const float ep = ex2_approx((0.25f * q - 1.f) * kexp);
                    if (valid) dpool += 0.25f * ep * c2.x * (c2.y * pgx + c2.z * pgy - c2.w);
                }
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    float d = 0.f;
                    if (ra[h].x != 0.f && valid) {
                        const float e = ex2_approx((acc[mi][ni][2 * h + j] - 1.f) * kexp);
                        const float row_term = ra[h].w * (rb[h].x * c0.x + rb[h].y * c0.y + rb[h].z * c0.z - rb[h].w);
                        const float col_term = c0.w * (c1.x * ra[h].y + c1.y * ra[h].z - c1.z);
                        d = e * (row_term + col_term) + dpool;
                        if (a.g_full != nullptr) d += a.g_full[((size_t)b * geo.P + pix[h]) * geo.N + n];
                    }
                    d2[h][j] = d;
                }
            }
#pragma unroll
            for (int h = 0; h < 2; h++)
                *reinterpret_cast<float2 *>(Ds + (r_lo + 8 * h) * DS + 32 * wn + 8 * ni + 2 * t) = make_float2(d2[h][0], d2[h][1]);
        }
    }
}

// backward, row-block CTAs: g_img_feat[c][p] = sum_n dS[p][n] mesh_feat[n][c]
constexpr int R_AS = 0;
constexpr int R_BS = R_AS + C * AS;          // 2 buffers
constexpr int R_DS = R_BS + 2 * BN * BS;
constexpr int R_RP = R_DS + BM * DS;         // [BM][8]
constexpr int R_CP = R_RP + BM * 8;          // 2 buffers [BN][8]
constexpr int R_IX = R_CP + 2 * BN * CPS;    // pix[BM], pool[BM] (int)
constexpr int R_TOTAL = R_IX + 2 * BM;

template <bool FUSE_COLS>
__global__ void __launch_bounds__(NT, 2) corr_bwd_rows_kernel(Geo geo, BwdArgs a)
{
    extern __shared__ __align__(16) float sm[];
    float *As = sm + R_AS, *Bs = sm + R_BS, *Ds = sm + R_DS, *Rp = sm + R_RP, *Cp = sm + R_CP;
    int *s_pix = reinterpret_cast<int *>(sm + R_IX), *s_pool = s_pix + BM;

    const int pblk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    if (pblk >= active_row_blocks(a.blocks + (size_t)b * ((geo.P >> 2) + 1))) return;
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3, wm = warp & 3, wn = warp >> 2;
    const float *img_b = a.img_feat + (size_t)b * C * geo.P;
    const float *mesh_b = a.mesh_feat + (size_t)b * geo.N * C;

    load_rowparams(geo, a, b, pblk, Rp, s_pix, s_pool);
    __syncthreads();
    load_A(geo, As, img_b, s_pix);
    load_B(geo, Bs, mesh_b, 0);
    cp_async_commit();
    load_colparams(geo, a, b, 0, Cp);

    float out[2][4][4];  // g_img tile: rows = pixels (32 per warp), cols = channels (32 per warp)
#pragma unroll
    for (int mi = 0; mi < 2; mi++)
#pragma unroll
        for (int ci = 0; ci < 4; ci++)
#pragma unroll
            for (int k = 0; k < 4; k++) out[mi][ci][k] = 0.f;

    for (int it = 0; it < geo.ntile; it++) {
        const int n0 = it * BN;
        float *Bt = Bs + (it & 1) * BN * BS, *Ct = Cp + (it & 1) * BN * CPS;
        cp_async_wait<0>();
        __syncthreads();  // tile `it` has landed; every warp is done with iteration it-1 (Ds, Bt reads)
        if (it + 1 < geo.ntile) {
            load_B(geo, Bs + ((it + 1) & 1) * BN * BS, mesh_b, n0 + BN);
            cp_async_commit();
            load_colparams(geo, a, b, n0 + BN, Cp + ((it + 1) & 1) * BN * CPS);
        }
        float acc[2][4][4];
        mma_S(acc, As, Bt, wm, wn, g, t);
        make_dS(geo, a, b, n0, acc, Rp, Ct, s_pix, s_pool, Ds, wm, wn, g, t);
        __syncthreads();
        // out[p][c] += dS[p][k=n] * mesh[k=n][c]   (A = Ds row-major, B[k][col] = Bt[k][col])
#pragma unroll
        for (int k0 = 0; k0 < BN; k0 += 8) {
            uint32_t af[2][4], bf[4][2];
#pragma unroll
            for (int mi = 0; mi < 2; mi++) {
                const float *p = Ds + (32 * wm + 16 * mi + g) * DS + k0 + t;
                af[mi][0] = f2tf32(p[0]); af[mi][1] = f2tf32(p[8 * DS]);
                af[mi][2] = f2tf32(p[4]); af[mi][3] = f2tf32(p[8 * DS + 4]);
            }
#pragma unroll
            for (int ci = 0; ci < 4; ci++) {
                const float *p = Bt + (k0 + t) * BS + 32 * wn + 8 * ci + g;
                bf[ci][0] = f2tf32(p[0]); bf[ci][1] = f2tf32(p[4 * BS]);
            }
#pragma unroll
            for (int mi = 0; mi < 2; mi++)
#pragma unroll
                for (int ci = 0; ci < 4; ci++) mma_tf32(out[mi][ci], af[mi], bf[ci]);
        }
        if (FUSE_COLS) {
            // the same dS tile also yields this row block's share of g_mesh_feat[n][c] = sum_p dS[p][n] img_feat[c][p]
            // (A[row = n][k = p] = Ds[k][row], B[k = p][col = c] = As[col][k]); reduced over the row blocks with global
            // float reductions into the zero-filled gradient -- the separate vertex-block kernel (a second recompute of
            // S and dS for every tile) is not launched
            float gm[4][4];
#pragma unroll
            for (int ci = 0; ci < 4; ci++)
#pragma unroll
                for (int k = 0; k < 4; k++) gm[ci][k] = 0.f;
#pragma unroll 4
            for (int k0 = 0; k0 < BM; k0 += 8) {
                uint32_t af[4], bf[4][2];
                const float *p = Ds + (k0 + t) * DS + 16 * wm + g;
                af[0] = f2tf32(p[0]); af[1] = f2tf32(p[8]);
                af[2] = f2tf32(p[4 * DS]); af[3] = f2tf32(p[4 * DS + 8]);
#pragma unroll
                for (int ci = 0; ci < 4; ci++) {
                    const float *q = As + (32 * wn + 8 * ci + g) * AS + k0 + t;
                    bf[ci][0] = f2tf32(q[0]); bf[ci][1] = f2tf32(q[4]);
                }
#pragma unroll
                for (int ci = 0; ci < 4; ci++) mma_tf32(gm[ci], af, bf[ci]);
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int n = n0 + 16 * wm + 8 * h + g;
                if (n < geo.N) {
#pragma unroll
                    for (int ci = 0; ci < 4; ci++) {
                        float *dst = a.g_mesh_feat + ((size_t)b * geo.N + n) * C + 32 * wn + 8 * ci + 2 * t;
                        atomicAdd(dst, gm[ci][2 * h]);
                        atomicAdd(dst + 1, gm[ci][2 * h + 1]);
                    }
                }
            }
        }
    }
    // g_img_feat[b][c][pixel(r)]
#pragma unroll
    for (int mi = 0; mi < 2; mi++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int pix = s_pix[32 * wm + 16 * mi + 8 * h + g];
            if (pix < 0) continue;
#pragma unroll
            for (int ci = 0; ci < 4; ci++)
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int c = 32 * wn + 8 * ci + 2 * t + j;
                    a.g_img_feat[((size_t)b * C + c) * geo.P + pix] = out[mi][ci][2 * h + j];
                }
        }
}

// backward, vertex-block CTAs: g_mesh_feat[n][c] = sum_p dS[p][n] img_feat[c][p]
constexpr int V_AS = 0;
constexpr int V_BS = V_AS + C * AS;
constexpr int V_DS = V_BS + BN * BS;
constexpr int V_RP = V_DS + BM * DS;
constexpr int V_CP = V_RP + BM * 8;
constexpr int V_IX = V_CP + BN * CPS;
constexpr int V_TOTAL = V_IX + 2 * BM;

__global__ void __launch_bounds__(NT, 2) corr_bwd_cols_kernel(Geo geo, BwdArgs a)
{
    extern __shared__ __align__(16) float sm[];
    float *As = sm + V_AS, *Bs = sm + V_BS, *Ds = sm + V_DS, *Rp = sm + V_RP, *Cp = sm + V_CP;
    int *s_pix = reinterpret_cast<int *>(sm + V_IX), *s_pool = s_pix + BM;

    const int nblk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, n0 = nblk * BN;
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3, wm = warp & 3, wn = warp >> 2;
    const float *img_b = a.img_feat + (size_t)b * C * geo.P;
    const float *mesh_b = a.mesh_feat + (size_t)b * geo.N * C;

    load_B(geo, Bs, mesh_b, n0);
    load_colparams(geo, a, b, n0, Cp);

    float out[4][4];  // g_mesh tile: rows = vertices (16 per warp: wm), cols = channels (32 per warp: wn)
#pragma unroll
    for (int ci = 0; ci < 4; ci++)
#pragma unroll
        for (int k = 0; k < 4; k++) out[ci][k] = 0.f;

    const int nact = active_row_blocks(a.blocks + (size_t)b * ((geo.P >> 2) + 1));
    for (int pblk = 0; pblk < nact; pblk++) {
        __syncthreads();  // previous iteration finished reading As / Ds / Rp
        load_rowparams(geo, a, b, pblk, Rp, s_pix, s_pool);
        __syncthreads();
        load_A(geo, As, img_b, s_pix);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        float acc[2][4][4];
        mma_S(acc, As, Bs, wm, wn, g, t);
        make_dS(geo, a, b, n0, acc, Rp, Cp, s_pix, s_pool, Ds, wm, wn, g, t);
        __syncthreads();
        // out[n][c] += dS^T[n][k=p] * img[k=p][c]   (A[row][k] = Ds[k][row], B[k][col] = As[col][k])
#pragma unroll 4
        for (int k0 = 0; k0 < BM; k0 += 8) {
            uint32_t af[4], bf[4][2];
            const float *p = Ds + (k0 + t) * DS + 16 * wm + g;
            af[0] = f2tf32(p[0]); af[1] = f2tf32(p[8]);
            af[2] = f2tf32(p[4 * DS]); af[3] = f2tf32(p[4 * DS + 8]);
#pragma unroll
            for (int ci = 0; ci < 4; ci++) {
                const float *q = As + (32 * wn + 8 * ci + g) * AS + k0 + t;
                bf[ci][0] = f2tf32(q[0]); bf[ci][1] = f2tf32(q[4]);
            }
#pragma unroll
            for (int ci = 0; ci < 4; ci++) mma_tf32(out[ci], af, bf[ci]);
        }
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int n = n0 + 16 * wm + 8 * h + g;
        if (n < geo.N) {
#pragma unroll
            for (int ci = 0; ci < 4; ci++) {
                const int c = 32 * wn + 8 * ci + 2 * t;
                *reinterpret_cast<float2 *>(a.g_mesh_feat + ((size_t)b * geo.N + n) * C + c) =
                    make_float2(out[ci][2 * h], out[ci][2 * h + 1]);
            }
        }
    }
}

// ---- foreground block list + constants of the dropped blocks ----------------------------------
// one CTA per image: ordered compaction of the 2x2 blocks with at least one foreground pixel; mean vertex
__global__ void __launch_bounds__(NT) corr_blocklist_kernel(Geo geo, const float *__restrict__ mask_down,
                                                            const float *__restrict__ pred_v, int *__restrict__ blocks,
                                                            float *__restrict__ vmean)
{
    __shared__ int s_cnt[NT / 32];
    __shared__ float s_red[3 * (NT / 32)];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = geo.P >> 2, w2 = geo.wf >> 1;
    const float *m = mask_down + (size_t)b * geo.P;
    int *blk = blocks + (size_t)b * (nblk + 1);
    int base = 0;
    for (int i0 = 0; i0 < nblk; i0 += NT) {
        const int id = i0 + tid;
        bool keep = false;
        if (id < nblk) {
            const int by = id / w2, bx = id - by * w2, p = 2 * by * geo.wf + 2 * bx;
            keep = m[p] != 0.f || m[p + 1] != 0.f || m[p + geo.wf] != 0.f || m[p + geo.wf + 1] != 0.f;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_cnt[warp] = __popc(bal);
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; w++) {
            if (w == warp) off = total;
            total += s_cnt[w];
        }
        if (keep) blk[1 + base + off + __popc(bal & ((1u << lane) - 1u))] = id;
        base += total;
        __syncthreads();
    }
    if (tid == 0) blk[0] = base;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int n = tid; n < geo.N; n += NT) {
        const float *v = pred_v + ((size_t)b * geo.N + n) * 3;
        sx += v[0]; sy += v[1]; sz += v[2];
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    if (lane == 0) { s_red[warp] = sx; s_red[NT / 32 + warp] = sy; s_red[2 * (NT / 32) + warp] = sz; }
    __syncthreads();
    if (tid < 3) {
        float t = 0.f;
        for (int w = 0; w < NT / 32; w++) t += s_red[tid * (NT / 32) + w];
        vmean[b * 4 + tid] = t / (float)geo.N;
    }
}

// one warp per 2x2 block that is NOT on the list: uniform row softmax -> match = mean vertex, rsum = N;
// pointcorr rows (pooled / full) = -1e5 exactly (correspondence.py:44,48)
__global__ void __launch_bounds__(NT) corr_fill_kernel(Geo geo, const float *__restrict__ mask_down,
                                                       const float *__restrict__ vmean, float *__restrict__ pc_full,
                                                       float *__restrict__ pc_pool, float *__restrict__ match,
                                                       float *__restrict__ rsum)
{
    const int b = blockIdx.y, lane = threadIdx.x & 31, id = blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    const int nblk = geo.P >> 2, w2 = geo.wf >> 1;
    if (id >= nblk) return;
    const float *m = mask_down + (size_t)b * geo.P;
    const int by = id / w2, bx = id - by * w2, p0 = 2 * by * geo.wf + 2 * bx;
    if (m[p0] != 0.f || m[p0 + 1] != 0.f || m[p0 + geo.wf] != 0.f || m[p0 + geo.wf + 1] != 0.f) return;
    if (pc_pool != nullptr) {
        float *row = pc_pool + ((size_t)b * nblk + id) * geo.N;
        for (int n = lane; n < geo.N; n += 32) row[n] = -1e5f;
    }
    const int px[4] = { p0, p0 + 1, p0 + geo.wf, p0 + geo.wf + 1 };
    if (pc_full != nullptr) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float *row = pc_full + ((size_t)b * geo.P + px[k]) * geo.N;
            for (int n = lane; n < geo.N; n += 32) row[n] = -1e5f;
        }
    }
    if (lane < 4) {
        const size_t bp = (size_t)b * geo.P + px[lane];
        match[bp * 3 + 0] = vmean[b * 4 + 0];
        match[bp * 3 + 1] = vmean[b * 4 + 1];
        match[bp * 3 + 2] = vmean[b * 4 + 2];
        rsum[bp] = (float)geo.N;
    }
}

static size_t blocks_bytes(int B, int P) { return ((size_t)B * ((P >> 2) + 1) * sizeof(int) + 255) / 256 * 256; }
static size_t vmean_bytes(int B) { return ((size_t)B * 4 * sizeof(float) + 255) / 256 * 256; }

static bool make_geo(Geo &g, int B, int hf, int wf, int N, int Cc, float tau)
{
    if (B <= 0 || B > 65535 || N <= 0 || Cc != C) return false;
    if (!(wf == 8 || wf == 16 || wf == 32 || wf == 64) || hf <= 0) return false;
    if ((hf * wf) % BM != 0 || ((BM / wf) & 1)) return false;
    g.B = B; g.P = hf * wf; g.N = N; g.hf = hf; g.wf = wf; g.tau = tau;
    g.npblk = g.P / BM;
    g.ntile = (N + BN - 1) / BN;
    return true;
}

}  // namespace corr

// scp_corr_tc.cu: the training-mode forward on tcgen05 (S in tensor memory, soft-max statistics in the GEMM epilogue)
namespace corr_tc {
bool eligible(int B, int hf, int wf, int N, int Cc);
size_t workspace_bytes(int B, int hf, int wf, int N);
int forward(const float *img_feat, const float *mesh_feat, const float *mask_down, const float *pred_v,
            const float *meshgrid, float tau, int B, int hf, int wf, int N, float *pc_pool, float *match,
            float *imatch, float *rsum, float *csum, float *A_pool, float *csum_pool, const int *blocks,
            const float *vmean, void *ws, cudaStream_t st);
}  // namespace corr_tc
}  // namespace scp

using namespace scp::corr;

// SCP_CORR_FWD=legacy keeps the mma.sync forward for every call (A/B runs, parity tests of both paths)
static bool tc_forward_enabled()
{
    const char *m = getenv("SCP_CORR_FWD");
    return !(m && m[0] == 'l');
}
static size_t legacy_ws_bytes(int B, int hf, int wf, int N)
{
    // foreground block lists, mean vertices, full-res + pooled column partials
    return blocks_bytes(B, hf * wf) + vmean_bytes(B) + 2 * (size_t)B * ((size_t)hf * wf / BM) * N * 4 * sizeof(float);
}

extern "C" size_t scp_corr_workspace_bytes(int B, int hf, int wf, int N)
{
    if (B <= 0 || hf <= 0 || wf <= 0 || N <= 0) return 0;
    size_t n = legacy_ws_bytes(B, hf, wf, N);
    if (scp::corr_tc::eligible(B, hf, wf, N, C)) {   // the tensor-core forward keeps the block lists / mean vertices in front
        const size_t t = blocks_bytes(B, hf * wf) + vmean_bytes(B) + scp::corr_tc::workspace_bytes(B, hf, wf, N);
        if (t > n) n = t;
    }
    return n;
}

extern "C" size_t scp_corr_backward_workspace_bytes(int B, int hf, int wf, int N)
{
    if (B <= 0 || hf <= 0 || wf <= 0 || N <= 0) return 0;
    return blocks_bytes(B, hf * wf) + vmean_bytes(B);
}

extern "C" int scp_corr_match_forward(const float *img_feat, const float *mesh_feat, const float *mask_down,
                                      const float *pred_v, const float *meshgrid, float tau, int B, int hf, int wf,
                                      int N, int Cc, float *pointcorr_full, float *pointcorr_pool, float *match,
                                      float *imatch, float *rsum, float *csum, float *A_pool, float *csum_pool,
                                      void *workspace, size_t workspace_bytes, void *stream)
{
    Geo geo;
    if (!make_geo(geo, B, hf, wf, N, Cc, tau)) {
        scp::set_last_error("scp_corr_match_forward: unsupported shape (B=%d hf=%d wf=%d N=%d C=%d; need C=64, "
                            "wf in {8,16,32,64}, hf*wf %% 128 == 0)", B, hf, wf, N, Cc);
        return -1;
    }
    if (!(tau > 0.f) || tau > 40.f) {
        // both soft-maxes use the fixed reference point exp(tau * (S - 1)) (valid for L2-normalised features, |S| <= 1):
        // for tau * 2 * log2(e) > 126 an entry with S = -1 underflows ex2.approx.ftz to 0
        scp::set_last_error("scp_corr_match_forward: tau = %g outside (0, 40] (fixed-reference-point soft-max; the shipped "
                            "configs use 10)", (double)tau);
        return -1;
    }
    if (!workspace || workspace_bytes < scp_corr_workspace_bytes(B, hf, wf, N)) {
        scp::set_last_error("scp_corr_match_forward: workspace too small");
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = F_TOTAL * sizeof(float);
    cudaFuncSetAttribute(corr_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (A_pool != nullptr && (pointcorr_pool == nullptr || csum_pool == nullptr)) {
        scp::set_last_error("scp_corr_match_forward: A_pool needs pointcorr_pool and csum_pool");
        return -1;
    }
    int *blocks = (int *)workspace;
    float *vmean = (float *)((char *)workspace + blocks_bytes(B, geo.P));
    float *part = (float *)((char *)vmean + vmean_bytes(B));
    float *part_pool = A_pool != nullptr ? part + (size_t)B * geo.npblk * N * 4 : nullptr;
    corr_blocklist_kernel<<<B, NT, 0, st>>>(geo, mask_down, pred_v, blocks, vmean);
    corr_fill_kernel<<<dim3(((geo.P >> 2) + NT / 32 - 1) / (NT / 32), B), NT, 0, st>>>(geo, mask_down, vmean, pointcorr_full,
                                                                                 pointcorr_pool, match, rsum);
    if (pointcorr_full == nullptr && tc_forward_enabled() && scp::corr_tc::eligible(B, hf, wf, N, Cc)) {
        // training path: similarity on tcgen05, accumulator in tensor memory, both soft-maxes in the GEMM epilogue
        const int rc = scp::corr_tc::forward(img_feat, mesh_feat, mask_down, pred_v, meshgrid, tau, B, hf, wf, N, pointcorr_pool,
                                             match, imatch, rsum, csum, A_pool, csum_pool, blocks, vmean, (char *)part, st);
        if (rc != 0) return rc;
        return scp::check_launch("scp_corr_match_forward (tcgen05)");
    }
    corr_fwd_kernel<<<dim3(geo.npblk, B), NT, smem, st>>>(geo, img_feat, mesh_feat, mask_down, pred_v, meshgrid,
                                                         pointcorr_full, pointcorr_pool, match, rsum, part, part_pool,
                                                         blocks);
    corr_colreduce_kernel<<<dim3((N + 127) / 128, B), 128, 0, st>>>(geo, part, meshgrid, imatch, csum, blocks);
    if (A_pool != nullptr)
        corr_colreduce_kernel<<<dim3((N + 127) / 128, B), 128, 0, st>>>(geo, part_pool, meshgrid, A_pool, csum_pool, blocks);
    return scp::check_launch("scp_corr_match_forward");
}

extern "C" int scp_corr_match_backward(const float *img_feat, const float *mesh_feat, const float *mask_down,
                                       const float *pred_v, const float *meshgrid, float tau, int B, int hf, int wf,
                                       int N, int Cc, const float *match, const float *imatch, const float *rsum,
                                       const float *csum, const float *g_match, const float *g_imatch,
                                       const float *g_pointcorr_pool, const float *g_pointcorr_full,
                                       const float *A_pool, const float *csum_pool, const float *g_A_pool,
                                       float *g_img_feat, float *g_mesh_feat, void *workspace, size_t workspace_bytes,
                                       void *stream)
{
    Geo geo;
    if (!make_geo(geo, B, hf, wf, N, Cc, tau)) {
        scp::set_last_error("scp_corr_match_backward: unsupported shape");
        return -1;
    }
    if (!workspace || workspace_bytes < blocks_bytes(B, geo.P) + vmean_bytes(B)) {
        scp::set_last_error("scp_corr_match_backward: workspace too small");
        return -1;
    }
    BwdArgs a;
    a.img_feat = img_feat; a.mesh_feat = mesh_feat; a.mask_down = mask_down; a.pred_v = pred_v; a.meshgrid = meshgrid;
    a.match = match; a.imatch = imatch; a.rsum = rsum; a.csum = csum;
    a.g_match = g_match; a.g_imatch = g_imatch; a.g_pool = g_pointcorr_pool; a.g_full = g_pointcorr_full;
    a.A_pool = A_pool; a.csum_pool = csum_pool; a.g_A_pool = (A_pool && csum_pool) ? g_A_pool : nullptr;
    a.g_img_feat = g_img_feat; a.g_mesh_feat = g_mesh_feat;
    cudaStream_t st = (cudaStream_t)stream;
    int *blocks = (int *)workspace;   // rebuilt here: the backward keeps no state besides the saved tensors
    corr_blocklist_kernel<<<B, NT, 0, st>>>(geo, mask_down, pred_v, blocks, (float *)((char *)workspace + blocks_bytes(B, geo.P)));
    a.blocks = blocks;
    cudaMemsetAsync(g_img_feat, 0, (size_t)B * C * geo.P * sizeof(float), st);   // background pixels: zero gradient
    const size_t smem_r = R_TOTAL * sizeof(float), smem_v = V_TOTAL * sizeof(float);
    // SCP_CORR_BWD=split: separate vertex-block kernel (deterministic summation order) instead of the fused reductions
    const char *mode = getenv("SCP_CORR_BWD");
    if (mode && mode[0] == 's') {
        cudaFuncSetAttribute(corr_bwd_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r);
        cudaFuncSetAttribute(corr_bwd_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v);
        corr_bwd_rows_kernel<false><<<dim3(geo.npblk, B), NT, smem_r, st>>>(geo, a);
        corr_bwd_cols_kernel<<<dim3(geo.ntile, B), NT, smem_v, st>>>(geo, a);
    } else {
        cudaFuncSetAttribute(corr_bwd_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r);
        cudaMemsetAsync(g_mesh_feat, 0, (size_t)B * N * C * sizeof(float), st);
        corr_bwd_rows_kernel<true><<<dim3(geo.npblk, B), NT, smem_r, st>>>(geo, a);
    }
    return scp::check_launch("scp_corr_match_backward");
}
