// Pre-training cycle loss, target-row part (PretrainedCorrespondence.compute_cycle_loss of the reference,
// model/module/pretrained_corr.py:120-139), fused forward + backward.
//
// For every image pair (src, tgt) only the k pseudo-matched target pixels j are consumed (:137), and with
//   Pm = softmax over pixels of tau * pointcorr_src (gated by depth_weight_src >= 0.5), A = grid . Pm  (from the
//   correspondence kernel), Pi_j = softmax over vertices of tau * pointcorr_tgt[j, :] (gated by depth_weight_tgt)
// the cycle point is  match_j = sum_n A[:, n] w_n Pi_j[n] / (sum_n w_n Pi_j[n] + 1e-5),  w = gate_src * gate_tgt
// (note A: the 1024 x 1024 `corr` of :130-131 is never formed).  The loss is
// mean(|match_j - pts_src_j|_2 * mask_j).
//
// One CTA per pair, one warp per gathered row: the row is read straight from the per-image pooled pointcorr
// (no (2B, k, N) gather), softmax statistics and the three weighted sums are warp-shuffle reductions, and the
// backward recomputes them and scatters d(row) with reductions into a zero-filled pointcorr gradient (rows picked by
// both pairings of an image add up) -- replaces index_select / softmax / gate / bmm / div / norm and their
// autograd (index_add, softmax backward, four bmm) of the op-by-op formulation.  HBM-bound: 2B * k rows of N floats.
#include "../../include/scp_b200.h"
#include "scp_common.cuh"

namespace scp {
namespace cycle {

constexpr int NT = 256, NW = NT / 32;
constexpr float LOG2E = 1.4426950408889634f;

struct Args {
    const float *pc, *A, *dw;                 // [B,P4,N], [B,2,N], [B,N]
    const long long *src_idx, *tgt_idx, *rows;   // [NP], [NP], [NP,k]
    const float *pts, *mask_k;                // [NP,2,k], [NP,k]
    float tau;
    int B, P4, N, NP, k;
};

struct RowStats {
    float m, S, numx, numy, den, mx, my, d;
};

// softmax statistics and the cycle point of one gathered row (whole warp)
__device__ __forceinline__ RowStats row_forward(const Args &a, const float *__restrict__ row, const float *sAx,
                                                const float *sAy, const float *sw, float px, float py, int lane)
{
    RowStats r;
    float m = -3.0e38f;
    for (int n = lane; n < a.N; n += 32) m = fmaxf(m, __ldg(row + n));
    r.m = warp_max(m);
    const float kk = a.tau * LOG2E;
    float S = 0.f, Sx = 0.f, Sy = 0.f, Sw = 0.f;
    for (int n = lane; n < a.N; n += 32) {
        const float e = ex2_approx((__ldg(row + n) - r.m) * kk);
        S += e; Sx += e * sAx[n]; Sy += e * sAy[n]; Sw += e * sw[n];
    }
    r.S = warp_sum(S);
    const float inv = 1.f / r.S;
    r.numx = warp_sum(Sx) * inv;
    r.numy = warp_sum(Sy) * inv;
    r.den = warp_sum(Sw) * inv + 1e-5f;
    r.mx = r.numx / r.den;
    r.my = r.numy / r.den;
    const float ex = r.mx - px, ey = r.my - py;
    r.d = sqrtf(ex * ex + ey * ey);
    return r;
}

__device__ __forceinline__ void load_pair(const Args &a, int pair, int &src, int &tgt, float *sAx, float *sAy, float *sw)
{
    src = (int)a.src_idx[pair];
    tgt = (int)a.tgt_idx[pair];
    for (int n = threadIdx.x; n < a.N; n += NT) {
        const float w = (a.dw[(size_t)src * a.N + n] >= 0.5f && a.dw[(size_t)tgt * a.N + n] >= 0.5f) ? 1.f : 0.f;
        sw[n] = w;
        sAx[n] = a.A[((size_t)src * 2 + 0) * a.N + n] * w;
        sAy[n] = a.A[((size_t)src * 2 + 1) * a.N + n] * w;
    }
}

__global__ void __launch_bounds__(NT) cycle_rows_fwd_kernel(Args a, float *__restrict__ pair_loss, float *__restrict__ match)
{
    extern __shared__ float sm[];
    float *sAx = sm, *sAy = sm + a.N, *sw = sm + 2 * a.N;
    __shared__ float s_part[NW];
    const int pair = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int src, tgt;
    load_pair(a, pair, src, tgt, sAx, sAy, sw);
    __syncthreads();
    float acc = 0.f;
    for (int j = warp; j < a.k; j += NW) {
        const size_t pj = (size_t)pair * a.k + j;
        const float px = a.pts[((size_t)pair * 2 + 0) * a.k + j], py = a.pts[((size_t)pair * 2 + 1) * a.k + j];
        const float *row = a.pc + ((size_t)tgt * a.P4 + a.rows[pj]) * a.N;
        const RowStats r = row_forward(a, row, sAx, sAy, sw, px, py, lane);
        if (lane == 0) {
            match[((size_t)pair * 2 + 0) * a.k + j] = r.mx;
            match[((size_t)pair * 2 + 1) * a.k + j] = r.my;
        }
        acc += r.d * a.mask_k[pj];
    }
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < NW; w++) t += s_part[w];
        pair_loss[pair] = t;
    }
}

__global__ void __launch_bounds__(NT) cycle_rows_bwd_kernel(Args a, const float *__restrict__ g_pair,
                                                            float *__restrict__ g_pc, float *__restrict__ g_A)
{
    extern __shared__ float sm[];
    float *sAx = sm, *sAy = sm + a.N, *sw = sm + 2 * a.N, *gAx = sm + 3 * a.N, *gAy = sm + 4 * a.N;
    const int pair = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int src, tgt;
    load_pair(a, pair, src, tgt, sAx, sAy, sw);
    for (int n = threadIdx.x; n < a.N; n += NT) gAx[n] = gAy[n] = 0.f;
    __syncthreads();
    const float g = g_pair[pair];
    const float kk = a.tau * LOG2E;
    for (int j = warp; j < a.k; j += NW) {
        const size_t pj = (size_t)pair * a.k + j;
        const float mk = a.mask_k[pj];
        if (mk == 0.f || g == 0.f) continue;                       // masked rows carry no gradient
        const float px = a.pts[((size_t)pair * 2 + 0) * a.k + j], py = a.pts[((size_t)pair * 2 + 1) * a.k + j];
        const size_t roff = ((size_t)tgt * a.P4 + a.rows[pj]) * a.N;
        const float *row = a.pc + roff;
        const RowStats r = row_forward(a, row, sAx, sAy, sw, px, py, lane);
        if (r.d == 0.f) continue;                                    // norm backward at 0: zero subgradient
        const float dmx = g * mk * (r.mx - px) / r.d, dmy = g * mk * (r.my - py) / r.d;
        const float dnx = dmx / r.den, dny = dmy / r.den;
        const float dden = -(dmx * r.numx + dmy * r.numy) / (r.den * r.den);
        const float dot = dnx * r.numx + dny * r.numy + dden * (r.den - 1e-5f);   // sum_n Pi_n dPi_n
        const float inv = 1.f / r.S;
        float *grow = g_pc + roff;
        for (int n = lane; n < a.N; n += 32) {
            const float pi = ex2_approx((__ldg(row + n) - r.m) * kk) * inv;
            const float w = sw[n];
            const float dpi = dnx * sAx[n] + dny * sAy[n] + dden * w;
            atomicAdd(grow + n, a.tau * pi * (dpi - dot));
            if (w != 0.f) {
                atomicAdd(gAx + n, pi * dnx);
                atomicAdd(gAy + n, pi * dny);
            }
        }
    }
    __syncthreads();
    for (int n = threadIdx.x; n < a.N; n += NT) {
        if (gAx[n] != 0.f) atomicAdd(g_A + ((size_t)src * 2 + 0) * a.N + n, gAx[n]);
        if (gAy[n] != 0.f) atomicAdd(g_A + ((size_t)src * 2 + 1) * a.N + n, gAy[n]);
    }
}

static bool fill(Args &a, const float *pc, const float *A, const float *dw, const long long *src_idx,
                 const long long *tgt_idx, const long long *rows, const float *pts, const float *mask_k, float tau, int B,
                 int P4, int N, int NP, int k)
{
    if (!pc || !A || !dw || !src_idx || !tgt_idx || !rows || !pts || !mask_k) return false;
    if (B <= 0 || P4 <= 0 || N <= 0 || N > 8192 || NP <= 0 || k <= 0) return false;
    a = Args{ pc, A, dw, src_idx, tgt_idx, rows, pts, mask_k, tau, B, P4, N, NP, k };
    return true;
}

}  // namespace cycle
}  // namespace scp

using namespace scp::cycle;

extern "C" int scp_cycle_rows_forward(const float *pointcorr_pool, const float *A_pool, const float *depth_weight,
                                      const long long *src_idx, const long long *tgt_idx, const long long *rows,
                                      const float *pts_src, const float *mask_k, float tau, int B, int P4, int N, int NP,
                                      int k, float *pair_loss, float *match, void *stream)
{
    Args a;
    if (!fill(a, pointcorr_pool, A_pool, depth_weight, src_idx, tgt_idx, rows, pts_src, mask_k, tau, B, P4, N, NP, k) ||
        !pair_loss || !match) {
        scp::set_last_error("scp_cycle_rows_forward: bad arguments (B=%d P4=%d N=%d NP=%d k=%d)", B, P4, N, NP, k);
        return -1;
    }
    const size_t smem = (size_t)3 * N * sizeof(float);
    cudaFuncSetAttribute(cycle_rows_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cycle_rows_fwd_kernel<<<NP, NT, smem, (cudaStream_t)stream>>>(a, pair_loss, match);
    return scp::check_launch("scp_cycle_rows_forward");
}

extern "C" int scp_cycle_rows_backward(const float *pointcorr_pool, const float *A_pool, const float *depth_weight,
                                       const long long *src_idx, const long long *tgt_idx, const long long *rows,
                                       const float *pts_src, const float *mask_k, float tau, int B, int P4, int N, int NP,
                                       int k, const float *g_pair_loss, float *g_pointcorr_pool, float *g_A_pool,
                                       void *stream)
{
    Args a;
    if (!fill(a, pointcorr_pool, A_pool, depth_weight, src_idx, tgt_idx, rows, pts_src, mask_k, tau, B, P4, N, NP, k) ||
        !g_pair_loss || !g_pointcorr_pool || !g_A_pool) {
        scp::set_last_error("scp_cycle_rows_backward: bad arguments");
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(g_pointcorr_pool, 0, (size_t)B * P4 * N * sizeof(float), st);
    cudaMemsetAsync(g_A_pool, 0, (size_t)B * 2 * N * sizeof(float), st);
    const size_t smem = (size_t)5 * N * sizeof(float);
    cudaFuncSetAttribute(cycle_rows_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cycle_rows_bwd_kernel<<<NP, NT, smem, st>>>(a, g_pair_loss, g_pointcorr_pool, g_A_pool);
    return scp::check_launch("scp_cycle_rows_backward");
}
