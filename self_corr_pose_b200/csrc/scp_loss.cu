// Image-space loss terms of the training step, fused: silhouette pyramid, texture, depth and 3D-match losses
// (model/util/loss_utils.py:236-252, :273-284, :317-320 of the reference) in three launches -- forward values
// per image, and one backward kernel that writes the gradients straight into the (B,4,H,W) gradient tensors the
// SoftRas backward consumes and into the low-resolution 3D match of the correspondence kernel.  Replaces ~150
// element-wise / pooling / reduction launches of the op-by-op formulation.
//
// HBM-bound: every map is read once per kernel with 16-byte loads; a thread owns 16 consecutive pixels of one
// image row, which is exactly one block of the coarsest level of the row-wise silhouette pyramid (the 3-D maps make
// F.interpolate(mode='area') pool along the last axis only, loss_utils.py:240-242).
#include "../../include/scp_b200.h"
#include "scp_common.cuh"

namespace scp {
namespace loss {

constexpr int SEG = 16, NT = 256;

struct View {            // one (B, [3,] H, W) map: plane contiguous, channel stride H*W, batch stride `bs` floats
    const float *p;
    long long bs;
};
struct GView {
    float *p;
    long long bs;
};
enum { IMG = 0, MASK, DEPTH, MASK_RENDER, TEX_RENDER, TEX_MASK, DEPTH_RENDER, DEPTH_MASK, MATCH_GT, MATCH_MASK, MATCH_FULL, NIN };
enum { G_MASK_RENDER = 0, G_TEX_RENDER, G_TEX_MASK, G_DEPTH_RENDER, G_MATCH_FULL, NOUT };

struct Args {
    View in[NIN];
    const float *match_lr;   // [B, hf*wf, 3] (hf > 0) -- nearest-upsampled on the fly; else in[MATCH_FULL] is used
    int B, H, W, hf, wf, use_depth;
};

__device__ __forceinline__ void load16(const float *p, float (&v)[SEG])
{
    const float4 *q = reinterpret_cast<const float4 *>(p);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float4 a = __ldg(q + i);
        v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
    }
}
__device__ __forceinline__ void store16(float *p, const float (&v)[SEG])
{
    float4 *q = reinterpret_cast<float4 *>(p);
#pragma unroll
    for (int i = 0; i < 4; i++) q[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

template <int N, class T>
__device__ __forceinline__ void block_sum(T (&v)[N], T *smem /* N * 8 */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if (lane == 0) smem[k * (NT / 32) + warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < N; k++) {
            T t = lane < NT / 32 ? smem[k * (NT / 32) + lane] : T(0);
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            v[k] = t;
        }
    }
    __syncthreads();
}

// batch-global sums of the depth scale (loss_utils.py:276): ws[0..3] = sum pred, n pred, sum gt, n gt
__global__ void __launch_bounds__(NT) depth_sums_kernel(Args a, double *ws)
{
    __shared__ double red[4 * (NT / 32)];
    const int b = blockIdx.y, hw = a.H * a.W;
    const long base = ((long)blockIdx.x * NT + threadIdx.x) * SEG;
    double s[4] = { 0., 0., 0., 0. };
    if (base < hw) {
        float dr[SEG], dm[SEG], d[SEG], m[SEG];
        load16(a.in[DEPTH_RENDER].p + b * a.in[DEPTH_RENDER].bs + base, dr);
        load16(a.in[DEPTH_MASK].p + b * a.in[DEPTH_MASK].bs + base, dm);
        load16(a.in[DEPTH].p + b * a.in[DEPTH].bs + base, d);
        load16(a.in[MASK].p + b * a.in[MASK].bs + base, m);
        float sp = 0.f, np = 0.f, sg = 0.f, ng = 0.f;
#pragma unroll
        for (int j = 0; j < SEG; j++) {
            if (dm[j] != 0.f) { sp += dr[j]; np += 1.f; }
            if (m[j] * d[j] != 0.f) { sg += d[j]; ng += 1.f; }
        }
        s[0] = sp; s[1] = np; s[2] = sg; s[3] = ng;
    }
    block_sum<4>(s, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 4; k++) atomicAdd(ws + k, s[k]);
    }
}

__device__ __forceinline__ float depth_scale(const double *ws)
{
    return (float)((ws[0] / ws[1]) / (ws[2] / ws[3]));
}

// silhouette pyramid on one 16-pixel block: returns sum over levels of (block size) * mean^2 and, per pixel, the sum
// over levels of the block mean containing it (the gradient of the former is twice the latter)
__device__ __forceinline__ float pyramid16(const float (&d)[SEG], float (&gsum)[SEG])
{
    float tot = 0.f;
    float m1[8], m2[4], m3[2];
#pragma unroll
    for (int j = 0; j < 16; j++) tot += d[j] * d[j];
#pragma unroll
    for (int j = 0; j < 8; j++) { m1[j] = 0.5f * (d[2 * j] + d[2 * j + 1]); tot += 2.f * m1[j] * m1[j]; }
#pragma unroll
    for (int j = 0; j < 4; j++) { m2[j] = 0.5f * (m1[2 * j] + m1[2 * j + 1]); tot += 4.f * m2[j] * m2[j]; }
#pragma unroll
    for (int j = 0; j < 2; j++) { m3[j] = 0.5f * (m2[2 * j] + m2[2 * j + 1]); tot += 8.f * m3[j] * m3[j]; }
    const float m4 = 0.5f * (m3[0] + m3[1]);
    tot += 16.f * m4 * m4;
#pragma unroll
    for (int j = 0; j < 16; j++) gsum[j] = d[j] + m1[j >> 1] + m2[j >> 2] + m3[j >> 3] + m4;
    return tot;
}

// BWD = false: per-image loss sums (losses[b][0..3] += ..., ws[4 + b] += depth-scale coupling term)
// BWD = true : gradients, given the upstream gradient of every per-image loss value
template <bool BWD>
__global__ void __launch_bounds__(NT) image_loss_kernel(Args a, float *losses, double *ws, const float *g_losses,
                                                        GView g0, GView g1, GView g2, GView g3, GView g4, float *g_match_lr)
{
    __shared__ double redd[NT / 32];
    __shared__ float redf[4 * (NT / 32)];
    __shared__ float s_k;
    const int b = blockIdx.y, hw = a.H * a.W;
    const long base = ((long)blockIdx.x * NT + threadIdx.x) * SEG;
    const bool act = base < hw;
    const float inv_hw = 1.f / (float)hw;
    const long P1 = hw;   // channel stride
    float gl[4] = { 0.f, 0.f, 0.f, 0.f };
    if (BWD) {
#pragma unroll
        for (int k = 0; k < 4; k++) gl[k] = g_losses[b * 4 + k] * inv_hw;
        if (a.use_depth) {   // dL/dscale = sum_b g[b][depth] * T_b, then d scale / d pred_i = sel_i / (Np * Sg / Ng)
            double part[1] = { 0. };
            for (int i = threadIdx.x; i < a.B; i += NT) part[0] += (double)g_losses[i * 4 + 2] * ws[4 + i];
            block_sum<1>(part, redd);
            if (threadIdx.x == 0) s_k = (float)(part[0] / (ws[1] * (ws[2] / ws[3])));
            __syncthreads();
        }
    }
    float acc[4] = { 0.f, 0.f, 0.f, 0.f };
    double tb[1] = { 0. };
    if (act) {
        float mask[SEG];
        load16(a.in[MASK].p + b * a.in[MASK].bs + base, mask);
        // ---- silhouette pyramid (loss_utils.py:236-244)
        {
            float mr[SEG], d[SEG], gs[SEG];
            load16(a.in[MASK_RENDER].p + b * a.in[MASK_RENDER].bs + base, mr);
#pragma unroll
            for (int j = 0; j < SEG; j++) d[j] = mr[j] - mask[j];
            const float tot = pyramid16(d, gs);
            if (!BWD) acc[0] = 0.2f * tot;
            else {
#pragma unroll
                for (int j = 0; j < SEG; j++) gs[j] *= 0.4f * gl[0];
                store16(g0.p + b * g0.bs + base, gs);
            }
        }
        // ---- texture (loss_utils.py:246-252)
        {
            float tm[SEG], gtm[SEG];
            load16(a.in[TEX_MASK].p + b * a.in[TEX_MASK].bs + base, tm);
#pragma unroll
            for (int j = 0; j < SEG; j++) gtm[j] = 0.f;
            float lt = 0.f;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                float im[SEG], tr[SEG], gtr[SEG];
                load16(a.in[IMG].p + b * a.in[IMG].bs + c * P1 + base, im);
                load16(a.in[TEX_RENDER].p + b * a.in[TEX_RENDER].bs + c * P1 + base, tr);
#pragma unroll
                for (int j = 0; j < SEG; j++) {
                    const float fg = mask[j] > 0.f ? 1.f : 0.f;
                    const float gt = im[j] * fg, white = 1.f - fg + gt;
                    const float e = gt - tr[j] * tm[j], w = white - tr[j];
                    if (!BWD) lt += 0.75f * e * e + fabsf(w) * (1.f / 3.f);
                    else {
                        const float sg = w > 0.f ? -1.f : (w < 0.f ? 1.f : 0.f);   // d|white - t| / dt
                        gtr[j] = gl[1] * (-1.5f * e * tm[j] + sg * (1.f / 3.f));
                        gtm[j] += gl[1] * (-1.5f * e * tr[j]);
                    }
                }
                if (BWD) store16(g1.p + b * g1.bs + c * P1 + base, gtr);
            }
            if (!BWD) acc[1] = lt;
            else store16(g2.p + b * g2.bs + base, gtm);
        }
        // ---- depth (loss_utils.py:273-284)
        if (a.use_depth) {
            float dr[SEG], dm[SEG], d[SEG], gd[SEG];
            load16(a.in[DEPTH_RENDER].p + b * a.in[DEPTH_RENDER].bs + base, dr);
            load16(a.in[DEPTH_MASK].p + b * a.in[DEPTH_MASK].bs + base, dm);
            load16(a.in[DEPTH].p + b * a.in[DEPTH].bs + base, d);
            const float sc = depth_scale(ws);
            float ld = 0.f, t = 0.f;
#pragma unroll
            for (int j = 0; j < SEG; j++) {
                const bool drop = (mask[j] * dm[j] == 0.f) || (d[j] == 0.f);
                const float diff = drop ? 0.f : dr[j] - sc * d[j];
                const float sq = diff * diff;
                const float dsq = sq < 1.f ? 2.f * diff : 0.f;   // d min(diff^2, 1) / d diff  (relu'(0) = 0)
                if (!BWD) { ld += fminf(sq, 1.f); t += dsq * -d[j]; }
                else gd[j] = gl[2] * dsq + (dm[j] != 0.f ? s_k : 0.f);
            }
            if (!BWD) { acc[2] = ld; tb[0] = (double)t * inv_hw; }
            else store16(g3.p + b * g3.bs + base, gd);
        }
        // ---- 3D match (loss_utils.py:317-320); match is the nearest-upsampled (hf, wf) map (correspondence.py:71)
        {
            float mm[SEG];
            load16(a.in[MATCH_MASK].p + b * a.in[MATCH_MASK].bs + base, mm);
            const int y = (int)(base / a.W), x0 = (int)(base - (long)y * a.W);
            const bool lr = a.hf > 0;
            const int sx = lr ? a.W / a.wf : 1, sy = lr ? a.H / a.hf : 1;
            float gt[3][SEG], mf[3][SEG];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                load16(a.in[MATCH_GT].p + b * a.in[MATCH_GT].bs + c * P1 + base, gt[c]);
                if (!lr) load16(a.in[MATCH_FULL].p + b * a.in[MATCH_FULL].bs + c * P1 + base, mf[c]);
            }
            const float *mrow = lr ? a.match_lr + ((long)b * a.hf * a.wf + (long)(y / sy) * a.wf) * 3 : nullptr;
            float *grow = (BWD && lr) ? g_match_lr + ((long)b * a.hf * a.wf + (long)(y / sy) * a.wf) * 3 : nullptr;
            float lmt = 0.f;
            int cur = -1;
            float m3[3] = { 0.f, 0.f, 0.f }, gacc[3] = { 0.f, 0.f, 0.f };
#pragma unroll
            for (int j = 0; j < SEG; j++) {
                if (lr) {
                    const int cx = (x0 + j) / sx;
                    if (cx != cur) {
                        if (BWD && cur >= 0) {
#pragma unroll
                            for (int c = 0; c < 3; c++) { atomicAdd(grow + cur * 3 + c, gacc[c]); gacc[c] = 0.f; }
                        }
                        cur = cx;
#pragma unroll
                        for (int c = 0; c < 3; c++) m3[c] = __ldg(mrow + cx * 3 + c);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 3; c++) m3[c] = mf[c][j];
                }
                const bool valid = mm[j] > 0.f && mask[j] > 0.f;
                const float e0 = m3[0] - gt[0][j], e1 = m3[1] - gt[1][j], e2 = m3[2] - gt[2][j];
                const float n = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
                if (!BWD) lmt += valid ? n : 0.f;
                else {
                    const float s = (valid && n > 0.f) ? gl[3] / n : 0.f;
                    if (lr) { gacc[0] += s * e0; gacc[1] += s * e1; gacc[2] += s * e2; }
                    else { mf[0][j] = s * e0; mf[1][j] = s * e1; mf[2][j] = s * e2; }
                }
            }
            if (!BWD) acc[3] = lmt;
            else if (lr) {
                if (cur >= 0) {
#pragma unroll
                    for (int c = 0; c < 3; c++) atomicAdd(grow + cur * 3 + c, gacc[c]);
                }
            } else {
#pragma unroll
                for (int c = 0; c < 3; c++) store16(g4.p + b * g4.bs + c * P1 + base, mf[c]);
            }
        }
    }
    if (!BWD) {
        block_sum<4>(acc, redf);
        block_sum<1>(tb, redd);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < 4; k++) atomicAdd(losses + b * 4 + k, acc[k] * inv_hw);
            if (a.use_depth) atomicAdd(ws + 4 + b, tb[0]);
        }
    }
}

static bool fill_args(Args &a, const void *const *maps, const long long *bstrides, const float *match_lr, int B, int H,
                      int W, int hf, int wf, int use_depth)
{
    if (B <= 0 || H <= 0 || W <= 0 || W % SEG != 0 || B > 65535) return false;
    if (hf > 0 && (wf <= 0 || H % hf != 0 || W % wf != 0 || !match_lr)) return false;
    for (int i = 0; i < NIN; i++) {
        a.in[i].p = (const float *)maps[i];
        a.in[i].bs = bstrides[i];
        const bool needed = i != MATCH_FULL ? (use_depth || (i != DEPTH && i != DEPTH_RENDER && i != DEPTH_MASK)) : hf <= 0;
        if (needed && (!a.in[i].p || ((uintptr_t)a.in[i].p & 15) || (a.in[i].bs & 3))) return false;
    }
    a.match_lr = match_lr;
    a.B = B; a.H = H; a.W = W; a.hf = hf; a.wf = wf; a.use_depth = use_depth;
    return true;
}

}  // namespace loss
}  // namespace scp

using namespace scp::loss;

extern "C" size_t scp_image_losses_workspace_bytes(int B) { return B > 0 ? (size_t)(4 + B) * sizeof(double) : 0; }

extern "C" int scp_image_losses_forward(const void *const *maps, const long long *bstrides, const float *match_lr, int B,
                                        int H, int W, int hf, int wf, int use_depth, float *losses, void *workspace,
                                        void *stream)
{
    Args a;
    if (!fill_args(a, maps, bstrides, match_lr, B, H, W, hf, wf, use_depth) || !losses || !workspace) {
        scp::set_last_error("scp_image_losses_forward: unsupported arguments (B=%d H=%d W=%d hf=%d wf=%d; W %% 16 == 0, "
                            "16-byte aligned maps required)", B, H, W, hf, wf);
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    double *ws = (double *)workspace;
    cudaMemsetAsync(ws, 0, scp_image_losses_workspace_bytes(B), st);
    cudaMemsetAsync(losses, 0, (size_t)B * 4 * sizeof(float), st);
    const dim3 grid((unsigned)(((long)H * W / SEG + NT - 1) / NT), B);
    if (use_depth) depth_sums_kernel<<<grid, NT, 0, st>>>(a, ws);
    GView z{ nullptr, 0 };
    image_loss_kernel<false><<<grid, NT, 0, st>>>(a, losses, ws, nullptr, z, z, z, z, z, nullptr);
    return scp::check_launch("scp_image_losses_forward");
}

extern "C" int scp_image_losses_backward(const void *const *maps, const long long *bstrides, const float *match_lr, int B,
                                         int H, int W, int hf, int wf, int use_depth, const float *g_losses,
                                         const void *workspace, void *const *g_maps, const long long *g_bstrides,
                                         float *g_match_lr, void *stream)
{
    Args a;
    if (!fill_args(a, maps, bstrides, match_lr, B, H, W, hf, wf, use_depth) || !g_losses || !workspace || !g_maps) {
        scp::set_last_error("scp_image_losses_backward: unsupported arguments");
        return -1;
    }
    GView g[NOUT];
    for (int i = 0; i < NOUT; i++) {
        g[i].p = (float *)g_maps[i];
        g[i].bs = g_bstrides[i];
        const bool needed = i == G_MATCH_FULL ? hf <= 0 : (i != G_DEPTH_RENDER || use_depth);
        if (needed && (!g[i].p || ((uintptr_t)g[i].p & 15) || (g[i].bs & 3))) {
            scp::set_last_error("scp_image_losses_backward: gradient map %d missing or misaligned", i);
            return -1;
        }
    }
    if (hf > 0 && !g_match_lr) {
        scp::set_last_error("scp_image_losses_backward: g_match_lr missing");
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (hf > 0) cudaMemsetAsync(g_match_lr, 0, (size_t)B * hf * wf * 3 * sizeof(float), st);
    const dim3 grid((unsigned)(((long)H * W / SEG + NT - 1) / NT), B);
    image_loss_kernel<true><<<grid, NT, 0, st>>>(a, nullptr, (double *)workspace, g_losses, g[0], g[1], g[2], g[3], g[4],
                                                 g_match_lr);
    return scp::check_launch("scp_image_losses_backward");
}
