// Channels-last (NHWC) fp32 glue of the convolutional encoder: the three element-wise /
// windowed operators between the cuDNN convolutions whose ATen NHWC kernels run 5-10x below the HBM roofline at the
// training shape (torch.profiler of one step, B = 64: max-pool 0.56 + 0.96 ms, bilinear up-sampling 0.98 + 0.88 ms,
// F.normalize of the pixel features with its strided divisions and the NHWC -> (B,C,P) copy ~1.5 ms), forward + backward:
//
//   maxpool3x3s2   torchvision resnet18's stem pooling (image_encoder.py:122 -> resnet.maxpool: kernel 3, stride 2, pad 1):
//                  first maximum in (kh, kw) scan order wins, NaN propagates (at::native max_pool_forward_nhwc); the
//                  forward keeps the winning window position as one byte per element, the backward GATHERS (each input
//                  element looks at the <= 4 windows that cover it): no atomics, deterministic
//   upsample2x     F.interpolate(mode='bilinear', align_corners=False) of the feature decoder (image_encoder.py:170-178),
//                  at::native::upsample_bilinear2d's source-index and weight formulas; the backward is the exact-2x gather
//                  (an input pixel receives from 4 x 4 output pixels with weights {1/4, 3/4, 3/4 | 1, 1/4})
//   l2norm         F.normalize(feat, p=2, dim=1) of the pixel features (encoder.py:36): reads the projection's NHWC output
//                  [B][P][C], writes unit vectors as the contiguous (B, C, P) matrix the correspondence kernel consumes
//                  (transposed through shared memory) + 1/norm per pixel; the backward returns the NHWC gradient
//
// All HBM-bound: every element is read once and written once per pass (the up-sampling backward re-reads its 4x4 taps
// through L1/L2).  Threads own 4 consecutive channels (16-byte accesses); C % 4 == 0.
#include <float.h>
#include <stdint.h>

#include "../../include/scp_b200.h"
#include "scp_common.cuh"

namespace scp {
namespace nhwc {

// ---- max-pool 3x3 / stride 2 / pad 1 ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float4 *__restrict__ x, float4 *__restrict__ y,
                                                          uchar4 *__restrict__ idx, int H, int W, int C4, int OH, int OW, long total)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int c = (int)(t % C4);
    long r = t / C4;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const long b = r / OH;
    float m[4] = { -INFINITY, -INFINITY, -INFINITY, -INFINITY };
    unsigned char k[4] = { 0, 0, 0, 0 };
    bool first = true;
#pragma unroll
    for (int kh = 0; kh < 3; kh++) {
        const int ih = 2 * oh - 1 + kh;
        if (ih < 0 || ih >= H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; kw++) {
            const int iw = 2 * ow - 1 + kw;
            if (iw < 0 || iw >= W) continue;
            const float4 v = __ldg(x + ((b * H + ih) * W + iw) * C4 + c);
            const float vv[4] = { v.x, v.y, v.z, v.w };
            if (first) {          // at::native starts the index at the window's first element and maxval at -inf
                k[0] = k[1] = k[2] = k[3] = (unsigned char)(kh * 3 + kw);
                first = false;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (vv[j] > m[j] || vv[j] != vv[j]) {      // `(val > maxval) || isnan(val)`: first maximum wins, NaN propagates
                    m[j] = vv[j];
                    k[j] = (unsigned char)(kh * 3 + kw);
                }
            }
        }
    }
    y[t] = make_float4(m[0], m[1], m[2], m[3]);
    idx[t] = make_uchar4(k[0], k[1], k[2], k[3]);
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float4 *__restrict__ gy, const uchar4 *__restrict__ idx,
                                                          float4 *__restrict__ gx, int H, int W, int C4, int OH, int OW, long total)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int c = (int)(t % C4);
    long r = t / C4;
    const int iw = (int)(r % W); r /= W;
    const int ih = (int)(r % H);
    const long b = r / H;
    float a[4] = { 0.f, 0.f, 0.f, 0.f };
    // windows covering (ih, iw): oh with 2 oh - 1 <= ih <= 2 oh + 1
    const int oh0 = max((ih) / 2, 0), oh1 = min((ih + 1) / 2, OH - 1);
    const int ow0 = max((iw) / 2, 0), ow1 = min((iw + 1) / 2, OW - 1);
    for (int oh = oh0; oh <= oh1; oh++) {
        const int kh = ih - (2 * oh - 1);
        if (kh < 0 || kh > 2) continue;
        for (int ow = ow0; ow <= ow1; ow++) {
            const int kw = iw - (2 * ow - 1);
            if (kw < 0 || kw > 2) continue;
            const long o = ((b * OH + oh) * OW + ow) * C4 + c;
            const uchar4 q = __ldg(idx + o);
            const unsigned char code = (unsigned char)(kh * 3 + kw);
            if (q.x != code && q.y != code && q.z != code && q.w != code) continue;
            const float4 g = __ldg(gy + o);
            if (q.x == code) a[0] += g.x;
            if (q.y == code) a[1] += g.y;
            if (q.z == code) a[2] += g.z;
            if (q.w == code) a[3] += g.w;
        }
    }
    gx[t] = make_float4(a[0], a[1], a[2], a[3]);
}

// ---- bilinear up-sampling, align_corners = False -------------------------------------------------------------------
__device__ __forceinline__ void src_index(float scale, int dst, int in, int &i0, int &step, float &l0, float &l1)
{
    // area_pixel_compute_source_index + the index / lambda statements of upsample_bilinear2d_nhwc_out_frame
    float s = scale * ((float)dst + 0.5f) - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = (int)s;
    step = i0 < in - 1 ? 1 : 0;
    l1 = s - (float)i0;
    l0 = 1.f - l1;
}

__global__ void __launch_bounds__(256) upsample_fwd_kernel(const float4 *__restrict__ x, float4 *__restrict__ y, int H, int W,
                                                           int C4, int OH, int OW, float sh, float sw, long total)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int c = (int)(t % C4);
    long r = t / C4;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const long b = r / OH;
    int h1, hp, w1, wp;
    float h0l, h1l, w0l, w1l;
    src_index(sh, oh, H, h1, hp, h0l, h1l);
    src_index(sw, ow, W, w1, wp, w0l, w1l);
    const float4 *p = x + ((b * H + h1) * W + w1) * C4 + c;
    const float4 a = __ldg(p), bb = __ldg(p + (long)wp * C4), cc = __ldg(p + (long)hp * W * C4),
                 d = __ldg(p + ((long)hp * W + wp) * C4);
    float4 o;
    o.x = h0l * (w0l * a.x + w1l * bb.x) + h1l * (w0l * cc.x + w1l * d.x);
    o.y = h0l * (w0l * a.y + w1l * bb.y) + h1l * (w0l * cc.y + w1l * d.y);
    o.z = h0l * (w0l * a.z + w1l * bb.z) + h1l * (w0l * cc.z + w1l * d.z);
    o.w = h0l * (w0l * a.w + w1l * bb.w) + h1l * (w0l * cc.w + w1l * d.w);
    y[t] = o;
}

// taps of input index k along one axis for an exact 2x up-sampling: output 2k-1 (1/4), 2k (3/4; 1 at k = 0),
// 2k+1 (3/4; 1 at k = n-1), 2k+2 (1/4)
__device__ __forceinline__ int taps2x(int k, int n, int (&d)[4], float (&w)[4])
{
    int m = 0;
    if (k >= 1) { d[m] = 2 * k - 1; w[m] = 0.25f; m++; }
    d[m] = 2 * k; w[m] = k == 0 ? 1.f : 0.75f; m++;
    d[m] = 2 * k + 1; w[m] = k == n - 1 ? 1.f : 0.75f; m++;
    if (k + 1 <= n - 1) { d[m] = 2 * k + 2; w[m] = 0.25f; m++; }
    return m;
}

__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const float4 *__restrict__ gy, float4 *__restrict__ gx, int H, int W,
                                                             int C4, long total)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int c = (int)(t % C4);
    long r = t / C4;
    const int iw = (int)(r % W); r /= W;
    const int ih = (int)(r % H);
    const long b = r / H;
    const int OH = 2 * H, OW = 2 * W;
    int dh[4], dw[4];
    float wh[4], ww[4];
    const int nh = taps2x(ih, H, dh, wh), nw = taps2x(iw, W, dw, ww);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < nh; i++) {
        for (int j = 0; j < nw; j++) {
            const float4 g = __ldg(gy + ((b * OH + dh[i]) * OW + dw[j]) * C4 + c);
            const float w = wh[i] * ww[j];
            a.x += w * g.x; a.y += w * g.y; a.z += w * g.z; a.w += w * g.w;
        }
    }
    gx[t] = a;
}

// ---- L2 normalisation over channels, NHWC in -> (B, C, P) out -----------------------------------------------------
// CTA = 32 pixels x C channels (C <= 128, C % 4 == 0); warp w handles pixels w, w + 8, ...
constexpr int LN_PIX = 32;
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const float *__restrict__ x, float *__restrict__ y, float *__restrict__ inv,
                                                         int P, int C, float eps)
{
    extern __shared__ float tile[];   // [C][LN_PIX + 1]
    const int b = blockIdx.y, p0 = blockIdx.x * LN_PIX, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *xb = x + ((size_t)b * P + p0) * C;
    for (int pp = warp; pp < LN_PIX; pp += 8) {
        const bool ok = p0 + pp < P;
        float v[4] = { 0.f, 0.f, 0.f, 0.f };
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = lane + 32 * j;
            if (ok && c < C) { v[j] = xb[(size_t)pp * C + c]; ss += v[j] * v[j]; }
        }
        ss = warp_sum(ss);
        const float n = sqrtf(ss);
        const float r = 1.f / fmaxf(n, eps);
        if (lane == 0 && ok) inv[(size_t)b * P + p0 + pp] = n < eps ? -r : r;   // sign = "the clamp was active"
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = lane + 32 * j;
            if (c < C) tile[c * (LN_PIX + 1) + pp] = v[j] * r;
        }
    }
    __syncthreads();
    float *yb = y + (size_t)b * C * P + p0;
    for (int i = threadIdx.x; i < C * LN_PIX; i += blockDim.x) {
        const int c = i / LN_PIX, pp = i - c * LN_PIX;
        if (p0 + pp < P) yb[(size_t)c * P + pp] = tile[c * (LN_PIX + 1) + pp];
    }
}

// gx[p][c] = |inv[p]| * (gy[c][p] - y[c][p] * sum_c gy y)      (clamped pixels: gx = gy / eps, no projection)
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const float *__restrict__ gy, const float *__restrict__ y,
                                                         const float *__restrict__ inv, float *__restrict__ gx, int P, int C)
{
    extern __shared__ float tile[];   // g[C][33], y[C][33]
    float *tg = tile, *ty = tile + C * (LN_PIX + 1);
    const int b = blockIdx.y, p0 = blockIdx.x * LN_PIX, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *gb = gy + (size_t)b * C * P + p0, *yb = y + (size_t)b * C * P + p0;
    for (int i = threadIdx.x; i < C * LN_PIX; i += blockDim.x) {
        const int c = i / LN_PIX, pp = i - c * LN_PIX;
        const bool ok = p0 + pp < P;
        tg[c * (LN_PIX + 1) + pp] = ok ? gb[(size_t)c * P + pp] : 0.f;
        ty[c * (LN_PIX + 1) + pp] = ok ? yb[(size_t)c * P + pp] : 0.f;
    }
    __syncthreads();
    float *xb = gx + ((size_t)b * P + p0) * C;
    for (int pp = warp; pp < LN_PIX; pp += 8) {
        if (p0 + pp >= P) continue;
        float g[4], yy[4], dot = 0.f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = lane + 32 * j;
            g[j] = c < C ? tg[c * (LN_PIX + 1) + pp] : 0.f;
            yy[j] = c < C ? ty[c * (LN_PIX + 1) + pp] : 0.f;
            dot += g[j] * yy[j];
        }
        dot = warp_sum(dot);
        const float r = inv[(size_t)b * P + p0 + pp];
        const float s = fabsf(r), proj = r < 0.f ? 0.f : dot;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = lane + 32 * j;
            if (c < C) xb[(size_t)pp * C + c] = s * (g[j] - yy[j] * proj);
        }
    }
}

static inline unsigned blocks(long total) { return (unsigned)((total + 255) / 256); }

}  // namespace nhwc
}  // namespace scp

using namespace scp::nhwc;

extern "C" int scp_nhwc_maxpool3x3s2_forward(const float *x, float *y, unsigned char *idx, int B, int H, int W, int C, void *stream)
{
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3) || !x || !y || !idx) { scp::set_last_error("scp_nhwc_maxpool3x3s2_forward: bad argument"); return -1; }
    const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
    const long total = (long)B * OH * OW * (C / 4);
    maxpool_fwd_kernel<<<blocks(total), 256, 0, (cudaStream_t)stream>>>((const float4 *)x, (float4 *)y, (uchar4 *)idx, H, W, C / 4, OH, OW, total);
    return scp::check_launch("scp_nhwc_maxpool3x3s2_forward");
}

extern "C" int scp_nhwc_maxpool3x3s2_backward(const float *gy, const unsigned char *idx, float *gx, int B, int H, int W, int C, void *stream)
{
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3) || !gy || !gx || !idx) { scp::set_last_error("scp_nhwc_maxpool3x3s2_backward: bad argument"); return -1; }
    const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
    const long total = (long)B * H * W * (C / 4);
    maxpool_bwd_kernel<<<blocks(total), 256, 0, (cudaStream_t)stream>>>((const float4 *)gy, (const uchar4 *)idx, (float4 *)gx, H, W, C / 4, OH, OW, total);
    return scp::check_launch("scp_nhwc_maxpool3x3s2_backward");
}

extern "C" int scp_nhwc_upsample_bilinear_forward(const float *x, float *y, int B, int H, int W, int C, int OH, int OW, void *stream)
{
    if (B <= 0 || H <= 0 || W <= 0 || OH <= 0 || OW <= 0 || C <= 0 || (C & 3) || !x || !y) { scp::set_last_error("scp_nhwc_upsample_bilinear_forward: bad argument"); return -1; }
    const long total = (long)B * OH * OW * (C / 4);
    upsample_fwd_kernel<<<blocks(total), 256, 0, (cudaStream_t)stream>>>((const float4 *)x, (float4 *)y, H, W, C / 4, OH, OW,
                                                                          (float)H / (float)OH, (float)W / (float)OW, total);
    return scp::check_launch("scp_nhwc_upsample_bilinear_forward");
}

extern "C" int scp_nhwc_upsample2x_bilinear_backward(const float *gy, float *gx, int B, int H, int W, int C, void *stream)
{
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3) || !gy || !gx) { scp::set_last_error("scp_nhwc_upsample2x_bilinear_backward: bad argument"); return -1; }
    const long total = (long)B * H * W * (C / 4);
    upsample2x_bwd_kernel<<<blocks(total), 256, 0, (cudaStream_t)stream>>>((const float4 *)gy, (float4 *)gx, H, W, C / 4, total);
    return scp::check_launch("scp_nhwc_upsample2x_bilinear_backward");
}

extern "C" int scp_nhwc_l2norm_forward(const float *x, float *y, float *inv_norm, int B, int P, int C, float eps, void *stream)
{
    if (B <= 0 || P <= 0 || C <= 0 || C > 128 || !x || !y || !inv_norm) { scp::set_last_error("scp_nhwc_l2norm_forward: bad argument (C <= 128)"); return -1; }
    const dim3 grid((P + LN_PIX - 1) / LN_PIX, B);
    l2norm_fwd_kernel<<<grid, 256, (size_t)C * (LN_PIX + 1) * sizeof(float), (cudaStream_t)stream>>>(x, y, inv_norm, P, C, eps);
    return scp::check_launch("scp_nhwc_l2norm_forward");
}

extern "C" int scp_nhwc_l2norm_backward(const float *gy, const float *y, const float *inv_norm, float *gx, int B, int P, int C, void *stream)
{
    if (B <= 0 || P <= 0 || C <= 0 || C > 128 || !gy || !y || !inv_norm || !gx) { scp::set_last_error("scp_nhwc_l2norm_backward: bad argument (C <= 128)"); return -1; }
    const dim3 grid((P + LN_PIX - 1) / LN_PIX, B);
    l2norm_bwd_kernel<<<grid, 256, (size_t)2 * C * (LN_PIX + 1) * sizeof(float), (cudaStream_t)stream>>>(gy, y, inv_norm, gx, P, C);
    return scp::check_launch("scp_nhwc_l2norm_backward");
}
