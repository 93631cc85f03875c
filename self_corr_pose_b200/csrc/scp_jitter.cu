// Photometric jitter + ImageNet normalisation of the encoder input in (at most) two passes over the batch.
//
// Replaces `self.resnet_transform(self.random_jitter(img))` of Encoder.encode_img (model/module/encoder.py:30-32 of the
// reference): torchvision ColorJitter(0.2, 0.2, 0.2, 0.05) on a batched tensor = brightness / contrast / saturation / hue
// in a random order with ONE parameter set for the whole batch, then Normalize(mean, std).  torchvision evaluates this as
// ~70 element-wise launches over (B,3,H,W) plus, inside the HSV -> RGB step, an einsum that cuBLAS runs as B*H*W tiny
// batched products (7 ms per call at B = 64 on a B200; the encoder runs it twice per step).  Here: one reduction pass for
// the per-image grey mean the contrast step needs (over the image as transformed by the steps that precede it) and one
// pass that applies all four steps and the normalisation.  HBM-bound: 12 B read (x2 with contrast) + 12 B written per pixel.
//
// The arithmetic restates torchvision/transforms/_functional_tensor.py statement by statement (_blend, rgb_to_grayscale,
// _rgb2hsv, _hsv2rgb, adjust_*), with explicit round-to-nearest multiplies / adds so that nvcc does not contract them into
// FMAs: the result equals torchvision's on the same GPU up to the grey-mean reduction order.
#include "../../include/scp_b200.h"
#include "scp_common.cuh"

namespace scp {
namespace jitter {

struct Params {
    int order[4];          // torchvision fn ids in application order: 0 brightness, 1 contrast, 2 saturation, 3 hue; -1 = skip
    float r1[3], r2[3];    // blend ratios of brightness / contrast / saturation: ratio and (1 - ratio) as torch rounds them
    float hue;
    float mean[3], std[3];
};

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }

// rgb_to_grayscale: (0.2989 r + 0.587 g + 0.114 b), left to right
__device__ __forceinline__ float gray(float r, float g, float b)
{
    return add(add(mul(0.2989f, r), mul(0.587f, g)), mul(0.114f, b));
}

// _blend(img1, img2, ratio) = clamp(ratio * img1 + (1 - ratio) * img2, 0, 1)
__device__ __forceinline__ float blend(float x, float y, float r1, float r2) { return clamp01(add(mul(r1, x), mul(r2, y))); }

__device__ __forceinline__ void hue_shift(float &r, float &g, float &b, float hue)
{
    // _rgb2hsv
    const float maxc = fmaxf(fmaxf(r, g), b), minc = fminf(fminf(r, g), b);
    const bool eqc = maxc == minc;
    const float cr = sub(maxc, minc);
    const float s = cr / (eqc ? 1.f : maxc);
    const float crd = eqc ? 1.f : cr;
    const float rc = sub(maxc, r) / crd, gc = sub(maxc, g) / crd, bc = sub(maxc, b) / crd;
    const float hr = (maxc == r) ? sub(bc, gc) : 0.f;
    const float hg = ((maxc == g) && (maxc != r)) ? sub(add(2.f, rc), bc) : 0.f;
    const float hb = ((maxc != g) && (maxc != r)) ? sub(add(4.f, gc), rc) : 0.f;
    float h = add(add(hr, hg), hb);
    h = fmodf(add(mul(h, (float)(1.0 / 6.0)), 1.f), 1.f);   // torch divides by a python scalar as a multiply by its reciprocal
    // h = (h + hue_factor) % 1.0  (python-style remainder)
    h = add(h, hue);
    h = sub(h, floorf(h));
    if (h >= 1.f) h = 0.f;                      // remainder never returns the divisor
    const float v = maxc;
    // _hsv2rgb
    const float h6 = mul(h, 6.f);
    const float fi = floorf(h6);
    const float f = sub(h6, fi);
    int i = (int)fi;
    i %= 6;
    const float p = clamp01(mul(v, sub(1.f, s)));
    const float q = clamp01(mul(v, sub(1.f, mul(s, f))));
    const float t = clamp01(mul(v, sub(1.f, mul(s, sub(1.f, f)))));
    switch (i) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
    }
}

// applies steps order[0..n) to one pixel; `gmean` = grey mean of this image for the contrast step
__device__ __forceinline__ void apply(const Params &P, int n, float gmean, float &r, float &g, float &b)
{
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (k >= n) break;
        const int op = P.order[k];
        if (op == 0) {                          // adjust_brightness: blend with zeros
            r = blend(r, 0.f, P.r1[0], P.r2[0]); g = blend(g, 0.f, P.r1[0], P.r2[0]); b = blend(b, 0.f, P.r1[0], P.r2[0]);
        } else if (op == 1) {                   // adjust_contrast: blend with the image's grey mean
            r = blend(r, gmean, P.r1[1], P.r2[1]); g = blend(g, gmean, P.r1[1], P.r2[1]); b = blend(b, gmean, P.r1[1], P.r2[1]);
        } else if (op == 2) {                   // adjust_saturation: blend with the pixel's grey value
            const float y = gray(r, g, b);
            r = blend(r, y, P.r1[2], P.r2[2]); g = blend(g, y, P.r1[2], P.r2[2]); b = blend(b, y, P.r1[2], P.r2[2]);
        } else if (op == 3) {
            hue_shift(r, g, b, P.hue);
        }
    }
}

// pass A: per-image sum of the grey value after the steps that precede the contrast step (n_before of them)
__global__ void __launch_bounds__(256)
gray_sum_kernel(const float *__restrict__ img, int HW, Params P, const Params *__restrict__ Pdev, int n_before,
                double *__restrict__ gsum)
{
    if (Pdev != nullptr) {      // parameters in device memory (CUDA-graph replays re-read them): find the contrast step here
        P = *Pdev;
        n_before = -1;
        for (int k = 3; k >= 0; k--)
            if (P.order[k] == 1) n_before = k;
        if (n_before < 0) return;
    }
    const int b = blockIdx.y;
    const float *base = img + (long)b * 3 * HW;
    double acc = 0.0;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < HW; i += gridDim.x * blockDim.x * 4) {
        const float4 R = *reinterpret_cast<const float4 *>(base + i), G = *reinterpret_cast<const float4 *>(base + HW + i),
                     Bv = *reinterpret_cast<const float4 *>(base + 2 * HW + i);
        float r[4] = { R.x, R.y, R.z, R.w }, g[4] = { G.x, G.y, G.z, G.w }, bl[4] = { Bv.x, Bv.y, Bv.z, Bv.w };
#pragma unroll
        for (int j = 0; j < 4; j++) {
            apply(P, n_before, 0.f, r[j], g[j], bl[j]);
            acc += (double)gray(r[j], g[j], bl[j]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; w++) s += part[w];
        atomicAdd(gsum + b, s);
    }
}

// pass B: all steps + (x - mean) / std
__global__ void __launch_bounds__(256)
jitter_norm_kernel(const float *__restrict__ img, float *__restrict__ out, int HW, Params P, const Params *__restrict__ Pdev,
                   const double *__restrict__ gsum, int nhwc_out)
{
    if (Pdev != nullptr) P = *Pdev;
    const int b = blockIdx.y;
    const float *base = img + (long)b * 3 * HW;
    float *ob = out + (long)b * (nhwc_out == 3 ? 8 : nhwc_out == 2 ? 4 : 3) * HW;
    const float gmean = gsum ? (float)(gsum[b] / (double)HW) : 0.f;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= HW) return;
    const float4 R = *reinterpret_cast<const float4 *>(base + i), G = *reinterpret_cast<const float4 *>(base + HW + i),
                 Bv = *reinterpret_cast<const float4 *>(base + 2 * HW + i);
    float r[4] = { R.x, R.y, R.z, R.w }, g[4] = { G.x, G.y, G.z, G.w }, bl[4] = { Bv.x, Bv.y, Bv.z, Bv.w };
#pragma unroll
    for (int j = 0; j < 4; j++) {
        apply(P, 4, gmean, r[j], g[j], bl[j]);
        r[j] = sub(r[j], P.mean[0]) / P.std[0];
        g[j] = sub(g[j], P.mean[1]) / P.std[1];
        bl[j] = sub(bl[j], P.mean[2]) / P.std[2];
    }
    if (nhwc_out == 3) {   // channels-last padded to EIGHT channels [B][HW][8]
        float4 *o4 = reinterpret_cast<float4 *>(ob + 8 * (long)i);
#pragma unroll
        for (int j = 0; j < 4; j++) { o4[2 * j] = make_float4(r[j], g[j], bl[j], 0.f); o4[2 * j + 1] = make_float4(0.f, 0.f, 0.f, 0.f); }
        return;
    }
    if (nhwc_out == 2) {   // channels-last with a zero fourth channel [B][HW][4]: one 16-byte store per pixel; the stem
        float4 *o4 = reinterpret_cast<float4 *>(ob + 4 * (long)i);   // convolution then runs cuDNN's vectorised NHWC kernels
#pragma unroll
        for (int j = 0; j < 4; j++) o4[j] = make_float4(r[j], g[j], bl[j], 0.f);
        return;
    }
    if (nhwc_out) {     // channels-last output [B][HW][3]: 4 pixels = 12 consecutive floats
        float4 *o4 = reinterpret_cast<float4 *>(ob + 3 * (long)i);
        o4[0] = make_float4(r[0], g[0], bl[0], r[1]);
        o4[1] = make_float4(g[1], bl[1], r[2], g[2]);
        o4[2] = make_float4(bl[2], r[3], g[3], bl[3]);
        return;
    }
    *reinterpret_cast<float4 *>(ob + i) = make_float4(r[0], r[1], r[2], r[3]);
    *reinterpret_cast<float4 *>(ob + HW + i) = make_float4(g[0], g[1], g[2], g[3]);
    *reinterpret_cast<float4 *>(ob + 2 * HW + i) = make_float4(bl[0], bl[1], bl[2], bl[3]);
}

}  // namespace jitter
}  // namespace scp

extern "C" size_t scp_color_jitter_workspace_bytes(int B) { return B > 0 ? (size_t)B * sizeof(double) : 0; }

extern "C" int scp_color_jitter_normalize(const float *img, float *out, int B, int HW, const int *order, const float *ratios,
                                          float hue, const float *mean, const float *std, int nhwc_out, void *workspace,
                                          size_t workspace_bytes, void *stream)
{
    using namespace scp::jitter;
    if (!img || !out || !order || !ratios || !mean || !std || B <= 0 || HW <= 0 || HW % 4 != 0) {
        scp::set_last_error("scp_color_jitter_normalize: bad arguments (B=%d HW=%d; HW must be a multiple of 4)", B, HW);
        return -1;
    }
    if (!workspace || workspace_bytes < scp_color_jitter_workspace_bytes(B)) {
        scp::set_last_error("scp_color_jitter_normalize: workspace too small");
        return -1;
    }
    Params P;
    int n_before = -1;
    for (int k = 0; k < 4; k++) {
        P.order[k] = order[k];
        if (order[k] < -1 || order[k] > 3) { scp::set_last_error("scp_color_jitter_normalize: bad step id %d", order[k]); return -1; }
        if (order[k] == 1 && n_before < 0) n_before = k;
    }
    for (int k = 0; k < 3; k++) {
        P.r1[k] = ratios[2 * k];
        P.r2[k] = ratios[2 * k + 1];
        P.mean[k] = mean[k];
        P.std[k] = std[k];
    }
    P.hue = hue;
    cudaStream_t st = (cudaStream_t)stream;
    double *gsum = nullptr;
    if (n_before >= 0) {
        gsum = (double *)workspace;
        cudaMemsetAsync(gsum, 0, (size_t)B * sizeof(double), st);
        const int bx = (HW / 4 + 255) / 256;
        gray_sum_kernel<<<dim3(bx < 32 ? bx : 32, B), 256, 0, st>>>(img, HW, P, nullptr, n_before, gsum);
    }
    jitter_norm_kernel<<<dim3((HW / 4 + 255) / 256, B), 256, 0, st>>>(img, out, HW, P, nullptr, gsum, nhwc_out);
    return scp::check_launch("scp_color_jitter_normalize");
}

// Same operation with the parameter block in DEVICE memory (layout = scp_jitter_params of the header): the launch itself
// carries no per-step values, so a CUDA graph that contains it can be replayed with new parameters (the caller refreshes
// the block, e.g. through a captured copy from pinned host memory).
extern "C" int scp_color_jitter_normalize_dparams(const float *img, float *out, int B, int HW, const void *params_dev,
                                                  int nhwc_out, void *workspace, size_t workspace_bytes, void *stream)
{
    using namespace scp::jitter;
    static_assert(sizeof(Params) == sizeof(scp_jitter_params), "parameter block layout");
    if (!img || !out || !params_dev || B <= 0 || HW <= 0 || HW % 4 != 0) {
        scp::set_last_error("scp_color_jitter_normalize_dparams: bad arguments (B=%d HW=%d)", B, HW);
        return -1;
    }
    if (!workspace || workspace_bytes < scp_color_jitter_workspace_bytes(B)) {
        scp::set_last_error("scp_color_jitter_normalize_dparams: workspace too small");
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    double *gsum = (double *)workspace;
    const Params *pd = (const Params *)params_dev;
    Params dummy = {};
    cudaMemsetAsync(gsum, 0, (size_t)B * sizeof(double), st);
    const int bx = (HW / 4 + 255) / 256;
    gray_sum_kernel<<<dim3(bx < 32 ? bx : 32, B), 256, 0, st>>>(img, HW, dummy, pd, -1, gsum);
    jitter_norm_kernel<<<dim3((HW / 4 + 255) / 256, B), 256, 0, st>>>(img, out, HW, dummy, pd, gsum, nhwc_out);
    return scp::check_launch("scp_color_jitter_normalize_dparams");
}
