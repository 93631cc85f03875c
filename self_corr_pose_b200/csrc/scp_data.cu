// Data path of the training loader on the GPU: what Wild6DDataset.__getitem__
// (data/dataset_wild6d.py:115-182 of the reference) does per frame on a CPU worker AFTER the image files are decoded --
// silhouette bounding box, randomly scaled crop box, crop intrinsics, and the three torchvision `resized_crop` calls
// (RGB bilinear on a float64 tensor, mask and depth nearest on float32 tensors) -- as two kernels over a whole batch of
// decoded uint8 / uint16 frames resident in HBM.
//
//   bbox_crop_kernel     one CTA per frame: min / max of the mask's foreground coordinates (:131-137), the crop box
//                        (:138-139, python int() truncation of float64 products), focal length / principal point of the
//                        crop in float64 (:143-147)
//   resized_crop_kernel  one thread per output pixel: RGB = bilinear resize of the zero-padded crop (with or without
//                        torchvision's antialias filter), mask / depth = nearest resize with torch's float32 index rule
//
// Arithmetic follows torch's CPU kernels so that the batch equals the reference loader's: the image is u8 / 255.0 in
// float64 (`img * 1.0` -> ToTensor -> `/ 255.`), interpolated in float64 and rounded to float32 once at the end (the
// reference rounds in Trainer.batch_reshape's `.float()`); nearest source index = min(floorf(dst * (float)in / out), in - 1)
// (UpSampleKernel.cpp nearest_idx); bilinear without antialias: src = max((dst + 0.5) * in / out - 0.5, 0)
// (area_pixel_compute_source_index); with antialias the separable triangle filter of _upsample_bilinear2d_aa
// (support = max(scale, 1), window [int(c - s + 0.5), int(c + s + 0.5)) clipped to the crop, weights normalised per axis).
// A crop box that leaves the frame reads zeros there (torchvision's crop pads with 0).
// HBM-bound: reads the crop's share of the u8 / u16 frames, writes 5 fp32 planes of S x S per frame.
#include <stdint.h>

#include "../../include/scp_b200.h"
#include "scp_common.cuh"

namespace scp {
namespace data {

__global__ void __launch_bounds__(1024) bbox_crop_kernel(const uint8_t *__restrict__ mask, const double *__restrict__ rand_scale,
                                                         const double *__restrict__ intr, int H, int W, int S, int no_stretch,
                                                         int *__restrict__ crop, long long *__restrict__ center,
                                                         long long *__restrict__ length, double *__restrict__ foc_crop,
                                                         double *__restrict__ pp_crop, int *__restrict__ status)
{
    const int b = blockIdx.x;
    const uint8_t *m = mask + (size_t)b * H * W;
    int xmin = W, xmax = -1, ymin = H, ymax = -1;
    // 16 mask bytes per load; a row never straddles a load when W % 16 == 0, else the scalar path
    if ((W & 15) == 0) {
        const uint4 *m4 = reinterpret_cast<const uint4 *>(m);
        const int n4 = H * W / 16, w4 = W / 16;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) {
            const uint4 q = __ldg(m4 + i);
            if ((q.x | q.y | q.z | q.w) == 0u) continue;
            const int y = i / w4, x0 = (i - y * w4) * 16;
            const uint32_t wds[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (wds[k] == 0u) continue;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if ((wds[k] >> (8 * j)) & 0xffu) {
                        const int x = x0 + 4 * k + j;
                        xmin = min(xmin, x); xmax = max(xmax, x);
                    }
                }
            }
            ymin = min(ymin, y); ymax = max(ymax, y);
        }
    } else {
        for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
            if (m[i]) {
                const int y = i / W, x = i - y * W;
                xmin = min(xmin, x); xmax = max(xmax, x); ymin = min(ymin, y); ymax = max(ymax, y);
            }
        }
    }
    __shared__ int s[4];
    if (threadIdx.x == 0) { s[0] = W; s[1] = -1; s[2] = H; s[3] = -1; }
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&s[0], xmin); atomicMax(&s[1], xmax); atomicMin(&s[2], ymin); atomicMax(&s[3], ymax);
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    xmin = s[0]; xmax = s[1]; ymin = s[2]; ymax = s[3];
    if (xmax < 0) {   // empty silhouette: the reference raises on `.max()` of an empty array; reported through status
        status[b] = 1;
        crop[4 * b] = crop[4 * b + 1] = crop[4 * b + 2] = crop[4 * b + 3] = 0;
        center[2 * b] = center[2 * b + 1] = length[2 * b] = length[2 * b + 1] = 0;
        foc_crop[2 * b] = foc_crop[2 * b + 1] = pp_crop[2 * b] = pp_crop[2 * b + 1] = 0.0;
        return;
    }
    status[b] = 0;
    const long long cx = (xmax + xmin) / 2, cy = (ymax + ymin) / 2;            // `//` of non-negative integers
    const long long lx0 = (xmax - xmin) / 2, ly0 = (ymax - ymin) / 2;
    const double r0 = rand_scale[2 * b], r1 = rand_scale[2 * b + 1];
    long long lx, ly;
    if (no_stretch) {
        const long long ml = lx0 > ly0 ? lx0 : ly0;
        lx = ly = (long long)(r0 * (double)ml);                                 // python int(): truncation
    } else {
        lx = (long long)(r0 * (double)lx0);
        ly = (long long)(r1 * (double)ly0);
    }
    center[2 * b] = cx; center[2 * b + 1] = cy;
    length[2 * b] = lx; length[2 * b + 1] = ly;
    crop[4 * b + 0] = (int)(cy - ly);      // top
    crop[4 * b + 1] = (int)(cx - lx);      // left
    crop[4 * b + 2] = (int)(2 * ly);       // height
    crop[4 * b + 3] = (int)(2 * lx);       // width
    // crop_factor = [maxw / 2 / length[0], maxh / 2 / length[1]]  (python floats; division by zero -> inf like numpy would warn)
    const double fx = intr[4 * b], fy = intr[4 * b + 1], px = intr[4 * b + 2], py = intr[4 * b + 3];
    const double cfx = (double)S / 2.0 / (double)lx, cfy = (double)S / 2.0 / (double)ly;
    foc_crop[2 * b] = fx * cfx;
    foc_crop[2 * b + 1] = fy * cfy;
    pp_crop[2 * b] = (px - (double)(cx - lx)) * cfx;
    pp_crop[2 * b + 1] = (py - (double)(cy - ly)) * cfy;
}

// torch's antialias window of one output index along one axis (compute_indices_weights_aa / _compute_indices_min_size_weights_aa)
struct Window {
    int lo, n;          // first source index inside the crop, number of taps
    double center, inv, total;
};
__device__ __forceinline__ Window aa_window(int dst, double scale, int in)
{
    Window w;
    const double support = scale >= 1.0 ? scale : 1.0;
    w.inv = scale >= 1.0 ? 1.0 / scale : 1.0;
    w.center = scale * ((double)dst + 0.5);
    w.lo = max((int)(w.center - support + 0.5), 0);
    w.n = min((int)(w.center + support + 0.5), in) - w.lo;
    w.total = 0.0;
    for (int j = 0; j < w.n; j++) {
        const double x = fabs(((double)(j + w.lo) - w.center + 0.5) * w.inv);
        w.total += x < 1.0 ? 1.0 - x : 0.0;
    }
    return w;
}
__device__ __forceinline__ double aa_weight(const Window &w, int j)
{
    const double x = fabs(((double)(j + w.lo) - w.center + 0.5) * w.inv);
    const double t = x < 1.0 ? 1.0 - x : 0.0;
    return w.total != 0.0 ? t / w.total : t;
}

__global__ void __launch_bounds__(256) resized_crop_kernel(const uint8_t *__restrict__ img, const uint8_t *__restrict__ mask,
                                                           const uint16_t *__restrict__ depth, const int *__restrict__ crop,
                                                           int H, int W, int S, int bgr, int antialias,
                                                           float *__restrict__ img_out, float *__restrict__ mask_out,
                                                           float *__restrict__ depth_out)
{
    const int b = blockIdx.y;
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= S * S) return;
    const int oy = o / S, ox = o - oy * S;
    const int top = crop[4 * b], left = crop[4 * b + 1], ch = crop[4 * b + 2], cw = crop[4 * b + 3];
    const size_t plane = (size_t)S * S;
    if (ch <= 0 || cw <= 0) {   // empty crop (flagged by bbox_crop_kernel): zeros
        if (img_out) for (int c = 0; c < 3; c++) img_out[((size_t)b * 3 + c) * plane + o] = 0.f;
        if (mask_out) mask_out[(size_t)b * plane + o] = 0.f;
        if (depth_out) depth_out[(size_t)b * plane + o] = 0.f;
        return;
    }
    const uint8_t *im = img + (size_t)b * H * W * 3;
    // ---- nearest (mask, depth): float32 index arithmetic as in torch's nearest_idx ----
    {
        const float sy = (float)ch / (float)S, sx = (float)cw / (float)S;
        const int iy = ch == S ? oy : min((int)floorf((float)oy * sy), ch - 1);
        const int ix = cw == S ? ox : min((int)floorf((float)ox * sx), cw - 1);
        const int y = top + iy, x = left + ix;
        const bool inside = y >= 0 && y < H && x >= 0 && x < W;
        if (mask_out) mask_out[(size_t)b * plane + o] = inside && mask[(size_t)b * H * W + (size_t)y * W + x] ? 1.f : 0.f;
        if (depth_out) depth_out[(size_t)b * plane + o] = inside ? (float)depth[(size_t)b * H * W + (size_t)y * W + x] : 0.f;
    }
    if (!img_out) return;
    // ---- bilinear (RGB) in float64 ----
    const double sy = (double)ch / (double)S, sx = (double)cw / (double)S;
    double acc[3] = { 0.0, 0.0, 0.0 };
    auto texel = [&](int iy, int ix, double w) {
        const int y = top + iy, x = left + ix;
        if (y < 0 || y >= H || x < 0 || x >= W) return;          // zero padding
        const uint8_t *p = im + ((size_t)y * W + x) * 3;
        const double r = (double)p[bgr ? 2 : 0] / 255.0, g = (double)p[1] / 255.0, bl = (double)p[bgr ? 0 : 2] / 255.0;
        acc[0] += w * r; acc[1] += w * g; acc[2] += w * bl;
    };
    if (antialias) {
        const Window wy = aa_window(oy, sy, ch), wx = aa_window(ox, sx, cw);
        for (int j = 0; j < wy.n; j++) {
            const double a = aa_weight(wy, j);
            double row[3] = { 0.0, 0.0, 0.0 };
            for (int i = 0; i < wx.n; i++) {                          // horizontal pass first, as torch's separable kernel
                const double w = aa_weight(wx, i);
                const int y = top + wy.lo + j, x = left + wx.lo + i;
                if (y < 0 || y >= H || x < 0 || x >= W) continue;
                const uint8_t *p = im + ((size_t)y * W + x) * 3;
                row[0] += w * ((double)p[bgr ? 2 : 0] / 255.0);
                row[1] += w * ((double)p[1] / 255.0);
                row[2] += w * ((double)p[bgr ? 0 : 2] / 255.0);
            }
            acc[0] += a * row[0]; acc[1] += a * row[1]; acc[2] += a * row[2];
        }
    } else {
        double fy = sy * ((double)oy + 0.5) - 0.5, fx = sx * ((double)ox + 0.5) - 0.5;
        fy = fy < 0.0 ? 0.0 : fy;
        fx = fx < 0.0 ? 0.0 : fx;
        const int y0 = min((int)fy, ch - 1), x0 = min((int)fx, cw - 1);
        const int y1 = min(y0 + 1, ch - 1), x1 = min(x0 + 1, cw - 1);
        const double ly = fy - (double)y0, lx = fx - (double)x0;
        texel(y0, x0, (1.0 - ly) * (1.0 - lx));
        texel(y0, x1, (1.0 - ly) * lx);
        texel(y1, x0, ly * (1.0 - lx));
        texel(y1, x1, ly * lx);
    }
#pragma unroll
    for (int c = 0; c < 3; c++) img_out[((size_t)b * 3 + c) * plane + o] = (float)acc[c];
}

}  // namespace data
}  // namespace scp

extern "C" int scp_data_bbox_crop(const unsigned char *mask, const double *rand_scale, const double *intr, int B, int H, int W,
                                  int img_size, int no_stretch, int *crop, long long *center, long long *length,
                                  double *foc_crop, double *pp_crop, int *status, void *stream)
{
    if (B <= 0 || H <= 0 || W <= 0 || img_size <= 0) { scp::set_last_error("scp_data_bbox_crop: bad shape"); return -1; }
    if (!mask || !rand_scale || !intr || !crop || !center || !length || !foc_crop || !pp_crop || !status) {
        scp::set_last_error("scp_data_bbox_crop: null pointer");
        return -1;
    }
    scp::data::bbox_crop_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(mask, rand_scale, intr, H, W, img_size, no_stretch, crop,
                                                                       center, length, foc_crop, pp_crop, status);
    return scp::check_launch("scp_data_bbox_crop");
}

extern "C" int scp_data_resized_crop(const unsigned char *img, const unsigned char *mask, const unsigned short *depth,
                                     const int *crop, int B, int H, int W, int img_size, int bgr, int antialias,
                                     float *img_out, float *mask_out, float *depth_out, void *stream)
{
    if (B <= 0 || H <= 0 || W <= 0 || img_size <= 0) { scp::set_last_error("scp_data_resized_crop: bad shape"); return -1; }
    if (!crop || (img_out && !img) || (mask_out && !mask) || (depth_out && !depth)) {
        scp::set_last_error("scp_data_resized_crop: null pointer");
        return -1;
    }
    const dim3 grid((img_size * img_size + 255) / 256, B);
    scp::data::resized_crop_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, mask, depth, crop, H, W, img_size, bgr, antialias,
                                                                            img_out, mask_out, depth_out);
    return scp::check_launch("scp_data_resized_crop");
}
