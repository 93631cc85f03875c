// SoftRas soft rasteriser for sm_100a: tile-binned, shared-memory staged, forward + backward.
//
// Replaces third-party/softras/soft_renderer/cuda/soft_rasterize_cuda_kernel.cu of the reference
// (pre-pass :245-305, forward :308-483, backward :486-668).  The reference runs one thread per
// pixel over ALL faces; here one CTA owns a 16x16 pixel tile, culls the face list against the
// tile (ordered stream compaction, so every pixel still visits its faces in ascending face
// index -- needed for the hard z-buffer tie-break and for a reproducible online softmax),
// stages packed 192-byte face records in shared memory and lets each thread (one pixel, warps
// cover 8x4 pixel blocks) evaluate only the surviving faces.  The backward pass reduces the
// per-(pixel,face) gradients inside the warp (shuffles), then inside the CTA (shared-memory
// accumulators) and issues one global atomic per face value per tile instead of one per pixel.
#include <stdarg.h>

#include "../../include/scp_b200.h"
#include "scp_common.cuh"

namespace scp {
namespace softras {

constexpr int TILE = 16;        // pixels per tile side
constexpr int NTHREADS = 256;   // one thread per tile pixel
constexpr int REC = 48;         // floats per packed face record (12 x float4)
constexpr int REC4 = REC / 4;
constexpr int LIST_CAP = 128;   // records staged per round
constexpr int NGRAD = 18;       // 9 face-coordinate + 9 vertex-texture gradients per face

struct Params {
    int B, nf, T, R, is, tiles_x;
    float near_, far_, eps, sigma, gamma, threshold, margin;
    int dist_mode, rgb_mode, alpha_mode, tex_mode, double_side;
};

// Packed per-face record (float4 slots):
//  0: bbox  (xmin-m, xmax+m, ymin-m, ymax+m)        m = sqrt(dist_eps*sigma)
//  1: inv[0..3]   2: inv[4..7]   3: inv[8], x0, y0, z0
//  4: x1, y1, z1, x2            5: y2, z2, front-facing flag, face index (int bits)
//  6..8: E_k = sym[row k] - sym[row k+1] (3 floats) and den_k = E_k[k] - E_k[k+1]
//  9: tex[0..3]  10: tex[4..7]  11: tex[8], obt0, obt1, obt2
struct Face {
    float bx0, bx1, by0, by1;
    float inv[9];
    float x[3], y[3], z[3];
    float front;
    int idx;
    float E[3][3], den[3];
    float tex[9];
    float obt[3];
};

__device__ __forceinline__ void load_bbox(const float *r, float &bx0, float &bx1, float &by0, float &by1)
{
    const float4 q = *reinterpret_cast<const float4 *>(r);
    bx0 = q.x; bx1 = q.y; by0 = q.z; by1 = q.w;
}

__device__ __forceinline__ void load_face(const float *r, Face &f)
{
    const float4 *q = reinterpret_cast<const float4 *>(r);
    float4 a = q[1], b = q[2], c = q[3], d = q[4], e = q[5];
    f.inv[0] = a.x; f.inv[1] = a.y; f.inv[2] = a.z; f.inv[3] = a.w;
    f.inv[4] = b.x; f.inv[5] = b.y; f.inv[6] = b.z; f.inv[7] = b.w;
    f.inv[8] = c.x; f.x[0] = c.y; f.y[0] = c.z; f.z[0] = c.w;
    f.x[1] = d.x; f.y[1] = d.y; f.z[1] = d.z; f.x[2] = d.w;
    f.y[2] = e.x; f.z[2] = e.y; f.front = e.z; f.idx = __float_as_int(e.w);
    a = q[6]; b = q[7]; c = q[8];
    f.E[0][0] = a.x; f.E[0][1] = a.y; f.E[0][2] = a.z; f.den[0] = a.w;
    f.E[1][0] = b.x; f.E[1][1] = b.y; f.E[1][2] = b.z; f.den[1] = b.w;
    f.E[2][0] = c.x; f.E[2][1] = c.y; f.E[2][2] = c.z; f.den[2] = c.w;
    a = q[9]; b = q[10]; c = q[11];
    f.tex[0] = a.x; f.tex[1] = a.y; f.tex[2] = a.z; f.tex[3] = a.w;
    f.tex[4] = b.x; f.tex[5] = b.y; f.tex[6] = b.z; f.tex[7] = b.w;
    f.tex[8] = c.x; f.obt[0] = c.y; f.obt[1] = c.z; f.obt[2] = c.w;
}

// ---- pack kernel: per-face pre-pass (kernel.cu:245-305) + bbox + record ------------------
__global__ void __launch_bounds__(256) pack_kernel(Params p, const float *__restrict__ faces,
                                                   const float *__restrict__ textures, float *faces_info,
                                                   int compute_info, float4 *__restrict__ bbox,
                                                   float *__restrict__ rec)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)p.B * p.nf) return;
    const float *f = faces + i * 9;
    const float x0 = f[0], y0 = f[1], z0 = f[2], x1 = f[3], y1 = f[4], z1 = f[5], x2 = f[6], y2 = f[7], z2 = f[8];
    float inv[9], sym[9], obt[3] = { 0.f, 0.f, 0.f };
    float *info = faces_info + i * 27;
    if (compute_info) {
        const float adj[9] = { y1 - y2, x2 - x1, x1 * y2 - x2 * y1,
                               y2 - y0, x0 - x2, x2 * y0 - x0 * y2,
                               y0 - y1, x1 - x0, x0 * y1 - x1 * y0 };
        float det = x2 * (y0 - y1) + x0 * (y1 - y2) + x1 * (y2 - y0);
        det = det > 0.f ? fmaxf(det, 1e-10f) : fminf(det, -1e-10f);
#pragma unroll
        for (int k = 0; k < 9; k++) inv[k] = adj[k] / det;
        const float px[3] = { x0, x1, x2 }, py[3] = { y0, y1, y2 };
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int k = 0; k < 3; k++) sym[3 * j + k] = px[j] * px[k] + py[j] * py[k] + 1.f;
        bool found = false;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
            const bool o = (px[k1] - px[k]) * (px[k2] - px[k]) + (py[k1] - py[k]) * (py[k2] - py[k]) < 0.f;
            if (o && !found) { obt[k] = 1.f; found = true; }
        }
#pragma unroll
        for (int k = 0; k < 9; k++) { info[k] = inv[k]; info[9 + k] = sym[k]; }
#pragma unroll
        for (int k = 0; k < 3; k++) info[18 + k] = obt[k];
    } else {
#pragma unroll
        for (int k = 0; k < 9; k++) { inv[k] = info[k]; sym[k] = info[9 + k]; }
#pragma unroll
        for (int k = 0; k < 3; k++) obt[k] = info[18 + k];
    }
    const float m = p.margin;
    const float4 bb = make_float4(fminf(fminf(x0, x1), x2) - m, fmaxf(fmaxf(x0, x1), x2) + m,
                                  fminf(fminf(y0, y1), y2) - m, fmaxf(fmaxf(y0, y1), y2) + m);
    bbox[i] = bb;
    float E[3][3], den[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int k1 = (k + 1) % 3;
#pragma unroll
        for (int c = 0; c < 3; c++) E[k][c] = sym[3 * k + c] - sym[3 * k1 + c];
        den[k] = E[k][k] - E[k][k1];
    }
    const float front = ((y2 - y0) * (x1 - x0) < (y1 - y0) * (x2 - x0)) ? 1.f : 0.f;
    float tex[9];
#pragma unroll
    for (int k = 0; k < 9; k++) tex[k] = 0.f;
    const int nt = p.T * 3;
    if (nt <= 9) {
        const float *t = textures + i * nt;
        for (int k = 0; k < nt; k++) tex[k] = t[k];
    }
    float4 *r = reinterpret_cast<float4 *>(rec + i * REC);
    r[0] = bb;
    r[1] = make_float4(inv[0], inv[1], inv[2], inv[3]);
    r[2] = make_float4(inv[4], inv[5], inv[6], inv[7]);
    r[3] = make_float4(inv[8], x0, y0, z0);
    r[4] = make_float4(x1, y1, z1, x2);
    r[5] = make_float4(y2, z2, front, __int_as_float((int)(i % p.nf)));
    r[6] = make_float4(E[0][0], E[0][1], E[0][2], den[0]);
    r[7] = make_float4(E[1][0], E[1][1], E[1][2], den[1]);
    r[8] = make_float4(E[2][0], E[2][1], E[2][2], den[2]);
    r[9] = make_float4(tex[0], tex[1], tex[2], tex[3]);
    r[10] = make_float4(tex[4], tex[5], tex[6], tex[7]);
    r[11] = make_float4(tex[8], obt[0], obt[1], obt[2]);
}

// ---- per-(pixel, face) fragment ----------------------------------------------------------
struct Frag {
    float w[3];   // raw barycentric
    float c[3];   // barycentric of the closest point on the selected edge (euclid) / w (bary)
    float sign, dx, dy, dis, frag;
};

// parameter t of the foot of the perpendicular on edge K (vertices K, K+1), kernel.cu:81-87
template <int K>
__device__ __forceinline__ float edge_param(const Face &f, const float *w)
{
    constexpr int K1 = (K + 1) % 3;
    return (w[0] * f.E[K][0] + w[1] * f.E[K][1] + w[2] * f.E[K][2] - f.E[K][K1]) / f.den[K];
}

template <int K>
__device__ __forceinline__ void edge_point_inside(const Face &f, const float *w, float &best, float &bx, float &by,
                                                  float *c)
{
    constexpr int K1 = (K + 1) % 3, K2 = (K + 2) % 3;
    float t[3];
    t[K] = edge_param<K>(f, w);
    t[K1] = 1.f - t[K];
    t[K2] = 0.f;
    const float d0 = t[0] - w[0], d1 = t[1] - w[1], d2 = t[2] - w[2];
    const float ex = d0 * f.x[0] + d1 * f.x[1] + d2 * f.x[2];
    const float ey = d0 * f.y[0] + d1 * f.y[1] + d2 * f.y[2];
    const float dd = ex * ex + ey * ey;
    if (dd < best) { best = dd; bx = ex; by = ey; c[0] = t[0]; c[1] = t[1]; c[2] = t[2]; }
}

template <int K>
__device__ __forceinline__ void edge_point_outside(const Face &f, const float *w, float *c)
{
    constexpr int K1 = (K + 1) % 3, K2 = (K + 2) % 3;
    const float t = edge_param<K>(f, w);
    c[K] = fminf(fmaxf(t, 0.f), 1.f);
    c[K1] = fminf(fmaxf(1.f - t, 0.f), 1.f);
    c[K2] = 0.f;
}

// Euclidean point-to-triangle vector, kernel.cu:61-151
__device__ __forceinline__ void euclid(const Face &f, float xp, float yp, Frag &o)
{
    const float *w = o.w;
    if (w[0] > 0.f && w[1] > 0.f && w[2] > 0.f && w[0] < 1.f && w[1] < 1.f && w[2] < 1.f) {
        float best = 100000000.f, bx = 0.f, by = 0.f;
        o.c[0] = o.c[1] = o.c[2] = 0.f;
        edge_point_inside<0>(f, w, best, bx, by, o.c);
        edge_point_inside<1>(f, w, best, bx, by, o.c);
        edge_point_inside<2>(f, w, best, bx, by, o.c);
        o.dx = bx; o.dy = by; o.sign = 1.f;
        return;
    }
    int v0;
    if (w[1] <= 0.f && w[2] <= 0.f) {
        v0 = 0;
        if (f.obt[0] == 1.f && (xp - f.x[0]) * (f.x[2] - f.x[0]) + (yp - f.y[0]) * (f.y[2] - f.y[0]) > 0.f) v0 = 2;
    } else if (w[2] <= 0.f && w[0] <= 0.f) {
        v0 = 1;
        if (f.obt[1] == 1.f && (xp - f.x[1]) * (f.x[0] - f.x[1]) + (yp - f.y[1]) * (f.y[0] - f.y[1]) > 0.f) v0 = 0;
    } else if (w[0] <= 0.f && w[1] <= 0.f) {
        v0 = 2;
        if (f.obt[2] == 1.f && (xp - f.x[2]) * (f.x[1] - f.x[2]) + (yp - f.y[2]) * (f.y[1] - f.y[2]) > 0.f) v0 = 1;
    } else if (w[0] <= 0.f) v0 = 1;
    else if (w[1] <= 0.f) v0 = 2;
    else if (w[2] <= 0.f) v0 = 0;
    else  // a component >= 1 with the others > 0 (rounding only; the reference indexes with -1 here)
        v0 = (w[0] >= w[1] && w[0] >= w[2]) ? 1 : (w[1] >= w[2] ? 2 : 0);
    if (v0 == 0) edge_point_outside<0>(f, w, o.c);
    else if (v0 == 1) edge_point_outside<1>(f, w, o.c);
    else edge_point_outside<2>(f, w, o.c);
    const float d0 = o.c[0] - w[0], d1 = o.c[1] - w[1], d2 = o.c[2] - w[2];
    o.dx = d0 * f.x[0] + d1 * f.x[1] + d2 * f.x[2];
    o.dy = d0 * f.y[0] + d1 * f.y[1] + d2 * f.y[2];
    o.sign = -1.f;
}

__device__ __forceinline__ bool inside_closed(const float *w)
{
    return w[0] <= 1.f && w[0] >= 0.f && w[1] <= 1.f && w[1] >= 0.f && w[2] <= 1.f && w[2] >= 0.f;
}

// fragment probability; false = face skipped for this pixel (kernel.cu:375-404)
template <bool FAST>
__device__ __forceinline__ bool eval_frag(const Params &p, const Face &f, float xp, float yp, Frag &o)
{
    o.w[0] = f.inv[0] * xp + f.inv[1] * yp + f.inv[2];
    o.w[1] = f.inv[3] * xp + f.inv[4] * yp + f.inv[5];
    o.w[2] = f.inv[6] * xp + f.inv[7] * yp + f.inv[8];
    const int dist_mode = FAST ? SCP_DIST_EUCLIDEAN : p.dist_mode;
    o.sign = 0.f; o.dx = o.dy = o.dis = 0.f;
    if (dist_mode == SCP_DIST_EUCLIDEAN) {
        euclid(f, xp, yp, o);
        o.dis = o.dx * o.dx + o.dy * o.dy;
        if (o.sign < 0.f && o.dis >= p.threshold) return false;
        o.frag = 1.f / (1.f + expf(-o.sign * o.dis / p.sigma));
    } else if (dist_mode == SCP_DIST_BARYCENTRIC) {
        const float *w = o.w;
        const float d = w[0] > w[1] ? (w[1] > w[2] ? w[2] : w[1]) : (w[0] > w[2] ? w[2] : w[0]);
        o.dis = d > 0.f ? d * d : -(d * d);
        o.c[0] = w[0]; o.c[1] = w[1]; o.c[2] = w[2];
        if (-o.dis >= p.threshold) return false;
        o.frag = 1.f / (1.f + expf(-o.dis / p.sigma));
    } else {
        if (!inside_closed(o.w)) return false;
        o.frag = 1.f;
    }
    return true;
}

// clamp + renormalise barycentrics, then perspective-correct depth (kernel.cu:53-58, :421-424)
__device__ __forceinline__ float clip_and_depth(const Face &f, const float *w, float *wc)
{
#pragma unroll
    for (int k = 0; k < 3; k++) wc[k] = fmaxf(fminf(w[k], 1.f), 0.f);
    const float s = fmaxf(wc[0] + wc[1] + wc[2], 1e-5f);
#pragma unroll
    for (int k = 0; k < 3; k++) wc[k] /= s;
    return 1.f / (wc[0] / f.z[0] + wc[1] / f.z[1] + wc[2] / f.z[2]);
}

// surface-texture texel index (kernel.cu:181-188), clamped into the table (the reference can read
// one texel past it when a clipped weight is exactly 1)
__device__ __forceinline__ int surface_texel(const float *wc, int R)
{
    const int wx = (int)(wc[0] * R), wy = (int)(wc[1] * R);
    int idx = ((wc[0] + wc[1]) * R - wx - wy <= 1.f) ? wy * R + wx : (R - 1 - wy) * R + (R - 1 - wx);
    return min(max(idx, 0), R * R - 1);
}

template <bool FAST>
__device__ __forceinline__ void sample_color(const Params &p, const Face &f, const float *wc,
                                             const float *__restrict__ textures, int b, float *col)
{
    const int tex_mode = p.tex_mode;
    if (tex_mode == SCP_TEX_VERTEX) {
#pragma unroll
        for (int k = 0; k < 3; k++) col[k] = wc[0] * f.tex[k] + wc[1] * f.tex[3 + k] + wc[2] * f.tex[6 + k];
    } else if (p.T == 1) {
#pragma unroll
        for (int k = 0; k < 3; k++) col[k] = f.tex[k];
    } else {
        const float *t = textures + ((size_t)b * p.nf + f.idx) * p.T * 3 + surface_texel(wc, p.R) * 3;
#pragma unroll
        for (int k = 0; k < 3; k++) col[k] = t[k];
    }
}

// ---- ordered tile culling -----------------------------------------------------------------
// Scans faces [base, base+256) of image b against the tile rectangle and writes the surviving
// face indices to s_list in ascending order.  Returns the count (uniform across the CTA).
__device__ __forceinline__ int cull_chunk(const Params &p, const float4 *__restrict__ bbox, int b, int base,
                                          float x_lo, float x_hi, float y_lo, float y_hi, int *s_list,
                                          int *s_warp_cnt)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fidx = base + tid;
    bool hit = false;
    if (fidx < p.nf) {
        const float4 bb = bbox[(size_t)b * p.nf + fidx];
        hit = !(x_lo > bb.y || x_hi < bb.x || y_lo > bb.w || y_hi < bb.z);
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_warp_cnt[warp] = __popc(ballot);
    __syncthreads();
    int off = 0, total = 0;
#pragma unroll
    for (int k = 0; k < NTHREADS / 32; k++) {
        const int c = s_warp_cnt[k];
        if (k < warp) off += c;
        total += c;
    }
    if (hit) s_list[off + __popc(ballot & ((1u << lane) - 1u))] = fidx;
    __syncthreads();
    return total;
}

__device__ __forceinline__ void stage_records(const float *__restrict__ rec, int b, int nf, const int *s_list,
                                              int start, int m, float *s_rec)
{
    const float4 *src = reinterpret_cast<const float4 *>(rec);
    float4 *dst = reinterpret_cast<float4 *>(s_rec);
    for (int i = threadIdx.x; i < m * REC4; i += NTHREADS) {
        const int j = i / REC4, q = i - j * REC4;
        dst[i] = src[((size_t)b * nf + s_list[start + j]) * REC4 + q];
    }
}

struct Pixel {
    int px, py, pn;
    bool valid;
    float xp, yp;
};

__device__ __forceinline__ void tile_setup(const Params &p, Pixel &px, float &x_lo, float &x_hi, float &y_lo,
                                           float &y_hi)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = (blockIdx.x % p.tiles_x) * TILE, ty0 = (blockIdx.x / p.tiles_x) * TILE;
    px.px = tx0 + (warp & 1) * 8 + (lane & 7);
    px.py = ty0 + (warp >> 1) * 4 + (lane >> 3);
    px.valid = px.px < p.is && px.py < p.is;
    px.pn = px.py * p.is + px.px;
    const float fis = (float)p.is;
    // pixel centres, row 0 = top (kernel.cu:343-346); numerators are exact integers
    px.xp = (float)(2 * px.px + 1 - p.is) / fis;
    px.yp = (float)(2 * (p.is - 1 - px.py) + 1 - p.is) / fis;
    const int tx1 = min(tx0 + TILE - 1, p.is - 1), ty1 = min(ty0 + TILE - 1, p.is - 1);
    x_lo = (float)(2 * tx0 + 1 - p.is) / fis;
    x_hi = (float)(2 * tx1 + 1 - p.is) / fis;
    y_hi = (float)(2 * (p.is - 1 - ty0) + 1 - p.is) / fis;
    y_lo = (float)(2 * (p.is - 1 - ty1) + 1 - p.is) / fis;
}

// ---- forward ------------------------------------------------------------------------------
template <int RGB, bool FAST>
__global__ void __launch_bounds__(NTHREADS) forward_kernel(Params p, const float4 *__restrict__ bbox,
                                                          const float *__restrict__ rec,
                                                          const float *__restrict__ textures,
                                                          float *__restrict__ aggrs_info,
                                                          float *__restrict__ soft_colors)
{
    __shared__ __align__(16) float s_rec[LIST_CAP * REC];
    __shared__ int s_list[NTHREADS];
    __shared__ int s_warp_cnt[NTHREADS / 32];

    const int b = blockIdx.y;
    const size_t plane = (size_t)p.is * p.is;
    Pixel px;
    float x_lo, x_hi, y_lo, y_hi;
    tile_setup(p, px, x_lo, x_hi, y_lo, y_hi);
    const int alpha_mode = FAST ? SCP_ALPHA_PROD : p.alpha_mode;

    float col[3] = { 0.f, 0.f, 0.f };
    float alpha = alpha_mode == SCP_ALPHA_PROD ? 1.f : 0.f;
    float sm_sum = expf(p.eps / p.gamma), sm_max = p.eps;
    float depth_min = 10000000.f;
    int face_min = -1;
    if (px.valid) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float bg = soft_colors[((size_t)b * 4 + k) * plane + px.pn];
            col[k] = RGB == SCP_RGB_HARD ? bg : bg * sm_sum;
        }
    }

    for (int base = 0; base < p.nf; base += NTHREADS) {
        const int n = cull_chunk(p, bbox, b, base, x_lo, x_hi, y_lo, y_hi, s_list, s_warp_cnt);
        for (int start = 0; start < n; start += LIST_CAP) {
            const int m = min(LIST_CAP, n - start);
            stage_records(rec, b, p.nf, s_list, start, m, s_rec);
            __syncthreads();
            if (px.valid) {
                for (int j = 0; j < m; j++) {
                    const float *r = s_rec + j * REC;
                    float bx0, bx1, by0, by1;
                    load_bbox(r, bx0, bx1, by0, by1);
                    if (px.xp > bx1 || px.xp < bx0 || px.yp > by1 || px.yp < by0) continue;
                    Face f;
                    load_face(r, f);
                    Frag fr;
                    if (!eval_frag<FAST>(p, f, px.xp, px.yp, fr)) continue;

                    // alpha accumulates before the depth test (kernel.cu:408-417)
                    if (alpha_mode == SCP_ALPHA_PROD) alpha *= 1.f - fr.frag;
                    else if (alpha_mode == SCP_ALPHA_SUM) alpha += fr.frag;
                    else if (fr.frag > 0.5f) alpha = 1.f;

                    float wc[3];
                    const float zp = clip_and_depth(f, fr.w, wc);
                    if (zp < p.near_ || zp > p.far_) continue;

                    if (RGB == SCP_RGB_HARD) {
                        if (zp < depth_min && inside_closed(fr.w) && (p.double_side || f.front != 0.f)) {
                            depth_min = zp;
                            face_min = f.idx;
                            sample_color<FAST>(p, f, wc, textures, b, col);
                        }
                    } else if (f.front != 0.f || p.double_side) {
                        const float zn = (p.far_ - zp) / (p.far_ - p.near_);
                        float rescale = 1.f;
                        if (zn > sm_max) {
                            rescale = expf((sm_max - zn) / p.gamma);
                            sm_max = zn;
                        }
                        const float ez = expf((zn - sm_max) / p.gamma);
                        sm_sum = rescale * sm_sum + ez * fr.frag;
                        float c[3];
                        sample_color<FAST>(p, f, wc, textures, b, c);
#pragma unroll
                        for (int k = 0; k < 3; k++) col[k] = rescale * col[k] + ez * fr.frag * c[k];
                    }
                }
            }
            __syncthreads();
        }
    }

    if (!px.valid) return;
    float a_out;
    if (alpha_mode == SCP_ALPHA_PROD) a_out = 1.f - alpha;
    else if (alpha_mode == SCP_ALPHA_SUM) a_out = alpha / p.nf;
    else a_out = alpha;
    soft_colors[((size_t)b * 4 + 3) * plane + px.pn] = a_out;
    if (RGB == SCP_RGB_HARD) {
        if (face_min != -1) {
#pragma unroll
            for (int k = 0; k < 3; k++) soft_colors[((size_t)b * 4 + k) * plane + px.pn] = col[k];
        }
        aggrs_info[((size_t)b * 2 + 0) * plane + px.pn] = depth_min;
        aggrs_info[((size_t)b * 2 + 1) * plane + px.pn] = (float)face_min;
    } else {
#pragma unroll
        for (int k = 0; k < 3; k++) soft_colors[((size_t)b * 4 + k) * plane + px.pn] = col[k] / sm_sum;
        aggrs_info[((size_t)b * 2 + 0) * plane + px.pn] = sm_sum;
        aggrs_info[((size_t)b * 2 + 1) * plane + px.pn] = sm_max;
    }
}

// ---- backward -----------------------------------------------------------------------------
template <int RGB, bool FAST>
__global__ void __launch_bounds__(NTHREADS) backward_kernel(Params p, const float4 *__restrict__ bbox,
                                                           const float *__restrict__ rec,
                                                           const float *__restrict__ textures,
                                                           const float *__restrict__ soft_colors,
                                                           const float *__restrict__ aggrs_info,
                                                           const float *__restrict__ grad_soft_colors,
                                                           float *grad_faces, float *grad_textures)
{
    __shared__ __align__(16) float s_rec[LIST_CAP * REC];
    __shared__ float s_grad[LIST_CAP * NGRAD];
    __shared__ int s_list[NTHREADS];
    __shared__ int s_warp_cnt[NTHREADS / 32];

    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const size_t plane = (size_t)p.is * p.is;
    Pixel px;
    float x_lo, x_hi, y_lo, y_hi;
    tile_setup(p, px, x_lo, x_hi, y_lo, y_hi);
    const int alpha_mode = FAST ? SCP_ALPHA_PROD : p.alpha_mode;
    const int dist_mode = FAST ? SCP_DIST_EUCLIDEAN : p.dist_mode;
    // vertex textures and 1-texel surface textures accumulate through shared memory; larger
    // surface tables (never used by this repo's renders) go straight to global atomics
    const bool tex_in_rec = p.T * 3 <= 9;

    float g[4] = { 0.f, 0.f, 0.f, 0.f }, out[4] = { 0.f, 0.f, 0.f, 0.f };
    float sm_sum = 1.f, sm_max = 0.f;
    if (px.valid) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            g[k] = grad_soft_colors[((size_t)b * 4 + k) * plane + px.pn];
            out[k] = soft_colors[((size_t)b * 4 + k) * plane + px.pn];
        }
        sm_sum = aggrs_info[((size_t)b * 2 + 0) * plane + px.pn];
        sm_max = aggrs_info[((size_t)b * 2 + 1) * plane + px.pn];
    }
    // a tile whose incoming gradient is identically zero contributes nothing
    const bool any_grad = __syncthreads_or(px.valid && (g[0] != 0.f || g[1] != 0.f || g[2] != 0.f || g[3] != 0.f));
    if (!any_grad) return;

    for (int base = 0; base < p.nf; base += NTHREADS) {
        const int n = cull_chunk(p, bbox, b, base, x_lo, x_hi, y_lo, y_hi, s_list, s_warp_cnt);
        for (int start = 0; start < n; start += LIST_CAP) {
            const int m = min(LIST_CAP, n - start);
            stage_records(rec, b, p.nf, s_list, start, m, s_rec);
            for (int i = tid; i < m * NGRAD; i += NTHREADS) s_grad[i] = 0.f;
            __syncthreads();
            for (int j = 0; j < m; j++) {
                const float *r = s_rec + j * REC;
                float gv[9], gt[9];
#pragma unroll
                for (int k = 0; k < 9; k++) gv[k] = gt[k] = 0.f;
                bool active = false;
                float bx0, bx1, by0, by1;
                load_bbox(r, bx0, bx1, by0, by1);
                if (px.valid && !(px.xp > bx1 || px.xp < bx0 || px.yp > by1 || px.yp < by0)) {
                    Face f;
                    load_face(r, f);
                    Frag fr;
                    if (eval_frag<FAST>(p, f, px.xp, px.yp, fr)) {
                        float Gxy = 0.f;
                        float Ga = g[3];
                        if (alpha_mode == SCP_ALPHA_SUM) Ga /= p.nf;
                        else if (alpha_mode == SCP_ALPHA_PROD) Ga *= (1.f - out[3]) / fmaxf(1.f - fr.frag, 1e-6f);
                        Gxy += Ga;
                        float wc[3];
                        const float zp = clip_and_depth(f, fr.w, wc);
                        if (!(zp < p.near_ || zp > p.far_)) {  // else: face dropped incl. its alpha gradient
                            active = true;
                            if (RGB == SCP_RGB_HARD) {
                                if ((float)f.idx == sm_max) {
                                    if (p.tex_mode == SCP_TEX_VERTEX) {
#pragma unroll
                                        for (int v = 0; v < 3; v++)
#pragma unroll
                                            for (int k = 0; k < 3; k++) gt[3 * v + k] = wc[v] * g[k];
                                    } else if (p.T == 1) {
#pragma unroll
                                        for (int k = 0; k < 3; k++) gt[k] = g[k];
                                    } else {
                                        float *dst = grad_textures + ((size_t)b * p.nf + f.idx) * p.T * 3 +
                                                     surface_texel(wc, p.R) * 3;
#pragma unroll
                                        for (int k = 0; k < 3; k++) atomicAdd(dst + k, g[k]);
                                    }
                                }
                            } else if (f.front != 0.f || p.double_side) {
                                const float zn = (p.far_ - zp) / (p.far_ - p.near_);
                                const float s = fr.frag * expf((zn - sm_max) / p.gamma) / sm_sum;
                                float c[3];
                                sample_color<FAST>(p, f, wc, textures, b, c);
                                float Q = 0.f;
#pragma unroll
                                for (int k = 0; k < 3; k++) Q += g[k] * (c[k] - out[k]);
                                if (p.tex_mode == SCP_TEX_VERTEX) {
#pragma unroll
                                    for (int v = 0; v < 3; v++)
#pragma unroll
                                        for (int k = 0; k < 3; k++) gt[3 * v + k] = s * (wc[v] * g[k]);
                                } else if (p.T == 1) {
#pragma unroll
                                    for (int k = 0; k < 3; k++) gt[k] = s * g[k];
                                } else {
                                    float *dst = grad_textures + ((size_t)b * p.nf + f.idx) * p.T * 3 +
                                                 surface_texel(wc, p.R) * 3;
#pragma unroll
                                    for (int k = 0; k < 3; k++) atomicAdd(dst + k, s * g[k]);
                                }
                                Q *= s;
                                Gxy += Q / fr.frag;
                                const float Gz = Q / p.gamma / (p.near_ - p.far_) * zp * zp;
#pragma unroll
                                for (int v = 0; v < 3; v++) gv[3 * v + 2] = Gz * wc[v] / f.z[v] / f.z[v];
                            }
                            Gxy *= fr.frag * (1.f - fr.frag) / p.sigma;  // sigmoid'
                            if (dist_mode == SCP_DIST_EUCLIDEAN) {
#pragma unroll
                                for (int v = 0; v < 3; v++) {
                                    const float s2 = 2.f * fr.sign * Gxy * fr.c[v];
                                    gv[3 * v + 0] = s2 * fr.dx;
                                    gv[3 * v + 1] = s2 * fr.dy;
                                }
                            } else if (dist_mode == SCP_DIST_BARYCENTRIC) {
                                // kernel.cu:161-175
                                const float *t = fr.c;
                                const int q = t[0] > t[1] ? (t[1] > t[2] ? 2 : 1) : (t[0] > t[2] ? 2 : 0);
                                const float sc = 2.f * sqrtf(fabsf(fr.dis));
#pragma unroll
                                for (int l = 0; l < 2; l++) {
                                    const float iq = q == 0 ? f.inv[l] : (q == 1 ? f.inv[3 + l] : f.inv[6 + l]);
#pragma unroll
                                    for (int v = 0; v < 3; v++) {
                                        const float acc = -iq * f.inv[3 * v + 0] * px.xp +
                                                          -iq * f.inv[3 * v + 1] * px.yp + -iq * f.inv[3 * v + 2];
                                        gv[3 * v + l] = acc * Gxy * sc;
                                    }
                                }
                            }
                        }
                    }
                }
                // warp reduction, then one shared-memory atomic per value per warp
                if (__any_sync(0xffffffffu, active)) {
                    float *acc = s_grad + j * NGRAD;
#pragma unroll
                    for (int k = 0; k < 9; k++) {
                        if (RGB == SCP_RGB_HARD && (k % 3) == 2) continue;  // no depth gradient in hard mode
                        const float v = warp_sum(gv[k]);
                        if (lane == 0 && v != 0.f) atomicAdd(acc + k, v);
                    }
                    if (tex_in_rec) {
#pragma unroll
                        for (int k = 0; k < 9; k++) {
                            const float v = warp_sum(gt[k]);
                            if (lane == 0 && v != 0.f) atomicAdd(acc + 9 + k, v);
                        }
                    }
                }
            }
            __syncthreads();
            // flush: one global atomic per touched value per tile
            for (int i = tid; i < m * NGRAD; i += NTHREADS) {
                const float v = s_grad[i];
                if (v != 0.f) {
                    const int j = i / NGRAD, k = i - j * NGRAD;
                    const size_t fg = (size_t)b * p.nf + s_list[start + j];
                    if (k < 9) atomicAdd(grad_faces + fg * 9 + k, v);
                    else if (k - 9 < p.T * 3) atomicAdd(grad_textures + fg * p.T * 3 + (k - 9), v);
                }
            }
            __syncthreads();
        }
    }
}

// ---- host side ----------------------------------------------------------------------------
static bool make_params(Params &p, int B, int nf, int T, int is, float near_, float far_, float eps, float sigma,
                        int dist_mode, float dist_eps, float gamma, int rgb_mode, int alpha_mode, int tex_mode,
                        int double_side)
{
    if (B <= 0 || nf <= 0 || T <= 0 || is <= 0 || B > 65535) return false;
    if (dist_mode < 0 || dist_mode > 2 || rgb_mode < 0 || rgb_mode > 1 || alpha_mode < 0 || alpha_mode > 2 ||
        tex_mode < 0 || tex_mode > 1)
        return false;
    if (tex_mode == SCP_TEX_VERTEX && T != 3) return false;
    p.B = B; p.nf = nf; p.T = T; p.R = (int)sqrt((double)T); p.is = is;
    p.tiles_x = (is + TILE - 1) / TILE;
    p.near_ = near_; p.far_ = far_; p.eps = eps; p.sigma = sigma; p.gamma = gamma;
    p.threshold = dist_eps * sigma;
    p.margin = sqrtf(p.threshold);
    p.dist_mode = dist_mode; p.rgb_mode = rgb_mode; p.alpha_mode = alpha_mode; p.tex_mode = tex_mode;
    p.double_side = double_side;
    return true;
}

static size_t bbox_bytes(int B, int nf) { return (((size_t)B * nf * sizeof(float4)) + 255) / 256 * 256; }
static size_t rec_bytes(int B, int nf) { return (size_t)B * nf * REC * sizeof(float); }

}  // namespace softras
}  // namespace scp

using namespace scp::softras;

extern "C" size_t scp_softras_workspace_bytes(int B, int nf)
{
    if (B <= 0 || nf <= 0) return 0;
    return bbox_bytes(B, nf) + rec_bytes(B, nf);
}

extern "C" int scp_softras_forward(const float *faces, const float *textures, float *faces_info, float *aggrs_info,
                                   float *soft_colors, int B, int nf, int T, int image_size, float near_,
                                   float far_, float eps, float sigma_val, int func_id_dist, float dist_eps,
                                   float gamma_val, int func_id_rgb, int func_id_alpha, int texture_sample_type,
                                   int double_side, void *workspace, size_t workspace_bytes, void *stream)
{
    Params p;
    if (!make_params(p, B, nf, T, image_size, near_, far_, eps, sigma_val, func_id_dist, dist_eps, gamma_val,
                     func_id_rgb, func_id_alpha, texture_sample_type, double_side)) {
        scp::set_last_error("scp_softras_forward: unsupported arguments (B=%d nf=%d T=%d is=%d modes %d/%d/%d/%d)", B,
                            nf, T, image_size, func_id_dist, func_id_rgb, func_id_alpha, texture_sample_type);
        return -1;
    }
    if (!workspace || workspace_bytes < scp_softras_workspace_bytes(B, nf)) {
        scp::set_last_error("scp_softras_forward: workspace too small");
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float4 *bbox = (float4 *)workspace;
    float *rec = (float *)((char *)workspace + bbox_bytes(B, nf));
    const long nfaces = (long)B * nf;
    pack_kernel<<<(unsigned)((nfaces + 255) / 256), 256, 0, st>>>(p, faces, textures, faces_info, 1, bbox, rec);
    const dim3 grid(p.tiles_x * p.tiles_x, B);
    const bool fast = func_id_dist == SCP_DIST_EUCLIDEAN && func_id_alpha == SCP_ALPHA_PROD;
    if (func_id_rgb == SCP_RGB_HARD) {
        if (fast) forward_kernel<SCP_RGB_HARD, true><<<grid, NTHREADS, 0, st>>>(p, bbox, rec, textures, aggrs_info, soft_colors);
        else forward_kernel<SCP_RGB_HARD, false><<<grid, NTHREADS, 0, st>>>(p, bbox, rec, textures, aggrs_info, soft_colors);
    } else {
        if (fast) forward_kernel<SCP_RGB_SOFTMAX, true><<<grid, NTHREADS, 0, st>>>(p, bbox, rec, textures, aggrs_info, soft_colors);
        else forward_kernel<SCP_RGB_SOFTMAX, false><<<grid, NTHREADS, 0, st>>>(p, bbox, rec, textures, aggrs_info, soft_colors);
    }
    return scp::check_launch("scp_softras_forward");
}

extern "C" int scp_softras_backward(const float *faces, const float *textures, const float *soft_colors,
                                    const float *faces_info, const float *aggrs_info, float *grad_faces,
                                    float *grad_textures, const float *grad_soft_colors, int B, int nf, int T,
                                    int image_size, float near_, float far_, float eps, float sigma_val,
                                    int func_id_dist, float dist_eps, float gamma_val, int func_id_rgb,
                                    int func_id_alpha, int texture_sample_type, int double_side, void *workspace,
                                    size_t workspace_bytes, void *stream)
{
    Params p;
    if (!make_params(p, B, nf, T, image_size, near_, far_, eps, sigma_val, func_id_dist, dist_eps, gamma_val,
                     func_id_rgb, func_id_alpha, texture_sample_type, double_side)) {
        scp::set_last_error("scp_softras_backward: unsupported arguments");
        return -1;
    }
    if (!workspace || workspace_bytes < scp_softras_workspace_bytes(B, nf)) {
        scp::set_last_error("scp_softras_backward: workspace too small");
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float4 *bbox = (float4 *)workspace;
    float *rec = (float *)((char *)workspace + bbox_bytes(B, nf));
    const long nfaces = (long)B * nf;
    pack_kernel<<<(unsigned)((nfaces + 255) / 256), 256, 0, st>>>(p, faces, textures, const_cast<float *>(faces_info),
                                                                 0, bbox, rec);
    const dim3 grid(p.tiles_x * p.tiles_x, B);
    const bool fast = func_id_dist == SCP_DIST_EUCLIDEAN && func_id_alpha == SCP_ALPHA_PROD;
    if (func_id_rgb == SCP_RGB_HARD) {
        if (fast) backward_kernel<SCP_RGB_HARD, true><<<grid, NTHREADS, 0, st>>>(p, bbox, rec, textures, soft_colors, aggrs_info, grad_soft_colors, grad_faces, grad_textures);
        else backward_kernel<SCP_RGB_HARD, false><<<grid, NTHREADS, 0, st>>>(p, bbox, rec, textures, soft_colors, aggrs_info, grad_soft_colors, grad_faces, grad_textures);
    } else {
        if (fast) backward_kernel<SCP_RGB_SOFTMAX, true><<<grid, NTHREADS, 0, st>>>(p, bbox, rec, textures, soft_colors, aggrs_info, grad_soft_colors, grad_faces, grad_textures);
        else backward_kernel<SCP_RGB_SOFTMAX, false><<<grid, NTHREADS, 0, st>>>(p, bbox, rec, textures, soft_colors, aggrs_info, grad_soft_colors, grad_faces, grad_textures);
    }
    return scp::check_launch("scp_softras_backward");
}
