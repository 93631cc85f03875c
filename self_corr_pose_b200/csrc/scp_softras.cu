// SoftRas soft rasteriser for sm_100a: tile-binned, shared-memory staged, forward + backward.
//
// Replaces third-party/softras/soft_renderer/cuda/soft_rasterize_cuda_kernel.cu of the reference
// (pre-pass :245-305, forward :308-483, backward :486-668).  The reference runs one thread per
// pixel over ALL faces; here one CTA owns a 16x16 pixel tile, culls the face list against the
// tile (ordered stream compaction, so every pixel still visits its faces in ascending face
// index -- needed for the hard z-buffer tie-break and for a reproducible online softmax),
// then every warp (an 8x4 pixel block, one pixel per lane) culls that list once more against its
// own footprint and evaluates only the surviving faces, reading packed 192-byte face records
// through L1 (all lanes read the same record: one transaction).  After the tile list is built the
// warps run independently -- no CTA barrier in the pair loop.  The backward pass reduces the
// per-(pixel,face) gradients across the warp with a folded butterfly (20 shuffles for 18 values,
// each lane ending up with one total) and issues one global reduction per value per
// (warp, face) instead of one per (pixel, face).
#include <stdarg.h>
#include <stdlib.h>

#include "../../include/scp_b200.h"
#include "scp_common.cuh"

namespace scp {
namespace softras {

constexpr int TILE = 16;        // pixels per tile side
constexpr int NTHREADS = 256;   // one thread per tile pixel
constexpr int REC = 48;         // floats per packed face record (12 x float4)
constexpr int REC4 = REC / 4;
constexpr int NWARPS = NTHREADS / 32;
constexpr int SCAN = 4;         // faces tested per thread per culling round
constexpr int LIST_CAP = 2048;  // tile face list capacity (indices) per phase

// Occupancy tunables (compile-time, for A/B variant builds: python -m self_corr_pose_b200.build --variant NAME -D...;
// the defaults are the measured round-1 choice): resident CTAs per SM the register allocator must allow for the
// tile-centric kernels (3 -> 85 registers at 256 threads) and for the face-centric backward (8 x 2 warps at 128).
#ifndef SCP_SOFTRAS_TILE_CTAS
#define SCP_SOFTRAS_TILE_CTAS 3
#endif
#ifndef SCP_SOFTRAS_FACE_WARPS
#define SCP_SOFTRAS_FACE_WARPS 2
#endif
#ifndef SCP_SOFTRAS_FACE_CTAS
#define SCP_SOFTRAS_FACE_CTAS 10
#endif
// The face-centric backward keeps its face record (48 floats) in shared memory instead of registers, so that
// SCP_SOFTRAS_FACE_CTAS can rise above 8 without spilling (95 registers, 20 resident warps per SM).  Round-2 A/B on a B200
// (B = 64, 256 px, 2556 faces; gpurun_out/r2_call1.log): soft-texture backward 3.04 -> 2.87 ms, depth backward 1.26 -> 1.18 ms,
// parity tests unchanged -> default.  (The other two round-1 candidates lost: 16x2 pixel blocks 3.04 -> 3.13 ms, two pixels
// per lane in the forward 2.31 -> 2.74 ms.)
#ifndef SCP_SOFTRAS_FACE_SMEM
#define SCP_SOFTRAS_FACE_SMEM 1
#endif
// Pixel block a warp of the face-centric backward covers per iteration: BW x (32 / BW) pixels (default 8 x 4).  16 x 2
// wastes fewer lanes on the 14-pixel boxes of the sigma = 1e-4 renders and on the 30-pixel boxes of the soft-texture
// render (87 / 94 % instead of 77 / 88 % of the lanes inside the box) and reads 64-byte row segments; candidate, untimed.
#ifndef SCP_SOFTRAS_FACE_BW
#define SCP_SOFTRAS_FACE_BW 8
#endif
// Next-round candidate, NOT validated on a GPU yet (default off): forward traversal with TWO pixels per lane (4 warps per
// 16x16 tile, a warp owns an 8x8 block, lane = column + rows r and r+4): one record load per (warp, face) serves 64
// pixels instead of 32 -- halves the L1 data-path traffic that bounds forward_kernel and gives
// every lane two independent dependency chains.  Same per-pixel arithmetic and face order as forward_kernel.
#ifndef SCP_SOFTRAS_FWD_2PX
#define SCP_SOFTRAS_FWD_2PX 0
#endif

struct Params {
    int B, nf, T, R, is, tiles_x;
    float near_, far_, eps, sigma, gamma, threshold, margin;
    float inv_sigma, inv_gamma, inv_depth_range;
    int dist_mode, rgb_mode, alpha_mode, tex_mode, double_side;
};

// Packed per-face record (float4 slots):
//  0: bbox  (xmin-m, xmax+m, ymin-m, ymax+m)        m = sqrt(dist_eps*sigma)
//  1: inv[0..3]   2: inv[4..7]   3: inv[8], x0, y0, 1/z0
//  4: x1, y1, 1/z1, x2          5: y2, 1/z2, front-facing flag, face index (int bits)
//  6..8: E_k = sym[row k] - sym[row k+1] (3 floats) and 1/den_k, den_k = E_k[k] - E_k[k+1]
//  9: tex[0..3]  10: tex[4..7]  11: tex[8], obt0, obt1, obt2
struct Face {
    float bx0, bx1, by0, by1;
    float inv[9];
    float x[3], y[3], rz[3];
    float front;
    int idx;
    float E[3][3], rden[3];
    float tex[9];
    float obt[3];
};

__device__ __forceinline__ void load_bbox(const float *r, float &bx0, float &bx1, float &by0, float &by1)
{
    const float4 q = __ldg(reinterpret_cast<const float4 *>(r));
    bx0 = q.x; bx1 = q.y; by0 = q.z; by1 = q.w;
}

__device__ __forceinline__ void load_face(const float *r, Face &f)
{
    const float4 *q = reinterpret_cast<const float4 *>(r);
    float4 a = __ldg(q + 1), b = __ldg(q + 2), c = __ldg(q + 3), d = __ldg(q + 4), e = __ldg(q + 5);
    f.inv[0] = a.x; f.inv[1] = a.y; f.inv[2] = a.z; f.inv[3] = a.w;
    f.inv[4] = b.x; f.inv[5] = b.y; f.inv[6] = b.z; f.inv[7] = b.w;
    f.inv[8] = c.x; f.x[0] = c.y; f.y[0] = c.z; f.rz[0] = c.w;
    f.x[1] = d.x; f.y[1] = d.y; f.rz[1] = d.z; f.x[2] = d.w;
    f.y[2] = e.x; f.rz[2] = e.y; f.front = e.z; f.idx = __float_as_int(e.w);
    a = __ldg(q + 6); b = __ldg(q + 7); c = __ldg(q + 8);
    f.E[0][0] = a.x; f.E[0][1] = a.y; f.E[0][2] = a.z; f.rden[0] = a.w;
    f.E[1][0] = b.x; f.E[1][1] = b.y; f.E[1][2] = b.z; f.rden[1] = b.w;
    f.E[2][0] = c.x; f.E[2][1] = c.y; f.E[2][2] = c.z; f.rden[2] = c.w;
    a = __ldg(q + 9); b = __ldg(q + 10); c = __ldg(q + 11);
    f.tex[0] = a.x; f.tex[1] = a.y; f.tex[2] = a.z; f.tex[3] = a.w;
    f.tex[4] = b.x; f.tex[5] = b.y; f.tex[6] = b.z; f.tex[7] = b.w;
    f.tex[8] = c.x; f.obt[0] = c.y; f.obt[1] = c.z; f.obt[2] = c.w;
}

// ---- pack kernel: per-face pre-pass (kernel.cu:245-305) + bbox + record ------------------
__global__ void __launch_bounds__(256) pack_kernel(Params p, const float *__restrict__ faces,
                                                   const float *__restrict__ textures, float *faces_info,
                                                   int compute_info, float4 *__restrict__ bbox,
                                                   float *__restrict__ rec, int *img_bbox)
{
    const long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = i0 < (long)p.B * p.nf;
    const long i = in_range ? i0 : (long)p.B * p.nf - 1;  // out-of-range lanes shadow the last face (no stores)
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
        if (in_range) {
#pragma unroll
            for (int k = 0; k < 9; k++) { info[k] = inv[k]; info[9 + k] = sym[k]; }
#pragma unroll
            for (int k = 0; k < 3; k++) info[18 + k] = obt[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < 9; k++) { inv[k] = info[k]; sym[k] = info[9 + k]; }
#pragma unroll
        for (int k = 0; k < 3; k++) obt[k] = info[18 + k];
    }
    const float m = p.margin;
    const float4 bb = make_float4(fminf(fminf(x0, x1), x2) - m, fmaxf(fmaxf(x0, x1), x2) + m,
                                  fminf(fminf(y0, y1), y2) - m, fmaxf(fmaxf(y0, y1), y2) + m);
    {
        // whole-mesh screen bbox per image: min over faces of (xmin, -xmax, ymin, -ymax) as ordered ints
        const int bimg = (int)(i / p.nf);
        int key[4] = { __float_as_int(bb.x), __float_as_int(-bb.y), __float_as_int(bb.z), __float_as_int(-bb.w) };
#pragma unroll
        for (int k = 0; k < 4; k++) key[k] = key[k] >= 0 ? key[k] : key[k] ^ 0x7fffffff;
        const int b_first = __shfl_sync(0xffffffffu, bimg, 0);
        if (__all_sync(0xffffffffu, bimg == b_first)) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) key[k] = min(key[k], __shfl_xor_sync(0xffffffffu, key[k], o));
            }
            if ((threadIdx.x & 31) == 0) {
#pragma unroll
                for (int k = 0; k < 4; k++) atomicMin(img_bbox + 4 * bimg + k, key[k]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) atomicMin(img_bbox + 4 * bimg + k, key[k]);
        }
    }
    if (!in_range) return;
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
    r[3] = make_float4(inv[8], x0, y0, 1.f / z0);
    r[4] = make_float4(x1, y1, 1.f / z1, x2);
    r[5] = make_float4(y2, 1.f / z2, front, __int_as_float((int)(i % p.nf)));
    r[6] = make_float4(E[0][0], E[0][1], E[0][2], 1.f / den[0]);
    r[7] = make_float4(E[1][0], E[1][1], E[1][2], 1.f / den[1]);
    r[8] = make_float4(E[2][0], E[2][1], E[2][2], 1.f / den[2]);
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
    return (w[0] * f.E[K][0] + w[1] * f.E[K][1] + w[2] * f.E[K][2] - f.E[K][K1]) * f.rden[K];
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
        o.frag = __fdividef(1.f, 1.f + __expf(-o.sign * o.dis * p.inv_sigma));
    } else if (dist_mode == SCP_DIST_BARYCENTRIC) {
        const float *w = o.w;
        const float d = w[0] > w[1] ? (w[1] > w[2] ? w[2] : w[1]) : (w[0] > w[2] ? w[2] : w[0]);
        o.dis = d > 0.f ? d * d : -(d * d);
        o.c[0] = w[0]; o.c[1] = w[1]; o.c[2] = w[2];
        if (-o.dis >= p.threshold) return false;
        o.frag = __fdividef(1.f, 1.f + __expf(-o.dis * p.inv_sigma));
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
    for (int k = 0; k < 3; k++) wc[k] = __saturatef(w[k]);
    const float rs = __fdividef(1.f, fmaxf(wc[0] + wc[1] + wc[2], 1e-5f));
#pragma unroll
    for (int k = 0; k < 3; k++) wc[k] *= rs;
    return __fdividef(1.f, wc[0] * f.rz[0] + wc[1] * f.rz[1] + wc[2] * f.rz[2]);
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
// Scans faces [base, base + SCAN*256) of image b against the tile rectangle and APPENDS the
// surviving face indices to s_list[n0..] in ascending order.  Returns the new count (uniform).
__device__ __forceinline__ int cull_round(const Params &p, const float4 *__restrict__ bbox, int b, int base,
                                          float x_lo, float x_hi, float y_lo, float y_hi, int n0, int *s_list,
                                          int *s_cnt)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned ballots[SCAN];
#pragma unroll
    for (int k = 0; k < SCAN; k++) {
        const int fidx = base + k * NTHREADS + tid;
        bool hit = false;
        if (fidx < p.nf) {
            const float4 bb = __ldg(bbox + (size_t)b * p.nf + fidx);
            hit = !(x_lo > bb.y || x_hi < bb.x || y_lo > bb.w || y_hi < bb.z);
        }
        ballots[k] = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_cnt[k * NWARPS + warp] = __popc(ballots[k]);
    }
    __syncthreads();
    // exclusive prefix over the (k, warp) grid in face order
    int total = 0, my_off[SCAN];
#pragma unroll
    for (int k = 0; k < SCAN; k++) {
#pragma unroll
        for (int w = 0; w < NWARPS; w++) {
            if (w == warp) my_off[k] = total;
            total += s_cnt[k * NWARPS + w];
        }
    }
#pragma unroll
    for (int k = 0; k < SCAN; k++) {
        if (ballots[k] & (1u << lane))
            s_list[n0 + my_off[k] + __popc(ballots[k] & ((1u << lane) - 1u))] = base + k * NTHREADS + tid;
    }
    __syncthreads();
    return n0 + total;
}

struct Pixel {
    int px, py, pn;
    bool valid;
    float xp, yp;
    // footprint of this lane's warp (pixel centres of its 8x4 block, clipped to the image)
    float wx_lo, wx_hi, wy_lo, wy_hi;
};

__device__ __forceinline__ float centre_x(int px, int is) { return (float)(2 * px + 1 - is) / (float)is; }
// row 0 = top (kernel.cu:343-346); numerators are exact integers
__device__ __forceinline__ float centre_y(int py, int is) { return (float)(2 * (is - 1 - py) + 1 - is) / (float)is; }

__device__ __forceinline__ void tile_setup(const Params &p, Pixel &px, float &x_lo, float &x_hi, float &y_lo,
                                           float &y_hi)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = (blockIdx.x % p.tiles_x) * TILE, ty0 = (blockIdx.x / p.tiles_x) * TILE;
    const int bx = tx0 + (warp & 1) * 8, by = ty0 + (warp >> 1) * 4;
    px.px = bx + (lane & 7);
    px.py = by + (lane >> 3);
    px.valid = px.px < p.is && px.py < p.is;
    px.pn = px.py * p.is + px.px;
    px.xp = centre_x(px.px, p.is);
    px.yp = centre_y(px.py, p.is);
    px.wx_lo = centre_x(bx, p.is);
    px.wx_hi = centre_x(min(bx + 7, p.is - 1), p.is);
    px.wy_hi = centre_y(by, p.is);
    px.wy_lo = centre_y(min(by + 3, p.is - 1), p.is);
    x_lo = centre_x(tx0, p.is);
    x_hi = centre_x(min(tx0 + TILE - 1, p.is - 1), p.is);
    y_hi = centre_y(ty0, p.is);
    y_lo = centre_y(min(ty0 + TILE - 1, p.is - 1), p.is);
}

// whole-mesh screen bbox of image b (written by pack_kernel): tiles outside it have no work
__device__ __forceinline__ bool tile_outside_mesh(const int *__restrict__ img_bbox, int b, float x_lo, float x_hi,
                                                  float y_lo, float y_hi)
{
    const int4 q = __ldg(reinterpret_cast<const int4 *>(img_bbox) + b);
    // ordered-int encoding of floats: (xmin, -xmax, ymin, -ymax) minima
    const float mx0 = __int_as_float(q.x >= 0 ? q.x : q.x ^ 0x7fffffff);
    const float mx1 = -__int_as_float(q.y >= 0 ? q.y : q.y ^ 0x7fffffff);
    const float my0 = __int_as_float(q.z >= 0 ? q.z : q.z ^ 0x7fffffff);
    const float my1 = -__int_as_float(q.w >= 0 ? q.w : q.w ^ 0x7fffffff);
    return x_lo > mx1 || x_hi < mx0 || y_lo > my1 || y_hi < my0;
}

// ---- forward ------------------------------------------------------------------------------
struct FwdState {
    float col[3];
    float alpha, sm_sum, sm_max, depth_min;
    int face_min;
    float col2[3];   // RGB_DUAL: colour of the hard z-buffer pass (second texture set)
};

// third aggregation mode of the fused depth + NOCS traversal: softmax RGB of `textures` AND hard RGB of `textures2`
// over the same fragments (the two renders share sigma, the distance function and the alpha aggregation; gamma is
// unused by the hard mode, soft_rasterize_cuda_kernel.cu:428-435)
constexpr int RGB_DUAL = 2;

template <int RGB, bool FAST>
__device__ __forceinline__ void forward_pair(const Params &p, const float *__restrict__ r, const Pixel &px,
                                             const float *__restrict__ textures, const float *__restrict__ textures2,
                                             int b, FwdState &s)
{
    float bx0, bx1, by0, by1;
    load_bbox(r, bx0, bx1, by0, by1);
    if (px.xp > bx1 || px.xp < bx0 || px.yp > by1 || px.yp < by0) return;
    Face f;
    load_face(r, f);
    Frag fr;
    if (!eval_frag<FAST>(p, f, px.xp, px.yp, fr)) return;
    const int alpha_mode = FAST ? SCP_ALPHA_PROD : p.alpha_mode;

    // alpha accumulates before the depth test (kernel.cu:408-417)
    if (alpha_mode == SCP_ALPHA_PROD) s.alpha *= 1.f - fr.frag;
    else if (alpha_mode == SCP_ALPHA_SUM) s.alpha += fr.frag;
    else if (fr.frag > 0.5f) s.alpha = 1.f;

    float wc[3];
    const float zp = clip_and_depth(f, fr.w, wc);
    if (zp < p.near_ || zp > p.far_) return;

    if (RGB == SCP_RGB_HARD || RGB == RGB_DUAL) {
        if (zp < s.depth_min && inside_closed(fr.w) && (p.double_side || f.front != 0.f)) {
            s.depth_min = zp;
            s.face_min = f.idx;
            if (RGB == RGB_DUAL) {   // vertex colours of the second texture set, fetched only for z-buffer winners
                const float *t2 = textures2 + ((size_t)b * p.nf + f.idx) * 9;
#pragma unroll
                for (int k = 0; k < 3; k++) s.col2[k] = wc[0] * __ldg(t2 + k) + wc[1] * __ldg(t2 + 3 + k) + wc[2] * __ldg(t2 + 6 + k);
            } else {
                sample_color<FAST>(p, f, wc, textures, b, s.col);
            }
        }
    }
    if ((RGB == SCP_RGB_SOFTMAX || RGB == RGB_DUAL) && (f.front != 0.f || p.double_side)) {
        const float zn = (p.far_ - zp) * p.inv_depth_range;
        float rescale = 1.f;
        if (zn > s.sm_max) {
            rescale = __expf((s.sm_max - zn) * p.inv_gamma);
            s.sm_max = zn;
        }
        const float ez = __expf((zn - s.sm_max) * p.inv_gamma) * fr.frag;
        s.sm_sum = rescale * s.sm_sum + ez;
        float c[3];
        sample_color<FAST>(p, f, wc, textures, b, c);
#pragma unroll
        for (int k = 0; k < 3; k++) s.col[k] = rescale * s.col[k] + ez * c[k];
    }
}

template <int RGB, bool FAST>
__global__ void __launch_bounds__(NTHREADS, SCP_SOFTRAS_TILE_CTAS) forward_kernel(Params p, const float4 *__restrict__ bbox,
                                                          const float *__restrict__ rec,
                                                          const int *__restrict__ img_bbox,
                                                          const float *__restrict__ textures,
                                                          float *__restrict__ aggrs_info,
                                                          float *__restrict__ soft_colors,
                                                          const float *__restrict__ textures2,
                                                          float *__restrict__ aggrs_info2,
                                                          float *__restrict__ soft_colors2)
{
    __shared__ int s_list[LIST_CAP];
    __shared__ int s_cnt[SCAN * NWARPS];

    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const size_t plane = (size_t)p.is * p.is;
    Pixel px;
    float x_lo, x_hi, y_lo, y_hi;
    tile_setup(p, px, x_lo, x_hi, y_lo, y_hi);
    const int alpha_mode = FAST ? SCP_ALPHA_PROD : p.alpha_mode;

    FwdState s;
    s.col[0] = s.col[1] = s.col[2] = 0.f;
    s.alpha = alpha_mode == SCP_ALPHA_PROD ? 1.f : 0.f;
    s.sm_sum = __expf(p.eps * p.inv_gamma);
    s.sm_max = p.eps;
    s.depth_min = 10000000.f;
    s.face_min = -1;
    if (px.valid) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float bg = soft_colors[((size_t)b * 4 + k) * plane + px.pn];
            s.col[k] = RGB == SCP_RGB_HARD ? bg : bg * s.sm_sum;
        }
    }

    if (!tile_outside_mesh(img_bbox, b, x_lo, x_hi, y_lo, y_hi)) {
        int base = 0;
        while (base < p.nf) {
            int n = 0;
            while (base < p.nf && n + SCAN * NTHREADS <= LIST_CAP) {
                n = cull_round(p, bbox, b, base, x_lo, x_hi, y_lo, y_hi, n, s_list, s_cnt);
                base += SCAN * NTHREADS;
            }
            // warp-private pass: cull the tile list against the warp's 8x4 block, 32 faces at a time
            for (int i0 = 0; i0 < n; i0 += 32) {
                const int fi = i0 + lane < n ? s_list[i0 + lane] : -1;
                bool hit = false;
                if (fi >= 0) {
                    const float4 bb = __ldg(bbox + (size_t)b * p.nf + fi);
                    hit = !(px.wx_lo > bb.y || px.wx_hi < bb.x || px.wy_lo > bb.w || px.wy_hi < bb.z);
                }
                unsigned todo = __ballot_sync(0xffffffffu, hit);
                while (todo) {
                    const int j = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int fsel = __shfl_sync(0xffffffffu, fi, j);
                    if (px.valid) forward_pair<RGB, FAST>(p, rec + ((size_t)b * p.nf + fsel) * REC, px, textures, textures2, b, s);
                }
            }
            if (base < p.nf) __syncthreads();  // the list is about to be rebuilt
        }
    }

    if (!px.valid) return;
    float a_out;
    if (alpha_mode == SCP_ALPHA_PROD) a_out = 1.f - s.alpha;
    else if (alpha_mode == SCP_ALPHA_SUM) a_out = s.alpha / p.nf;
    else a_out = s.alpha;
    soft_colors[((size_t)b * 4 + 3) * plane + px.pn] = a_out;
    if (RGB == RGB_DUAL) {   // hard pass -> second output set (its colours keep the pre-filled background when no face hit)
        soft_colors2[((size_t)b * 4 + 3) * plane + px.pn] = a_out;
        if (s.face_min != -1) {
#pragma unroll
            for (int k = 0; k < 3; k++) soft_colors2[((size_t)b * 4 + k) * plane + px.pn] = s.col2[k];
        }
        aggrs_info2[((size_t)b * 2 + 0) * plane + px.pn] = s.depth_min;
        aggrs_info2[((size_t)b * 2 + 1) * plane + px.pn] = (float)s.face_min;
    }
    if (RGB == SCP_RGB_HARD) {
        if (s.face_min != -1) {
#pragma unroll
            for (int k = 0; k < 3; k++) soft_colors[((size_t)b * 4 + k) * plane + px.pn] = s.col[k];
        }
        aggrs_info[((size_t)b * 2 + 0) * plane + px.pn] = s.depth_min;
        aggrs_info[((size_t)b * 2 + 1) * plane + px.pn] = (float)s.face_min;
    } else {
        const float rs = 1.f / s.sm_sum;
#pragma unroll
        for (int k = 0; k < 3; k++) soft_colors[((size_t)b * 4 + k) * plane + px.pn] = s.col[k] * rs;
        aggrs_info[((size_t)b * 2 + 0) * plane + px.pn] = s.sm_sum;
        aggrs_info[((size_t)b * 2 + 1) * plane + px.pn] = s.sm_max;
    }
}

#if SCP_SOFTRAS_FWD_2PX
// ---- forward, two pixels per lane (candidate, see SCP_SOFTRAS_FWD_2PX above) ------------------------------------------
constexpr int NT2 = 128, NW2 = NT2 / 32, SCAN2 = SCAN * NTHREADS / NT2;     // same 1024 faces per culling round

// cull_round for a CTA of NT2 threads (ascending face order preserved)
__device__ __forceinline__ int cull_round2(const Params &p, const float4 *__restrict__ bbox, int b, int base,
                                           float x_lo, float x_hi, float y_lo, float y_hi, int n0, int *s_list,
                                           int *s_cnt)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned ballots[SCAN2];
#pragma unroll
    for (int k = 0; k < SCAN2; k++) {
        const int fidx = base + k * NT2 + tid;
        bool hit = false;
        if (fidx < p.nf) {
            const float4 bb = __ldg(bbox + (size_t)b * p.nf + fidx);
            hit = !(x_lo > bb.y || x_hi < bb.x || y_lo > bb.w || y_hi < bb.z);
        }
        ballots[k] = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_cnt[k * NW2 + warp] = __popc(ballots[k]);
    }
    __syncthreads();
    int total = 0, my_off[SCAN2];
#pragma unroll
    for (int k = 0; k < SCAN2; k++) {
#pragma unroll
        for (int w = 0; w < NW2; w++) {
            if (w == warp) my_off[k] = total;
            total += s_cnt[k * NW2 + w];
        }
    }
#pragma unroll
    for (int k = 0; k < SCAN2; k++) {
        if (ballots[k] & (1u << lane))
            s_list[n0 + my_off[k] + __popc(ballots[k] & ((1u << lane) - 1u))] = base + k * NT2 + tid;
    }
    __syncthreads();
    return n0 + total;
}

// forward_pair without the loads: the per-(pixel, face) update of the aggregation state
template <int RGB, bool FAST>
__device__ __forceinline__ void forward_eval(const Params &p, const Face &f, float xp, float yp,
                                             const float *__restrict__ textures, const float *__restrict__ textures2,
                                             int b, FwdState &s)
{
    Frag fr;
    if (!eval_frag<FAST>(p, f, xp, yp, fr)) return;
    const int alpha_mode = FAST ? SCP_ALPHA_PROD : p.alpha_mode;
    if (alpha_mode == SCP_ALPHA_PROD) s.alpha *= 1.f - fr.frag;
    else if (alpha_mode == SCP_ALPHA_SUM) s.alpha += fr.frag;
    else if (fr.frag > 0.5f) s.alpha = 1.f;

    float wc[3];
    const float zp = clip_and_depth(f, fr.w, wc);
    if (zp < p.near_ || zp > p.far_) return;

    if (RGB == SCP_RGB_HARD || RGB == RGB_DUAL) {
        if (zp < s.depth_min && inside_closed(fr.w) && (p.double_side || f.front != 0.f)) {
            s.depth_min = zp;
            s.face_min = f.idx;
            if (RGB == RGB_DUAL) {
                const float *t2 = textures2 + ((size_t)b * p.nf + f.idx) * 9;
#pragma unroll
                for (int k = 0; k < 3; k++) s.col2[k] = wc[0] * __ldg(t2 + k) + wc[1] * __ldg(t2 + 3 + k) + wc[2] * __ldg(t2 + 6 + k);
            } else {
                sample_color<FAST>(p, f, wc, textures, b, s.col);
            }
        }
    }
    if ((RGB == SCP_RGB_SOFTMAX || RGB == RGB_DUAL) && (f.front != 0.f || p.double_side)) {
        const float zn = (p.far_ - zp) * p.inv_depth_range;
        float rescale = 1.f;
        if (zn > s.sm_max) {
            rescale = __expf((s.sm_max - zn) * p.inv_gamma);
            s.sm_max = zn;
        }
        const float ez = __expf((zn - s.sm_max) * p.inv_gamma) * fr.frag;
        s.sm_sum = rescale * s.sm_sum + ez;
        float c[3];
        sample_color<FAST>(p, f, wc, textures, b, c);
#pragma unroll
        for (int k = 0; k < 3; k++) s.col[k] = rescale * s.col[k] + ez * c[k];
    }
}

template <int RGB>
__device__ __forceinline__ void init_state(const Params &p, FwdState &s, int alpha_mode, bool valid,
                                           const float *__restrict__ soft_colors, int b, size_t plane, int pn)
{
    s.col[0] = s.col[1] = s.col[2] = 0.f;
    s.alpha = alpha_mode == SCP_ALPHA_PROD ? 1.f : 0.f;
    s.sm_sum = __expf(p.eps * p.inv_gamma);
    s.sm_max = p.eps;
    s.depth_min = 10000000.f;
    s.face_min = -1;
    if (valid) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float bg = soft_colors[((size_t)b * 4 + k) * plane + pn];
            s.col[k] = RGB == SCP_RGB_HARD ? bg : bg * s.sm_sum;
        }
    }
}

// epilogue of forward_kernel for one pixel
template <int RGB>
__device__ __forceinline__ void store_pixel(const Params &p, const FwdState &s, int alpha_mode, int b, size_t plane, int pn,
                                            float *__restrict__ aggrs_info, float *__restrict__ soft_colors,
                                            float *__restrict__ aggrs_info2, float *__restrict__ soft_colors2)
{
    float a_out;
    if (alpha_mode == SCP_ALPHA_PROD) a_out = 1.f - s.alpha;
    else if (alpha_mode == SCP_ALPHA_SUM) a_out = s.alpha / p.nf;
    else a_out = s.alpha;
    soft_colors[((size_t)b * 4 + 3) * plane + pn] = a_out;
    if (RGB == RGB_DUAL) {
        soft_colors2[((size_t)b * 4 + 3) * plane + pn] = a_out;
        if (s.face_min != -1) {
#pragma unroll
            for (int k = 0; k < 3; k++) soft_colors2[((size_t)b * 4 + k) * plane + pn] = s.col2[k];
        }
        aggrs_info2[((size_t)b * 2 + 0) * plane + pn] = s.depth_min;
        aggrs_info2[((size_t)b * 2 + 1) * plane + pn] = (float)s.face_min;
    }
    if (RGB == SCP_RGB_HARD) {
        if (s.face_min != -1) {
#pragma unroll
            for (int k = 0; k < 3; k++) soft_colors[((size_t)b * 4 + k) * plane + pn] = s.col[k];
        }
        aggrs_info[((size_t)b * 2 + 0) * plane + pn] = s.depth_min;
        aggrs_info[((size_t)b * 2 + 1) * plane + pn] = (float)s.face_min;
    } else {
        const float rs = 1.f / s.sm_sum;
#pragma unroll
        for (int k = 0; k < 3; k++) soft_colors[((size_t)b * 4 + k) * plane + pn] = s.col[k] * rs;
        aggrs_info[((size_t)b * 2 + 0) * plane + pn] = s.sm_sum;
        aggrs_info[((size_t)b * 2 + 1) * plane + pn] = s.sm_max;
    }
}

template <int RGB, bool FAST>
__global__ void __launch_bounds__(NT2, 4) forward_kernel2(Params p, const float4 *__restrict__ bbox,
                                                        const float *__restrict__ rec,
                                                        const int *__restrict__ img_bbox,
                                                        const float *__restrict__ textures,
                                                        float *__restrict__ aggrs_info,
                                                        float *__restrict__ soft_colors,
                                                        const float *__restrict__ textures2,
                                                        float *__restrict__ aggrs_info2,
                                                        float *__restrict__ soft_colors2)
{
    __shared__ int s_list[LIST_CAP];
    __shared__ int s_cnt[SCAN2 * NW2];

    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t plane = (size_t)p.is * p.is;
    const int alpha_mode = FAST ? SCP_ALPHA_PROD : p.alpha_mode;
    // tile and the warp's 8x8 block; lane = column (lane & 7) and the two rows (lane >> 3) and (lane >> 3) + 4
    const int tx0 = (blockIdx.x % p.tiles_x) * TILE, ty0 = (blockIdx.x / p.tiles_x) * TILE;
    const int bx = tx0 + (warp & 1) * 8, by = ty0 + (warp >> 1) * 8;
    const int pxx = bx + (lane & 7), pyA = by + (lane >> 3), pyB = pyA + 4;
    const bool validA = pxx < p.is && pyA < p.is, validB = pxx < p.is && pyB < p.is;
    const int pnA = pyA * p.is + pxx, pnB = pyB * p.is + pxx;
    const float xp = centre_x(pxx, p.is), ypA = centre_y(pyA, p.is), ypB = centre_y(pyB, p.is);
    const float wx_lo = centre_x(bx, p.is), wx_hi = centre_x(min(bx + 7, p.is - 1), p.is);
    const float wy_hi = centre_y(by, p.is), wy_lo = centre_y(min(by + 7, p.is - 1), p.is);
    const float x_lo = centre_x(tx0, p.is), x_hi = centre_x(min(tx0 + TILE - 1, p.is - 1), p.is);
    const float y_hi = centre_y(ty0, p.is), y_lo = centre_y(min(ty0 + TILE - 1, p.is - 1), p.is);

    FwdState sA, sB;
    init_state<RGB>(p, sA, alpha_mode, validA, soft_colors, b, plane, pnA);
    init_state<RGB>(p, sB, alpha_mode, validB, soft_colors, b, plane, pnB);

    if (!tile_outside_mesh(img_bbox, b, x_lo, x_hi, y_lo, y_hi)) {
        int base = 0;
        while (base < p.nf) {
            int n = 0;
            while (base < p.nf && n + SCAN2 * NT2 <= LIST_CAP) {
                n = cull_round2(p, bbox, b, base, x_lo, x_hi, y_lo, y_hi, n, s_list, s_cnt);
                base += SCAN2 * NT2;
            }
            for (int i0 = 0; i0 < n; i0 += 32) {
                const int fi = i0 + lane < n ? s_list[i0 + lane] : -1;
                bool hit = false;
                if (fi >= 0) {
                    const float4 bb = __ldg(bbox + (size_t)b * p.nf + fi);
                    hit = !(wx_lo > bb.y || wx_hi < bb.x || wy_lo > bb.w || wy_hi < bb.z);
                }
                unsigned todo = __ballot_sync(0xffffffffu, hit);
                while (todo) {
                    const int j = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int fsel = __shfl_sync(0xffffffffu, fi, j);
                    const float *r = rec + ((size_t)b * p.nf + fsel) * REC;
                    float bx0, bx1, by0, by1;
                    load_bbox(r, bx0, bx1, by0, by1);
                    const bool in_x = !(xp > bx1 || xp < bx0);
                    const bool inA = validA && in_x && !(ypA > by1 || ypA < by0);
                    const bool inB = validB && in_x && !(ypB > by1 || ypB < by0);
                    if (inA || inB) {
                        Face f;
                        load_face(r, f);
                        if (inA) forward_eval<RGB, FAST>(p, f, xp, ypA, textures, textures2, b, sA);
                        if (inB) forward_eval<RGB, FAST>(p, f, xp, ypB, textures, textures2, b, sB);
                    }
                }
            }
            if (base < p.nf) __syncthreads();  // the list is about to be rebuilt
        }
    }
    if (validA) store_pixel<RGB>(p, sA, alpha_mode, b, plane, pnA, aggrs_info, soft_colors, aggrs_info2, soft_colors2);
    if (validB) store_pixel<RGB>(p, sB, alpha_mode, b, plane, pnB, aggrs_info, soft_colors, aggrs_info2, soft_colors2);
}
#endif  // SCP_SOFTRAS_FWD_2PX

// ---- backward -----------------------------------------------------------------------------
// Folded butterfly: 18 per-lane values -> after 20 shuffles each lane holds the warp total of ONE
// value (index returned in `idx`; lanes holding padding get idx >= 18 and a zero total).
__device__ __forceinline__ float fold18(float (&v)[18], int lane, int &idx)
{
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
    float a[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {  // 18 -> 9
        const float send = b4 ? v[i] : v[9 + i], keep = b4 ? v[9 + i] : v[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    float c[5];
#pragma unroll
    for (int i = 0; i < 5; i++) {  // 9 -> 5 (upper half: entries 5..8, padded)
        const float hi = i < 4 ? a[5 + i] : 0.f;
        const float send = b3 ? a[i] : hi, keep = b3 ? hi : a[i];
        c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    float d[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {  // 5 -> 3 (upper: 3..4, padded)
        const float hi = i < 2 ? c[3 + i] : 0.f;
        const float send = b2 ? c[i] : hi, keep = b2 ? hi : c[i];
        d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    float e[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {  // 3 -> 2 (upper: 2, padded)
        const float hi = i < 1 ? d[2 + i] : 0.f;
        const float send = b1 ? d[i] : hi, keep = b1 ? hi : d[i];
        e[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    const float send = b0 ? e[0] : e[1], keep = b0 ? e[1] : e[0];
    const float tot = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    // position inside each level; invalid (padding) when a level index runs past its size
    const int i1 = b0 ? 1 : 0;                 // within e (size 2)
    const int i2 = (b1 ? 2 : 0) + i1;          // within d (size 3)
    const int i3 = (b2 ? 3 : 0) + i2;          // within c (size 5)
    const int i4 = (b3 ? 5 : 0) + i3;          // within a (size 9)
    const bool ok = (!b1 || i1 < 1) && (!b2 || i2 < 2) && (!b3 || i3 < 4);
    idx = ok ? (b4 ? 9 : 0) + i4 : 18;
    return tot;
}

// gradient contribution of one (pixel, face) pair, ADDED into gv[0..8] (face coordinates) / gv[9..17] (vertex colours)
template <int RGB, bool FAST>
__device__ __forceinline__ bool backward_pair_loaded(const Params &p, const Face &f, const Pixel &px,
                                                     const float *__restrict__ textures, int b, const float *g,
                                                     const float *out, float sm_sum, float sm_max, float *grad_textures,
                                                     float (&gv)[18])
{
    Frag fr;
    if (!eval_frag<FAST>(p, f, px.xp, px.yp, fr)) return false;
    const int alpha_mode = FAST ? SCP_ALPHA_PROD : p.alpha_mode;
    const int dist_mode = FAST ? SCP_DIST_EUCLIDEAN : p.dist_mode;

    float Gxy = g[3];
    if (alpha_mode == SCP_ALPHA_SUM) Gxy /= p.nf;
    else if (alpha_mode == SCP_ALPHA_PROD) Gxy *= __fdividef(1.f - out[3], fmaxf(1.f - fr.frag, 1e-6f));
    float wc[3];
    const float zp = clip_and_depth(f, fr.w, wc);
    if (zp < p.near_ || zp > p.far_) return false;  // face dropped incl. its alpha gradient (kernel.cu:599)

    float *gt = gv + 9;
    if (RGB == SCP_RGB_HARD) {
        if ((float)f.idx == sm_max) {  // aggrs_info[1] holds the winning face index in hard mode
            if (p.tex_mode == SCP_TEX_VERTEX) {
#pragma unroll
                for (int v = 0; v < 3; v++)
#pragma unroll
                    for (int k = 0; k < 3; k++) gt[3 * v + k] += wc[v] * g[k];
            } else if (p.T == 1) {
#pragma unroll
                for (int k = 0; k < 3; k++) gt[k] += g[k];
            } else {
                float *dst = grad_textures + ((size_t)b * p.nf + f.idx) * p.T * 3 + surface_texel(wc, p.R) * 3;
#pragma unroll
                for (int k = 0; k < 3; k++) atomicAdd(dst + k, g[k]);
            }
        }
    } else if (f.front != 0.f || p.double_side) {
        const float zn = (p.far_ - zp) * p.inv_depth_range;
        const float s = __fdividef(fr.frag * __expf((zn - sm_max) * p.inv_gamma), sm_sum);
        float c[3];
        sample_color<FAST>(p, f, wc, textures, b, c);
        float Q = 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) Q += g[k] * (c[k] - out[k]);
        if (p.tex_mode == SCP_TEX_VERTEX) {
#pragma unroll
            for (int v = 0; v < 3; v++)
#pragma unroll
                for (int k = 0; k < 3; k++) gt[3 * v + k] += s * (wc[v] * g[k]);
        } else if (p.T == 1) {
#pragma unroll
            for (int k = 0; k < 3; k++) gt[k] += s * g[k];
        } else {
            float *dst = grad_textures + ((size_t)b * p.nf + f.idx) * p.T * 3 + surface_texel(wc, p.R) * 3;
#pragma unroll
            for (int k = 0; k < 3; k++) atomicAdd(dst + k, s * g[k]);
        }
        Q *= s;
        Gxy += __fdividef(Q, fr.frag);
        const float Gz = Q * p.inv_gamma / (p.near_ - p.far_) * zp * zp;
#pragma unroll
        for (int v = 0; v < 3; v++) gv[3 * v + 2] += Gz * wc[v] * f.rz[v] * f.rz[v];
    }
    Gxy *= fr.frag * (1.f - fr.frag) * p.inv_sigma;  // sigmoid'
    if (dist_mode == SCP_DIST_EUCLIDEAN) {
#pragma unroll
        for (int v = 0; v < 3; v++) {
            const float s2 = 2.f * fr.sign * Gxy * fr.c[v];
            gv[3 * v + 0] += s2 * fr.dx;
            gv[3 * v + 1] += s2 * fr.dy;
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
                const float acc = -iq * f.inv[3 * v + 0] * px.xp + -iq * f.inv[3 * v + 1] * px.yp +
                                  -iq * f.inv[3 * v + 2];
                gv[3 * v + l] += acc * Gxy * sc;
            }
        }
    }
    return true;
}

template <int RGB, bool FAST>
__device__ __forceinline__ bool backward_pair(const Params &p, const float *__restrict__ r, const Pixel &px,
                                              const float *__restrict__ textures, int b, const float *g,
                                              const float *out, float sm_sum, float sm_max, float *grad_textures,
                                              float (&gv)[18])
{
    float bx0, bx1, by0, by1;
    load_bbox(r, bx0, bx1, by0, by1);
    if (px.xp > bx1 || px.xp < bx0 || px.yp > by1 || px.yp < by0) return false;
    Face f;
    load_face(r, f);
    return backward_pair_loaded<RGB, FAST>(p, f, px, textures, b, g, out, sm_sum, sm_max, grad_textures, gv);
}

// ---- backward, face-centric traversal ------------------------------------------------------------------------
// One warp per (image, face): the packed record is loaded ONCE into registers, the warp walks the face's inflated
// bounding box in 8x4 pixel blocks (one pixel per lane), accumulates the 18 gradients in registers across ALL blocks and
// reduces them across the warp once at the end -- no per-(warp, face) record loads / butterfly / global reductions as in
// the tile-centric kernel (41 % of its instructions, profiles/r1_softras_backward_tile_source_page.csv.gz), no atomics on
// grad_faces / grad_textures at all (one warp owns a face), deterministic summation order.  The per-pixel operands
// (incoming gradient, colours, aggregates: 10 floats) are re-read per block through L1.
constexpr int FACE_WARPS = SCP_SOFTRAS_FACE_WARPS;    // small CTAs: a warp that finishes its face early frees its slot (face sizes vary)
template <int RGB, bool FAST>
__global__ void __launch_bounds__(FACE_WARPS * 32, SCP_SOFTRAS_FACE_CTAS) backward_face_kernel(Params p, const float4 *__restrict__ bbox,
                                                                         const float *__restrict__ rec,
                                                                         const float *__restrict__ textures,
                                                                         const float *__restrict__ soft_colors,
                                                                         const float *__restrict__ aggrs_info,
                                                                         const float *__restrict__ grad_soft_colors,
                                                                         float *grad_faces, float *grad_textures)
{
    const int lane = threadIdx.x & 31;
    const long fg = (long)blockIdx.x * FACE_WARPS + (threadIdx.x >> 5);     // flat (image, face) index
    if (fg >= (long)p.B * p.nf) return;
    const int b = (int)(fg / p.nf);
    const float4 bb = __ldg(bbox + fg);
    // pixel range whose centres can lie inside the inflated bbox (one pixel of slack; backward_pair re-tests exactly):
    // xp = (2 px + 1 - is) / is,  yp = (is - 1 - 2 py) / is  (row 0 = top)
    const float is = (float)p.is;
    const int ix0 = max(0, (int)floorf((bb.x * is + is - 1.f) * 0.5f) - 1);
    const int ix1 = min(p.is - 1, (int)ceilf((bb.y * is + is - 1.f) * 0.5f) + 1);
    const int iy0 = max(0, (int)floorf((is - 1.f - bb.w * is) * 0.5f) - 1);
    const int iy1 = min(p.is - 1, (int)ceilf((is - 1.f - bb.z * is) * 0.5f) + 1);
    if (ix0 > ix1 || iy0 > iy1) return;
#if SCP_SOFTRAS_FACE_SMEM
    __shared__ Face s_face[FACE_WARPS];
    Face &f = s_face[threadIdx.x >> 5];
    if (lane == 0) load_face(rec + fg * REC, f);
    __syncwarp();
#else
    Face f;
    load_face(rec + fg * REC, f);
#endif
    const int ntex = p.T * 3 <= 9 ? p.T * 3 : 0;
    const size_t plane = (size_t)p.is * p.is;
    const float *gsc = grad_soft_colors + (size_t)b * 4 * plane, *sc = soft_colors + (size_t)b * 4 * plane;
    const float *ag = aggrs_info + (size_t)b * 2 * plane;
    float gv[18];
#pragma unroll
    for (int k = 0; k < 18; k++) gv[k] = 0.f;
    bool any = false;
    constexpr int BW = SCP_SOFTRAS_FACE_BW, BH = 32 / BW;
    static_assert(BW == 8 || BW == 16 || BW == 32, "block width");
#if SCP_SOFTRAS_FACE_LINEAR
    // the bounding box's pixels in row-major order, 32 consecutive ones per round: every lane holds a pixel of the box
    // except in the last round (8x4 blocks leave 19-25 % of the lanes outside a 35x37 / 18x20 pixel box)
    const int bw = ix1 - ix0 + 1, npx = bw * (iy1 - iy0 + 1);
    const float inv_bw = 1.f / (float)bw;
    (void)BH;
    for (int i0 = 0; i0 < npx; i0 += 32) {
        {
            Pixel px;
            const int i = i0 + lane;
            const int row = (int)(((float)i + 0.5f) * inv_bw);   // exact: the quotient is >= 0.5 / 256 away from an integer
            px.px = ix0 + (i - row * bw);
            px.py = iy0 + row;
            px.valid = i < npx;
#else
    for (int y0 = iy0; y0 <= iy1; y0 += BH) {
        for (int x0 = ix0; x0 <= ix1; x0 += BW) {
            Pixel px;
            px.px = x0 + (lane % BW);
            px.py = y0 + (lane / BW);
            px.valid = px.px <= ix1 && px.py <= iy1;
#endif
            px.pn = px.py * p.is + px.px;
            px.xp = centre_x(px.px, p.is);
            px.yp = centre_y(px.py, p.is);
            const bool in_box = px.valid && !(px.xp > bb.y || px.xp < bb.x || px.yp > bb.w || px.yp < bb.z);
            // all ten per-pixel operands are requested at once (one memory latency per block instead of two)
            float g[4] = { 0.f, 0.f, 0.f, 0.f }, out[4] = { 0.f, 0.f, 0.f, 0.f }, sm_sum = 1.f, sm_max = 0.f;
            if (in_box) {
#pragma unroll
                for (int k = 0; k < 4; k++) g[k] = __ldg(gsc + k * plane + px.pn);
#pragma unroll
                for (int k = 0; k < 4; k++) out[k] = __ldg(sc + k * plane + px.pn);
                sm_sum = __ldg(ag + px.pn);
                sm_max = __ldg(ag + plane + px.pn);
            }
            const bool lane_grad = in_box && (g[0] != 0.f || g[1] != 0.f || g[2] != 0.f || g[3] != 0.f);
            if (lane_grad)
                any |= backward_pair_loaded<RGB, FAST>(p, f, px, textures, b, g, out, sm_sum, sm_max, grad_textures, gv);
        }
    }
    if (!__any_sync(0xffffffffu, any)) return;
    int idx;
    const float tot = fold18(gv, lane, idx);
    if (tot != 0.f) {    // single writer per face: plain read-modify-write into the caller's zero-filled buffers
        if (idx < 9) grad_faces[fg * 9 + idx] += tot;
        else if (idx - 9 < ntex) grad_textures[fg * ntex + (idx - 9)] += tot;
    }
}

template <int RGB, bool FAST>
__global__ void __launch_bounds__(NTHREADS, SCP_SOFTRAS_TILE_CTAS) backward_kernel(Params p, const float4 *__restrict__ bbox,
                                                           const float *__restrict__ rec,
                                                           const int *__restrict__ img_bbox,
                                                           const float *__restrict__ textures,
                                                           const float *__restrict__ soft_colors,
                                                           const float *__restrict__ aggrs_info,
                                                           const float *__restrict__ grad_soft_colors,
                                                           float *grad_faces, float *grad_textures)
{
    __shared__ int s_list[LIST_CAP];
    __shared__ int s_cnt[SCAN * NWARPS];

    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const size_t plane = (size_t)p.is * p.is;
    Pixel px;
    float x_lo, x_hi, y_lo, y_hi;
    tile_setup(p, px, x_lo, x_hi, y_lo, y_hi);
    if (tile_outside_mesh(img_bbox, b, x_lo, x_hi, y_lo, y_hi)) return;
    // vertex textures and 1-texel surface textures ride in the 18-value fold; larger surface tables
    // (never used by this repo's renders) go straight to global atomics inside backward_pair
    const int ntex = p.T * 3 <= 9 ? p.T * 3 : 0;

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
    const bool lane_grad = px.valid && (g[0] != 0.f || g[1] != 0.f || g[2] != 0.f || g[3] != 0.f);
    // a tile whose incoming gradient is identically zero contributes nothing
    if (!__syncthreads_or(lane_grad)) return;
    const bool warp_grad = __any_sync(0xffffffffu, lane_grad);

    int base = 0;
    while (base < p.nf) {
        int n = 0;
        while (base < p.nf && n + SCAN * NTHREADS <= LIST_CAP) {
            n = cull_round(p, bbox, b, base, x_lo, x_hi, y_lo, y_hi, n, s_list, s_cnt);
            base += SCAN * NTHREADS;
        }
        if (warp_grad) {
            for (int i0 = 0; i0 < n; i0 += 32) {
                const int fi = i0 + lane < n ? s_list[i0 + lane] : -1;
                bool hit = false;
                if (fi >= 0) {
                    const float4 bb = __ldg(bbox + (size_t)b * p.nf + fi);
                    hit = !(px.wx_lo > bb.y || px.wx_hi < bb.x || px.wy_lo > bb.w || px.wy_hi < bb.z);
                }
                unsigned todo = __ballot_sync(0xffffffffu, hit);
                while (todo) {
                    const int j = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int fsel = __shfl_sync(0xffffffffu, fi, j);
                    float gv[18];
#pragma unroll
                    for (int k = 0; k < 18; k++) gv[k] = 0.f;
                    bool active = false;
                    if (lane_grad)
                        active = backward_pair<RGB, FAST>(p, rec + ((size_t)b * p.nf + fsel) * REC, px, textures, b, g,
                                                          out, sm_sum, sm_max, grad_textures, gv);
                    if (__any_sync(0xffffffffu, active)) {
                        int idx;
                        const float tot = fold18(gv, lane, idx);
                        if (tot != 0.f) {
                            const size_t fg = (size_t)b * p.nf + fsel;
                            if (idx < 9) atomicAdd(grad_faces + fg * 9 + idx, tot);
                            else if (idx - 9 < ntex) atomicAdd(grad_textures + fg * ntex + (idx - 9), tot);
                        }
                    }
                }
            }
        }
        if (base < p.nf) __syncthreads();
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
    p.inv_sigma = (float)(1.0 / (double)sigma);
    p.inv_gamma = (float)(1.0 / (double)gamma);
    p.inv_depth_range = (float)(1.0 / ((double)far_ - (double)near_));
    p.dist_mode = dist_mode; p.rgb_mode = rgb_mode; p.alpha_mode = alpha_mode; p.tex_mode = tex_mode;
    p.double_side = double_side;
    return true;
}

static size_t bbox_bytes(int B, int nf) { return (((size_t)B * nf * sizeof(float4)) + 255) / 256 * 256; }
static size_t rec_bytes(int B, int nf) { return (size_t)B * nf * REC * sizeof(float); }
static size_t imgbb_bytes(int B) { return (((size_t)B * 4 * sizeof(int)) + 255) / 256 * 256; }

}  // namespace softras
}  // namespace scp

using namespace scp::softras;

extern "C" size_t scp_softras_workspace_bytes(int B, int nf)
{
    if (B <= 0 || nf <= 0) return 0;
    return bbox_bytes(B, nf) + rec_bytes(B, nf) + imgbb_bytes(B);
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
    int *img_bbox = (int *)((char *)rec + rec_bytes(B, nf));
    const long nfaces = (long)B * nf;
    cudaMemsetAsync(img_bbox, 0x7f, (size_t)B * 4 * sizeof(int), st);
    pack_kernel<<<(unsigned)((nfaces + 255) / 256), 256, 0, st>>>(p, faces, textures, faces_info, 1, bbox, rec,
                                                                 img_bbox);
    const dim3 grid(p.tiles_x * p.tiles_x, B);
    const bool fast = func_id_dist == SCP_DIST_EUCLIDEAN && func_id_alpha == SCP_ALPHA_PROD;
    if (func_id_rgb == SCP_RGB_HARD) {
#if SCP_SOFTRAS_FWD_2PX
        if (fast) forward_kernel2<SCP_RGB_HARD, true><<<grid, NT2, 0, st>>>(p, bbox, rec, img_bbox, textures, aggrs_info, soft_colors, nullptr, nullptr, nullptr);
#else
        if (fast) forward_kernel<SCP_RGB_HARD, true><<<grid, NTHREADS, 0, st>>>(p, bbox, rec, img_bbox, textures, aggrs_info, soft_colors, nullptr, nullptr, nullptr);
#endif
        else forward_kernel<SCP_RGB_HARD, false><<<grid, NTHREADS, 0, st>>>(p, bbox, rec, img_bbox, textures, aggrs_info, soft_colors, nullptr, nullptr, nullptr);
    } else {
#if SCP_SOFTRAS_FWD_2PX
        if (fast) forward_kernel2<SCP_RGB_SOFTMAX, true><<<grid, NT2, 0, st>>>(p, bbox, rec, img_bbox, textures, aggrs_info, soft_colors, nullptr, nullptr, nullptr);
#else
        if (fast) forward_kernel<SCP_RGB_SOFTMAX, true><<<grid, NTHREADS, 0, st>>>(p, bbox, rec, img_bbox, textures, aggrs_info, soft_colors, nullptr, nullptr, nullptr);
#endif
        else forward_kernel<SCP_RGB_SOFTMAX, false><<<grid, NTHREADS, 0, st>>>(p, bbox, rec, img_bbox, textures, aggrs_info, soft_colors, nullptr, nullptr, nullptr);
    }
    return scp::check_launch("scp_softras_forward");
}

extern "C" int scp_softras_forward_dual(const float *faces, const float *textures_soft, const float *textures_hard,
                                        float *faces_info, float *aggrs_info_soft, float *soft_colors_soft,
                                        float *aggrs_info_hard, float *soft_colors_hard, int B, int nf, int image_size,
                                        float near_, float far_, float eps, float sigma_val, float dist_eps,
                                        float gamma_val, int double_side, void *workspace, size_t workspace_bytes,
                                        void *stream)
{
    Params p;
    if (!make_params(p, B, nf, 3, image_size, near_, far_, eps, sigma_val, SCP_DIST_EUCLIDEAN, dist_eps, gamma_val,
                     SCP_RGB_SOFTMAX, SCP_ALPHA_PROD, SCP_TEX_VERTEX, double_side) ||
        !textures_hard || !aggrs_info_hard || !soft_colors_hard) {
        scp::set_last_error("scp_softras_forward_dual: unsupported arguments (B=%d nf=%d is=%d)", B, nf, image_size);
        return -1;
    }
    if (!workspace || workspace_bytes < scp_softras_workspace_bytes(B, nf)) {
        scp::set_last_error("scp_softras_forward_dual: workspace too small");
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float4 *bbox = (float4 *)workspace;
    float *rec = (float *)((char *)workspace + bbox_bytes(B, nf));
    int *img_bbox = (int *)((char *)rec + rec_bytes(B, nf));
    const long nfaces = (long)B * nf;
    cudaMemsetAsync(img_bbox, 0x7f, (size_t)B * 4 * sizeof(int), st);
    pack_kernel<<<(unsigned)((nfaces + 255) / 256), 256, 0, st>>>(p, faces, textures_soft, faces_info, 1, bbox, rec,
                                                                 img_bbox);
    const dim3 grid(p.tiles_x * p.tiles_x, B);
#if SCP_SOFTRAS_FWD_2PX
    forward_kernel2<RGB_DUAL, true><<<grid, NT2, 0, st>>>(p, bbox, rec, img_bbox, textures_soft, aggrs_info_soft,
                                                          soft_colors_soft, textures_hard, aggrs_info_hard,
                                                          soft_colors_hard);
#else
    forward_kernel<RGB_DUAL, true><<<grid, NTHREADS, 0, st>>>(p, bbox, rec, img_bbox, textures_soft, aggrs_info_soft,
                                                              soft_colors_soft, textures_hard, aggrs_info_hard,
                                                              soft_colors_hard);
#endif
    return scp::check_launch("scp_softras_forward_dual");
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
    int *img_bbox = (int *)((char *)rec + rec_bytes(B, nf));
    const long nfaces = (long)B * nf;
    cudaMemsetAsync(img_bbox, 0x7f, (size_t)B * 4 * sizeof(int), st);
    pack_kernel<<<(unsigned)((nfaces + 255) / 256), 256, 0, st>>>(p, faces, textures, const_cast<float *>(faces_info),
                                                                 0, bbox, rec, img_bbox);
    const bool fast = func_id_dist == SCP_DIST_EUCLIDEAN && func_id_alpha == SCP_ALPHA_PROD;
    // Traversal: face-centric (one warp per face, record and gradient accumulators in registers, no atomics) for the
    // softmax-RGB renders -- the two backward launches of a training step; tile-centric (same traversal as the forward,
    // gradients reduced with global float reductions) for hard RGB, whose geometry gradient is the ill-conditioned alpha
    // term alone: there the two instantiations differ by the compiler's FMA contraction choices at
    // the 1e-2 level, and the tile kernel is the one pinned to the oracle.  SCP_SOFTRAS_BWD=tile / face forces one.
    const char *mode = getenv("SCP_SOFTRAS_BWD");
    const bool use_tile = mode && mode[0] == 't' ? true : (mode && mode[0] == 'f' ? false : func_id_rgb == SCP_RGB_HARD);
    if (use_tile) {
        const dim3 grid(p.tiles_x * p.tiles_x, B);
        if (func_id_rgb == SCP_RGB_HARD) {
            if (fast) backward_kernel<SCP_RGB_HARD, true><<<grid, NTHREADS, 0, st>>>(p, bbox, rec, img_bbox, textures, soft_colors, aggrs_info, grad_soft_colors, grad_faces, grad_textures);
            else backward_kernel<SCP_RGB_HARD, false><<<grid, NTHREADS, 0, st>>>(p, bbox, rec, img_bbox, textures, soft_colors, aggrs_info, grad_soft_colors, grad_faces, grad_textures);
        } else {
            if (fast) backward_kernel<SCP_RGB_SOFTMAX, true><<<grid, NTHREADS, 0, st>>>(p, bbox, rec, img_bbox, textures, soft_colors, aggrs_info, grad_soft_colors, grad_faces, grad_textures);
            else backward_kernel<SCP_RGB_SOFTMAX, false><<<grid, NTHREADS, 0, st>>>(p, bbox, rec, img_bbox, textures, soft_colors, aggrs_info, grad_soft_colors, grad_faces, grad_textures);
        }
    } else {
        const unsigned grid = (unsigned)((nfaces + FACE_WARPS - 1) / FACE_WARPS);
        const int nt = FACE_WARPS * 32;
        if (func_id_rgb == SCP_RGB_HARD) {
            if (fast) backward_face_kernel<SCP_RGB_HARD, true><<<grid, nt, 0, st>>>(p, bbox, rec, textures, soft_colors, aggrs_info, grad_soft_colors, grad_faces, grad_textures);
            else backward_face_kernel<SCP_RGB_HARD, false><<<grid, nt, 0, st>>>(p, bbox, rec, textures, soft_colors, aggrs_info, grad_soft_colors, grad_faces, grad_textures);
        } else {
            if (fast) backward_face_kernel<SCP_RGB_SOFTMAX, true><<<grid, nt, 0, st>>>(p, bbox, rec, textures, soft_colors, aggrs_info, grad_soft_colors, grad_faces, grad_textures);
            else backward_face_kernel<SCP_RGB_SOFTMAX, false><<<grid, nt, 0, st>>>(p, bbox, rec, textures, soft_colors, aggrs_info, grad_soft_colors, grad_faces, grad_textures);
        }
    }
    return scp::check_launch("scp_softras_backward");
}
