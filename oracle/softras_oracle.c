/*
 * oracle/softras_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp32 with the reference's fp64 promotions) of the
 * SoftRas soft-rasterize operator used by kywind/self-corr-pose:
 *   pre-pass   third-party/softras/soft_renderer/cuda/soft_rasterize_cuda_kernel.cu:245-305
 *   forward    ...soft_rasterize_cuda_kernel.cu:308-483  (helpers :24-58, :61-158, :178-194)
 *   backward   ...soft_rasterize_cuda_kernel.cu:486-668  (helpers :161-175, :197-217)
 * Buffer initialisation contract (caller pre-fills soft_colors with the
 * background colour, zeros everything else):
 *   third-party/softras/soft_renderer/functional/soft_rasterize.py:35,47-53,88-89
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product path never does.
 *
 * Parity pin: this restatement is checked (tests/test_oracle_softras.py) against
 * oracle/_ref/libsoftras_ref_cpu.so, which is the reference's own kernel source
 * compiled for the host through oracle/ref_shim.h (see oracle/Makefile).
 *
 * Written as one "pair evaluation" (pixel x face) shared by forward and
 * backward, instead of the reference's two duplicated kernel bodies.
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

enum { DIST_HARD = 0, DIST_BARY = 1, DIST_EUCLID = 2 };
enum { RGB_HARD = 0, RGB_SOFTMAX = 1 };
enum { ALPHA_HARD = 0, ALPHA_SUM = 1, ALPHA_PROD = 2 };
enum { TEX_SURFACE = 0, TEX_VERTEX = 1 };

typedef struct {
    int B, nf, T, R, is;
    float near_, far_, eps, sigma, dist_eps, gamma;
    int dist_mode, rgb_mode, alpha_mode, tex_mode, double_side;
} sr_params;

/* ---- small helpers ------------------------------------------------------ */

static inline float f_max3(float a, float b, float c) { float m = a > b ? a : b; return m > c ? m : c; }
static inline float f_min3(float a, float b, float c) { float m = a < b ? a : b; return m < c ? m : c; }

/* pixel centre in NDC, row 0 = top (kernel.cu:343-346); evaluated in double, rounded once */
static inline void pixel_centre(int pn, int is, float *xp, float *yp)
{
    const int yi = is - 1 - pn / is;
    const int xi = pn % is;
    *yp = (float)((2. * yi + 1. - is) / is);
    *xp = (float)((2. * xi + 1. - is) / is);
}

/* bbox reject with margin (kernel.cu:32-38) */
static inline int outside_bbox(float x, float y, const float *f, float margin)
{
    return x > f_max3(f[0], f[3], f[6]) + margin || x < f_min3(f[0], f[3], f[6]) - margin ||
           y > f_max3(f[1], f[4], f[7]) + margin || y < f_min3(f[1], f[4], f[7]) - margin;
}

/* front-facing test (kernel.cu:41-44) */
static inline int front_facing(const float *f)
{
    return (f[7] - f[1]) * (f[3] - f[0]) < (f[4] - f[1]) * (f[6] - f[0]);
}

static inline int inside_closed(const float *w)
{
    return w[0] <= 1 && w[0] >= 0 && w[1] <= 1 && w[1] >= 0 && w[2] <= 1 && w[2] >= 0;
}

/* clamp to [0,1] then renormalise (kernel.cu:53-58) */
static inline void clip_bary(float *w)
{
    for (int k = 0; k < 3; k++) {
        double c = w[k] < 1. ? (double)w[k] : 1.;
        c = c > 0. ? c : 0.;
        w[k] = (float)c;
    }
    const float s = w[0] + w[1] + w[2];
    const float wsum = (float)((double)s > 1e-5 ? (double)s : 1e-5);
    for (int k = 0; k < 3; k++) w[k] /= wsum;
}

/* foot of the perpendicular on edge (v0,v1) in barycentric form, from the Gram rows */
static inline void edge_foot(const float *sym, const float *w, int v0, float *t)
{
    const int v1 = (v0 + 1) % 3, v2 = (v0 + 2) % 3;
    float a[3];
    a[0] = sym[3 * v0 + 0] - sym[3 * v1 + 0];
    a[1] = sym[3 * v0 + 1] - sym[3 * v1 + 1];
    a[2] = sym[3 * v0 + 2] - sym[3 * v1 + 2];
    t[v0] = (w[0] * a[0] + w[1] * a[1] + w[2] * a[2] - a[v1]) / (a[v0] - a[v1]);
    t[v1] = 1 - t[v0];
    t[v2] = 0;
}

/* Euclidean point-to-triangle vector (kernel.cu:61-151).
 * Outputs sign (+1 inside / -1 outside), (dx,dy), and t[] = (closest-point barycentric) - w. */
static void euclid_distance(const float *f, const float *info, const float *w,
                            float xp, float yp, float *sign, float *dx, float *dy, float *t)
{
    const float *sym = info + 9;
    const float *obt = info + 18;
    if (w[0] > 0 && w[1] > 0 && w[2] > 0 && w[0] < 1 && w[1] < 1 && w[2] < 1) {
        float best = 100000000.f, bx = 0, by = 0;
        for (int k = 0; k < 3; k++) {
            float tk[3];
            edge_foot(sym, w, k, tk);
            tk[0] -= w[0]; tk[1] -= w[1]; tk[2] -= w[2];
            const float ex = tk[0] * f[0] + tk[1] * f[3] + tk[2] * f[6];
            const float ey = tk[0] * f[1] + tk[1] * f[4] + tk[2] * f[7];
            const float d2 = ex * ex + ey * ey;
            if (d2 < best) { best = d2; bx = ex; by = ey; t[0] = tk[0]; t[1] = tk[1]; t[2] = tk[2]; }
        }
        *dx = bx; *dy = by; *sign = 1;
        return;
    }
    int v0 = -1;
    if (w[1] <= 0 && w[2] <= 0) {
        v0 = 0;
        if (obt[0] == 1 && (xp - f[0]) * (f[6] - f[0]) + (yp - f[1]) * (f[7] - f[1]) > 0) v0 = 2;
    } else if (w[2] <= 0 && w[0] <= 0) {
        v0 = 1;
        if (obt[1] == 1 && (xp - f[3]) * (f[0] - f[3]) + (yp - f[4]) * (f[1] - f[4]) > 0) v0 = 0;
    } else if (w[0] <= 0 && w[1] <= 0) {
        v0 = 2;
        if (obt[2] == 1 && (xp - f[6]) * (f[3] - f[6]) + (yp - f[7]) * (f[4] - f[7]) > 0) v0 = 1;
    } else if (w[0] <= 0) v0 = 1;
    else if (w[1] <= 0) v0 = 2;
    else if (w[2] <= 0) v0 = 0;
    if (v0 < 0) {
        /* w has a component >= 1 with the others > 0: only reachable through rounding
         * (the components sum to 1).  The reference indexes with v0 = -1 here (UB);
         * we take the edge opposite the largest component. */
        v0 = (w[0] >= w[1] && w[0] >= w[2]) ? 1 : (w[1] >= w[2] ? 2 : 0);
    }
    edge_foot(sym, w, v0, t);
    for (int k = 0; k < 3; k++) {
        float c = t[k] > 0.f ? t[k] : 0.f;
        c = c < 1.f ? c : 1.f;
        t[k] = c - w[k];
    }
    *dx = t[0] * f[0] + t[1] * f[3] + t[2] * f[6];
    *dy = t[0] * f[1] + t[1] * f[4] + t[2] * f[7];
    *sign = -1;
}

/* signed squared smallest barycentric (kernel.cu:154-158) */
static inline float bary_distance(const float *w)
{
    float d = w[0] > w[1] ? (w[1] > w[2] ? w[2] : w[1]) : (w[0] > w[2] ? w[2] : w[0]);
    return d > 0 ? (float)pow(d, 2) : (float)-pow(d, 2);
}

/* surface-texture texel index (kernel.cu:181-188), clamped to the table */
static inline int surface_texel(const float *w, int wx, int wy, int R)
{
    int idx = (w[0] + w[1]) * R - wx - wy <= 1 ? wy * R + wx : (R - 1 - wy) * R + (R - 1 - wx);
    if (idx < 0) idx = 0;
    if (idx > R * R - 1) idx = R * R - 1;
    return idx;
}

/* texture sampling (kernel.cu:178-194) */
static inline float sample_tex(const float *tex, const float *w, int R, int k, int mode)
{
    if (mode == TEX_SURFACE) {
        const int wx = (int)(w[0] * R), wy = (int)(w[1] * R);
        /* texel index clamped into the R*R table: the reference reads one texel past the face's
         * table when a clipped weight is exactly 1 (e.g. R = 1, the mask render); not reproduced. */
        return tex[surface_texel(w, wx, wy, R) * 3 + k];
    }
    return w[0] * tex[k] + w[1] * tex[3 + k] + w[2] * tex[6 + k];
}

/* d(colour)/d(texel j) (kernel.cu:197-217) */
static inline float sample_tex_grad(float g, const float *w, int R, int j, int mode)
{
    if (mode == TEX_SURFACE) {
        const int wx = (int)(w[0] * R), wy = (int)(w[1] * R);
        return j == surface_texel(w, wx, wy, R) ? g : 0.f;
    }
    return w[j] * g;
}

/* One (pixel, face) fragment: everything both passes need.  Returns 0 when the
 * face is skipped for this pixel (bbox / distance reject). */
typedef struct {
    float w[3];      /* raw barycentric */
    float t[3];      /* closest-point barycentric minus w (euclid) or w (bary) */
    float sign, dx, dy, dis, frag;
} sr_pair;

static int eval_pair(const sr_params *p, const float *f, const float *info, float xp, float yp,
                     float threshold, float margin, sr_pair *o)
{
    if (outside_bbox(xp, yp, f, margin)) return 0;
    o->w[0] = info[0] * xp + info[1] * yp + info[2];
    o->w[1] = info[3] * xp + info[4] * yp + info[5];
    o->w[2] = info[6] * xp + info[7] * yp + info[8];
    o->sign = 0; o->dx = o->dy = o->dis = 0;
    if (p->dist_mode == DIST_HARD) {
        if (!inside_closed(o->w)) return 0;
        o->frag = 1.f;
    } else if (p->dist_mode == DIST_BARY) {
        o->dis = bary_distance(o->w);
        o->t[0] = o->w[0]; o->t[1] = o->w[1]; o->t[2] = o->w[2];
        if (-o->dis >= threshold) return 0;
        o->frag = (float)(1. / (1. + expf(-o->dis / p->sigma)));
    } else {
        euclid_distance(f, info, o->w, xp, yp, &o->sign, &o->dx, &o->dy, o->t);
        o->dis = o->dx * o->dx + o->dy * o->dy;
        if (o->sign < 0 && o->dis >= threshold) return 0;
        o->frag = (float)(1. / (1. + expf(-o->sign * o->dis / p->sigma)));
    }
    return 1;
}

/* ---- pre-pass (kernel.cu:245-305) --------------------------------------- */

void scp_oracle_softras_prepass(const float *faces, float *faces_info, int B, int nf)
{
    for (long i = 0; i < (long)B * nf; i++) {
        const float *f = faces + i * 9;
        float *inv = faces_info + i * 27, *sym = inv + 9, *obt = inv + 18;
        const float x0 = f[0], y0 = f[1], x1 = f[3], y1 = f[4], x2 = f[6], y2 = f[7];
        const float adj[9] = {
            y1 - y2, x2 - x1, x1 * y2 - x2 * y1,
            y2 - y0, x0 - x2, x2 * y0 - x0 * y2,
            y0 - y1, x1 - x0, x0 * y1 - x1 * y0 };
        float det = x2 * (y0 - y1) + x0 * (y1 - y2) + x1 * (y2 - y0);
        /* clamp |det| >= 1e-10 keeping the sign; the literal is double in the reference */
        det = det > 0 ? (float)((double)det > 1e-10 ? (double)det : 1e-10)
                      : (float)((double)det < -1e-10 ? (double)det : -1e-10);
        for (int k = 0; k < 9; k++) inv[k] = adj[k] / det;
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++)
                sym[3 * j + k] = f[3 * j] * f[3 * k] + f[3 * j + 1] * f[3 * k + 1] + 1;
        /* only the FIRST obtuse corner is flagged; buffer is zero-initialised by the caller */
        const float px[3] = { x0, x1, x2 }, py[3] = { y0, y1, y2 };
        for (int k = 0; k < 3; k++) {
            const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
            if ((px[k1] - px[k]) * (px[k2] - px[k]) + (py[k1] - py[k]) * (py[k2] - py[k]) < 0) {
                obt[k] = 1;
                break;
            }
        }
    }
}

/* ---- forward (kernel.cu:308-483) ---------------------------------------- */

static void forward_pixel(const sr_params *p, int bn, int pn, const float *faces, const float *textures,
                          const float *faces_info, float *aggrs_info, float *soft_colors)
{
    const int is = p->is, nf = p->nf;
    const size_t plane = (size_t)is * is;
    float xp, yp;
    pixel_centre(pn, is, &xp, &yp);
    const float threshold = p->dist_eps * p->sigma;
    const float margin = sqrtf(threshold);

    float col[4] = { 1.f, 1.f, 1.f, 0.f };
    if (p->alpha_mode == ALPHA_PROD) col[3] = 1.f;
    float sm_sum = expf(p->eps / p->gamma);
    float sm_max = p->eps;
    for (int k = 0; k < 3; k++) {
        const float bg = soft_colors[((size_t)bn * 4 + k) * plane + pn];
        col[k] = p->rgb_mode == RGB_HARD ? bg : bg * sm_sum;
    }
    float depth_min = 10000000.f;
    int face_min = -1;

    for (int fn = 0; fn < nf; fn++) {
        const float *f = faces + ((size_t)bn * nf + fn) * 9;
        const float *tex = textures + ((size_t)bn * nf + fn) * p->T * 3;
        const float *info = faces_info + ((size_t)bn * nf + fn) * 27;
        sr_pair pr;
        if (!eval_pair(p, f, info, xp, yp, threshold, margin, &pr)) continue;

        /* alpha is accumulated BEFORE the depth test (kernel.cu:408-417) */
        if (p->alpha_mode == ALPHA_HARD) { if (pr.frag > 0.5) col[3] = 1.f; }
        else if (p->alpha_mode == ALPHA_SUM) col[3] += pr.frag;
        else col[3] = (float)((double)col[3] * (1. - pr.frag));

        float wc[3] = { pr.w[0], pr.w[1], pr.w[2] };
        clip_bary(wc);
        const float zp = (float)(1. / (wc[0] / f[2] + wc[1] / f[5] + wc[2] / f[8]));
        if (zp < p->near_ || zp > p->far_) continue;

        if (p->rgb_mode == RGB_HARD) {
            if (zp < depth_min && inside_closed(pr.w) && (p->double_side || front_facing(f))) {
                depth_min = zp;
                face_min = fn;
                for (int k = 0; k < 3; k++) col[k] = sample_tex(tex, wc, p->R, k, p->tex_mode);
            }
        } else if (front_facing(f) || p->double_side) {
            const float zn = (p->far_ - zp) / (p->far_ - p->near_);
            float rescale = 1.f;
            if (zn > sm_max) {
                rescale = expf((sm_max - zn) / p->gamma);
                sm_max = zn;
            }
            const float ez = expf((zn - sm_max) / p->gamma);
            sm_sum = rescale * sm_sum + ez * pr.frag;
            for (int k = 0; k < 3; k++) {
                const float c = sample_tex(tex, wc, p->R, k, p->tex_mode);
                col[k] = rescale * col[k] + ez * pr.frag * c;
            }
        }
    }

    float *a_out = soft_colors + ((size_t)bn * 4 + 3) * plane + pn;
    if (p->alpha_mode == ALPHA_HARD) *a_out = col[3];
    else if (p->alpha_mode == ALPHA_SUM) *a_out = col[3] / nf;
    else *a_out = (float)(1. - col[3]);

    if (p->rgb_mode == RGB_HARD) {
        if (face_min != -1)
            for (int k = 0; k < 3; k++) soft_colors[((size_t)bn * 4 + k) * plane + pn] = col[k];
        aggrs_info[((size_t)bn * 2 + 0) * plane + pn] = depth_min;
        aggrs_info[((size_t)bn * 2 + 1) * plane + pn] = (float)face_min;
    } else {
        for (int k = 0; k < 3; k++) soft_colors[((size_t)bn * 4 + k) * plane + pn] = col[k] / sm_sum;
        aggrs_info[((size_t)bn * 2 + 0) * plane + pn] = sm_sum;
        aggrs_info[((size_t)bn * 2 + 1) * plane + pn] = sm_max;
    }
}

/* ---- backward (kernel.cu:486-668) --------------------------------------- */

static void backward_pixel(const sr_params *p, int bn, int pn, const float *faces, const float *textures,
                           const float *soft_colors, const float *faces_info, const float *aggrs_info,
                           float *grad_faces, float *grad_textures, const float *grad_soft_colors)
{
    const int is = p->is, nf = p->nf;
    const size_t plane = (size_t)is * is;
    float xp, yp;
    pixel_centre(pn, is, &xp, &yp);
    const float threshold = p->dist_eps * p->sigma;
    const float margin = sqrtf(threshold);
    const float sm_sum = aggrs_info[((size_t)bn * 2 + 0) * plane + pn];
    const float sm_max = aggrs_info[((size_t)bn * 2 + 1) * plane + pn]; /* = face index in hard mode */
    float g[4], out[4];
    for (int k = 0; k < 4; k++) {
        g[k] = grad_soft_colors[((size_t)bn * 4 + k) * plane + pn];
        out[k] = soft_colors[((size_t)bn * 4 + k) * plane + pn];
    }

    for (int fn = 0; fn < nf; fn++) {
        const float *f = faces + ((size_t)bn * nf + fn) * 9;
        const float *tex = textures + ((size_t)bn * nf + fn) * p->T * 3;
        const float *info = faces_info + ((size_t)bn * nf + fn) * 27;
        sr_pair pr;
        if (!eval_pair(p, f, info, xp, yp, threshold, margin, &pr)) continue;

        float *gf = grad_faces + ((size_t)bn * nf + fn) * 9;
        float *gt = grad_textures + ((size_t)bn * nf + fn) * p->T * 3;
        float gv[3][3] = { { 0 } };
        float Gxy = 0;

        float Ga = g[3];
        if (p->alpha_mode == ALPHA_SUM) Ga /= nf;
        else if (p->alpha_mode == ALPHA_PROD) {
            const float om = 1 - pr.frag;
            const double den = (double)om > 1e-6 ? (double)om : 1e-6;
            Ga = (float)((double)Ga * ((double)(1 - out[3]) / den));
        }
        Gxy += Ga;

        float wc[3] = { pr.w[0], pr.w[1], pr.w[2] };
        clip_bary(wc);
        const float zp = (float)(1. / (wc[0] / f[2] + wc[1] / f[5] + wc[2] / f[8]));
        if (zp < p->near_ || zp > p->far_) continue;  /* drops this face's alpha gradient too */

        if (p->rgb_mode == RGB_HARD) {
            if ((float)fn == sm_max)
                for (int k = 0; k < 3; k++)
                    for (int j = 0; j < p->T; j++)
                        gt[3 * j + k] += sample_tex_grad(g[k], wc, p->R, j, p->tex_mode);
        } else if (front_facing(f) || p->double_side) {
            float Q = 0.f;
            const float zn = (p->far_ - zp) / (p->far_ - p->near_);
            const float s = pr.frag * expf((zn - sm_max) / p->gamma) / sm_sum;
            for (int k = 0; k < 3; k++) {
                for (int j = 0; j < p->T; j++)
                    gt[3 * j + k] += s * sample_tex_grad(g[k], wc, p->R, j, p->tex_mode);
                const float c = sample_tex(tex, wc, p->R, k, p->tex_mode);
                Q += g[k] * (c - out[k]);
            }
            Q *= s;
            Gxy += Q / pr.frag;
            const float Gz = Q / p->gamma / (p->near_ - p->far_) * zp * zp;
            gv[0][2] = Gz * wc[0] / f[2] / f[2];
            gv[1][2] = Gz * wc[1] / f[5] / f[5];
            gv[2][2] = Gz * wc[2] / f[8] / f[8];
        }

        Gxy *= pr.frag * (1 - pr.frag) / p->sigma;
        if (p->dist_mode == DIST_BARY) {
            /* kernel.cu:161-175 */
            const float *t = pr.t;
            const int q = t[0] > t[1] ? (t[1] > t[2] ? 2 : 1) : (t[0] > t[2] ? 2 : 0);
            for (int l = 0; l < 2; l++)
                for (int k = 0; k < 3; k++) {
                    float acc = 0;
                    for (int c = 0; c < 3; c++)
                        acc += -info[3 * q + l] * info[3 * k + c] * (c == 0 ? xp : (c == 1 ? yp : 1));
                    float v = acc * Gxy;
                    v = (float)((double)v * (pr.dis > 0 ? 2. * sqrtf(pr.dis) : 2. * sqrtf(-pr.dis)));
                    gv[k][l] = v;
                }
        } else if (p->dist_mode == DIST_EUCLID) {
            for (int k = 0; k < 3; k++) {
                gv[k][0] = 2 * pr.sign * Gxy * (pr.t[k] + pr.w[k]) * pr.dx;
                gv[k][1] = 2 * pr.sign * Gxy * (pr.t[k] + pr.w[k]) * pr.dy;
            }
        }
        for (int k = 0; k < 3; k++)
            for (int l = 0; l < 3; l++) gf[3 * k + l] += gv[k][l];
    }
}

/* ---- exported entry points ---------------------------------------------- */

static sr_params make_params(int B, int nf, int T, int is, float near_, float far_, float eps, float sigma,
                             int dist_mode, float dist_eps, float gamma, int rgb_mode, int alpha_mode,
                             int tex_mode, int double_side)
{
    sr_params p;
    p.B = B; p.nf = nf; p.T = T; p.R = (int)sqrt((double)T); p.is = is;
    p.near_ = near_; p.far_ = far_; p.eps = eps; p.sigma = sigma; p.dist_eps = dist_eps; p.gamma = gamma;
    p.dist_mode = dist_mode; p.rgb_mode = rgb_mode; p.alpha_mode = alpha_mode; p.tex_mode = tex_mode;
    p.double_side = double_side;
    return p;
}

/* mirrors forward_soft_rasterize(...) of soft_rasterize_cuda.cpp:59-91; nthreads<=0 -> all cores */
int scp_oracle_softras_forward(const float *faces, const float *textures, float *faces_info, float *aggrs_info,
                               float *soft_colors, int B, int nf, int T, int is, float near_, float far_,
                               float eps, float sigma, int dist_mode, float dist_eps, float gamma, int rgb_mode,
                               int alpha_mode, int tex_mode, int double_side, int nthreads)
{
    const sr_params p = make_params(B, nf, T, is, near_, far_, eps, sigma, dist_mode, dist_eps, gamma, rgb_mode,
                                    alpha_mode, tex_mode, double_side);
    scp_oracle_softras_prepass(faces, faces_info, B, nf);
    const long total = (long)B * is * is;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 256)
#endif
    for (long i = 0; i < total; i++)
        forward_pixel(&p, (int)(i / ((long)is * is)), (int)(i % ((long)is * is)), faces, textures, faces_info,
                      aggrs_info, soft_colors);
    (void)nthreads;
    return 0;
}

/* mirrors backward_soft_rasterize(...) of soft_rasterize_cuda.cpp:94-132.  Sequential over pixels
 * inside one image (deterministic accumulation order); images run in parallel. */
int scp_oracle_softras_backward(const float *faces, const float *textures, const float *soft_colors,
                                const float *faces_info, const float *aggrs_info, float *grad_faces,
                                float *grad_textures, const float *grad_soft_colors, int B, int nf, int T, int is,
                                float near_, float far_, float eps, float sigma, int dist_mode, float dist_eps,
                                float gamma, int rgb_mode, int alpha_mode, int tex_mode, int double_side,
                                int nthreads)
{
    const sr_params p = make_params(B, nf, T, is, near_, far_, eps, sigma, dist_mode, dist_eps, gamma, rgb_mode,
                                    alpha_mode, tex_mode, double_side);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int bn = 0; bn < B; bn++)
        for (int pn = 0; pn < is * is; pn++)
            backward_pixel(&p, bn, pn, faces, textures, soft_colors, faces_info, aggrs_info, grad_faces,
                           grad_textures, grad_soft_colors);
    (void)nthreads;
    return 0;
}
