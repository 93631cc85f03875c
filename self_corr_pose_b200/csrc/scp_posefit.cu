// Inference pose fit (Tester.pose_fitting -> estimateSimilarityTransform of the reference,
// model/util/umeyama.py:9-41,95-201): the two point-cloud passes of the batched RANSAC + Umeyama fit.
//
// The reference runs, per image, 100 RANSAC rounds in Python; every round evaluates the candidate transform on ALL n
// correspondences (evaluateModel, :143-159: residual norm over the points).  The batched host formulation
// (model/util/umeyama.py::fit_similarity_batch) fits all rounds of all images at once and then needs
//   (1) the residual table r[l][h] = || tgt_l - (s R)_lh src_l - t_lh ||_F over the image's real correspondences, and
//   (2) for the winning round of every image: the inlier set (per-point residual below the pass threshold), its share,
//       and the moments of the closed-form fit over the inliers (means, centred covariance, centred source variance).
// In torch ops (1) materialises an (L, H, n, 3) tensor -- 384 MB for 32 images of 10 k correspondences -- and (2) another
// dozen (L, n, 3) temporaries; here (1) reads the two point clouds once per 512-point chunk with the 100 candidates of the
// image in registers (thread = candidate, the points broadcast from shared memory), and (2) is one CTA per image making
// the two passes the closed form needs (means first, centred sums second) with fixed-order block reductions: results are
// bit-reproducible run to run.  The 3x3 SVDs stay a batched torch call on the host side.
#include "../../include/scp_b200.h"
#include "scp_common.cuh"

namespace scp {
namespace posefit {

constexpr int NT = 128;          // residual table: threads = candidates (H <= 128)
constexpr int CHUNK = 512;       // points per CTA of the residual table
constexpr int MT = 256;          // moments kernel: threads per image

// partial[l][c][h] = sum over the chunk's real points of | tgt - (A_lh src + t_lh) |^2
__global__ void __launch_bounds__(NT) residual_table_kernel(const float *__restrict__ src, const float *__restrict__ tgt,
                                                            const int *__restrict__ counts, const float *__restrict__ hypA,
                                                            const float *__restrict__ hypT, int n_max, int H, int nchunk,
                                                            float *__restrict__ partial)
{
    __shared__ float s_pts[CHUNK * 6];
    const int c = blockIdx.x, l = blockIdx.y, tid = threadIdx.x;
    const int n = min(counts[l], n_max), i0 = c * CHUNK, m = max(0, min(CHUNK, n - i0));
    const float *sp = src + ((size_t)l * n_max + i0) * 3, *tp = tgt + ((size_t)l * n_max + i0) * 3;
    for (int k = tid; k < 3 * m; k += NT) {
        s_pts[k] = sp[k];
        s_pts[CHUNK * 3 + k] = tp[k];
    }
    __syncthreads();
    if (tid >= H) return;
    const float *A = hypA + ((size_t)l * H + tid) * 9, *T = hypT + ((size_t)l * H + tid) * 3;
    const float a0 = A[0], a1 = A[1], a2 = A[2], a3 = A[3], a4 = A[4], a5 = A[5], a6 = A[6], a7 = A[7], a8 = A[8];
    const float t0 = T[0], t1 = T[1], t2 = T[2];
    float acc = 0.f;
    for (int i = 0; i < m; i++) {
        const float x = s_pts[3 * i], y = s_pts[3 * i + 1], z = s_pts[3 * i + 2];
        const float dx = s_pts[CHUNK * 3 + 3 * i] - (a0 * x + a1 * y + a2 * z + t0);
        const float dy = s_pts[CHUNK * 3 + 3 * i + 1] - (a3 * x + a4 * y + a5 * z + t1);
        const float dz = s_pts[CHUNK * 3 + 3 * i + 2] - (a6 * x + a7 * y + a8 * z + t2);
        acc += dx * dx + dy * dy + dz * dz;
    }
    partial[((size_t)l * nchunk + c) * H + tid] = acc;
}

// fixed-order block sum of K values per thread -> every thread gets the totals
template <int K>
__device__ __forceinline__ void block_sum(float (&v)[K], float *s_red /* [MT / 32][K] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = warp_sum(v[k]);
    __syncthreads();                 // s_red may still be read from the previous call
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) s_red[warp * K + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) {
        float t = 0.f;
        for (int w = 0; w < MT / 32; w++) t += s_red[w * K + k];
        v[k] = t;
    }
}

// per-point residual of the winning transform against the pass threshold: ONE statement for both passes of the moments
// kernel (the explicit fused multiply-adds fix the rounding, so the two passes select the same points)
__device__ __forceinline__ bool is_inlier(const float (&a)[12], float thr, float x, float y, float z, float u0, float u1,
                                          float u2)
{
    const float dx = u0 - fmaf(a[0], x, fmaf(a[1], y, fmaf(a[2], z, a[9])));
    const float dy = u1 - fmaf(a[3], x, fmaf(a[4], y, fmaf(a[5], z, a[10])));
    const float dz = u2 - fmaf(a[6], x, fmaf(a[7], y, fmaf(a[8], z, a[11])));
    return sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz))) < thr;
}

// One CTA per image.  Winning transform (A = s R, t), pass threshold, `found` flag (a round was accepted).
// out[l][0..17] = n_used, inlier count, mean src (3), mean tgt (3), centred covariance sum (9: cov[i][j] =
// sum (tgt_i - mt_i)(src_j - ms_j)), centred source square sum.  The point set is the inliers when the image is accepted
// (found and at least 10 % inliers, umeyama.py:29-31), all real points otherwise (keeps the closed form finite; the
// caller discards that result).
__global__ void __launch_bounds__(MT) inlier_moments_kernel(const float *__restrict__ src, const float *__restrict__ tgt,
                                                            const int *__restrict__ counts, const float *__restrict__ bestA,
                                                            const float *__restrict__ bestT, const float *__restrict__ pass_t,
                                                            const unsigned char *__restrict__ found, int n_max,
                                                            float *__restrict__ out)
{
    __shared__ float s_red[(MT / 32) * 10];
    const int l = blockIdx.x, tid = threadIdx.x;
    const int n = min(counts[l], n_max);
    const float *sp = src + (size_t)l * n_max * 3, *tp = tgt + (size_t)l * n_max * 3;
    const float *A = bestA + (size_t)l * 9, *T = bestT + (size_t)l * 3;
    const float a[12] = { A[0], A[1], A[2], A[3], A[4], A[5], A[6], A[7], A[8], T[0], T[1], T[2] };
    const float thr = pass_t[l];

    // pass 1: inlier count and the coordinate sums of both candidate point sets
    float v[7] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f }, w[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
    for (int i = tid; i < n; i += MT) {
        const float x = sp[3 * i], y = sp[3 * i + 1], z = sp[3 * i + 2];
        const float u0 = tp[3 * i], u1 = tp[3 * i + 1], u2 = tp[3 * i + 2];
        const bool in = is_inlier(a, thr, x, y, z, u0, u1, u2);
        w[0] += x; w[1] += y; w[2] += z; w[3] += u0; w[4] += u1; w[5] += u2;
        if (in) { v[0] += 1.f; v[1] += x; v[2] += y; v[3] += z; v[4] += u0; v[5] += u1; v[6] += u2; }
    }
    block_sum<7>(v, s_red);
    block_sum<6>(w, s_red);
    const float n_in = v[0];
    const bool accepted = found[l] != 0 && n > 0 && n_in / (float)n >= 0.1f;
    const float cnt = accepted ? n_in : (float)n;
    float ms[3], mt[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        ms[k] = (accepted ? v[1 + k] : w[k]) / cnt;
        mt[k] = (accepted ? v[4 + k] : w[3 + k]) / cnt;
    }
    // pass 2: centred covariance and source variance over the chosen set
    float q[10] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
    for (int i = tid; i < n; i += MT) {
        const float x = sp[3 * i], y = sp[3 * i + 1], z = sp[3 * i + 2];
        const float u0 = tp[3 * i], u1 = tp[3 * i + 1], u2 = tp[3 * i + 2];
        const bool use = !accepted || is_inlier(a, thr, x, y, z, u0, u1, u2);
        if (use) {
            const float xs = x - ms[0], ys = y - ms[1], zs = z - ms[2];
            const float xt = u0 - mt[0], yt = u1 - mt[1], zt = u2 - mt[2];
            q[0] += xt * xs; q[1] += xt * ys; q[2] += xt * zs;
            q[3] += yt * xs; q[4] += yt * ys; q[5] += yt * zs;
            q[6] += zt * xs; q[7] += zt * ys; q[8] += zt * zs;
            q[9] += xs * xs + ys * ys + zs * zs;
        }
    }
    block_sum<10>(q, s_red);
    if (tid == 0) {
        float *o = out + (size_t)l * 18;
        o[0] = cnt;
        o[1] = n_in;
#pragma unroll
        for (int k = 0; k < 3; k++) { o[2 + k] = ms[k]; o[5 + k] = mt[k]; }
#pragma unroll
        for (int k = 0; k < 10; k++) o[8 + k] = q[k];
    }
}

}  // namespace posefit
}  // namespace scp

using namespace scp::posefit;

extern "C" size_t scp_posefit_chunks(int n_max) { return n_max <= 0 ? 0 : (size_t)((n_max + CHUNK - 1) / CHUNK); }

extern "C" int scp_posefit_residual_table(const float *src, const float *tgt, const int *counts, const float *hyp_A,
                                          const float *hyp_t, int L, int n_max, int H, float *partial, void *stream)
{
    if (L <= 0 || L > 65535 || n_max <= 0 || H <= 0 || H > NT || !src || !tgt || !counts || !hyp_A || !hyp_t || !partial) {
        scp::set_last_error("scp_posefit_residual_table: bad arguments (L=%d n_max=%d H=%d; H <= %d)", L, n_max, H, NT);
        return -1;
    }
    const int nchunk = (int)scp_posefit_chunks(n_max);
    residual_table_kernel<<<dim3(nchunk, L), NT, 0, (cudaStream_t)stream>>>(src, tgt, counts, hyp_A, hyp_t, n_max, H, nchunk,
                                                                           partial);
    return scp::check_launch("scp_posefit_residual_table");
}

extern "C" int scp_posefit_inlier_moments(const float *src, const float *tgt, const int *counts, const float *best_A,
                                          const float *best_t, const float *pass_t, const unsigned char *found, int L,
                                          int n_max, float *out, void *stream)
{
    if (L <= 0 || n_max <= 0 || !src || !tgt || !counts || !best_A || !best_t || !pass_t || !found || !out) {
        scp::set_last_error("scp_posefit_inlier_moments: bad arguments (L=%d n_max=%d)", L, n_max);
        return -1;
    }
    inlier_moments_kernel<<<L, MT, 0, (cudaStream_t)stream>>>(src, tgt, counts, best_A, best_t, pass_t, found, n_max, out);
    return scp::check_launch("scp_posefit_inlier_moments");
}
