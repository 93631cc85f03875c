// Symmetry regulariser: surface-sample reconstruction + rotation + brute-force 1-nearest-neighbour in ONE kernel.
//
// Replaces, for CanonicalMesh.compute_symmetry_loss (model/module/mesh.py:53-62 of the reference), the chain
//   pytorch3d.ops.sample_points_from_meshes (the gather / barycentric combination part; the random face and weight draws
//   stay with the caller) -> sample_pts.bmm(symm_rots) -> chamfer_distance_single_way -> pytorch3d knn_points(K=1)
//   (model/util/chamfer.py:152-156): dists = squared distance of every mesh vertex to its nearest rotated sample.
// The reference materialises (k*B, 10000, 3) samples, their rotation, and a (k*B, N, 10000) search; here a CTA keeps the
// mesh's vertices in shared memory, rebuilds 512 rotated samples at a time from (face index, barycentric weights) and
// every thread scans them for its own vertex.  HBM traffic: k*B*S*(8 + 12) bytes of sample descriptors per vertex chunk
// (4 chunks of 256 vertices at N = 995) + the vertices; the work is k*B*N*S distance evaluations (1.27 G at B = 64),
// FP32-issue bound.
#include "../../include/scp_b200.h"
#include "scp_common.cuh"

namespace scp {
namespace sym {

constexpr int THREADS = 256, TILE = 512;

struct Rot { float r[9]; };

__device__ __forceinline__ float3 sample_point(const float *sv, const int *__restrict__ faces, long long f, const float *__restrict__ w3,
                                               const Rot &R)
{
    const int a0 = faces[f * 3], a1 = faces[f * 3 + 1], a2 = faces[f * 3 + 2];
    const float w0 = w3[0], w1 = w3[1], w2 = w3[2];
    // pts = w0 v0 + w1 v1 + w2 v2 (sample_points_from_meshes), y = pts . R (row vector times matrix, mesh.py:60)
    const float px = w0 * sv[a0 * 3] + w1 * sv[a1 * 3] + w2 * sv[a2 * 3];
    const float py = w0 * sv[a0 * 3 + 1] + w1 * sv[a1 * 3 + 1] + w2 * sv[a2 * 3 + 1];
    const float pz = w0 * sv[a0 * 3 + 2] + w1 * sv[a1 * 3 + 2] + w2 * sv[a2 * 3 + 2];
    return make_float3(px * R.r[0] + py * R.r[3] + pz * R.r[6], px * R.r[1] + py * R.r[4] + pz * R.r[7],
                       px * R.r[2] + py * R.r[5] + pz * R.r[8]);
}

// grid = (ceil(N / THREADS), k * B); dynamic smem = N*3 floats (rounded to 16 B) + TILE float4
__global__ void __launch_bounds__(THREADS)
nn_fwd_kernel(const float *__restrict__ pred_v, const int *__restrict__ faces, const long long *__restrict__ face_idx,
              const float *__restrict__ w, const float *__restrict__ rots, int k, int N, int S, float *__restrict__ dist,
              int *__restrict__ nn_idx)
{
    extern __shared__ float4 smem4[];
    float4 *sy = smem4;                                     // TILE rotated samples
    float *sv = reinterpret_cast<float *>(smem4 + TILE);    // N * 3 vertices of this mesh
    const int m = blockIdx.y, b = m / k, tid = threadIdx.x;
    const float *v = pred_v + (long)b * N * 3;
    for (int t = tid; t < N * 3; t += THREADS) sv[t] = v[t];
    Rot R;
#pragma unroll
    for (int i = 0; i < 9; i++) R.r[i] = rots[(m - b * k) * 9 + i];
    __syncthreads();
    const int n = blockIdx.x * THREADS + tid;
    const int nc = n < N ? n : N - 1;
    const float x0 = sv[nc * 3], x1 = sv[nc * 3 + 1], x2 = sv[nc * 3 + 2];
    float best = 3.0e38f;
    int bi = 0;
    const long base = (long)m * S;
    for (int s0 = 0; s0 < S; s0 += TILE) {
        __syncthreads();
        for (int t = tid; t < TILE; t += THREADS) {
            const int s = s0 + t;
            float3 y = make_float3(1e18f, 1e18f, 1e18f);      // past the end: never the nearest
            if (s < S) y = sample_point(sv, faces, face_idx[base + s], w + (base + s) * 3, R);
            sy[t] = make_float4(y.x, y.y, y.z, 0.f);
        }
        __syncthreads();
        const int cnt = min(TILE, S - s0);
#pragma unroll 8
        for (int t = 0; t < cnt; t++) {
            const float4 y = sy[t];                           // all lanes read the same address: broadcast
            const float dx = x0 - y.x, dy = x1 - y.y, dz = x2 - y.z;
            const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            if (d < best) { best = d; bi = s0 + t; }          // strict: the lowest sample index wins ties
        }
    }
    if (n < N) {
        dist[(long)m * N + n] = best;
        nn_idx[(long)m * N + n] = bi;
    }
}

// one thread per (mesh, vertex): d(dist)/d(vertex) directly and, through the nearest sample's barycentric combination,
// into the three vertices of its face; all k replicas of image b accumulate into g_pred_v[b]
__global__ void nn_bwd_kernel(const float *__restrict__ pred_v, const int *__restrict__ faces, const long long *__restrict__ face_idx,
                              const float *__restrict__ w, const float *__restrict__ rots, const int *__restrict__ nn_idx,
                              const float *__restrict__ g_dist, int k, int N, int S, long total, float *__restrict__ g_pred_v)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int m = (int)(i / N), n = (int)(i - (long)m * N), b = m / k;
    const float *v = pred_v + (long)b * N * 3;
    Rot R;
#pragma unroll
    for (int j = 0; j < 9; j++) R.r[j] = rots[(m - b * k) * 9 + j];
    const long js = (long)m * S + nn_idx[i];
    const long long f = face_idx[js];
    const float *w3 = w + js * 3;
    const float3 y = sample_point(v, faces, f, w3, R);
    const float g = 2.f * g_dist[i];
    const float gx = g * (v[n * 3] - y.x), gy = g * (v[n * 3 + 1] - y.y), gz = g * (v[n * 3 + 2] - y.z);
    float *gv = g_pred_v + (long)b * N * 3;
    atomicAdd(gv + n * 3, gx);
    atomicAdd(gv + n * 3 + 1, gy);
    atomicAdd(gv + n * 3 + 2, gz);
    // d/d(pts) = -(gx,gy,gz) . R^T
    const float px = -(gx * R.r[0] + gy * R.r[1] + gz * R.r[2]);
    const float py = -(gx * R.r[3] + gy * R.r[4] + gz * R.r[5]);
    const float pz = -(gx * R.r[6] + gy * R.r[7] + gz * R.r[8]);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int a = faces[f * 3 + c];
        const float wc = w3[c];
        atomicAdd(gv + a * 3, wc * px);
        atomicAdd(gv + a * 3 + 1, wc * py);
        atomicAdd(gv + a * 3 + 2, wc * pz);
    }
}

}  // namespace sym
}  // namespace scp

using namespace scp::sym;

static bool sym_args_ok(const void *a, const void *b, const void *c, const void *d, const void *e, int B, int k, int N, int S)
{
    return a && b && c && d && e && B > 0 && k > 0 && N > 0 && S > 0;
}

extern "C" int scp_symmetry_nn_forward(const float *pred_v, const int *faces, const long long *face_idx, const float *w,
                                       const float *rots, int B, int k, int N, int S, float *dist, int *nn_idx, void *stream)
{
    if (!sym_args_ok(pred_v, faces, face_idx, w, rots, B, k, N, S) || !dist || !nn_idx) {
        scp::set_last_error("scp_symmetry_nn_forward: bad arguments (B=%d k=%d N=%d S=%d)", B, k, N, S);
        return -1;
    }
    const size_t smem = (size_t)TILE * sizeof(float4) + (size_t)N * 3 * sizeof(float);
    if (smem > 200 * 1024) {
        scp::set_last_error("scp_symmetry_nn_forward: %d vertices do not fit in shared memory", N);
        return -1;
    }
    if (smem > 48 * 1024) cudaFuncSetAttribute(nn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    nn_fwd_kernel<<<dim3((N + THREADS - 1) / THREADS, B * k), THREADS, smem, (cudaStream_t)stream>>>(
        pred_v, faces, face_idx, w, rots, k, N, S, dist, nn_idx);
    return scp::check_launch("scp_symmetry_nn_forward");
}

extern "C" int scp_symmetry_nn_backward(const float *pred_v, const int *faces, const long long *face_idx, const float *w,
                                        const float *rots, const int *nn_idx, const float *g_dist, int B, int k, int N, int S,
                                        float *g_pred_v, void *stream)
{
    if (!sym_args_ok(pred_v, faces, face_idx, w, rots, B, k, N, S) || !nn_idx || !g_dist || !g_pred_v) {
        scp::set_last_error("scp_symmetry_nn_backward: bad arguments (B=%d k=%d N=%d S=%d)", B, k, N, S);
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(g_pred_v, 0, (size_t)B * N * 3 * sizeof(float), st);
    const long total = (long)B * k * N;
    nn_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pred_v, faces, face_idx, w, rots, nn_idx, g_dist, k, N, S,
                                                                total, g_pred_v);
    return scp::check_launch("scp_symmetry_nn_backward");
}
