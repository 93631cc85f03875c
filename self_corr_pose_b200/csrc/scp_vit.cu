// Frozen DINO ViT-S/8 feature extractor for sm_100a: layer-9 key features of every image.
//
// Replaces DINO.forward (model/module/network/dino.py:102-109) -> VisionTransformer.get_specific_tokens
// (third-party/zsp/zsp/method/vision_transformer_flexible.py:249-262) of the reference.  Only what the
// consumer uses is computed: blocks 0..8 and the key projection of block 9 (the reference also runs
// blocks 10-11, materialises every attention map and feeds each image 4 times).
//
//   patch embedding, QKV / proj / fc1 / fc2 / K9 projections : persistent tcgen05 GEMM (scp_gemm.cuh), TMA-fed,
//        accumulators in TMEM, fused epilogues (bias, +pos-embed, head-major q/k/v split, residual add, exact
//        GELU, transposed feature store)
//   attention : flash-style kernel (scores never leave registers), bf16 tensor-core MMA, online softmax
//   LayerNorm : one warp per token, fp32 statistics, bf16 output feeding the next TMA load
// Activations: residual stream fp32, GEMM operands bf16, accumulation fp32.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "../../include/scp_b200.h"
#include "scp_common.cuh"
#include "scp_gemm.cuh"
#include "scp_fa.cuh"
#include "scp_fa2.cuh"
#include "scp_fa3.cuh"

namespace scp {
namespace vit {

constexpr int D = 384, HEADS = 6, HD = 64, MLP = 1536, PATCH = 8, KP = 3 * PATCH * PATCH;  // 192

typedef __nv_bfloat16 bf16;

// ---- small kernels --------------------------------------------------------------------------------
// im2col for the 8x8/8 patch convolution: A0[b*np + p][c*64 + dy*8 + dx] = img[b][c][8*py+dy][8*px+dx]
__global__ void im2col_kernel(const float *__restrict__ img, bf16 *__restrict__ out, int B, int H, int W)
{
    const int pw = W / PATCH, ph = H / PATCH;
    const long total = (long)B * ph * pw * 3 * PATCH;  // one thread per (patch, c, dy): 8 contiguous pixels
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int dy = i % PATCH;
    const int c = (i / PATCH) % 3;
    const long patch = i / (3 * PATCH);
    const int px = patch % pw, py = (patch / pw) % ph, b = patch / ((long)pw * ph);
    const float *src = img + (((long)b * 3 + c) * H + py * PATCH + dy) * W + px * PATCH;
    const float4 lo = *reinterpret_cast<const float4 *>(src), hi = *reinterpret_cast<const float4 *>(src + 4);
    __align__(16) bf16 v[8] = { __float2bfloat16(lo.x), __float2bfloat16(lo.y), __float2bfloat16(lo.z), __float2bfloat16(lo.w),
                               __float2bfloat16(hi.x), __float2bfloat16(hi.y), __float2bfloat16(hi.z), __float2bfloat16(hi.w) };
    *reinterpret_cast<uint4 *>(out + patch * KP + c * PATCH * PATCH + dy * PATCH) = *reinterpret_cast<const uint4 *>(v);
}

// CLS rows of the residual stream: x[b][0][:] = cls_token + pos_embed[0]
__global__ void cls_kernel(float *__restrict__ x, const float *__restrict__ cls_pos0, int B, int T)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * D) x[(long)(i / D) * T * D + (i % D)] = cls_pos0[i % D];
}

// LayerNorm over D = 384 (eps 1e-6), one warp per token, fp32 in -> bf16 out
__global__ void layernorm_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b,
                                 bf16 *__restrict__ y, long M)
{
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    const float4 *xr = reinterpret_cast<const float4 *>(x + row * D);
    float4 v[3];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        v[i] = xr[lane + 32 * i];
        s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mean = warp_sum(s) * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + 1e-6f);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int c = 4 * (lane + 32 * i);
        const float4 ww = *reinterpret_cast<const float4 *>(w + c), bb = *reinterpret_cast<const float4 *>(b + c);
        __align__(8) bf16 o[4] = { __float2bfloat16(v[i].x * rstd * ww.x + bb.x), __float2bfloat16(v[i].y * rstd * ww.y + bb.y),
                                  __float2bfloat16(v[i].z * rstd * ww.z + bb.z), __float2bfloat16(v[i].w * rstd * ww.w + bb.w) };
        *reinterpret_cast<uint2 *>(y + row * D + c) = *reinterpret_cast<const uint2 *>(o);
    }
}

// ---- x3 (fp32-class) variants: outputs are SPLIT bf16 pairs in the i32 layout of scp_gemm.cuh -------------------
// physical element index of logical column c (hi part; the lo part is 32 elements further)
__device__ __forceinline__ int i32_col(int c) { return ((c >> 5) << 6) + (c & 31); }

__global__ void im2col_x3_kernel(const float *__restrict__ img, bf16 *__restrict__ out, int B, int H, int W)
{
    const int pw = W / PATCH, ph = H / PATCH;
    const long total = (long)B * ph * pw * 3 * PATCH;  // one thread per (patch, c, dy): 8 contiguous pixels
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int dy = i % PATCH;
    const int c = (i / PATCH) % 3;
    const long patch = i / (3 * PATCH);
    const int px = patch % pw, py = (patch / pw) % ph, b = patch / ((long)pw * ph);
    const float *src = img + (((long)b * 3 + c) * H + py * PATCH + dy) * W + px * PATCH;
    const float4 a = *reinterpret_cast<const float4 *>(src), d = *reinterpret_cast<const float4 *>(src + 4);
    uint32_t h[4], l[4];
    h[0] = scp::gemm::split_bf16x2(a.x, a.y, l[0]);
    h[1] = scp::gemm::split_bf16x2(a.z, a.w, l[1]);
    h[2] = scp::gemm::split_bf16x2(d.x, d.y, l[2]);
    h[3] = scp::gemm::split_bf16x2(d.z, d.w, l[3]);
    bf16 *dst = out + patch * (2 * KP) + i32_col(c * PATCH * PATCH + dy * PATCH);
    *reinterpret_cast<uint4 *>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4 *>(dst + 32) = make_uint4(l[0], l[1], l[2], l[3]);
}

// LayerNorm over D = 384 (eps 1e-6), one warp per token, fp32 in -> split bf16 pairs out (row pitch 2 D)
__global__ void layernorm_x3_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b,
                                    bf16 *__restrict__ y, long M)
{
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    const float4 *xr = reinterpret_cast<const float4 *>(x + row * D);
    float4 v[3];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        v[i] = xr[lane + 32 * i];
        s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mean = warp_sum(s) * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + 1e-6f);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int c = 4 * (lane + 32 * i);
        const float4 ww = *reinterpret_cast<const float4 *>(w + c), bb = *reinterpret_cast<const float4 *>(b + c);
        uint32_t h[2], l[2];
        h[0] = scp::gemm::split_bf16x2(v[i].x * rstd * ww.x + bb.x, v[i].y * rstd * ww.y + bb.y, l[0]);
        h[1] = scp::gemm::split_bf16x2(v[i].z * rstd * ww.z + bb.z, v[i].w * rstd * ww.w + bb.w, l[1]);
        bf16 *dst = y + row * (2 * D) + i32_col(c);
        *reinterpret_cast<uint2 *>(dst) = make_uint2(h[0], h[1]);
        *reinterpret_cast<uint2 *>(dst + 32) = make_uint2(l[0], l[1]);
    }
}

// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16 output rounding of 4e-3): one exp,
// one reciprocal, five FMAs instead of erff's ~30-instruction branchy polynomial
__device__ __forceinline__ float erf_as(float x)
{
    const float a = fabsf(x);
    const float t = __fdividef(1.f, 1.f + 0.3275911f * a);
    const float poly = t * (0.254829592f + t * (-0.284496736f + t * (1.421413741f + t * (-1.453152027f + t * 1.061405429f))));
    const float y = 1.f - poly * __expf(-a * a);
    return copysignf(y, x);
}

// ---- GEMM epilogues (see scp_gemm.cuh: staged = lane is the column, coalesced along a row) ---------------
struct EpiPatch {  // tokens: x[b][1+p][:] = acc + bias + pos[p]
    static constexpr bool kStaged = true;
    static constexpr bool kTmaReduceAdd = false;
    static constexpr bool kMixed = false;
    float *x; const float *bias, *pos; int np, T;
    __device__ __forceinline__ void chunk(int row0, int nrows, int col, const float *stg, int lane) const
    {
        int b = row0 / np, p = row0 - b * np;
        const float bb = __ldg(bias + col);
        for (int r = 0; r < nrows; r++) {
            x[((long)b * T + 1 + p) * D + col] = stg[r * 33 + lane] + bb + __ldg(pos + (long)p * D + col);
            if (++p == np) { p = 0; b++; }
        }
    }
};

struct EpiQKV {  // head-major split: q/k[b][h][t][64] bf16, and v TRANSPOSED vT[b][h][64][Tp] (keys contiguous: the
                 // K-major B operand of the P.V tensor-core product); Tp = T rounded up to 8, pad columns stay zero
    static constexpr bool kStaged = true;
    static constexpr bool kTmaReduceAdd = false;
    static constexpr bool kMixed = true;
    bf16 *q, *k, *vt; const float *bias; int T, Tp;
    __device__ __forceinline__ bool direct(int col0) const { return col0 >= 2 * D; }
    __device__ __forceinline__ void chunk(int row0, int nrows, int col, const float *stg, int lane) const
    {
        const int which = col / D, c = col - which * D, h = c >> 6, d = c & 63;
        int b = row0 / T, t = row0 - b * T;
        bf16 *base = (which == 0 ? q : k) + (long)h * T * HD + d;
        const float bb = __ldg(bias + col);
#pragma unroll 4
        for (int r = 0; r < nrows; r++) {
            base[((long)b * HEADS * T + t) * HD] = __float2bfloat16(stg[r * 33 + lane] + bb);
            if (++t == T) { t = 0; b++; }
        }
    }
    // v columns, lane = row (token): consecutive lanes write consecutive tokens of one vT row
    __device__ __forceinline__ void operator()(int row, int col0, const float (&a)[32]) const
    {
        const int c = col0 - 2 * D, h = c >> 6, d0 = c & 63;
        const int b = row / T, t = row - b * T;
        bf16 *dst = vt + (((long)b * HEADS + h) * HD + d0) * Tp + t;
#pragma unroll
        for (int i = 0; i < 32; i++) dst[(long)i * Tp] = __float2bfloat16(a[i] + __ldg(bias + col0 + i));
    }
};

struct EpiQKVTokens {  // q | k token-major bf16 in one [M][768] matrix (bias added; register -> swizzled box -> TMA store, no
                       // per-image row arithmetic: GEMM rows ARE the token rows); v TRANSPOSED per head as in EpiQKV
    static constexpr bool kStaged = false;
    static constexpr bool kTmaReduceAdd = false;
    static constexpr bool kMixed = true;
    static constexpr bool kTmaStoreBf16 = true;
    bf16 *vt; const float *bias; int T, Tp;
    __device__ __forceinline__ float apply(float z) const { return z; }
    __device__ __forceinline__ bool direct(int col0) const { return col0 >= 2 * D; }
    __device__ __forceinline__ void operator()(int row, int col0, const float (&a)[32]) const
    {
        const int c = col0 - 2 * D, h = c >> 6, d0 = c & 63;
        const int b = row / T, t = row - b * T;
        bf16 *dst = vt + (((long)b * HEADS + h) * HD + d0) * Tp + t;
        const float4 *b4 = reinterpret_cast<const float4 *>(bias + col0);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float4 bb = __ldg(b4 + i);
            dst[(long)(4 * i) * Tp] = __float2bfloat16(a[4 * i] + bb.x);
            dst[(long)(4 * i + 1) * Tp] = __float2bfloat16(a[4 * i + 1] + bb.y);
            dst[(long)(4 * i + 2) * Tp] = __float2bfloat16(a[4 * i + 2] + bb.z);
            dst[(long)(4 * i + 3) * Tp] = __float2bfloat16(a[4 * i + 3] + bb.w);
        }
    }
};

struct EpiQKVTokens3 {  // x3 mode: q | k split pairs through the TMA-store epilogue into one [M][1536] matrix (i32 layout);
                        // v TRANSPOSED per head as two planes vt[2][B*6*64][Tp] (hi, lo)
    static constexpr bool kStaged = false;
    static constexpr bool kTmaReduceAdd = false;
    static constexpr bool kMixed = true;
    static constexpr bool kTmaStoreBf16 = true;
    bf16 *vt; const float *bias; int T, Tp; long plane;
    __device__ __forceinline__ float apply(float z) const { return z; }
    __device__ __forceinline__ bool direct(int col0) const { return col0 >= 2 * D; }
    __device__ __forceinline__ void operator()(int row, int col0, const float (&a)[32]) const
    {
        const int c = col0 - 2 * D, h = c >> 6, d0 = c & 63;
        const int b = row / T, t = row - b * T;
        bf16 *dst = vt + (((long)b * HEADS + h) * HD + d0) * Tp + t;
        const float4 *b4 = reinterpret_cast<const float4 *>(bias + col0);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float4 bb = __ldg(b4 + i);
            const float z[4] = { a[4 * i] + bb.x, a[4 * i + 1] + bb.y, a[4 * i + 2] + bb.z, a[4 * i + 3] + bb.w };
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const bf16 hi = __float2bfloat16(z[j]);
                dst[(long)(4 * i + j) * Tp] = hi;
                dst[plane + (long)(4 * i + j) * Tp] = __float2bfloat16(z[j] - __bfloat162float(hi));
            }
        }
    }
};

struct EpiQKVPlain {  // q/k/v[b][h][t][64] bf16 (layout of the mma.sync attention variant)
    static constexpr bool kStaged = true;
    static constexpr bool kTmaReduceAdd = false;
    static constexpr bool kMixed = false;
    bf16 *q, *k, *v; const float *bias; int T;
    __device__ __forceinline__ void chunk(int row0, int nrows, int col, const float *stg, int lane) const
    {
        const int which = col / D, c = col - which * D, h = c >> 6, d = c & 63;
        int b = row0 / T, t = row0 - b * T;
        bf16 *base = (which == 0 ? q : (which == 1 ? k : v)) + (long)h * T * HD + d;
        const float bb = __ldg(bias + col);
#pragma unroll 4
        for (int r = 0; r < nrows; r++) {
            base[((long)b * HEADS * T + t) * HD] = __float2bfloat16(stg[r * 33 + lane] + bb);
            if (++t == T) { t = 0; b++; }
        }
    }
};

struct EpiResidual {  // x[tile] += acc + bias: fp32 residual stream updated by TMA reduce-add (no SM loads)
    static constexpr bool kStaged = false;
    static constexpr bool kTmaReduceAdd = true;
    static constexpr bool kMixed = false;
    const float *bias;
    __device__ void operator()(int, int, const float (&)[32]) const {}
};

struct EpiGelu {  // h[row][:] = gelu_erf(acc + bias)  bf16, through the register -> swizzled box -> TMA store epilogue
    static constexpr bool kStaged = false;
    static constexpr bool kTmaReduceAdd = false;
    static constexpr bool kMixed = false;
    static constexpr bool kTmaStoreBf16 = true;
    const float *bias;
    // exact-erf GELU with erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7), folded so that the sign handling and
    // the 0.5 (1 + erf) disappear:  0.5 z (1 + erf(z / sqrt 2)) = max(z, 0) - |z| * [0.5 P(t)] * exp(-z^2 / 2),
    // t = 1 / (1 + p |z| / sqrt 2)  -- 14 instructions per element (2 on the MUFU pipe); the epilogue of this GEMM is
    // issue-bound (ncu: 71 % issue slots with two epilogue warps per scheduler)
    __device__ __forceinline__ float apply(float z) const
    {
        const float a = fabsf(z);
        const float t = __fdividef(1.f, fmaf(0.3275911f * 0.70710678118654752f, a, 1.f));
        const float hp = t * (0.5f * 0.254829592f + t * (0.5f * -0.284496736f + t * (0.5f * 1.421413741f +
                         t * (0.5f * -1.453152027f + t * (0.5f * 1.061405429f)))));
        const float w = z * 0.84932180028801904f;              // sqrt(log2(e) / 2): exp(-z^2/2) = 2^(-w^2)
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-w * w));
        return fmaf(-a, hp * e, fmaxf(z, 0.f));
    }
    __device__ void operator()(int, int, const float (&)[32]) const {}
};

struct EpiKeys {  // feat[b][col][t-1] = acc + bias for patch tokens (CLS dropped); (b, 384, hp, wp) fp32
    static constexpr bool kStaged = false;   // output is contiguous along the rows (tokens): lane = row
    static constexpr bool kTmaReduceAdd = false;
    static constexpr bool kMixed = false;
    float *feat; const float *bias; int T;
    bf16 *tokens;   // optional second copy, token-major bf16 [b][t-1][384]: the K-major operand of the arg-match GEMM
    int split;      // x3 mode: the token copy is written as split pairs, [b][t-1][768] in the i32 layout
    __device__ void operator()(int row, int col0, const float (&a)[32]) const
    {
        const int b = row / T, t = row - b * T;
        if (t == 0) return;
        float *dst = feat + ((long)b * D + col0) * (T - 1) + (t - 1);
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] = a[i] + bias[col0 + i];
#pragma unroll
        for (int i = 0; i < 32; i++) dst[(long)i * (T - 1)] = v[i];
        if (tokens != nullptr && split) {
            uint4 *q = reinterpret_cast<uint4 *>(tokens + ((long)b * (T - 1) + (t - 1)) * (2 * D) + 2 * col0);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint32_t h[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; e++) h[e] = scp::gemm::split_bf16x2(v[8 * i + 2 * e], v[8 * i + 2 * e + 1], l[e]);
                q[i] = make_uint4(h[0], h[1], h[2], h[3]);
                q[4 + i] = make_uint4(l[0], l[1], l[2], l[3]);
            }
        } else if (tokens != nullptr) {
            uint4 *q = reinterpret_cast<uint4 *>(tokens + ((long)b * (T - 1) + (t - 1)) * D + col0);
#pragma unroll
            for (int i = 0; i < 4; i++)
                q[i] = make_uint4(scp::gemm::pack_bf16x2(v[8 * i], v[8 * i + 1]), scp::gemm::pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                  scp::gemm::pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), scp::gemm::pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
        }
    }
};

// arg-max of the masked feature similarity (model/module/pretrained_corr.py:85-89): best[row] = max over the unmasked
// columns of (ordered similarity bits << 32 | ~column) -- larger similarity wins, the lower column wins ties
struct EpiArgmax {
    static constexpr bool kStaged = false;
    static constexpr bool kTmaReduceAdd = false;
    static constexpr bool kMixed = false;
    unsigned long long *best; const float *col_mask; int rows_per_batch, ncols;
    __device__ void operator()(int row, int col0, const float (&a)[32]) const
    {
        const int p = row / rows_per_batch;
        const float4 *cm = reinterpret_cast<const float4 *>(col_mask + (long)p * ncols + col0);
        float bv = 0.f;
        int bi = -1;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float4 m = __ldg(cm + i);
            const float mk[4] = { m.x, m.y, m.z, m.w };
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float x = a[4 * i + j];
                if (mk[j] > 0.f && (bi < 0 || x > bv)) { bv = x; bi = col0 + 4 * i + j; }
            }
        }
        if (bi >= 0) {
            const uint32_t u = __float_as_uint(bv);
            const uint32_t key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // order-preserving float -> uint
            atomicMax(best + row, ((unsigned long long)key << 32) | (uint32_t)(0xffffffffu - (uint32_t)bi));
        }
    }
};

struct EpiPlain {  // C[row][:] = acc (+ bias)   fp32, used by the exported test GEMM
    static constexpr bool kStaged = true;
    static constexpr bool kTmaReduceAdd = false;
    static constexpr bool kMixed = false;
    float *c; const float *bias; int ld;
    __device__ __forceinline__ void chunk(int row0, int nrows, int col, const float *stg, int lane) const
    {
        const float bb = bias ? __ldg(bias + col) : 0.f;
        for (int r = 0; r < nrows; r++) c[(long)(row0 + r) * ld + col] = stg[r * 33 + lane] + bb;
    }
};

// ---- attention: flash-style, bf16 mma.sync m16n8k16, 64 queries per CTA (4 warps x 16), 64-key tiles ------
constexpr int AQ = 64, AK = 64, APAD = 72;  // smem rows padded to 144 B -> conflict-free ldmatrix

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const bf16 *p)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const bf16 *p)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ void cp16(void *s, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(s)), "l"(g));
}

// q,k,v: [B*HEADS][T][64] bf16;  o: [B][T][HEADS*64] bf16;  scale = 64^-0.5 applied to q.k^T (vit:90)
__global__ void __launch_bounds__(128)
attention_kernel(const bf16 *__restrict__ q, const bf16 *__restrict__ k, const bf16 *__restrict__ v, bf16 *__restrict__ o,
                 int T, float scale_log2e)
{
    __shared__ __align__(16) bf16 sQ[AQ * APAD];
    __shared__ __align__(16) bf16 sK[2][AK * APAD];
    __shared__ __align__(16) bf16 sV[2][AK * APAD];

    const int bh = blockIdx.y, q0 = blockIdx.x * AQ, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const bf16 *qb = q + (long)bh * T * HD, *kb = k + (long)bh * T * HD, *vb = v + (long)bh * T * HD;
    const int ntiles = (T + AK - 1) / AK;

    // rows past T are clamped to the last row (their results are never stored / are masked)
    auto load_tile = [&](bf16 *dst, const bf16 *src, int r0) {
        for (int i = tid; i < AK * (HD / 8); i += 128) {
            const int r = i >> 3, c = (i & 7) * 8;
            cp16(dst + r * APAD + c, src + (long)min(r0 + r, T - 1) * HD + c);
        }
    };
    load_tile(sQ, qb, q0);
    load_tile(sK[0], kb, 0);
    load_tile(sV[0], vb, 0);
    asm volatile("cp.async.commit_group;");

    uint32_t qf[4][4];            // Q fragments for the 4 k16 steps over d
    float oacc[8][4];             // 16 x 64 output tile
    float m_run[2] = { -1e30f, -1e30f }, l_run[2] = { 0.f, 0.f };
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) oacc[i][j] = 0.f;

    for (int it = 0; it < ntiles; it++) {
        asm volatile("cp.async.wait_group 0;");
        __syncthreads();
        if (it + 1 < ntiles) {
            load_tile(sK[(it + 1) & 1], kb, (it + 1) * AK);
            load_tile(sV[(it + 1) & 1], vb, (it + 1) * AK);
            asm volatile("cp.async.commit_group;");
        }
        if (it == 0) {
#pragma unroll
            for (int ks = 0; ks < 4; ks++)
                ldsm_x4(qf[ks], sQ + (warp * 16 + (lane & 15)) * APAD + ks * 16 + (lane >> 4) * 8);
        }
        const bf16 *Kt = sK[it & 1], *Vt = sV[it & 1];

        // S = Q K^T : 16 x 64 per warp
        float sacc[8][4];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) sacc[i][j] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
#pragma unroll
            for (int np = 0; np < 4; np++) {   // pairs of n8 key tiles
                uint32_t kf[4];
                // matrices: (keys 16np+0..7, d lo), (keys +0..7, d hi), (keys +8..15, d lo), (keys +8..15, d hi)
                ldsm_x4(kf, Kt + (np * 16 + (lane & 7) + ((lane >> 4) << 3)) * APAD + ks * 16 + ((lane >> 3) & 1) * 8);
                mma_bf16(sacc[2 * np], qf[ks], kf[0], kf[1]);
                mma_bf16(sacc[2 * np + 1], qf[ks], kf[2], kf[3]);
            }
        }
        // mask keys past T, online softmax (rows g and g+8 of this warp's 16)
        const int key0 = it * AK;
        float mx[2] = { m_run[0], m_run[1] };
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int key = key0 + nt * 8 + 2 * t4 + (j & 1);
                const float s = key < T ? sacc[nt][j] * scale_log2e : -1e30f;
                sacc[nt][j] = s;
                mx[j >> 1] = fmaxf(mx[j >> 1], s);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; r++) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        }
        float corr[2], rs[2] = { 0.f, 0.f };
#pragma unroll
        for (int r = 0; r < 2; r++) { corr[r] = exp2f(m_run[r] - mx[r]); m_run[r] = mx[r]; }
        uint32_t pf[4][4];   // P as A fragments for the 4 k16 steps over keys
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            const float p0 = exp2f(sacc[nt][0] - mx[0]), p1 = exp2f(sacc[nt][1] - mx[0]);
            const float p2 = exp2f(sacc[nt][2] - mx[1]), p3 = exp2f(sacc[nt][3] - mx[1]);
            rs[0] += p0 + p1; rs[1] += p2 + p3;
            pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
            pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
        }
#pragma unroll
        for (int r = 0; r < 2; r++) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            oacc[nt][0] *= corr[0]; oacc[nt][1] *= corr[0]; oacc[nt][2] *= corr[1]; oacc[nt][3] *= corr[1];
        }
        // O += P V : V^T fragments through ldmatrix.trans
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {       // keys 16ks..16ks+15
#pragma unroll
            for (int dp = 0; dp < 4; dp++) {   // pairs of n8 tiles over d
                uint32_t vf[4];
                // matrices: (keys lo, d 16dp+0..7), (keys hi, d +0..7), (keys lo, d +8..15), (keys hi, d +8..15)
                ldsm_x4_t(vf, Vt + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * APAD + dp * 16 + (lane >> 4) * 8);
                mma_bf16(oacc[2 * dp], pf[ks], vf[0], vf[1]);
                mma_bf16(oacc[2 * dp + 1], pf[ks], vf[2], vf[3]);
            }
        }
    }
    // finalize: full row sums across the quad, normalise, store o[b][t][h*64 + d]
#pragma unroll
    for (int r = 0; r < 2; r++) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const int b = bh / HEADS, h = bh - b * HEADS;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int t = q0 + warp * 16 + g + 8 * r;
        if (t < T) {
            const float inv = 1.f / l_run[r];
            bf16 *dst = o + ((long)b * T + t) * D + h * HD + 2 * t4;
#pragma unroll
            for (int nt = 0; nt < 8; nt++)
                *reinterpret_cast<uint32_t *>(dst + nt * 8) = pack_bf16(oacc[nt][2 * r] * inv, oacc[nt][2 * r + 1] * inv);
        }
    }
}

}  // namespace vit
}  // namespace scp

using namespace scp::vit;

// ---- C ABI ------------------------------------------------------------------------------------------
extern "C" int scp_gemm_bf16_tn(const void *A, const void *W, const float *bias, float *C, int M, int N, int K,
                                void *stream)
{
    EpiPlain epi{ C, bias, N };
    int rc = scp::gemm::launch(A, K, W, K, M, N, K, epi, (cudaStream_t)stream);
    return rc ? rc : scp::check_launch("scp_gemm_bf16_tn");
}

extern "C" int scp_attention_bf16(const void *q, const void *k, const void *v, void *o, int B, int T, void *stream)
{
    if (B <= 0 || T <= 0) { scp::set_last_error("scp_attention_bf16: bad shape"); return -1; }
    const float scale_log2e = 0.125f * 1.4426950408889634f;
    attention_kernel<<<dim3((T + AQ - 1) / AQ, B * HEADS), 128, 0, (cudaStream_t)stream>>>(
        (const bf16 *)q, (const bf16 *)k, (const bf16 *)v, (bf16 *)o, T, scale_log2e);
    return scp::check_launch("scp_attention_bf16");
}

// SCP_VIT_ATTENTION: unset / "2" = fa2 (S double-buffered, P and O in TMEM), "1" = fa (first tcgen05 version),
// "m" = mma.sync variant (different V layout; only through scp_vit_s8_keys)
static int attention_variant()
{
    const char *e = getenv("SCP_VIT_ATTENTION");
    if (!e || !e[0]) return 2;
    return e[0] == 'm' ? 0 : (e[0] == '1' ? 1 : 2);
}

// token_major (fa2 only): q = k = the [B*T][768] q|k matrix of EpiQKVTokens
static int launch_fa(const bf16 *q, const bf16 *k, const bf16 *vt, bf16 *o, int B, int T, int Tp, cudaStream_t st,
                     bool token_major = false)
{
    const bool v2 = attention_variant() != 1;
    CUtensorMap tq, tk, tv;
    const uint64_t rows = token_major ? (uint64_t)B * T : (uint64_t)B * HEADS * T;
    const uint64_t inner = token_major ? 2 * D : HD;
    const uint32_t bq = v2 ? scp::fa2::BQ : scp::fa::BQ, bkv = v2 ? scp::fa2::BKV : scp::fa::BKV;
    if (token_major && !v2) {
        scp::set_last_error("tcgen05 attention: the token-major q|k layout needs variant 2");
        return -1;
    }
    if (!scp::gemm::make_tmap_bf16(&tq, q, inner, rows, inner, bq) || !scp::gemm::make_tmap_bf16(&tk, k, inner, rows, inner, bkv) ||
        !scp::gemm::make_tmap_bf16(&tv, vt, Tp, (uint64_t)B * HEADS * HD, Tp, 64)) {
        scp::set_last_error("tcgen05 attention: cuTensorMapEncodeTiled failed");
        return -1;
    }
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(scp::fa::fa_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, scp::fa::SMEM_BYTES);
        cudaFuncSetAttribute(scp::fa2::fa2_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, scp::fa2::SMEM_BYTES);
        attr_done = true;
    }
    const float scale_log2e = 0.125f * 1.4426950408889634f;
    if (v2)
        scp::fa2::fa2_fwd_kernel<<<dim3((T + bq - 1) / bq, B * HEADS), scp::fa2::NTHREADS, scp::fa2::SMEM_BYTES, st>>>(
            tq, tk, tv, o, T, scale_log2e, token_major ? 1 : 0);
    else
        scp::fa::fa_fwd_kernel<<<dim3((T + bq - 1) / bq, B * HEADS), scp::fa::NTHREADS, scp::fa::SMEM_BYTES, st>>>(
            tq, tk, tv, o, T, scale_log2e);
    return 0;
}

// q,k: [B*6][T][64] bf16; vt: [B*6][64][Tp] bf16 with Tp = ceil(T/8)*8 and zero padding; o: [B][T][384] bf16
extern "C" int scp_attention_tc5(const void *q, const void *k, const void *vt, void *o, int B, int T, void *stream)
{
    if (B <= 0 || T <= 0) { scp::set_last_error("scp_attention_tc5: bad shape"); return -1; }
    const int Tp = (T + 7) / 8 * 8;
    int rc = launch_fa((const bf16 *)q, (const bf16 *)k, (const bf16 *)vt, (bf16 *)o, B, T, Tp, (cudaStream_t)stream);
    return rc ? rc : scp::check_launch("scp_attention_tc5");
}

// q | k split token-major [B*T][1536] (i32 layout), vt planes [2][B*6*64][Tp], o split [B*T][768]
static int launch_fa3(const bf16 *qk, const bf16 *vt, bf16 *o, int B, int T, int Tp, cudaStream_t st)
{
    CUtensorMap tk, tv;
    const uint64_t rows = (uint64_t)B * T, inner = 4 * D;
    if (!scp::gemm::make_tmap_bf16(&tk, qk, inner, rows, inner, scp::fa3::BKV) ||
        !scp::gemm::make_tmap_bf16(&tv, vt, Tp, (uint64_t)2 * B * HEADS * HD, Tp, 64)) {
        scp::set_last_error("tcgen05 attention (x3): cuTensorMapEncodeTiled failed");
        return -1;
    }
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(scp::fa3::fa3_fwd_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, scp::fa3::SMEM_BYTES);
        cudaFuncSetAttribute(scp::fa3::fa3_fwd_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, scp::fa3::SMEM_BYTES);
        cudaFuncSetAttribute(scp::fa3::fa3_fwd_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, scp::fa3::SMEM_BYTES);
        attr_done = true;
    }
    const char *e = getenv("SCP_FA3_WARPS");
    const int fa3_warps = (e && atoi(e) == 8) ? 8 : 4;
    const float scale_log2e = 0.125f * 1.4426950408889634f;
    const dim3 grid((T + scp::fa3::BQ - 1) / scp::fa3::BQ, B * HEADS);
    if (fa3_warps == 8) {
        void *flags = nullptr;
        if (cudaGetSymbolAddress(&flags, scp::fa3::g_retry) != cudaSuccess ||
            cudaMemsetAsync(flags, 0, sizeof(int) * scp::fa3::NFLAGS, st) != cudaSuccess) {
            scp::set_last_error("tcgen05 attention (x3): retry flags");
            return -1;
        }
        // two threads per query row with a fixed reference point, then the robust form on the query tiles it flagged
        scp::fa3::fa3_fwd_kernel<8, false><<<grid, 64 + 32 * 8, scp::fa3::SMEM_BYTES, st>>>(qk, tk, tv, o, T, scale_log2e, B * HEADS * HD);
        scp::fa3::fa3_fwd_kernel<4, true><<<grid, 64 + 32 * 4, scp::fa3::SMEM_BYTES, st>>>(qk, tk, tv, o, T, scale_log2e, B * HEADS * HD);
    } else {
        scp::fa3::fa3_fwd_kernel<4, false><<<grid, 64 + 32 * 4, scp::fa3::SMEM_BYTES, st>>>(qk, tk, tv, o, T, scale_log2e, B * HEADS * HD);
    }
    return 0;
}

extern "C" int scp_attention_x3(const void *qk, const void *vt, void *o, int B, int T, void *stream)
{
    if (B <= 0 || T <= 0) { scp::set_last_error("scp_attention_x3: bad shape"); return -1; }
    const int Tp = (T + 7) / 8 * 8;
    int rc = launch_fa3((const bf16 *)qk, (const bf16 *)vt, (bf16 *)o, B, T, Tp, (cudaStream_t)stream);
    return rc ? rc : scp::check_launch("scp_attention_x3");
}

extern "C" int scp_gemm_bf16x3_tn(const void *A, const void *W, const float *bias, float *C, int M, int N, int K,
                                  void *stream)
{
    EpiPlain epi{ C, bias, N };
    int rc = scp::gemm::launch<EpiPlain, 3>(A, 2 * K, W, 2 * K, M, N, K, epi, (cudaStream_t)stream);
    return rc ? rc : scp::check_launch("scp_gemm_bf16x3_tn");
}

// bf16 element count multiplier of the activation buffers: split pairs in x3 mode
static inline size_t prec_mul(int precision) { return precision == SCP_VIT_X3 ? 2 : 1; }

extern "C" size_t scp_vit_workspace_bytes(int B, int H, int W, int precision)
{
    if (B <= 0 || H <= 0 || W <= 0 || H % PATCH || W % PATCH) return 0;
    const size_t np = (size_t)(H / PATCH) * (W / PATCH), T = np + 1, M = (size_t)B * T, pm = prec_mul(precision);
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t Tp = (T + 7) / 8 * 8;
    return al(M * D * 4) + al(M * D * 2 * pm) * 4 + al((size_t)B * D * Tp * 2 * pm) + al(M * MLP * 2 * pm) +
           al((size_t)B * np * KP * 2 * pm);
}

template <int NT>
static int vit_keys_impl(const scp_vit_weights *w, const float *img, float *feat, void *feat_tokens, int B, int H, int W,
                         int n_blocks, void *workspace, cudaStream_t st)
{
    constexpr size_t pm = NT == 3 ? 2 : 1;
    const int np = (H / PATCH) * (W / PATCH), T = np + 1;
    const long M = (long)B * T;
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    char *p = (char *)workspace;
    float *x = (float *)p; p += al(M * D * 4);
    bf16 *y = (bf16 *)p; p += al(M * D * 2 * pm);
    bf16 *qb = (bf16 *)p; p += al(M * D * 2 * pm);          // q | k: one [M][768 * pm] matrix spanning qb and kb
    bf16 *kb = (bf16 *)p; p += al(M * D * 2 * pm);
    const int Tp = (T + 7) / 8 * 8;
    bf16 *vb = (bf16 *)p; p += al((size_t)B * D * Tp * 2 * pm);   // V transposed per head: [B][6][64][Tp] (x3: hi plane, lo plane)
    cudaMemsetAsync(vb, 0, (size_t)B * D * Tp * 2 * pm, st);     // pad columns t in [T, Tp) must be zero
    bf16 *ob = (bf16 *)p; p += al(M * D * 2 * pm);
    bf16 *hb = (bf16 *)p; p += al(M * MLP * 2 * pm);
    bf16 *a0 = (bf16 *)p;
    constexpr int PD = (int)pm * D, PMLP = (int)pm * MLP, PKP = (int)pm * KP;   // physical row pitches

    int rc;
    // tokens
    {
        const long n = (long)B * np * 3 * PATCH;
        if (NT == 3) im2col_x3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(img, a0, B, H, W);
        else im2col_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(img, a0, B, H, W);
        cls_kernel<<<(B * D + 255) / 256, 256, 0, st>>>(x, w->cls_pos0, B, T);
        EpiPatch epi{ x, w->patch_b, w->pos, np, T };
        if ((rc = scp::gemm::launch<EpiPatch, NT>(a0, PKP, w->patch_w, PKP, B * np, D, KP, epi, st))) return rc;
    }
    const bool use_tc5_attention = attention_variant() != 0;
    const unsigned ln_grid = (unsigned)((M + 7) / 8);
    const float scale_log2e = 0.125f * 1.4426950408889634f;
    auto layernorm = [&](const float *lw, const float *lb) {
        if (NT == 3) layernorm_x3_kernel<<<ln_grid, 256, 0, st>>>(x, lw, lb, y, M);
        else layernorm_kernel<<<ln_grid, 256, 0, st>>>(x, lw, lb, y, M);
    };
    for (int i = 0; i < n_blocks; i++) {
        const scp_vit_block &bw = w->blocks[i];
        layernorm(bw.ln1_w, bw.ln1_b);
        if constexpr (NT == 3) {          // split q | k token-major, split V^T planes, fa3
            EpiQKVTokens3 eq{ vb, bw.qkv_b, T, Tp, (long)B * D * Tp };
            if ((rc = scp::gemm::launch<EpiQKVTokens3, 3>(y, PD, bw.qkv_w, PD, (int)M, 3 * D, D, eq, st, qb, 4 * D))) return rc;
            if ((rc = launch_fa3(qb, vb, ob, B, T, Tp, st))) return rc;
        } else if (attention_variant() == 2) {   // fa2: q | k token-major in one matrix (qb and kb are adjacent), V^T per head
            EpiQKVTokens eq{ vb, bw.qkv_b, T, Tp };
            if ((rc = scp::gemm::launch(y, D, bw.qkv_w, D, (int)M, 3 * D, D, eq, st, qb, 2 * D))) return rc;
            if ((rc = launch_fa(qb, qb, vb, ob, B, T, Tp, st, true))) return rc;
        } else if (use_tc5_attention) {   // first tcgen05 version: head-major q, k; V transposed per head
            EpiQKV eq{ qb, kb, vb, bw.qkv_b, T, Tp };
            if ((rc = scp::gemm::launch(y, D, bw.qkv_w, D, (int)M, 3 * D, D, eq, st))) return rc;
            if ((rc = launch_fa(qb, kb, vb, ob, B, T, Tp, st))) return rc;
        } else {                   // mma.sync variant (SCP_VIT_ATTENTION=mma)
            EpiQKVPlain eq{ qb, kb, vb, bw.qkv_b, T };
            if ((rc = scp::gemm::launch(y, D, bw.qkv_w, D, (int)M, 3 * D, D, eq, st))) return rc;
            attention_kernel<<<dim3((T + AQ - 1) / AQ, B * HEADS), 128, 0, st>>>(qb, kb, vb, ob, T, scale_log2e);
        }
        EpiResidual ep{ bw.proj_b };
        if ((rc = scp::gemm::launch<EpiResidual, NT>(ob, PD, bw.proj_w, PD, (int)M, D, D, ep, st, x, D))) return rc;
        layernorm(bw.ln2_w, bw.ln2_b);
        EpiGelu eg{ bw.fc1_b };
        if ((rc = scp::gemm::launch<EpiGelu, NT>(y, PD, bw.fc1_w, PD, (int)M, MLP, D, eg, st, hb, PMLP))) return rc;
        EpiResidual e2{ bw.fc2_b };
        if ((rc = scp::gemm::launch<EpiResidual, NT>(hb, PMLP, bw.fc2_w, PMLP, (int)M, D, MLP, e2, st, x, D))) return rc;
    }
    // key projection of block `n_blocks` (rows D..2D-1 of its qkv weight) on norm1(x)
    {
        const scp_vit_block &bw = w->blocks[n_blocks];
        layernorm(bw.ln1_w, bw.ln1_b);
        EpiKeys ek{ feat, bw.qkv_b + D, T, (bf16 *)feat_tokens, NT == 3 };
        if ((rc = scp::gemm::launch<EpiKeys, NT>(y, PD, (const bf16 *)bw.qkv_w + (size_t)D * PD, PD, (int)M, D, D, ek, st))) return rc;
    }
    return scp::check_launch("scp_vit_s8_keys");
}

extern "C" int scp_vit_s8_keys(const scp_vit_weights *w, const float *img, float *feat, void *feat_tokens, int B, int H,
                               int W, int n_blocks, int precision, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!w || B <= 0 || H % PATCH || W % PATCH || n_blocks < 0 || n_blocks >= SCP_VIT_MAX_BLOCKS ||
        (precision != SCP_VIT_BF16 && precision != SCP_VIT_X3)) {
        scp::set_last_error("scp_vit_s8_keys: bad arguments (B=%d H=%d W=%d n_blocks=%d precision=%d)", B, H, W, n_blocks,
                            precision);
        return -1;
    }
    if (!workspace || workspace_bytes < scp_vit_workspace_bytes(B, H, W, precision)) {
        scp::set_last_error("scp_vit_s8_keys: workspace too small");
        return -1;
    }
    return precision == SCP_VIT_X3
               ? vit_keys_impl<3>(w, img, feat, feat_tokens, B, H, W, n_blocks, workspace, (cudaStream_t)stream)
               : vit_keys_impl<1>(w, img, feat, feat_tokens, B, H, W, n_blocks, workspace, (cudaStream_t)stream);
}

// fw/bw arg-max matching of DINO features without the (pairs, np, np) similarity tensor: for pair p and every pixel r
// of image a_idx[p]: best[p][r] = arg-max over the pixels c of image w_idx[p] with w_mask[p][c] > 0 of
// <tokens[a_idx[p]][r], tokens[w_idx[p]][c]> (fp32 accumulation on the tcgen05 GEMM; operands bf16, or split bf16 pairs =
// fp32-class products with precision = SCP_VIT_X3, where tokens is [B][np][768] in the i32 layout).
extern "C" int scp_dino_argmatch(const void *tokens, const long long *a_idx, const long long *w_idx, const float *w_mask,
                                 int B, int np, int NP, int precision, unsigned long long *best, void *stream)
{
    if (!tokens || !a_idx || !w_idx || !w_mask || !best || B <= 0 || NP <= 0 || np <= 0 || np % scp::gemm::BM != 0 ||
        (precision != SCP_VIT_BF16 && precision != SCP_VIT_X3)) {
        scp::set_last_error("scp_dino_argmatch: bad arguments (B=%d np=%d NP=%d precision=%d; np must be a multiple of %d)",
                            B, np, NP, precision, scp::gemm::BM);
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(best, 0, (size_t)NP * np * sizeof(unsigned long long), st);
    EpiArgmax epi{ best, w_mask, np, np };
    int rc;
    if (precision == SCP_VIT_X3)
        rc = scp::gemm::launch<EpiArgmax, 3>(tokens, 2 * D, tokens, 2 * D, NP * np, np, D, epi, st, nullptr, 0, a_idx, w_idx, np,
                                             (long)B * np, (long)B * np);
    else
        rc = scp::gemm::launch(tokens, D, tokens, D, NP * np, np, D, epi, st, nullptr, 0, a_idx, w_idx, np, (long)B * np,
                               (long)B * np);
    return rc ? rc : scp::check_launch("scp_dino_argmatch");
}
