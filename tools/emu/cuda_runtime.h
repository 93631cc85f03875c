// Host emulation of the small CUDA subset the SoftRas kernels use -- TEST INFRASTRUCTURE ONLY (tools/emu/build_emu.py).
//
// Every CUDA thread of a CTA is an OS thread; CTAs run one after another.  __syncthreads is a pthread barrier over the
// CTA, warp collectives (__ballot_sync, __shfl_sync, ...) are write / barrier / read / barrier over the warp's 32 threads,
// so they require what the kernels guarantee anyway: all 32 lanes of a warp reach every collective (full masks, collectives
// only under warp-uniform control flow) and threads that return early do so warp- (or CTA-) uniformly before any later
// barrier.  `__shared__` becomes `static` (one CTA at a time).  Fast-math intrinsics map to the exact libm functions, so
// results differ from the GPU's in the last bits only.  Purpose: functional validation of kernel variants (traversal,
// indexing, reductions) against the C oracle when no GPU is at hand -- not timing, not bit parity.
#pragma once
#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __shared__ static
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) alignas(n)

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{ x, y, z, w }; }
static inline float2 make_float2(float x, float y) { return float2{ x, y }; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{ x, y, z, w }; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{ x, y, z, w }; }

typedef void *cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }

namespace scp_emu {
struct Cta {
    unsigned nthreads = 0;
    pthread_barrier_t block_bar;
    std::vector<pthread_barrier_t> warp_bar;
    std::vector<uint64_t> warp_slot;      // 32 slots per warp
    std::vector<unsigned char> dyn_smem;  // `extern __shared__` storage
    std::vector<uint32_t> warp_frag;      // 32 x 8 words per warp: operand fragments of an emulated mma.sync
    int block_or = 0;
};
extern thread_local Cta *cta;
extern std::mutex atomic_mutex;
void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body);
}  // namespace scp_emu

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

// ---- block / warp synchronisation ----
static inline void __syncthreads() { pthread_barrier_wait(&scp_emu::cta->block_bar); }
static inline int __syncthreads_or(int pred)
{
    scp_emu::Cta *c = scp_emu::cta;
    pthread_barrier_wait(&c->block_bar);
    if (threadIdx.x == 0) c->block_or = 0;
    pthread_barrier_wait(&c->block_bar);
    if (pred) { std::lock_guard<std::mutex> g(scp_emu::atomic_mutex); c->block_or = 1; }
    pthread_barrier_wait(&c->block_bar);
    return c->block_or;
}
static inline void scp_emu_warp_wait() { pthread_barrier_wait(&scp_emu::cta->warp_bar[threadIdx.x >> 5]); }
static inline void __syncwarp(unsigned = 0xffffffffu) { scp_emu_warp_wait(); }
static inline uint64_t scp_emu_exchange(uint64_t mine, int src_lane)
{
    uint64_t *slot = scp_emu::cta->warp_slot.data() + (threadIdx.x >> 5) * 32;
    slot[threadIdx.x & 31] = mine;
    scp_emu_warp_wait();
    const uint64_t got = slot[src_lane & 31];
    scp_emu_warp_wait();
    return got;
}
static inline unsigned __ballot_sync(unsigned, int pred)
{
    uint64_t *slot = scp_emu::cta->warp_slot.data() + (threadIdx.x >> 5) * 32;
    slot[threadIdx.x & 31] = pred ? 1u : 0u;
    scp_emu_warp_wait();
    unsigned m = 0;
    for (int i = 0; i < 32; i++) m |= (unsigned)(slot[i] & 1u) << i;
    scp_emu_warp_wait();
    return m;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
static inline uint32_t scp_emu_bits(float v) { uint32_t u; memcpy(&u, &v, 4); return u; }
static inline float scp_emu_float(uint32_t u) { float v; memcpy(&v, &u, 4); return v; }
static inline int __shfl_sync(unsigned, int v, int src) { return (int)scp_emu_exchange((uint32_t)v, src); }
static inline unsigned __shfl_sync(unsigned, unsigned v, int src) { return (unsigned)scp_emu_exchange(v, src); }
static inline float __shfl_sync(unsigned, float v, int src) { return scp_emu_float((uint32_t)scp_emu_exchange(scp_emu_bits(v), src)); }
static inline int __shfl_xor_sync(unsigned, int v, int m) { return (int)scp_emu_exchange((uint32_t)v, (threadIdx.x & 31) ^ m); }
static inline float __shfl_xor_sync(unsigned, float v, int m)
{
    return scp_emu_float((uint32_t)scp_emu_exchange(scp_emu_bits(v), (threadIdx.x & 31) ^ m));
}
static inline double __shfl_xor_sync(unsigned, double v, int m)
{
    uint64_t u;
    memcpy(&u, &v, 8);
    u = scp_emu_exchange(u, (threadIdx.x & 31) ^ m);
    memcpy(&v, &u, 8);
    return v;
}

static inline uint32_t (*scp_emu_warp_fragments())[8]
{
    return reinterpret_cast<uint32_t (*)[8]>(scp_emu::cta->warp_frag.data() + (threadIdx.x >> 5) * 256);
}

// ---- memory / arithmetic intrinsics ----
template <typename T> static inline T __ldg(const T *p) { return *p; }
#define __expf(x) expf(x)      // glibc declares a function of that name
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __saturatef(float x) { return x != x ? 0.f : fminf(fmaxf(x, 0.f), 1.f); }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { return scp_emu_bits(f); }
static inline float __uint_as_float(unsigned u) { return scp_emu_float(u); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline float atomicAdd(float *p, float v)
{
    std::lock_guard<std::mutex> g(scp_emu::atomic_mutex);
    const float old = *p;
    *p = old + v;
    return old;
}
static inline double atomicAdd(double *p, double v)
{
    std::lock_guard<std::mutex> g(scp_emu::atomic_mutex);
    const double old = *p;
    *p = old + v;
    return old;
}
static inline int atomicMin(int *p, int v)
{
    std::lock_guard<std::mutex> g(scp_emu::atomic_mutex);
    const int old = *p;
    if (v < old) *p = v;
    return old;
}
static inline int atomicMax(int *p, int v)
{
    std::lock_guard<std::mutex> g(scp_emu::atomic_mutex);
    const int old = *p;
    if (v > old) *p = v;
    return old;
}
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
#define SCP_EMU_LAUNCH(grid, block, smem, ...) scp_emu::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { __VA_ARGS__; })
