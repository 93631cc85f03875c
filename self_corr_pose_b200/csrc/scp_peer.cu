// Small-vector all-gather / all-reduce across the GPUs of one node over NVLink peer memory -- the statistics exchange of
// SyncBatchNorm (the reference converts every BatchNorm of the encoder, model/trainer.py:66: 20 layers x 2 encoder passes
// x (all_gather forward + all_reduce backward) = 80 collectives of <= 4 KB per training step).  NCCL runs each of them as a
// separate ring / tree launch (~25-60 us at 8 ranks); here one kernel per collective STORES the rank's vector straight into
// every peer's buffer (cudaIpc-mapped device memory, NVLink 5 / NVSwitch), raises a flag there, waits for the peers' flags in
// its own buffer and consumes the gathered vectors: one NVLink store round trip, no proxy, no ring.
//
// Protocol.  Every rank owns a buffer [flags: SLOTS x MAXW u32][data: SLOTS x MAXW x MAXN f32] that all peers map.  Collective
// number `seq` (a per-channel counter kept in device memory, so that CUDA-graph replays advance it) uses slot seq % SLOTS:
//   1. rank r writes its n floats into data[slot][r] of EVERY peer (and of itself);
//   2. __threadfence_system(), then flags[slot][r] := seq in every peer (st.release.sys);
//   3. wait until flags[slot][q] == seq for all q in the own buffer (ld.acquire.sys, bounded spin -> trap instead of a hang);
//   4. read data[slot][0..world) from the own buffer: gather, or the sum over ranks in rank order (identical on all ranks).
// A slot is reused SLOTS collectives later; a rank can only be one collective ahead of the slowest peer (step 3), so with
// SLOTS >= 2 nobody overwrites data that a peer still has to read.  All ranks must issue the same sequence of collectives
// on a channel, in stream order (two channels -- one per encoder pass -- keep concurrent streams apart).
#include <stdint.h>
#include <string.h>

#include "../../include/scp_b200.h"
#include "scp_common.cuh"

namespace scp {
namespace peer {

constexpr int SLOTS = SCP_PEER_SLOTS, MAXW = SCP_PEER_MAX_WORLD, MAXN = SCP_PEER_MAX_FLOATS;
constexpr int FLAG_WORDS = SLOTS * MAXW;                     // u32 flags at the start of the buffer
constexpr size_t DATA_OFF = FLAG_WORDS;                      // in 4-byte words
constexpr size_t BUFFER_WORDS = DATA_OFF + (size_t)SLOTS * MAXW * MAXN;

struct Peers { float *p[MAXW]; };

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256)
exchange_kernel(Peers peers, const float *__restrict__ src, int n, int rank, int world, unsigned *counter, float *__restrict__ dst,
                int reduce)
{
    __shared__ unsigned s_seq;
    const int tid = threadIdx.x;
    if (tid == 0) s_seq = *counter + 1;
    __syncthreads();
    const unsigned seq = s_seq;
    const int slot = seq % SLOTS;
    // 1. my vector -> slot [slot][rank] of every rank
    for (int q = 0; q < world; q++) {
        float *d = peers.p[q] + DATA_OFF + ((size_t)slot * MAXW + rank) * MAXN;
        for (int i = tid; i < n; i += blockDim.x) d[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. raise my flag everywhere, 3. wait for everybody's flag here
    if (tid < world) {
        st_release_sys(reinterpret_cast<unsigned *>(peers.p[tid]) + slot * MAXW + rank, seq);
        const unsigned *mine = reinterpret_cast<const unsigned *>(peers.p[rank]) + slot * MAXW + tid;
        unsigned spins = 0;
        while (ld_acquire_sys(mine) != seq) {
            if (++spins > (1u << 26)) __trap();            // a missing peer traps the kernel (after ~20 s) instead of hanging the GPU
        }
    }
    __syncthreads();
    // 4. consume
    const float *own = peers.p[rank] + DATA_OFF + (size_t)slot * MAXW * MAXN;
    if (reduce) {
        for (int i = tid; i < n; i += blockDim.x) {
            float s = 0.f;
            for (int q = 0; q < world; q++) s += __ldcg(own + (size_t)q * MAXN + i);
            dst[i] = s;
        }
    } else {
        for (int q = 0; q < world; q++)
            for (int i = tid; i < n; i += blockDim.x) dst[(size_t)q * n + i] = __ldcg(own + (size_t)q * MAXN + i);
    }
    if (tid == 0) *counter = seq;
}

}  // namespace peer
}  // namespace scp

using namespace scp::peer;

extern "C" size_t scp_peer_buffer_bytes(void) { return BUFFER_WORDS * 4; }

extern "C" int scp_peer_buffer_create(void **ptr, unsigned char *handle64)
{
    if (!ptr || !handle64) { scp::set_last_error("scp_peer_buffer_create: null argument"); return -1; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    cudaError_t e = cudaMalloc(ptr, BUFFER_WORDS * 4);
    if (e == cudaSuccess) e = cudaMemset(*ptr, 0, BUFFER_WORDS * 4);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle64), *ptr);
    if (e != cudaSuccess) {
        scp::set_last_error("scp_peer_buffer_create: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return (int)e;
    }
    return 0;
}

extern "C" int scp_peer_buffer_open(const unsigned char *handle64, void **ptr)
{
    if (!ptr || !handle64) { scp::set_last_error("scp_peer_buffer_open: null argument"); return -1; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        scp::set_last_error("scp_peer_buffer_open: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return (int)e;
    }
    return 0;
}

extern "C" int scp_peer_buffer_close(void *ptr, int own)
{
    cudaError_t e = own ? cudaFree(ptr) : cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    return 0;
}

extern "C" int scp_peer_exchange(void *const *peers, const float *src, int n, int rank, int world, unsigned *counter, float *dst,
                                 int reduce, void *stream)
{
    if (!peers || !src || !dst || !counter || n <= 0 || n > MAXN || world < 1 || world > MAXW || rank < 0 || rank >= world) {
        scp::set_last_error("scp_peer_exchange: bad arguments (n=%d, at most %d; world=%d, at most %d; rank=%d)", n, MAXN, world,
                            MAXW, rank);
        return -1;
    }
    Peers P;
    for (int q = 0; q < MAXW; q++) P.p[q] = q < world ? static_cast<float *>(peers[q]) : nullptr;
    exchange_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(P, src, n, rank, world, counter, dst, reduce);
    return scp::check_launch("scp_peer_exchange");
}
