// Runtime of the host emulation (see cuda_runtime.h in this directory): CTA-at-a-time launcher.
#include "cuda_runtime.h"

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
namespace scp_emu {
thread_local Cta *cta = nullptr;
std::mutex atomic_mutex;

void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body)
{
    const unsigned nt = block.x * block.y * block.z, nw = (nt + 31) / 32;
    if (block.y != 1 || block.z != 1 || nt % 32 != 0) {
        fprintf(stderr, "scp_emu: only 1-D blocks of whole warps are emulated\n");
        abort();
    }
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                Cta c;
                c.nthreads = nt;
                pthread_barrier_init(&c.block_bar, nullptr, nt);
                c.warp_bar.resize(nw);
                for (auto &b : c.warp_bar) pthread_barrier_init(&b, nullptr, 32);
                c.warp_slot.assign(nw * 32, 0);
                c.dyn_smem.assign(dyn_smem + 16, 0);
                c.warp_frag.assign(nw * 256, 0);
                std::vector<std::thread> threads;
                threads.reserve(nt);
                for (unsigned t = 0; t < nt; t++)
                    threads.emplace_back([&, t]() {
                        cta = &c;
                        threadIdx = uint3{ t, 0, 0 };
                        blockIdx = uint3{ bx, by, bz };
                        blockDim = block;
                        gridDim = grid;
                        body();
                    });
                for (auto &th : threads) th.join();
                pthread_barrier_destroy(&c.block_bar);
                for (auto &b : c.warp_bar) pthread_barrier_destroy(&b);
            }
}
}  // namespace scp_emu
