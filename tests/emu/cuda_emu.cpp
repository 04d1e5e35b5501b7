// Runtime of the CUDA-on-CPU shim (TEST INFRASTRUCTURE ONLY, see cuda_emu.h).
#include "cuda_emu.h"

#include <chrono>
#include <mutex>
double emu_now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

namespace emu {
Block g_block;
std::vector<unsigned char> g_dyn_smem;
thread_local uint3_emu t_threadIdx, t_blockIdx;
dim3 g_blockDim, g_gridDim;

static std::mutex g_launch_mutex;      // one "device": launches from several host threads run one after the other

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body)
{
    const int nt = (int)(block.x * block.y * block.z);
    if (nt <= 0 || grid.x * grid.y * grid.z == 0) return;
    std::lock_guard<std::mutex> guard(g_launch_mutex);
    g_blockDim = block;
    g_gridDim = grid;
    g_dyn_smem.assign(smem + 64, 0);
    g_block.n_threads = nt;
    g_block.bar = std::make_unique<std::barrier<>>(nt);
    const int nw = (nt + 31) / 32;
    g_block.warps.clear();
    g_block.warps.resize(nw);
    for (int w = 0; w < nw; ++w) {
        g_block.warps[w].n = std::min(32, nt - 32 * w);
        g_block.warps[w].bar = std::make_unique<std::barrier<>>(g_block.warps[w].n);
    }
    std::barrier<> end_bar(nt);
    std::vector<std::thread> th;
    th.reserve(nt);
    for (int t = 0; t < nt; ++t) {
        th.emplace_back([&, t]() {
            t_threadIdx.x = t % block.x;
            t_threadIdx.y = (t / block.x) % block.y;
            t_threadIdx.z = t / (block.x * block.y);
            for (unsigned bz = 0; bz < grid.z; ++bz)
                for (unsigned by = 0; by < grid.y; ++by)
                    for (unsigned bx = 0; bx < grid.x; ++bx) {
                        t_blockIdx.x = bx; t_blockIdx.y = by; t_blockIdx.z = bz;
                        body();
                        end_bar.arrive_and_wait();     // blocks run one after the other (static == __shared__)
                    }
        });
    }
    for (auto& x : th) x.join();
}
}  // namespace emu
