// Minimal CPU thread emulator for the CUDA kernels' LOGIC tests (test
// infrastructure only -- never loaded by the product).  Every CUDA thread of a
// CTA becomes a host thread; CTAs run one after another; __syncthreads and
// __syncwarp are std::barrier.  It exists because the build container has no
// GPU: index arithmetic, synchronisation placement and operand order of the
// kernels can be checked against the oracle before GPU time is spent.
#pragma once
#define OSC2_EMU 1

#include <atomic>
#include <barrier>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __shared__ static

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
inline thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
inline thread_local double *emu_dyn_smem = nullptr;
inline thread_local std::barrier<> *emu_block_barrier = nullptr;
inline thread_local std::barrier<> *emu_warp_barrier = nullptr;
inline thread_local std::vector<std::unique_ptr<std::barrier<>>> *emu_named = nullptr;
inline std::barrier<> *emu_named_barrier(int id) { return (*emu_named)[id].get(); }
inline thread_local double *emu_shfl_scratch = nullptr;   // [32] per warp
// __shfl_sync over the full warp: publish, read the source lane, and keep the slot until all have read
inline double emu_shfl(double v, int src_lane)
{
    emu_shfl_scratch[threadIdx.x % 32] = v;
    emu_warp_barrier->arrive_and_wait();
    const double out = emu_shfl_scratch[src_lane];
    emu_warp_barrier->arrive_and_wait();
    return out;
}

// __any_sync over the full warp
inline bool emu_any(bool pred)
{
    emu_shfl_scratch[threadIdx.x % 32] = pred ? 1.0 : 0.0;
    emu_warp_barrier->arrive_and_wait();
    bool out = false;
    for (int i = 0; i < 32; ++i) out = out || (emu_shfl_scratch[i] != 0.0);
    emu_warp_barrier->arrive_and_wait();
    return out;
}

#define OSC2_DYN_SMEM(name) double *name = emu_dyn_smem
#define OSC2_DYN_SMEM_ALIGNED(name) double *name = emu_dyn_smem

inline void __syncthreads() { emu_block_barrier->arrive_and_wait(); }
inline void __syncwarp() { emu_warp_barrier->arrive_and_wait(); }
using std::fabs;
using std::sqrt;

struct EmuLauncher {
    size_t max_smem() const { return 227 * 1024; }
    int sm_count() const { return 148; }
    long long launches = 0;

    template <class K, class... A>
    void launch(int /*kernel_id*/, K kernel, unsigned grid, unsigned block, size_t smem, A... args)
    {
        ++launches;
        std::vector<double> dyn((smem + 7) / 8 + 1);
        for (unsigned bid = 0; bid < grid; ++bid) {
            std::barrier<> bb((std::ptrdiff_t)block);
            const unsigned nwarps = (block + 31) / 32;
            std::vector<std::unique_ptr<std::barrier<>>> wb;
            for (unsigned w = 0; w < nwarps; ++w) {
                const unsigned lanes = (w + 1) * 32 <= block ? 32 : block - w * 32;
                wb.emplace_back(new std::barrier<>((std::ptrdiff_t)lanes));
            }
            std::vector<double> shfl((size_t)nwarps * 32, 0.0);
            std::vector<std::unique_ptr<std::barrier<>>> named;     // bar.sync / bar.arrive ids 0..15, all CTA-wide
            for (int i = 0; i < 16; ++i) named.emplace_back(new std::barrier<>((std::ptrdiff_t)block));
            // poison shared memory so that reads of unwritten slots show up
            for (auto &v : dyn) v = std::nan("");
            std::vector<std::thread> th;
            th.reserve(block);
            for (unsigned t = 0; t < block; ++t) {
                th.emplace_back([&, t]() {
                    threadIdx.x = t; blockIdx.x = bid; blockDim.x = block; gridDim.x = grid;
                    emu_dyn_smem = dyn.data();
                    emu_block_barrier = &bb;
                    emu_warp_barrier = wb[t / 32].get();
                    emu_named = &named;
                    emu_shfl_scratch = shfl.data() + (size_t)(t / 32) * 32;
                    kernel(args...);
                    // a thread that returns early must not block the others
                    bb.arrive_and_drop();
                    wb[t / 32]->arrive_and_drop();
                });
            }
            for (auto &x : th) x.join();
        }
    }
};
