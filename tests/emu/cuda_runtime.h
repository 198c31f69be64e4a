// TEST INFRASTRUCTURE ONLY -- never part of the product.  A stand-in for <cuda_runtime.h> that lets g++ compile the
// kernels' CUDA C++ SOURCE for the host, so their index / mask / ownership logic can be checked on a machine without
// a GPU (tests/test_emu_*.py).  A launch runs every CTA's threads as real host threads; the 32 lanes of a warp meet at
// a barrier inside every warp shuffle, which is all the lockstep those kernels rely on.  Arithmetic is IEEE with
// contraction off (-ffp-contract=off) and fmaf for the explicit fused forms, i.e. the rounding sequence of the GPU build.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __grid_constant__
#define __launch_bounds__(...)
#define __restrict__

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
inline float2 make_float2(float a, float b) { return float2{a, b}; }
inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }

typedef void *cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }

using std::max;
using std::min;

namespace emu {
struct Warp {
    std::barrier<> bar{32};
    uint64_t slot[32];
};
inline thread_local Warp *warp = nullptr;
inline thread_local int lane = 0;
inline long long launches = 0;     // kernels launched since load (the tests read it through emu_launches())

template <typename T>
T shuffle(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    std::memcpy(&warp->slot[lane], &v, sizeof(T));
    warp->bar.arrive_and_wait();
    T r = v;
    if (src_lane >= 0 && src_lane < 32) std::memcpy(&r, &warp->slot[src_lane], sizeof(T));
    warp->bar.arrive_and_wait();
    return r;
}
}  // namespace emu

inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

template <typename T> T __shfl_up_sync(unsigned, T v, int d) { return emu::shuffle(v, emu::lane - d); }
template <typename T> T __shfl_down_sync(unsigned, T v, int d) { return emu::shuffle(v, emu::lane + d); }
template <typename T> T __ldg(const T *p) { return *p; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }

namespace emu {
// kernel<<<grid, block, smem, stream>>>(args) is rewritten (tests/emu/build_emu.py) into launch(grid, block, [&]{ kernel(args); })
template <typename F>
void launch(dim3 grid, dim3 block, F body) {
    ++launches;
    const unsigned nthreads = block.x * block.y * block.z;
    const unsigned nwarps = (nthreads + 31) / 32;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                std::vector<std::unique_ptr<Warp>> warps;
                for (unsigned w = 0; w < nwarps; ++w) warps.emplace_back(new Warp);
                std::vector<std::thread> threads;
                threads.reserve(nthreads);
                for (unsigned t = 0; t < nthreads; ++t)
                    threads.emplace_back([&, t] {
                        threadIdx = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                        blockIdx = uint3{bx, by, bz};
                        blockDim = block;
                        gridDim = grid;
                        warp = warps[t / 32].get();
                        lane = (int)(t % 32);
                        body();
                    });
                for (auto &th : threads) th.join();
            }
}
}  // namespace emu
