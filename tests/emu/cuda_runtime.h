// TEST INFRASTRUCTURE ONLY -- never part of the product.  A stand-in for <cuda_runtime.h> that lets g++ compile the
// kernels' CUDA C++ SOURCE for the host, so their index / mask / ownership / ordering logic can be checked on a machine
// without a GPU (tests/test_emu_*.py).  Execution model: every thread of a CTA is a fiber (ucontext) on one host
// thread; fibers run until they reach a barrier -- every warp shuffle, __syncwarp and __syncthreads is one -- which is
// all the lockstep these kernels rely on.  CTAs are spread over a few host threads.  cp.async copies land as late as
// PTX allows (at the wait_group that covers them), so a missing wait shows up as wrong data.  Arithmetic is IEEE with
// contraction off (-ffp-contract=off) and fmaf for the explicit fused forms: the rounding sequence of the GPU build.
#pragma once
#include <ucontext.h>

#include "cuda.h"          // the emulated driver-API types (tests/emu/cuda.h)

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <memory>
#include <thread>
#include <vector>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __grid_constant__
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __restrict__
#define __align__(n) alignas(n)

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
inline float2 make_float2(float a, float b) { return float2{a, b}; }
inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
inline double2 make_double2(double a, double b) { return double2{a, b}; }

// ---- the slice of the runtime API the host side of the library touches (single device, everything in order)
typedef struct EmuStream_ *cudaStream_t;
typedef struct EmuEvent_ *cudaEvent_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
constexpr unsigned cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2;
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
struct cudaFuncAttributes { int numRegs; };
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -1; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
template <typename F> cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename F> cudaError_t cudaFuncGetAttributes(cudaFuncAttributes *a, F) { a->numRegs = 0; return cudaSuccess; }
template <typename T> cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)std::malloc(n); return cudaSuccess; }
inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemset(void *p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }

// driver entry points: the emulator hands out its own cuTensorMapEncodeTiled (tests/emu/cuda.h)
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
constexpr unsigned long long cudaEnableDefault = 0;
cudaError_t cudaGetDriverEntryPoint(const char *symbol, void **fn, unsigned long long flags, cudaDriverEntryPointQueryResult *status);

using std::max;
using std::min;
using std::signbit;

namespace emu {

struct Barrier {
    int expected = 0, arrived = 0;
    unsigned gen = 0;
};
struct Fiber {
    ucontext_t ctx;
    char *stack = nullptr;
    bool done = false;
    uint3 tid;
    int lane = 0;
    Barrier *warp = nullptr;
    uint64_t (*slots)[2] = nullptr;   // the warp's 32 shuffle mailboxes (up to 16 bytes each)
};
struct Block {
    ucontext_t sched;
    std::vector<Fiber> fibers;
    std::vector<Barrier> warps;
    std::vector<uint64_t> mail;       // 32 x 2 words per warp
    Barrier all;
    void (*body)(void *) = nullptr;
    void *arg = nullptr;
    unsigned long long events = 0;    // barrier releases + fiber exits: the deadlock detector watches it move
    alignas(128) unsigned char smem[232448];   // dynamic shared memory of the CTA (227 KB)
    ~Block() { for (auto &f : fibers) std::free(f.stack); }
};
inline thread_local Block *blk = nullptr;
inline thread_local Fiber *cur = nullptr;
inline std::atomic<long long> launches{0};   // kernels launched since load (the tests read it through emu_launches())
constexpr size_t STACK = 1u << 20;

inline void yield() { swapcontext(&cur->ctx, &blk->sched); }
inline void release(Barrier &b) { b.arrived = 0; ++b.gen; ++blk->events; }
inline void wait(Barrier &b) {
    const unsigned g = b.gen;
    if (++b.arrived == b.expected) release(b);
    else while (b.gen == g) yield();
}
inline void drop(Barrier &b) {           // a thread that exits no longer takes part
    --b.expected;
    if (b.expected > 0 && b.arrived == b.expected) release(b);
}
inline void cp_async_thread_exit();
inline void trampoline() {
    blk->body(blk->arg);
    cp_async_thread_exit();
    cur->done = true;
    ++blk->events;
    drop(*cur->warp);
    drop(blk->all);
}                                        // returning resumes uc_link = the scheduler

template <typename T>
T shuffle(T v, int src_lane) {
    static_assert(sizeof(T) <= 16, "shuffle payload");
    std::memcpy(cur->slots[cur->lane], &v, sizeof(T));
    wait(*cur->warp);
    T r = v;
    if (src_lane >= 0 && src_lane < 32) std::memcpy(&r, cur->slots[src_lane], sizeof(T));
    wait(*cur->warp);
    return r;
}

inline thread_local uint3 block_idx;
inline thread_local dim3 block_dim, grid_dim;

inline void run_block(Block &b, dim3 block, void (*body)(void *), void *arg) {
    const unsigned nthreads = block.x * block.y * block.z, nwarps = (nthreads + 31) / 32;
    blk = &b;
    b.body = body;
    b.arg = arg;
    if (b.fibers.size() < nthreads) {
        const size_t old = b.fibers.size();
        b.fibers.resize(nthreads);
        for (size_t t = old; t < nthreads; ++t) b.fibers[t].stack = (char *)std::malloc(STACK);
    }
    b.warps.assign(nwarps, Barrier{});
    b.mail.assign((size_t)nwarps * 64, 0);
    b.all = Barrier{};
    b.all.expected = (int)nthreads;
    for (unsigned t = 0; t < nthreads; ++t) {
        Fiber &f = b.fibers[t];
        f.done = false;
        f.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
        f.lane = (int)(t % 32);
        f.warp = &b.warps[t / 32];
        f.warp->expected++;
        f.slots = reinterpret_cast<uint64_t(*)[2]>(b.mail.data() + (size_t)(t / 32) * 64);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = STACK;
        f.ctx.uc_link = &b.sched;
        makecontext(&f.ctx, trampoline, 0);
    }
    unsigned live = nthreads;
    while (live) {
        const unsigned long long before = b.events;
        live = 0;
        for (unsigned t = 0; t < nthreads; ++t) {
            Fiber &f = b.fibers[t];
            if (f.done) continue;
            cur = &f;
            swapcontext(&b.sched, &f.ctx);
            if (!f.done) ++live;
        }
        if (live && b.events == before) {
            std::fprintf(stderr, "emu: deadlock -- %u threads of a CTA wait at barriers the others never reach\n", live);
            std::abort();
        }
    }
    cur = nullptr;
}

// CTA contexts (shared memory + fiber stacks) are expensive to set up: keep them in a free list across launches
struct BlockLease {
    static std::vector<Block *> &pool() { static std::vector<Block *> p; return p; }
    static std::atomic_flag &lock() { static std::atomic_flag f = ATOMIC_FLAG_INIT; return f; }
    Block *b = nullptr;
    BlockLease() {
        while (lock().test_and_set(std::memory_order_acquire)) { }
        if (!pool().empty()) { b = pool().back(); pool().pop_back(); }
        lock().clear(std::memory_order_release);
        if (b == nullptr) b = new Block;
    }
    ~BlockLease() {
        while (lock().test_and_set(std::memory_order_acquire)) { }
        pool().push_back(b);
        lock().clear(std::memory_order_release);
    }
};

// kernel<<<grid, block, smem, stream>>>(args) is rewritten (tests/emu/build_emu.py) into launch(grid, block, [&]{ kernel(args); })
template <typename F>
void launch(dim3 grid, dim3 block, F body) {
    ++launches;
    const unsigned long long nblocks = (unsigned long long)grid.x * grid.y * grid.z;
    if (nblocks == 0 || block.x * block.y * block.z == 0) return;
    std::atomic<unsigned long long> next{0};
    auto worker = [&] {
        BlockLease lease;                              // fiber stacks are reused by all the CTAs this host thread runs,
        Block *b = lease.b;                            // and by later launches
        for (;;) {
            const unsigned long long id = next.fetch_add(1);
            if (id >= nblocks) break;
            block_idx = uint3{(unsigned)(id % grid.x), (unsigned)((id / grid.x) % grid.y), (unsigned)(id / ((unsigned long long)grid.x * grid.y))};
            block_dim = block;
            grid_dim = grid;
            run_block(*b, block, [](void *p) { (*static_cast<F *>(p))(); }, &body);
        }
    };
    const unsigned hw = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    const unsigned nthr = (unsigned)std::min<unsigned long long>(hw, nblocks);
    if (nthr <= 1) { worker(); return; }
    std::vector<std::thread> pool;
    for (unsigned k = 0; k < nthr; ++k) pool.emplace_back(worker);
    for (auto &t : pool) t.join();
}
}  // namespace emu

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::block_idx)
#define blockDim (emu::block_dim)
#define gridDim (emu::grid_dim)

template <typename T> T __shfl_up_sync(unsigned, T v, int d) { return emu::shuffle(v, emu::cur->lane - d); }
template <typename T> T __shfl_down_sync(unsigned, T v, int d) { return emu::shuffle(v, emu::cur->lane + d); }
template <typename T> T __shfl_sync(unsigned, T v, int src) { return emu::shuffle(v, src); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::wait(*emu::cur->warp); }
inline void __syncthreads() { emu::wait(emu::blk->all); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
template <typename T> T __ldg(const T *p) { return *p; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicExch(unsigned long long *p, unsigned long long v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
// shared-window address of a pointer into the CTA's dynamic shared memory (what cp.async takes as its destination)
inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)((const unsigned char *)p - emu::blk->smem); }

namespace emu {
// what the inline PTX of the kernels does, statement by statement (tests/emu/build_emu.py maps each asm to one of these)
// cp.async is emulated as LATE as the PTX model allows: a copy lands only when a cp.async.wait_group covering its group
// executes (until then the destination keeps its old bytes), so a kernel that reads a ring slot before waiting for it
// fails the parity tests here too.  Groups are per thread, as in hardware.
struct PendingCopy { unsigned dst; const void *src; int bytes, src_bytes; };
struct AsyncState {
    static constexpr int CAP = 32;                    // committed groups a thread may have in flight
    std::vector<PendingCopy> ring[CAP];               // committed groups (circular, oldest at head); capacity is kept
    int head = 0, count = 0;
    std::vector<PendingCopy> open;                    // copies issued since the last commit
};
inline AsyncState &async_state() {
    static thread_local std::vector<AsyncState> per_fiber;            // indexed by the fiber's slot in its CTA
    const size_t id = (size_t)(cur - blk->fibers.data());
    if (per_fiber.size() <= id) per_fiber.resize(id + 1);
    return per_fiber[id];
}
inline void land(const PendingCopy &c) {
    unsigned char *d = blk->smem + c.dst;
    std::memset(d, 0, c.bytes);
    if (c.src_bytes > 0) std::memcpy(d, c.src, std::min(c.bytes, c.src_bytes));
}
inline void cp_async(unsigned dst, const void *src, int bytes, int src_bytes) {
    async_state().open.push_back(PendingCopy{dst, src, bytes, src_bytes});
}
inline void cp_async_wait(int newest_allowed_pending) {
    AsyncState &a = async_state();
    while (a.count > newest_allowed_pending) {
        for (const PendingCopy &c : a.ring[a.head]) land(c);
        a.ring[a.head].clear();
        a.head = (a.head + 1) % AsyncState::CAP;
        --a.count;
    }
}
inline void cp_async_commit() {
    AsyncState &a = async_state();
    if (a.count == AsyncState::CAP) cp_async_wait(AsyncState::CAP - 1);     // (hardware has its own limit; never reached here)
    a.ring[(a.head + a.count) % AsyncState::CAP].swap(a.open);
    ++a.count;
    a.open.clear();
}
inline void cp_async_thread_exit() {                  // a thread that ends leaves nothing behind for the fiber slot's next user
    AsyncState &a = async_state();
    for (auto &g : a.ring) g.clear();
    a.head = a.count = 0;
    a.open.clear();
}
inline unsigned long long ld_acquire(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline unsigned long long global_timer_ns() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}
inline void st_release(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
}  // namespace emu

namespace emu {
// ---- mbarrier and TMA (fd2d_chain.cu).  A barrier is its 8 bytes of CTA shared memory: {phase bit, expected arrivals,
// pending arrivals} and the transaction bytes still expected.  A phase completes when no arrival and no byte is
// pending.  try_wait spins by yielding to the CTA's other fibers: the producer warp runs on the same host thread.
// TMA copies land at issue (the data is there, the barrier says so later): what the emulation checks is the protocol --
// slot indices, parities, who waits for whom (a wait that can never be satisfied trips the deadlock detector) -- and
// the arithmetic, not the hardware's asynchrony.
struct MBar { uint32_t phase : 1, count : 15, pending : 16; int32_t tx; };
static_assert(sizeof(MBar) == 8, "an mbarrier is 8 bytes");
inline MBar &mbar_at(unsigned bar) { return *reinterpret_cast<MBar *>(blk->smem + bar); }
inline void mbar_settle(MBar &b) {
    if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.count; ++blk->events; }
}
inline void mbar_init(unsigned bar, unsigned count) { MBar &b = mbar_at(bar); b.phase = 0; b.count = count; b.pending = count; b.tx = 0; }
inline void mbar_arrive(unsigned bar, int leader) {
    if (!leader) return;
    MBar &b = mbar_at(bar);
    if (b.pending == 0) { std::fprintf(stderr, "emu: mbarrier over-arrived\n"); std::abort(); }
    --b.pending;
    mbar_settle(b);
}
inline void mbar_wait(unsigned bar, unsigned parity) {
    while (mbar_at(bar).phase == parity) yield();     // the phase with this parity has not completed yet
}
inline void st_shared_v2(unsigned addr, float x, float y) { float *d = reinterpret_cast<float *>(blk->smem + addr); d[0] = x; d[1] = y; }
// one box of one array: rows [row, row + box rows) x columns [col, col + box columns), zeros outside the tensor
inline void tma_box(unsigned dst, const CUtensorMap *m, int col, int row) {
    unsigned char *d = blk->smem + dst;
    const size_t row_bytes = (size_t)m->box[0] * m->elem;
    for (uint32_t r = 0; r < m->box[1]; ++r, d += row_bytes) {
        std::memset(d, 0, row_bytes);
        const long long rr = (long long)row + r;
        if (rr < 0 || rr >= (long long)m->dim[1]) continue;
        const long long c0 = std::max<long long>(col, 0), c1 = std::min<long long>((long long)col + m->box[0], (long long)m->dim[0]);
        if (c1 > c0) std::memcpy(d + (c0 - col) * m->elem, m->base + rr * m->row_stride + c0 * m->elem, (size_t)(c1 - c0) * m->elem);
    }
}
// the elected lane arms the barrier with the bytes of six boxes and issues them
inline void tma_issue6(unsigned dst0, unsigned bar, const CUtensorMap *m0, const CUtensorMap *m1, const CUtensorMap *m2,
                       const CUtensorMap *m3, const CUtensorMap *m4, const CUtensorMap *m5, int col, int bytes, int row,
                       unsigned dst1, unsigned dst2, unsigned dst3, unsigned dst4, unsigned dst5) {
    if (cur->lane != 0) return;                       // elect.sync
    MBar &b = mbar_at(bar);
    b.tx += bytes;
    if (b.pending == 0) { std::fprintf(stderr, "emu: mbarrier over-arrived (expect_tx)\n"); std::abort(); }
    --b.pending;
    const CUtensorMap *m[6] = {m0, m1, m2, m3, m4, m5};
    const unsigned dst[6] = {dst0, dst1, dst2, dst3, dst4, dst5};
    for (int a = 0; a < 6; ++a) {
        tma_box(dst[a], m[a], col, row);
        b.tx -= (int32_t)(m[a]->box[0] * m[a]->box[1] * m[a]->elem);
    }
    mbar_settle(b);
}
inline CUresult encode_tiled(CUtensorMap *m, CUtensorMapDataType, cuuint32_t rank, void *base, const cuuint64_t *gdim,
                             const cuuint64_t *gstride, const cuuint32_t *box, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
    if (rank != 2 || (reinterpret_cast<uintptr_t>(base) & 15u) != 0 || gstride[0] % 16 != 0 || box[0] > 256 || box[1] > 256) return 1;
    std::memset(m, 0, sizeof(*m));
    m->base = static_cast<const unsigned char *>(base);
    m->dim[0] = gdim[0]; m->dim[1] = gdim[1];
    m->row_stride = gstride[0];
    m->box[0] = box[0]; m->box[1] = box[1];
    m->elem = 4;
    return CUDA_SUCCESS;
}
}  // namespace emu

inline cudaError_t cudaGetDriverEntryPoint(const char *symbol, void **fn, unsigned long long, cudaDriverEntryPointQueryResult *status) {
    const bool known = std::strcmp(symbol, "cuTensorMapEncodeTiled") == 0;
    *fn = known ? reinterpret_cast<void *>(&emu::encode_tiled) : nullptr;
    if (status) *status = known ? cudaDriverEntryPointSuccess : cudaDriverEntryPointSymbolNotFound;
    return cudaSuccess;
}
