// cuda_emu.h — TEST INFRASTRUCTURE ONLY.  A tiny single-threaded SIMT emulator that lets the
// kernel sources in libhuffman_b200/csrc/cuda/ be compiled with g++ and executed on a box
// without a GPU, so that their LOGIC (tie-breaks, bit offsets, ownership of output bytes,
// speculative decode, error codes) can be checked against the oracle in the CPU test lane.
//
// It is not a CPU fallback: the product library (libhuffman_b200.so) is built by nvcc only,
// never contains this header, and fails with HUF_ERROR_FATAL when no B200 is present.  The
// emulated library is built by tests/emu/build_emu.py into tests/emu/_build/ and is only ever
// loaded by tests/test_emu_*.py.
//
// Model: every CUDA thread of one CTA is a ucontext fiber; CTAs run one after another.  A
// fiber runs until it reaches a warp/CTA collective, where it yields until its peers arrive.
// A collective that can never complete (divergent barrier) aborts with a message.
#pragma once

#include <ucontext.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <vector>

#define HUF_EMU 1

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __align__(x) alignas(x)
#define __shared__ static

struct uint4 {
    uint32_t x, y, z, w;
};
struct uint2 {
    uint32_t x, y;
};
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3() = default;
    dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

namespace hufemu {

struct Barrier {
    unsigned count = 0;
    unsigned gen = 0;
};

struct Warp {
    Barrier bar;
    uint64_t xchg[32];
};

constexpr size_t kStackBytes = 128 * 1024;

// Fiber stacks are recycled across launches (a fresh 1024-thread CTA would otherwise touch
// hundreds of megabytes per kernel).
inline std::vector<void *> &stack_pool()
{
    static std::vector<void *> pool;
    return pool;
}

struct Fiber {
    ucontext_t ctx;
    void *stack = nullptr;
    dim3 tidx;
    bool done = false;
    Warp *warp = nullptr;
};

struct Cta {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    Barrier bar;
    unsigned or_acc = 0, or_result = 0;
    ucontext_t sched;
    Fiber *cur = nullptr;
    dim3 bidx, bdim, gdim;
    std::vector<uint8_t> dyn;
    std::function<void()> body;
    uint64_t progress = 0;
};

inline Cta *&cta()
{
    static Cta *c = nullptr;
    return c;
}

inline void yield_fiber()
{
    Cta *c = cta();
    swapcontext(&c->cur->ctx, &c->sched);
}

inline void barrier_wait(Barrier &b, unsigned participants)
{
    Cta *c = cta();
    const unsigned gen = b.gen;
    if (++b.count == participants) {
        b.count = 0;
        b.gen++;
        c->progress++;
        return;
    }
    while (b.gen == gen) yield_fiber();
}

inline void fiber_entry()
{
    Cta *c = cta();
    c->body();
    c->cur->done = true;
    c->progress++;
    swapcontext(&c->cur->ctx, &c->sched);
}

// Run one kernel: grid CTAs of `block` threads with `smem` bytes of dynamic shared memory.
inline void launch(unsigned grid, unsigned block, size_t smem, const std::function<void()> &body)
{
    // one kernel at a time: the emulator's CTA state is global, and the multi-device lanes of the
    // library launch from one host thread per device
    static std::mutex launch_mu;
    std::lock_guard<std::mutex> launch_lock(launch_mu);
    for (unsigned b = 0; b < grid; b++) {
        Cta c;
        cta() = &c;
        c.bidx = dim3(b);
        c.bdim = dim3(block);
        c.gdim = dim3(grid);
        c.dyn.assign(smem + 64, 0);
        c.body = body;
        c.fibers.resize(block);
        c.warps.resize((block + 31) / 32);
        for (unsigned t = 0; t < block; t++) {
            Fiber &f = c.fibers[t];
            if (stack_pool().empty()) {
                f.stack = malloc(kStackBytes);
            } else {
                f.stack = stack_pool().back();
                stack_pool().pop_back();
            }
            f.tidx = dim3(t);
            f.warp = &c.warps[t / 32];
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = kStackBytes;
            f.ctx.uc_link = &c.sched;
            makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        }
        unsigned live = block;
        while (live) {
            const uint64_t before = c.progress;
            live = 0;
            for (unsigned t = 0; t < block; t++) {
                Fiber &f = c.fibers[t];
                if (f.done) continue;
                c.cur = &f;
                swapcontext(&c.sched, &f.ctx);
                if (!f.done) live++;
            }
            if (live && c.progress == before) {
                fprintf(stderr, "cuda_emu: deadlock in CTA %u (%u threads stuck at a barrier)\n", b, live);
                abort();
            }
        }
        for (Fiber &f : c.fibers) stack_pool().push_back(f.stack);
        cta() = nullptr;
    }
}

inline uint8_t *dyn_smem()
{
    uint8_t *p = cta()->dyn.data();
    return reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));
}

struct IdxProxy {
    int which;
    struct Dim {
        const IdxProxy *p;
        int axis;
        operator unsigned() const
        {
            Cta *c = cta();
            const dim3 &d = p->which == 0 ? c->cur->tidx : p->which == 1 ? c->bidx : p->which == 2 ? c->bdim : c->gdim;
            return axis == 0 ? d.x : axis == 1 ? d.y : d.z;
        }
    };
    Dim x{this, 0}, y{this, 1}, z{this, 2};
};

template <typename T>
inline uint64_t to_bits(T v)
{
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T>
inline T from_bits(uint64_t b)
{
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}

inline unsigned lane() { return cta()->cur->tidx.x & 31; }

template <typename T, typename F>
inline T warp_exchange(T v, F pick)
{
    Warp *w = cta()->cur->warp;
    const unsigned n = std::min(32u, cta()->bdim.x - (cta()->cur->tidx.x & ~31u));
    w->xchg[lane()] = to_bits(v);
    barrier_wait(w->bar, n);
    const T r = pick(w->xchg);
    barrier_wait(w->bar, n);
    return r;
}

}  // namespace hufemu

static const hufemu::IdxProxy threadIdx{0}, blockIdx{1}, blockDim{2}, gridDim{3};

// ---- collectives ----------------------------------------------------------------------------

inline void __syncthreads() { hufemu::barrier_wait(hufemu::cta()->bar, hufemu::cta()->bdim.x); }

inline int __syncthreads_or(int pred)
{
    hufemu::Cta *c = hufemu::cta();
    if (pred) c->or_acc = 1;
    __syncthreads();
    c->or_result = c->or_acc;
    __syncthreads();
    const int r = (int)c->or_result;
    c->or_acc = 0;  // every thread clears; harmless
    __syncthreads();
    return r;
}

inline int __syncthreads_count(int pred)
{
    static unsigned acc = 0, result = 0;  // CTAs run one after another
    if (pred) acc++;
    __syncthreads();
    result = acc;
    __syncthreads();
    const int r = (int)result;
    acc = 0;
    __syncthreads();
    return r;
}

inline void __syncwarp(unsigned = 0xffffffffu)
{
    hufemu::Warp *w = hufemu::cta()->cur->warp;
    const unsigned n = std::min(32u, hufemu::cta()->bdim.x - (hufemu::cta()->cur->tidx.x & ~31u));
    hufemu::barrier_wait(w->bar, n);
}

template <typename T>
inline T __shfl_sync(unsigned, T v, int src)
{
    return hufemu::warp_exchange(v, [&](uint64_t *x) { return hufemu::from_bits<T>(x[src & 31]); });
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned d)
{
    const unsigned l = hufemu::lane();
    return hufemu::warp_exchange(v, [&](uint64_t *x) { return l >= d ? hufemu::from_bits<T>(x[l - d]) : v; });
}
inline uint32_t __ballot_sync(unsigned, int pred)
{
    return hufemu::warp_exchange((uint32_t)(pred != 0), [&](uint64_t *x) {
        uint32_t m = 0;
        for (int l = 0; l < 32; l++) m |= (uint32_t)(x[l] & 1) << l;
        return m;
    });
}
inline uint32_t __match_any_sync(unsigned, uint32_t v)
{
    return hufemu::warp_exchange(v, [&](uint64_t *x) {
        uint32_t m = 0;
        for (int l = 0; l < 32; l++) m |= (uint32_t)((uint32_t)x[l] == v) << l;
        return m;
    });
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, unsigned d)
{
    const unsigned l = hufemu::lane();
    return hufemu::warp_exchange(v, [&](uint64_t *x) { return l + d < 32 ? hufemu::from_bits<T>(x[l + d]) : v; });
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int d)
{
    const unsigned l = hufemu::lane();
    return hufemu::warp_exchange(v, [&](uint64_t *x) { return hufemu::from_bits<T>(x[(l ^ d) & 31]); });
}

// ---- atomics (single OS thread: plain read-modify-write is atomic between yields) -----------

template <typename T>
inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T>
inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <typename T>
inline T atomicMax(T *p, T v) { T o = *p; *p = std::max(o, v); return o; }
template <typename T>
inline T atomicMin(T *p, T v) { T o = *p; *p = std::min(o, v); return o; }

// ---- integer intrinsics -----------------------------------------------------------------------

inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s)
{
    const uint64_t src = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t sel = (s >> (4 * i)) & 0xf;
        uint32_t byte = (uint32_t)(src >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh & 31));
}
inline uint32_t __funnelshift_rc(uint32_t lo, uint32_t hi, uint32_t sh)
{
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh < 32 ? sh : 32));
}
inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh)
{
    return (uint32_t)(((((uint64_t)hi << 32) | lo) << (sh & 31)) >> 32);
}
inline uint32_t __vcmpeq4(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int i = 0; i < 4; i++)
        if (((a >> (8 * i)) & 0xff) == ((b >> (8 * i)) & 0xff)) r |= 0xffu << (8 * i);
    return r;
}
inline uint32_t __vcmpleu4(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int i = 0; i < 4; i++)
        if (((a >> (8 * i)) & 0xff) <= ((b >> (8 * i)) & 0xff)) r |= 0xffu << (8 * i);
    return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __ffs(uint32_t v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline int __popc(uint32_t v) { return __builtin_popcount(v); }

template <typename A, typename B>
inline auto min(A a, B b) -> typename std::common_type<A, B>::type
{
    using T = typename std::common_type<A, B>::type;
    return (T)a < (T)b ? (T)a : (T)b;
}
template <typename A, typename B>
inline auto max(A a, B b) -> typename std::common_type<A, B>::type
{
    using T = typename std::common_type<A, B>::type;
    return (T)a > (T)b ? (T)a : (T)b;
}

// ---- the slice of the CUDA runtime that huf_b200.cu touches -----------------------------------

typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1 };
enum { cudaHostAllocDefault = 0, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
struct cudaDeviceProp {
    int major = 10, minor = 0, multiProcessorCount = 1;
    size_t sharedMemPerBlockOptin = 227 * 1024;
};

inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { *p = cudaDeviceProp(); return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (void *)1; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = (void *)1; return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -5; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
template <typename T>
inline cudaError_t cudaMalloc(T **p, size_t n)
{
    // poison fresh "device" memory so that reads of unwritten workspace show up in tests
    *p = (T *)malloc(n ? n : 1);
    if (*p) memset((void *)*p, 0xA5, n);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <typename T>
inline cudaError_t cudaMallocHost(T **p, size_t n) { *p = (T *)calloc(1, n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T>
inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned) { *p = (T *)malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
typedef void *cudaEvent_t;
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
template <typename K>
inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename K>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K, int, size_t) { *n = 1; return cudaSuccess; }

#define HUF_LAUNCH(kernel, grid, block, smem, stream, ...) \
    hufemu::launch((unsigned)(grid), (unsigned)(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })
