// common.cuh — shared device helpers for the sm_100a block codec kernels.
#pragma once

#include <cstdint>
#include <cstring>
#ifndef HUF_EMU  // tests/emu/cuda_emu.h stands in for the CUDA headers in the CPU test lane
#include <cuda_runtime.h>
#endif

namespace hufb200 {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// Stream layout constants (reference: src/encoder.c:325-342).
constexpr int kHdrFixed = 10;        // u64 orig_len + i16 tree_len
constexpr int kMaxTreeElems = 1025;  // 4 * 256 + 1 (Q1: one more than HUF_BTREE_LEN)
constexpr int kTreeStride = 1032;    // int16 elements reserved per block in the workspace
constexpr int kScanThreads = 1024;   // single-CTA scan kernels

// Error codes mirrored from huf_error_t so device code can report them.
enum : uint32_t {
    kOk = 0,
    kErrNoMem = 1,
    kErrInval = 2,
    kErrIO = 3,
    kErrFatal = 4,
    kErrOverflow = 5,
    kErrCorrupt = 6,
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_in_cta() { return threadIdx.x >> 5; }

// Streaming 16-byte load: read-only path, do not keep the line in L1.
__device__ __forceinline__ uint4 ld_stream_u4(const void *p)
{
#ifdef HUF_EMU
    return *reinterpret_cast<const uint4 *>(p);
#else
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
#endif
}

// Asynchronous 4-byte global -> shared copy (LDGSTS): many can be in flight per thread without
// holding registers; cp_async_wait_all() makes the issuing thread's copies visible to it.
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc)
{
#ifdef HUF_EMU
    *reinterpret_cast<uint32_t *>(smem_dst) = *reinterpret_cast<const uint32_t *>(gsrc);
#else
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gsrc) : "memory");
#endif
}

__device__ __forceinline__ void cp_async_wait_all()
{
#ifndef HUF_EMU
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// Per-thread cp.async groups: commit closes the group of copies issued since the last commit,
// wait_group<N> returns when at most N of the thread's most recent groups are still in flight.
__device__ __forceinline__ void cp_async_commit()
{
#ifndef HUF_EMU
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}

template <int N>
__device__ __forceinline__ void cp_async_wait_group()
{
#ifndef HUF_EMU
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

// Hint: bring the 128-byte line at p into L2 (no register, no stall).
__device__ __forceinline__ void prefetch_l2(const void *p)
{
#ifndef HUF_EMU
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// Shared memory through explicit shared-window addresses (32-bit registers on the device):
// the hot decode loop keeps its table and stage bases in plain registers, so the compiler has
// no generic pointer whose window base it would rebuild (S2R + LEA) inside the loop, and an
// aligned table base can be OR-ed into the index (one LOP3 instead of clamp + add).
#ifdef HUF_EMU
typedef uintptr_t saddr_t;
inline saddr_t smem_addr(const void *p) { return reinterpret_cast<uintptr_t>(p); }
inline saddr_t saddr_or(saddr_t base, uint32_t off) { return base + off; }
inline saddr_t saddr_pin(saddr_t a) { return a; }
inline uint32_t lds_u16(saddr_t a) { return *reinterpret_cast<const uint16_t *>(a); }
inline uint32_t lds_u32(saddr_t a) { return *reinterpret_cast<const uint32_t *>(a); }
// v = *a when pred is non-zero (predicated, no branch on the device)
inline void lds_u32_if(uint32_t &v, saddr_t a, uint32_t pred)
{
    if (pred) v = *reinterpret_cast<const uint32_t *>(a);
}
inline void lds_u32x2(saddr_t a, uint32_t &w0, uint32_t &w1)
{
    const uint32_t *p = reinterpret_cast<const uint32_t *>(a);
    w0 = p[0];
    w1 = p[1];
}
inline void lds_u32x3(saddr_t a, uint32_t &w0, uint32_t &w1, uint32_t &w2)
{
    const uint32_t *p = reinterpret_cast<const uint32_t *>(a);
    w0 = p[0];
    w1 = p[1];
    w2 = p[2];
}
inline void sts_u32(saddr_t a, uint32_t v) { *reinterpret_cast<uint32_t *>(a) = v; }
inline void sts_u8(saddr_t a, uint32_t v) { *reinterpret_cast<uint8_t *>(a) = (uint8_t)v; }
#else
typedef uint32_t saddr_t;
__device__ __forceinline__ saddr_t smem_addr(const void *p) { return (saddr_t)__cvta_generic_to_shared(p); }
// Pins an address in a register: the value becomes opaque to the compiler, which would
// otherwise rebuild it from the window base next to every use.
__device__ __forceinline__ saddr_t saddr_pin(saddr_t a)
{
    saddr_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(a));
    return r;
}
// base must be aligned beyond every bit of off
__device__ __forceinline__ saddr_t saddr_or(saddr_t base, uint32_t off) { return base | off; }
__device__ __forceinline__ uint32_t lds_u16(saddr_t a)
{
    uint32_t v;
    asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(saddr_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void lds_u32_if(uint32_t &v, saddr_t a, uint32_t pred)
{
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p ld.shared.u32 %0, [%1]; }" : "+r"(v) : "r"(a), "r"(pred));
}
__device__ __forceinline__ void lds_u32x2(saddr_t a, uint32_t &w0, uint32_t &w1)
{
    asm volatile("ld.shared.u32 %0, [%2]; ld.shared.u32 %1, [%2+4];" : "=r"(w0), "=r"(w1) : "r"(a));
}
__device__ __forceinline__ void lds_u32x3(saddr_t a, uint32_t &w0, uint32_t &w1, uint32_t &w2)
{
    asm volatile("ld.shared.u32 %0, [%3]; ld.shared.u32 %1, [%3+4]; ld.shared.u32 %2, [%3+8];"
                 : "=r"(w0), "=r"(w1), "=r"(w2)
                 : "r"(a));
}
__device__ __forceinline__ void sts_u32(saddr_t a, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u8(saddr_t a, uint32_t v)
{
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
#endif

// Bulk asynchronous copy global -> shared by the TMA unit (cp.async.bulk, SASS UBLKCP): one
// thread issues one instruction for the whole range, the bytes land without any register or
// load/store-unit traffic of the CTA, and an mbarrier in shared memory counts them in.
// Addresses and size are multiples of 16.  (tests/emu: a plain copy at request time.)
#ifdef HUF_EMU
inline void mbar_init(void *, uint32_t) {}
inline void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, void *)
{
    memcpy(smem_dst, gsrc, bytes);
}
inline void mbar_wait(void *, uint32_t) {}
inline void fence_async_proxy() {}
#else
__device__ __forceinline__ void mbar_init(void *mbar, uint32_t count)
{
    const uint32_t m = (uint32_t)__cvta_generic_to_shared(mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(m), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// Orders the CTA's earlier generic-proxy accesses of shared memory before later writes of the
// async proxy (the TMA unit) to the same bytes.
__device__ __forceinline__ void fence_async_proxy()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// One thread: expect `bytes` on the barrier and start the copy.
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, void *mbar)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const uint32_t m = (uint32_t)__cvta_generic_to_shared(mbar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gsrc), "r"(bytes), "r"(m)
                 : "memory");
}
// Every waiting thread: returns when the phase with the given parity has completed.
__device__ __forceinline__ void mbar_wait(void *mbar, uint32_t parity)
{
    const uint32_t m = (uint32_t)__cvta_generic_to_shared(mbar);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "HUF_MBAR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra HUF_MBAR_DONE;\n\t"
        "bra HUF_MBAR_WAIT;\n\t"
        "HUF_MBAR_DONE:\n\t"
        "}" ::"r"(m),
        "r"(parity)
        : "memory");
}
#endif

__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
    return v;
}

__device__ __forceinline__ uint32_t warp_max(uint32_t v)
{
#pragma unroll
    for (int d = 16; d; d >>= 1) v = max(v, __shfl_xor_sync(kFull, v, d));
    return v;
}

// Inclusive warp prefix sum.
template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v)
{
    const int l = lane_id();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(kFull, v, d);
        if (l >= d) v += o;
    }
    return v;
}

// Per-thread pieces of the single-CTA scans: a thread owns the contiguous range [lo, hi).
// Eight loads are issued before any use, so a range costs a few memory latencies, not one per item.
template <typename T>
__device__ __forceinline__ uint64_t range_sum(const T *in, uint64_t lo, uint64_t hi)
{
    uint64_t sum = 0;
    uint64_t i = lo;
    for (; i + 8 <= hi; i += 8) {
        T v[8];
#pragma unroll
        for (int q = 0; q < 8; q++) v[q] = in[i + q];
#pragma unroll
        for (int q = 0; q < 8; q++) sum += v[q];
    }
    for (; i < hi; i++) sum += in[i];
    return sum;
}

// out[i] = run + sum(in[lo..i)) for i in [lo, hi); returns run + sum(in[lo..hi)).
template <typename T>
__device__ __forceinline__ uint64_t range_excl_scan(const T *in, uint64_t *out, uint64_t lo, uint64_t hi,
                                                    uint64_t run)
{
    uint64_t i = lo;
    for (; i + 8 <= hi; i += 8) {
        T v[8];
#pragma unroll
        for (int q = 0; q < 8; q++) v[q] = in[i + q];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            out[i + q] = run;
            run += v[q];
        }
    }
    for (; i < hi; i++) {
        out[i] = run;
        run += in[i];
    }
    return run;
}

// Exclusive scan of in[0..n) by one CTA of kScanThreads threads with COALESCED accesses:
// out[i] = base + sum(in[0..i)), returns base + sum(in[0..n)) (valid in every thread).
// Every warp owns a contiguous run of elements and walks it in tiles of 32 (one element per
// lane).  (A contiguous range per THREAD makes every load and store of a warp touch 32
// different lines, and a single SM takes one line per cycle: 25 us for 16 K elements.)
template <typename T>
__device__ __forceinline__ uint64_t cta_excl_scan(const T *in, uint64_t *out, uint64_t n, uint64_t base,
                                                  uint64_t *warp_tot /* shared, kScanThreads / 32 + 1 */)
{
    constexpr int kW = kScanThreads / 32;
    const int w = warp_in_cta(), l = lane_id();
    const uint64_t span = ((n + (uint64_t)kW * 32 - 1) / ((uint64_t)kW * 32)) * 32;  // per warp, tiles of 32
    const uint64_t lo = n < (uint64_t)w * span ? n : (uint64_t)w * span;
    const uint64_t hi = n < lo + span ? n : lo + span;
    uint64_t s = 0;
    for (uint64_t i = lo + l; i < hi; i += 32) s += in[i];
    s = warp_sum(s);
    if (l == 0) warp_tot[w] = s;
    __syncthreads();
    if (w == 0) {
        const uint64_t t = warp_tot[l];
        const uint64_t ti = warp_incl_scan(t);
        warp_tot[l] = ti - t;
        if (l == 31) warp_tot[kW] = ti;
    }
    __syncthreads();
    uint64_t run = base + warp_tot[w];
    for (uint64_t t0 = lo; t0 < hi; t0 += 32) {
        const uint64_t i = t0 + l;
        const uint64_t v = i < hi ? (uint64_t)in[i] : 0;
        const uint64_t incl = warp_incl_scan(v);
        if (i < hi) out[i] = run + incl - v;
        run += __shfl_sync(kFull, incl, 31);
    }
    return base + warp_tot[kW];
}

}  // namespace hufb200
