// common.cuh — shared device helpers for the sm_100a block codec kernels.
#pragma once

#include <cstdint>
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

// Hint: bring the 128-byte line at p into L2 (no register, no stall).
__device__ __forceinline__ void prefetch_l2(const void *p)
{
#ifndef HUF_EMU
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

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

}  // namespace hufb200
