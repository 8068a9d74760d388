/*
 * ntcopy.c — bulk copy with streaming (non-temporal) stores for the staging copies of the host
 * lanes: neither side of those copies is read again by the CPU soon (the pinned buffer is read
 * by the DMA engine, the caller's buffer by the caller much later), so writing the destination
 * around the cache saves the read-for-ownership of every destination line -- a third of the
 * DRAM traffic of a copy, and the host memory system is what bounds huf_encode/huf_decode end
 * to end.  No reference counterpart (the reference copies byte by byte through src/bufio.c).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "internal.h"

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>

__attribute__((target("avx2"))) static void
copy_stream_avx2(uint8_t *dst, const uint8_t *src, size_t n)
{
    /* head: bring dst to 32-byte alignment */
    size_t head = (32 - ((uintptr_t)dst & 31)) & 31;
    if (head > n) {
        head = n;
    }
    memcpy(dst, src, head);
    dst += head;
    src += head;
    n -= head;
    while (n >= 128) {
        const __m256i a = _mm256_loadu_si256((const __m256i *)(src));
        const __m256i b = _mm256_loadu_si256((const __m256i *)(src + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i *)(src + 64));
        const __m256i d = _mm256_loadu_si256((const __m256i *)(src + 96));
        _mm256_stream_si256((__m256i *)(dst), a);
        _mm256_stream_si256((__m256i *)(dst + 32), b);
        _mm256_stream_si256((__m256i *)(dst + 64), c);
        _mm256_stream_si256((__m256i *)(dst + 96), d);
        src += 128;
        dst += 128;
        n -= 128;
    }
    _mm_sfence();
    memcpy(dst, src, n);
}
#endif

void
huf__copy_stream(void *dst, const void *src, size_t n)
{
#if defined(__x86_64__) && defined(__GNUC__)
    static int have = -1;
    if (have < 0) {
        __builtin_cpu_init();
        have = __builtin_cpu_supports("avx2") ? 1 : 0;
    }
    if (have && n >= 4096) {
        copy_stream_avx2(dst, src, n);
        return;
    }
#endif
    memcpy(dst, src, n);
}
