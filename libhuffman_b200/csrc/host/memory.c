/* memory.c — huf_malloc and huf_config_t lifetime
 * [ref: src/malloc.c:7-19, src/config.c:7-33]. */
#include <stdint.h>
#include <sys/mman.h>

#include "internal.h"

/* Zero-initialised allocation of num * size bytes (calloc semantics, like the reference). */
huf_error_t
huf_malloc(void **ptr, size_t size, size_t num)
{
    HUF_REQUIRE(ptr);
    *ptr = calloc(num, size);
    if (!*ptr) {
        return HUF_ERROR_MEMORY_ALLOCATION;
    }
#ifdef MADV_HUGEPAGE
    /* Large stream buffers are about to be first-touched by bulk copies: ask for huge pages on
     * the page-aligned interior (a hint; the block stays an ordinary free()-able allocation). */
    if (num && size && num * size >= ((size_t)8 << 20)) {
        const uintptr_t lo = ((uintptr_t)*ptr + 0x1fffff) & ~(uintptr_t)0x1fffff;
        const uintptr_t hi = ((uintptr_t)*ptr + num * size) & ~(uintptr_t)0x1fffff;
        if (hi > lo) {
            (void)madvise((void *)lo, hi - lo, MADV_HUGEPAGE);
        }
    }
#endif
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_config_init(huf_config_t **self)
{
    HUF_REQUIRE(self);
    return huf_malloc((void **)self, sizeof(huf_config_t), 1);
}

huf_error_t
huf_config_free(huf_config_t **self)
{
    HUF_REQUIRE(self);
    free(*self);
    *self = NULL;
    return HUF_ERROR_SUCCESS;
}
