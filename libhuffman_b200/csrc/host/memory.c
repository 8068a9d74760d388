/* memory.c — huf_malloc and huf_config_t lifetime
 * [ref: src/malloc.c:7-19, src/config.c:7-33]. */
#include "internal.h"

/* Zero-initialised allocation of num * size bytes (calloc semantics, like the reference). */
huf_error_t
huf_malloc(void **ptr, size_t size, size_t num)
{
    HUF_REQUIRE(ptr);
    *ptr = calloc(num, size);
    return *ptr ? HUF_ERROR_SUCCESS : HUF_ERROR_MEMORY_ALLOCATION;
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
