/*
 * histogram.c — the huf_histogram_t API object [ref: src/histogram.c:9-103].
 * Host-side object for API/link compatibility (exercised by the reference's
 * test/histogram_test.c); the codec's histogram is the K1 CUDA kernel.
 */
#include <string.h>

#include "internal.h"

huf_error_t
huf_histogram_init(huf_histogram_t **self, size_t iota, size_t length)
{
    HUF_REQUIRE(self);
    HUF_REQUIRE(iota);
    HUF_REQUIRE(length);
    HUF_TRY(huf_malloc((void **)self, sizeof(**self), 1));
    HUF_TRY(huf_malloc((void **)&(*self)->frequencies, sizeof(uint64_t), length));
    (*self)->iota = iota;
    (*self)->length = length;
    (*self)->start = (size_t)-1;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_histogram_free(huf_histogram_t **self)
{
    HUF_REQUIRE(self);
    if (*self) {
        free((*self)->frequencies);
        free(*self);
    }
    *self = NULL;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_histogram_reset(huf_histogram_t *self)
{
    HUF_REQUIRE(self);
    memset(self->frequencies, 0, self->length * sizeof(uint64_t));
    self->start = (size_t)-1;
    return HUF_ERROR_SUCCESS;
}

/* Count little-endian elements of `iota` bytes; a trailing partial element is ignored.
 * `start` tracks the smallest element value seen so far. */
huf_error_t
huf_histogram_populate(huf_histogram_t *self, void *buf, size_t len)
{
    HUF_REQUIRE(self);
    HUF_REQUIRE(buf);

    const uint8_t *p = buf;
    size_t width = self->iota > sizeof(uint64_t) ? sizeof(uint64_t) : self->iota;

    for (size_t n = len / self->iota; n; n--, p += self->iota) {
        uint64_t v = 0;
        memcpy(&v, p, width);
        self->frequencies[v]++;
        if (v < self->start) { /* start == (size_t)-1 when empty, so any v wins */
            self->start = (size_t)v;
        }
    }
    return HUF_ERROR_SUCCESS;
}
