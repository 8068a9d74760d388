/*
 * bufio.c — buffered byte I/O and the 8-bit MSB-first bit accumulator
 * [ref: src/bufio.c:18-32 (bits), src/bufio.c:37-320 (bytes)].
 *
 * Host-side API objects kept for link compatibility; the GPU codec stages whole block ranges
 * and does not push bytes through these one at a time.
 */
#include <string.h>

#include "internal.h"

/* Bits enter at bit 7 and move down; offset counts the free bits (8 = empty). */
void
huf_bit_write(huf_bit_read_writer_t *self, uint8_t bit)
{
    if (self->offset) {
        self->offset--;
    }
    self->bits |= (uint8_t)((bit & 1u) << self->offset);
}

void
huf_bit_read_writer_reset(huf_bit_read_writer_t *self)
{
    self->bits = 0;
    self->offset = 8;
}

huf_error_t
huf_bufio_read_writer_init(huf_bufio_read_writer_t **self, huf_read_writer_t *read_writer,
                           size_t size)
{
    HUF_REQUIRE(self);
    HUF_REQUIRE(read_writer);
    HUF_TRY(huf_malloc((void **)self, sizeof(**self), 1));
    if (size) { /* 0 = pass-through, no buffer */
        HUF_TRY(huf_malloc((void **)&(*self)->bytes, 1, size));
    }
    (*self)->capacity = size;
    (*self)->read_writer = read_writer;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_bufio_read_writer_free(huf_bufio_read_writer_t **self)
{
    HUF_REQUIRE(self);
    if (*self) {
        free((*self)->bytes);
        free(*self);
    }
    *self = NULL;
    return HUF_ERROR_SUCCESS;
}

static huf_error_t
drain(huf_bufio_read_writer_t *b)
{
    if (b->length) {
        huf_read_writer_t *rw = b->read_writer;
        HUF_TRY(rw->write(rw->stream, b->bytes, b->length));
        b->length = 0;
    }
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_bufio_read_writer_flush(huf_bufio_read_writer_t *self)
{
    HUF_REQUIRE(self);
    return drain(self);
}

huf_error_t
huf_bufio_write(huf_bufio_read_writer_t *self, const void *buf, size_t size)
{
    HUF_REQUIRE(self);
    HUF_REQUIRE(buf);
    if (self->capacity && self->length >= self->capacity) {
        HUF_TRY(drain(self));
    }
    if (self->capacity && size <= self->capacity - self->length) {
        memcpy(self->bytes + self->length, buf, size);
        self->length += size;
    } else if (size) {
        /* does not fit: push what is buffered, then hand the caller's bytes straight on */
        huf_read_writer_t *rw = self->read_writer;
        HUF_TRY(drain(self));
        HUF_TRY(rw->write(rw->stream, buf, size));
    }
    self->have_been_processed += size;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_bufio_read(huf_bufio_read_writer_t *self, void *buf, size_t size)
{
    HUF_REQUIRE(self);
    HUF_REQUIRE(buf);

    huf_read_writer_t *rw = self->read_writer;
    uint8_t *dst = buf;
    size_t want = size;
    size_t cached = self->length - self->offset;

    if (cached && want) {
        size_t n = cached < want ? cached : want;
        memcpy(dst, self->bytes + self->offset, n);
        self->offset += n;
        dst += n;
        want -= n;
    }
    if (want) {
        if (want >= self->capacity) {
            /* large (or unbuffered) request: bypass the buffer */
            size_t n = want;
            HUF_TRY(rw->read(rw->stream, dst, &n));
            self->length = self->offset = 0;
            if (n < want) {
                return HUF_ERROR_READ_WRITE;
            }
        } else {
            /* refill, then serve from the buffer */
            self->length = self->capacity;
            self->offset = 0;
            HUF_TRY(rw->read(rw->stream, self->bytes, &self->length));
            if (self->length < want) {
                return HUF_ERROR_READ_WRITE;
            }
            memcpy(dst, self->bytes, want);
            self->offset = want;
        }
    }
    self->have_been_processed += size;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_bufio_read_uint8(huf_bufio_read_writer_t *self, uint8_t *byte)
{
    HUF_REQUIRE(self);
    HUF_REQUIRE(byte);
    return huf_bufio_read(self, byte, 1);
}

huf_error_t
huf_bufio_write_uint8(huf_bufio_read_writer_t *self, uint8_t byte)
{
    HUF_REQUIRE(self);
    return huf_bufio_write(self, &byte, 1);
}
