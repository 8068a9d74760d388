/*
 * streams.c — the two stock huf_read_writer_t implementations
 * [ref: src/io.c:9-63 (fd), src/io.c:66-226 (mem)].
 *
 * Behaviour kept from the reference: read() reports a short or zero count at end of data
 * with HUF_ERROR_SUCCESS; the memory stream grows to max(2*capacity, 2*count) (the
 * 2 -> 16 case of test/io_test.c:49-62) and swaps the caller's buffer pointer; huf_memclose
 * leaves the data buffer to the caller.
 * Deliberate fixes: Q6 the fd is stored inside the stream object instead of
 * pointing at a dead stack slot; Q7 read(2) errors are detected through ssize_t; Q8 growth
 * never ends up smaller than used + count.
 */
#include <errno.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "internal.h"

/* ---- memory stream ---------------------------------------------------------------------- */

#define HUF_PIN_MIN_BYTES ((size_t)32 << 20) /* smaller buffers are not worth page-locking */

static void
mem_unpin(huf_memstream_t *m)
{
    if (m->pinned) {
        (void)huf_b200_host_unregister(m->pinned);
        m->pinned = NULL;
    }
}

void
huf__memstream_borrow(huf_memstream_t *m)
{
    /* HUF_B200_PIN_AFTER: page-lock from this use on (0: never); read per call, it is cheap */
    const char *env = getenv("HUF_B200_PIN_AFTER");
    int after = env ? atoi(env) : 4;

    if (after < 0) {
        after = 0;
    }
    size_t floor_bytes = HUF_PIN_MIN_BYTES;
    if ((env = getenv("HUF_B200_PIN_MIN_BYTES")) != NULL && atoll(env) > 0) {
        floor_bytes = (size_t)atoll(env); /* (tests) */
    }
    m->uses++;
    if (after && m->uses >= (unsigned)after && !m->pinned && *m->slot && m->room >= floor_bytes) {
        if (huf_b200_host_register(*m->slot, m->room) == HUF_ERROR_SUCCESS) {
            m->pinned = *m->slot;
        } else {
            m->uses = 0; /* (no GPU, no memory to lock: try again much later, not every call) */
        }
    }
}

huf_error_t
huf__memstream_reserve(huf_memstream_t *m, size_t extra, uint8_t **wptr)
{
    if (m->room - m->used < extra || !*m->slot) {
        size_t grown = m->room * 2;
        if (extra > grown) {
            grown = extra * 2;
        }
        if (grown < m->used + extra) { /* Q8 */
            grown = m->used + extra;
        }
        void *fresh = NULL;
        HUF_TRY(huf_malloc(&fresh, 1, grown));
        if (m->used) {
            memcpy(fresh, *m->slot, m->used);
        }
        mem_unpin(m); /* (the next codec call locks the new buffer if the stream is a busy one) */
        free(*m->slot);
        *m->slot = fresh;
        m->room = grown;
    }
    *wptr = (uint8_t *)*m->slot + m->used;
    return HUF_ERROR_SUCCESS;
}

static huf_error_t
mem_write(void *stream, const void *buf, size_t count)
{
    huf_memstream_t *m = stream;
    uint8_t *dst = NULL;

    if (!count) {
        return HUF_ERROR_SUCCESS;
    }
    HUF_TRY(huf__memstream_reserve(m, count, &dst));
    memcpy(dst, buf, count);
    m->used += count;
    return HUF_ERROR_SUCCESS;
}

static huf_error_t
mem_read(void *stream, void *buf, size_t *count)
{
    huf_memstream_t *m = stream;
    size_t left = m->used - m->rpos;
    size_t n = *count < left ? *count : left;

    if (n) {
        memcpy(buf, (const uint8_t *)*m->slot + m->rpos, n);
        m->rpos += n;
    }
    *count = n;
    return HUF_ERROR_SUCCESS;
}

huf_memstream_t *
huf__as_memstream(const huf_read_writer_t *rw)
{
    if (rw && rw->read == mem_read && rw->write == mem_write) {
        return rw->stream;
    }
    return NULL;
}

huf_error_t
huf_memopen(huf_read_writer_t **self, void **buf, size_t capacity)
{
    huf_memstream_t *m = NULL;

    HUF_REQUIRE(self);
    HUF_REQUIRE(buf);
    HUF_TRY(huf_malloc(buf, 1, capacity));
    HUF_TRY(huf_malloc((void **)self, sizeof(**self), 1));
    HUF_TRY(huf_malloc((void **)&m, sizeof(*m), 1));
    m->slot = buf;
    m->room = capacity;
    (*self)->stream = m;
    (*self)->write = mem_write;
    (*self)->read = mem_read;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_memlen(const huf_read_writer_t *self, size_t *len)
{
    HUF_REQUIRE(self);
    HUF_REQUIRE(len);
    *len = ((const huf_memstream_t *)self->stream)->used;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_memcap(const huf_read_writer_t *self, size_t *cap)
{
    HUF_REQUIRE(self);
    HUF_REQUIRE(cap);
    *cap = ((const huf_memstream_t *)self->stream)->room;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_memrewind(huf_read_writer_t *self)
{
    HUF_REQUIRE(self);
    huf_memstream_t *m = self->stream;
    m->used = 0;
    m->rpos = 0;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_memclose(huf_read_writer_t **self)
{
    HUF_REQUIRE(self);
    if (*self) {
        if ((*self)->read == mem_read) {
            mem_unpin((*self)->stream); /* the caller free()s the buffer: hand it back unlocked */
        }
        free((*self)->stream); /* the data buffer stays with the caller */
        free(*self);
    }
    *self = NULL;
    return HUF_ERROR_SUCCESS;
}

/* ---- file-descriptor stream ------------------------------------------------------------- */

typedef struct {
    int fd;
} huf_fdstream_t;

static huf_error_t
fd_write(void *stream, const void *buf, size_t count)
{
    const huf_fdstream_t *s = stream;
    const uint8_t *p = buf;

    while (count) {
        ssize_t n = write(s->fd, p, count);
        if (n < 0) {
            if (errno == EINTR) {
                continue;
            }
            return HUF_ERROR_READ_WRITE;
        }
        p += n;
        count -= (size_t)n;
    }
    return HUF_ERROR_SUCCESS;
}

static huf_error_t
fd_read(void *stream, void *buf, size_t *count)
{
    const huf_fdstream_t *s = stream;
    ssize_t n;

    do {
        n = read(s->fd, buf, *count);
    } while (n < 0 && errno == EINTR);
    if (n < 0) { /* Q7 */
        *count = 0;
        return HUF_ERROR_READ_WRITE;
    }
    *count = (size_t)n;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_fdopen(huf_read_writer_t **self, int fd)
{
    huf_fdstream_t *s = NULL;

    HUF_REQUIRE(self);
    HUF_TRY(huf_malloc((void **)self, sizeof(**self), 1));
    HUF_TRY(huf_malloc((void **)&s, sizeof(*s), 1));
    s->fd = fd; /* Q6 */
    (*self)->stream = s;
    (*self)->write = fd_write;
    (*self)->read = fd_read;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_fdclose(huf_read_writer_t **self)
{
    HUF_REQUIRE(self);
    if (*self) {
        free((*self)->stream);
        free(*self);
    }
    *self = NULL;
    return HUF_ERROR_SUCCESS;
}

/* ---- helper used by the codec ----------------------------------------------------------- */

huf_error_t
huf__read_fully(huf_read_writer_t *rw, void *dst, size_t want, size_t *got)
{
    uint8_t *p = dst;
    size_t have = 0;

    while (have < want) {
        size_t n = want - have;
        HUF_TRY(rw->read(rw->stream, p + have, &n));
        if (!n) {
            break; /* end of data */
        }
        have += n;
    }
    *got = have;
    return HUF_ERROR_SUCCESS;
}
