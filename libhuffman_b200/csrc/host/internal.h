/*
 * internal.h — private declarations shared by the host C layer.
 * Nothing here is part of the public ABI.
 */
#ifndef HUF_B200_HOST_INTERNAL_H
#define HUF_B200_HOST_INTERNAL_H

#include <huffman.h>
#include <huffman/b200.h>

/* Early-out helpers: every public function validates pointers the way the reference's
 * routine_param_m does (include/huffman/sys.h:38-44 of the reference). */
#define HUF_REQUIRE(p) do { if (!(p)) return HUF_ERROR_INVALID_ARGUMENT; } while (0)
#define HUF_TRY(expr) do { huf_error_t e__ = (expr); if (e__ != HUF_ERROR_SUCCESS) return e__; } while (0)

/* In-memory stream state (private, like the reference's huf_membuf_t, src/io.c:66-71). */
typedef struct huf_memstream {
    void **slot;   /* the caller's buffer pointer; rewritten when the buffer grows */
    size_t rpos;   /* read cursor */
    size_t used;   /* bytes written */
    size_t room;   /* allocated bytes */
    unsigned uses; /* codec calls that borrowed the buffer (huf__memstream_borrow) */
    void *pinned;  /* the buffer while it is page-locked for direct DMA, else NULL */
} huf_memstream_t;

/* Returns the memstream behind `rw` when `rw` was created by huf_memopen, else NULL.
 * Lets the codec move bytes in bulk instead of through per-call callbacks. */
huf_memstream_t *huf__as_memstream(const huf_read_writer_t *rw);

/* Make room for `extra` more bytes at the write end; returns the write pointer. */
huf_error_t huf__memstream_reserve(huf_memstream_t *m, size_t extra, uint8_t **wptr);

/* A codec call is about to use the stream's buffer in place: counts the use and page-locks a
 * large buffer that keeps being used (huf_b200_host_register; b200.h says when it pays). */
void huf__memstream_borrow(huf_memstream_t *m);

/* Read exactly up to `want` bytes by calling rw->read until it reports end of data. */
huf_error_t huf__read_fully(huf_read_writer_t *rw, void *dst, size_t want, size_t *got);

/* memcpy with streaming stores for large staging copies (ntcopy.c). */
void huf__copy_stream(void *dst, const void *src, size_t n);

/* Process-wide GPU context used by huf_encode/huf_decode (created lazily). */
huf_error_t huf__codec_context(huf_b200_ctx_t **ctx);
void huf__codec_context_unlock(void);

#endif
