/*
 * codec.c — huf_encode / huf_decode: the host orchestrators of the GPU block codec
 * [ref: src/encoder.c:261-388, src/decoder.c:201-287].
 *
 * The reference walks the stream one block at a time through per-byte callbacks.  Here the
 * host only moves bytes: pull the block range from the reader, stage it in HBM, run the
 * sm_100a kernels through the C-ABI shim (<huffman/b200.h>), push the result to the writer.
 * There is no CPU implementation of the codec in this library: without a usable B200 the
 * calls fail with HUF_ERROR_FATAL.
 *
 * Behaviour kept from the reference:
 *   - config is never modified; blocksize 0 means "one block of `length` bytes"
 *     (src/encoder.c:163-165); length 0 succeeds without output (Q12);
 *   - a reader that runs dry mid-block gives HUF_ERROR_READ_WRITE after the complete blocks
 *     before it were written (src/encoder.c:296-299, Q11);
 *   - decode consumes whole blocks while consumed < length (src/decoder.c:218) and reports the
 *     reference's error codes; output of the blocks before a failing one is still written.
 * Deliberate fix: NULL config/reader/writer return
 * HUF_ERROR_INVALID_ARGUMENT instead of crashing in cleanup.
 */
#include <pthread.h>
#include <string.h>

#include "internal.h"

/* Largest input range encoded per device round trip (whole blocks; one oversized block is
 * still processed in one piece). */
#define HUF_ENCODE_SPAN ((uint64_t)1 << 30)

/* ---- the process-wide GPU context -------------------------------------------------------- */

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static huf_b200_ctx_t *g_ctx;

/* Device buffers cached across calls: streaming callers (the Python compressor) issue many
 * small huf_encode calls and must not pay cudaMalloc each time. */
static struct {
    void *in, *out;
    uint64_t in_cap, out_cap;
} g_dev;

huf_error_t
huf__codec_context(huf_b200_ctx_t **ctx)
{
    pthread_mutex_lock(&g_lock);
    if (!g_ctx) {
        huf_error_t err = huf_b200_ctx_create(&g_ctx, -1);
        if (err != HUF_ERROR_SUCCESS) {
            g_ctx = NULL;
            pthread_mutex_unlock(&g_lock);
            return err;
        }
    }
    *ctx = g_ctx;
    return HUF_ERROR_SUCCESS; /* lock stays held until huf__codec_context_unlock */
}

void
huf__codec_context_unlock(void)
{
    pthread_mutex_unlock(&g_lock);
}

static huf_error_t
dev_reserve(void **slot, uint64_t *cap, uint64_t want)
{
    if (want <= *cap) {
        return HUF_ERROR_SUCCESS;
    }
    if (*slot) {
        huf_b200_dev_free(*slot);
        *slot = NULL;
        *cap = 0;
    }
    want += want / 8 + 4096;
    HUF_TRY(huf_b200_dev_alloc(slot, want));
    *cap = want;
    return HUF_ERROR_SUCCESS;
}

/* ---- encoder / decoder objects ----------------------------------------------------------- */

struct __huf_encoder {
    huf_config_t config; /* private copy, blocksize already defaulted */
};

struct __huf_decoder {
    huf_config_t config;
};

huf_error_t
huf_encoder_init(huf_encoder_t **self, const huf_config_t *config)
{
    HUF_REQUIRE(self);
    HUF_REQUIRE(config);
    HUF_TRY(huf_malloc((void **)self, sizeof(**self), 1));
    (*self)->config = *config;
    if (!(*self)->config.blocksize) {
        (*self)->config.blocksize = config->length;
    }
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_encoder_free(huf_encoder_t **self)
{
    HUF_REQUIRE(self);
    free(*self);
    *self = NULL;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_decoder_init(huf_decoder_t **self, const huf_config_t *config)
{
    HUF_REQUIRE(self);
    HUF_REQUIRE(config);
    HUF_TRY(huf_malloc((void **)self, sizeof(**self), 1));
    (*self)->config = *config;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_decoder_free(huf_decoder_t **self)
{
    HUF_REQUIRE(self);
    free(*self);
    *self = NULL;
    return HUF_ERROR_SUCCESS;
}

/* ---- moving bytes between streams and the device ----------------------------------------- */

/* Obtain `want` input bytes.  Own memory streams lend their buffer (no copy); any other reader
 * fills `*scratch`.  *got may be short at end of data. */
static huf_error_t
pull(huf_read_writer_t *reader, uint64_t want, uint8_t **scratch, const uint8_t **data,
     uint64_t *got)
{
    huf_memstream_t *m = huf__as_memstream(reader);

    if (m) {
        uint64_t left = m->used - m->rpos;
        *got = want < left ? want : left;
        *data = (const uint8_t *)*m->slot + m->rpos;
        m->rpos += *got;
        return HUF_ERROR_SUCCESS;
    }
    free(*scratch);
    *scratch = NULL;
    HUF_TRY(huf_malloc((void **)scratch, 1, want ? want : 1));
    size_t n = 0;
    HUF_TRY(huf__read_fully(reader, *scratch, want, &n));
    *data = *scratch;
    *got = n;
    return HUF_ERROR_SUCCESS;
}

/* Deliver `count` device bytes to the writer. */
static huf_error_t
push(huf_read_writer_t *writer, const void *d_src, uint64_t count)
{
    if (!count) {
        return HUF_ERROR_SUCCESS;
    }
    huf_memstream_t *m = huf__as_memstream(writer);
    if (m) {
        uint8_t *dst = NULL;
        HUF_TRY(huf__memstream_reserve(m, count, &dst));
        HUF_TRY(huf_b200_copy_d2h(dst, d_src, count));
        m->used += count;
        return HUF_ERROR_SUCCESS;
    }
    uint8_t *tmp = NULL;
    HUF_TRY(huf_malloc((void **)&tmp, 1, count));
    huf_error_t err = huf_b200_copy_d2h(tmp, d_src, count);
    if (err == HUF_ERROR_SUCCESS) {
        err = writer->write(writer->stream, tmp, count);
    }
    free(tmp);
    return err;
}

/* ---- huf_encode ----------------------------------------------------------------------------- */

static huf_error_t
encode_locked(huf_b200_ctx_t *ctx, const huf_config_t *cfg)
{
    const uint64_t blocksize = cfg->blocksize ? cfg->blocksize : cfg->length;
    uint64_t span = HUF_ENCODE_SPAN / blocksize * blocksize; /* whole blocks per round trip */
    uint64_t left = cfg->length;
    uint8_t *scratch = NULL;
    huf_error_t err = HUF_ERROR_SUCCESS;

    if (!span) {
        span = blocksize;
    }
    while (left && err == HUF_ERROR_SUCCESS) {
        const uint64_t want = left < span ? left : span;
        const uint8_t *data = NULL;
        uint64_t got = 0;

        err = pull(cfg->reader, want, &scratch, &data, &got);
        if (err != HUF_ERROR_SUCCESS) {
            break;
        }
        /* a short read ends the stream after the blocks that are complete */
        uint64_t usable = got;
        if (got < want) {
            usable = got / blocksize * blocksize;
        }
        if (usable) {
            const uint64_t bound = huf_b200_encode_bound(usable, blocksize);
            uint64_t out_len = 0;

            err = dev_reserve(&g_dev.in, &g_dev.in_cap, usable);
            if (err == HUF_ERROR_SUCCESS) {
                err = dev_reserve(&g_dev.out, &g_dev.out_cap, bound);
            }
            if (err == HUF_ERROR_SUCCESS) {
                err = huf_b200_copy_h2d(g_dev.in, data, usable);
            }
            if (err == HUF_ERROR_SUCCESS) {
                err = huf_b200_encode_async(ctx, g_dev.in, usable, blocksize, g_dev.out,
                                            g_dev.out_cap, HUF_B200_STREAM_PRIVATE);
            }
            if (err == HUF_ERROR_SUCCESS) {
                err = huf_b200_encode_finish(ctx, &out_len);
            }
            if (err == HUF_ERROR_SUCCESS) {
                err = push(cfg->writer, g_dev.out, out_len);
            }
        }
        if (err == HUF_ERROR_SUCCESS && got < want) {
            err = HUF_ERROR_READ_WRITE;
        }
        left -= want;
    }
    free(scratch);
    return err;
}

huf_error_t
huf_encode(const huf_config_t *config)
{
    HUF_REQUIRE(config);
    HUF_REQUIRE(config->reader);
    HUF_REQUIRE(config->writer);
    if (!config->length) {
        return HUF_ERROR_SUCCESS;
    }
    huf_b200_ctx_t *ctx = NULL;
    HUF_TRY(huf__codec_context(&ctx));
    huf_error_t err = encode_locked(ctx, config);
    huf__codec_context_unlock();
    return err;
}

/* ---- huf_decode ----------------------------------------------------------------------------- */

static huf_error_t
decode_locked(huf_b200_ctx_t *ctx, const huf_config_t *cfg)
{
    huf_memstream_t *m = huf__as_memstream(cfg->reader);
    uint8_t *scratch = NULL;
    const uint8_t *data = NULL;
    uint64_t avail = 0;
    uint64_t length = cfg->length;
    int at_eof = 0;
    huf_error_t err;

    /* Input: the reference pulls bytes as it needs them, so the last block may reach past
     * `length` when the reader holds more.  A memory stream lends everything it has; other
     * readers are asked for `length` bytes first and for more only if a block needs them. */
    if (m) {
        avail = m->used - m->rpos;
        data = (const uint8_t *)*m->slot + m->rpos;
        at_eof = 1;
    } else {
        err = pull(cfg->reader, length, &scratch, &data, &avail);
        if (err != HUF_ERROR_SUCCESS) {
            free(scratch);
            return err;
        }
        at_eof = avail < length;
    }

    uint64_t done_in = 0; /* compressed bytes already consumed by finished rounds */
    for (;;) {
        uint64_t est = 0, out_len = 0, consumed = 0;
        const uint64_t in_now = avail - done_in;
        const uint64_t len_now = length - done_in;

        err = dev_reserve(&g_dev.in, &g_dev.in_cap, in_now + 16);
        if (err == HUF_ERROR_SUCCESS) {
            err = huf_b200_copy_h2d(g_dev.in, data + done_in, in_now);
        }
        if (err == HUF_ERROR_SUCCESS) {
            err = huf_b200_decode_plan(ctx, g_dev.in, in_now, len_now, &est, NULL,
                                       HUF_B200_STREAM_PRIVATE);
        }
        if (err == HUF_ERROR_SUCCESS) {
            err = dev_reserve(&g_dev.out, &g_dev.out_cap, est + 16);
        }
        if (err == HUF_ERROR_SUCCESS) {
            err = huf_b200_decode_async(ctx, g_dev.in, in_now, len_now, g_dev.out, g_dev.out_cap,
                                        HUF_B200_STREAM_PRIVATE);
        }
        if (err != HUF_ERROR_SUCCESS) {
            break;
        }
        err = huf_b200_decode_finish(ctx, &out_len, &consumed);
        /* whatever was decoded before a failure is still delivered */
        huf_error_t werr = push(cfg->writer, g_dev.out, out_len);
        if (werr != HUF_ERROR_SUCCESS) {
            err = werr;
            break;
        }
        done_in += consumed;
        if (done_in >= length) {
            break; /* src/decoder.c:218: the loop condition is only checked between blocks */
        }
        if (err == HUF_ERROR_MEMORY_ALLOCATION && consumed) {
            continue; /* output buffer was the limit: resume behind the blocks delivered */
        }
        if (err == HUF_ERROR_READ_WRITE && !at_eof) {
            /* the failing block may simply continue in bytes not pulled yet */
            uint64_t more = avail > 65536 ? avail : 65536;
            uint8_t *bigger = NULL;
            huf_error_t e2 = huf_malloc((void **)&bigger, 1, avail + more);
            if (e2 != HUF_ERROR_SUCCESS) {
                err = e2;
                break;
            }
            memcpy(bigger, data, avail);
            size_t n = 0;
            e2 = huf__read_fully(cfg->reader, bigger + avail, more, &n);
            free(scratch);
            scratch = bigger;
            data = bigger;
            if (e2 != HUF_ERROR_SUCCESS) {
                err = e2;
                break;
            }
            at_eof = n < more;
            avail += n;
            if (n) {
                continue;
            }
        }
        break;
    }
    if (m) {
        m->rpos += done_in; /* consume exactly the whole blocks, like the unbuffered reference */
    }
    free(scratch);
    return err;
}

huf_error_t
huf_decode(const huf_config_t *config)
{
    HUF_REQUIRE(config);
    HUF_REQUIRE(config->reader);
    HUF_REQUIRE(config->writer);
    if (!config->length) {
        return HUF_ERROR_SUCCESS;
    }
    huf_b200_ctx_t *ctx = NULL;
    HUF_TRY(huf__codec_context(&ctx));
    huf_error_t err = decode_locked(ctx, config);
    huf__codec_context_unlock();
    return err;
}
