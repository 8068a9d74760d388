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
#include <stdlib.h>
#include <string.h>

#include "internal.h"

/* ---- the process-wide GPU contexts ------------------------------------------------------- */

/* One context on the current device, created lazily; or one per device listed in the
 * environment variable HUF_B200_DEVICES ("0,1,2,3" or "all"): huf_encode / huf_decode then split
 * a large call over those GPUs (contiguous block ranges / byte ranges).  The
 * reference has no configuration for this (huf_config_t must keep its layout), hence the
 * environment. */
#define HUF_MAX_DEVICES 16

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static huf_b200_ctx_t *g_ctx[HUF_MAX_DEVICES];
static int g_nctx;

static huf_error_t
contexts_create(void)
{
    const char *env = getenv("HUF_B200_DEVICES");
    int want[HUF_MAX_DEVICES];
    int n = 0;

    if (env && *env) {
        const int have = huf_b200_device_count();
        if (!strcmp(env, "all")) {
            for (int d = 0; d < have && n < HUF_MAX_DEVICES; d++) {
                want[n++] = d;
            }
        } else {
            const char *p = env;
            while (*p && n < HUF_MAX_DEVICES) {
                char *end = NULL;
                long d = strtol(p, &end, 10);
                if (end == p) {
                    break;
                }
                if (d >= 0 && d < have) {
                    want[n++] = (int)d;
                }
                p = *end == ',' ? end + 1 : end;
            }
        }
    }
    if (!n) {
        want[n++] = -1; /* the current device */
    }
    for (int i = 0; i < n; i++) {
        huf_error_t err = huf_b200_ctx_create(&g_ctx[i], want[i]);
        if (err != HUF_ERROR_SUCCESS) {
            while (i-- > 0) {
                huf_b200_ctx_destroy(&g_ctx[i]);
            }
            return err;
        }
    }
    g_nctx = n;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf__codec_context(huf_b200_ctx_t **ctx)
{
    pthread_mutex_lock(&g_lock);
    if (!g_nctx) {
        huf_error_t err = contexts_create();
        if (err != HUF_ERROR_SUCCESS) {
            pthread_mutex_unlock(&g_lock);
            return err;
        }
    }
    *ctx = g_ctx[0];
    return HUF_ERROR_SUCCESS; /* lock stays held until huf__codec_context_unlock */
}

void
huf__codec_context_unlock(void)
{
    pthread_mutex_unlock(&g_lock);
}

/* Calls at least this large are split over the listed devices (smaller ones are not worth the
 * second PCIe link); HUF_B200_MULTI_MIN overrides the threshold in bytes. */
static uint64_t
multi_threshold(void)
{
    const char *env = getenv("HUF_B200_MULTI_MIN");
    return env ? (uint64_t)strtoull(env, NULL, 10) : (uint64_t)64 << 20;
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

/* ---- streams as byte sources and sinks ------------------------------------------------------
 *
 * The host lanes of the shim (huf_b200_encode_host / huf_b200_decode_host) overlap the host
 * copies, both PCIe directions and the kernels; this file only tells them where the bytes are.
 * Own memory streams lend their buffers (bulk copies by the shim's copy threads, no per-call
 * callbacks); any other reader fills the pinned buffer it is handed through its read callback,
 * any other writer is handed the pinned result. */

static huf_error_t
reader_pull(void *arg, void *dst, uint64_t want, uint64_t *got)
{
    size_t n = 0;
    HUF_TRY(huf__read_fully(arg, dst, want, &n));
    *got = n;
    return HUF_ERROR_SUCCESS;
}

static huf_error_t
writer_push(void *arg, const void *src, uint64_t count)
{
    huf_read_writer_t *w = arg;
    return w->write(w->stream, src, count);
}

typedef struct {
    huf_memstream_t *m;
    uint64_t expect; /* bytes the call is expected to deliver in total (sizes the first growth) */
} mem_sink_t;

static huf_error_t
mem_reserve(void *arg, uint64_t count, void **dst)
{
    mem_sink_t *s = arg;
    uint8_t *p = NULL;
    uint64_t want = count;

    /* grow once to the expected size instead of doubling through it (every growth copies) */
    if (s->m->room - s->m->used < count && s->expect > count) {
        want = s->expect;
    }
    HUF_TRY(huf__memstream_reserve(s->m, want, &p));
    s->expect = s->expect > count ? s->expect - count : 0;
    *dst = p;
    return HUF_ERROR_SUCCESS;
}

static huf_error_t
mem_commit(void *arg, uint64_t count)
{
    mem_sink_t *s = arg;
    s->m->used += count;
    return HUF_ERROR_SUCCESS;
}

static huf_error_t
mem_room(void *arg, void **dst, uint64_t *avail)
{
    mem_sink_t *s = arg;
    *dst = *s->m->slot ? (uint8_t *)*s->m->slot + s->m->used : NULL;
    *avail = *s->m->slot ? s->m->room - s->m->used : 0;
    return HUF_ERROR_SUCCESS;
}

static void
make_source(huf_read_writer_t *reader, huf_b200_source_t *src)
{
    huf_memstream_t *m = huf__as_memstream(reader);

    memset(src, 0, sizeof(*src));
    if (m) {
        huf__memstream_borrow(m);
        src->data = (const uint8_t *)*m->slot + m->rpos;
        src->size = m->used - m->rpos;
        if (!src->data) {
            src->data = ""; /* an empty stream without a buffer still is a (dry) contiguous source */
        }
    } else {
        src->pull = reader_pull;
        src->arg = reader;
    }
}

static void
make_sink(huf_read_writer_t *writer, huf_b200_sink_t *dst, mem_sink_t *ms, uint64_t expect)
{
    memset(dst, 0, sizeof(*dst));
    ms->m = huf__as_memstream(writer);
    ms->expect = expect;
    if (ms->m) {
        huf__memstream_borrow(ms->m);
        dst->reserve = mem_reserve;
        dst->commit = mem_commit;
        dst->room = mem_room;
        dst->arg = ms;
    } else {
        dst->push = writer_push;
        dst->arg = writer;
    }
}

/* ---- huf_encode ----------------------------------------------------------------------------- */

huf_error_t
huf_encode(const huf_config_t *config)
{
    HUF_REQUIRE(config);
    HUF_REQUIRE(config->reader);
    HUF_REQUIRE(config->writer);
    if (!config->length) {
        return HUF_ERROR_SUCCESS; /* Q12 */
    }
    huf_b200_ctx_t *ctx = NULL;
    HUF_TRY(huf__codec_context(&ctx));

    const uint64_t blocksize = config->blocksize ? config->blocksize : config->length;
    huf_b200_source_t src;
    huf_b200_sink_t dst;
    mem_sink_t ms;
    uint64_t taken = 0;

    make_source(config->reader, &src);
    /* what the stream will be about: the bound only sizes a fresh buffer, pages never written
     * are never touched */
    uint64_t in_now = config->length;
    if (src.data && src.size < in_now) {
        in_now = src.size;
    }
    make_sink(config->writer, &dst, &ms, huf_b200_encode_bound(in_now, blocksize));
    huf_error_t err;
    if (g_nctx > 1 && src.data && dst.reserve && in_now >= config->length && in_now >= multi_threshold() &&
        in_now / blocksize >= (uint64_t)g_nctx) {
        /* whole input at hand, lending sink, enough blocks: one block range per GPU */
        err = huf_b200_encode_host_multi(g_ctx, g_nctx, src.data, in_now, blocksize, &dst, NULL);
        taken = err == HUF_ERROR_SUCCESS ? in_now : 0;
    } else {
        err = huf_b200_encode_host(ctx, &src, config->length, blocksize, &dst, &taken);
    }
    huf_memstream_t *m = huf__as_memstream(config->reader);
    if (m) {
        m->rpos += taken;
    }
    huf__codec_context_unlock();
    return err;
}

/* ---- huf_decode ----------------------------------------------------------------------------- */

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

    huf_b200_source_t src;
    huf_b200_sink_t dst;
    mem_sink_t ms;
    uint64_t consumed = 0;

    /* The reference pulls bytes as it needs them, so the last block may reach past `length`
     * when the reader holds more: a memory stream lends everything it has, other readers are
     * asked for `length` bytes first and for more only if a block needs them. */
    make_source(config->reader, &src);
    make_sink(config->writer, &dst, &ms, config->length + config->length / 2);
    huf_error_t err;
    if (g_nctx > 1 && src.data && dst.reserve && src.size >= multi_threshold()) {
        /* one byte range of the stream per GPU */
        err = huf_b200_decode_host_multi(g_ctx, g_nctx, src.data, src.size, config->length, &dst, &consumed);
    } else {
        err = huf_b200_decode_host(ctx, &src, config->length, &dst, &consumed);
    }
    huf_memstream_t *m = huf__as_memstream(config->reader);
    if (m) {
        m->rpos += consumed; /* consume exactly the whole blocks, like the unbuffered reference */
    }
    huf__codec_context_unlock();
    return err;
}
