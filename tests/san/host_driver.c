/* TEST INFRASTRUCTURE: exercises the host C layer (streams, bufio, unit-level objects, the
 * <huffman/sys.h> macros) under ASan/UBSan.  The codec itself needs the GPU and is not reached:
 * huf_encode must fail cleanly with HUF_ERROR_FATAL here. */
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <huffman.h>
#include <huffman/sys.h>

#define CHECK(x) do { if (!(x)) { fprintf(stderr, "CHECK failed at line %d: %s\n", __LINE__, #x); return 1; } } while (0)

static huf_error_t with_macros(int *p, int v)
{
    routine_m();
    void *buf = NULL;
    routine_param_m(p);
    routine_inrange_m(v, 1, 10);
    if (huf_malloc(void_pptr_m(&buf), 1, 16) != HUF_ERROR_SUCCESS) {
        routine_error_m(HUF_ERROR_MEMORY_ALLOCATION);
    }
    *p = v;
    routine_success_m();
    routine_ensure_m();
    free(buf);
    routine_defer_m();
}

int main(void)
{
    /* memory stream: growth (Q8: 10 bytes in a 10-byte buffer plus 15 more), read clamps, rewind */
    huf_read_writer_t *m = NULL;
    void *buf = NULL;
    CHECK(huf_memopen(&m, &buf, 10) == 0);
    CHECK(m->write(m->stream, "0123456789", 10) == 0);
    CHECK(m->write(m->stream, "abcdefghijklmno", 15) == 0);
    size_t n = 0, cap = 0;
    CHECK(huf_memlen(m, &n) == 0 && n == 25);
    CHECK(huf_memcap(m, &cap) == 0 && cap >= 25);
    char out[64];
    n = 64;
    CHECK(m->read(m->stream, out, &n) == 0 && n == 25 && !memcmp(out, "0123456789abcdefghijklmno", 25));
    n = 64;
    CHECK(m->read(m->stream, out, &n) == 0 && n == 0);
    CHECK(huf_memrewind(m) == 0 && huf_memlen(m, &n) == 0 && n == 0);

    /* buffered reader/writer on top of it */
    huf_bufio_read_writer_t *bw = NULL;
    CHECK(huf_bufio_read_writer_init(&bw, m, 8) == 0);
    for (int i = 0; i < 100; i++) CHECK(huf_bufio_write_uint8(bw, (uint8_t)i) == 0);
    CHECK(huf_bufio_write(bw, "tail", 4) == 0);
    CHECK(huf_bufio_read_writer_flush(bw) == 0);
    CHECK(huf_memlen(m, &n) == 0 && n == 104);
    CHECK(huf_bufio_read_writer_free(&bw) == 0 && bw == NULL);
    huf_bufio_read_writer_t *br = NULL;
    CHECK(huf_bufio_read_writer_init(&br, m, 16) == 0);
    uint8_t b = 0;
    for (int i = 0; i < 100; i++) CHECK(huf_bufio_read_uint8(br, &b) == 0 && b == (uint8_t)i);
    CHECK(huf_bufio_read(br, out, 4) == 0 && !memcmp(out, "tail", 4));
    CHECK(huf_bufio_read(br, out, 1) == HUF_ERROR_READ_WRITE);
    CHECK(huf_bufio_read_writer_free(&br) == 0);

    /* the codec needs the device: clean failure, nothing leaked */
    huf_read_writer_t *w = NULL;
    void *wbuf = NULL;
    CHECK(huf_memopen(&w, &wbuf, 0) == 0);
    huf_config_t *cfg = NULL;
    CHECK(huf_config_init(&cfg) == 0);
    cfg->length = 104;
    cfg->reader = m;
    cfg->writer = w;
    CHECK(huf_encode(cfg) == HUF_ERROR_FATAL);
    CHECK(huf_decode(cfg) == HUF_ERROR_FATAL);
    CHECK(huf_encode(NULL) == HUF_ERROR_INVALID_ARGUMENT);
    CHECK(huf_config_free(&cfg) == 0);
    CHECK(huf_memclose(&w) == 0 && huf_memclose(&m) == 0);
    free(wbuf);
    free(buf);

    /* fd stream */
    char path[] = "/tmp/huf_san_XXXXXX";
    int fd = mkstemp(path);
    CHECK(fd >= 0);
    huf_read_writer_t *f = NULL;
    CHECK(huf_fdopen(&f, fd) == 0);
    CHECK(f->write(f->stream, "fd stream", 9) == 0);
    CHECK(lseek(fd, 0, SEEK_SET) == 0);
    n = 64;
    CHECK(f->read(f->stream, out, &n) == 0 && n == 9);
    CHECK(huf_fdclose(&f) == 0);
    close(fd);
    unlink(path);

    /* unit-level objects */
    huf_histogram_t *h = NULL;
    CHECK(huf_histogram_init(&h, 1, 512) == 0);
    CHECK(huf_histogram_populate(h, "abracadabra", 11) == 0);
    huf_tree_t *t = NULL;
    CHECK(huf_tree_init(&t) == 0);
    CHECK(huf_tree_from_histogram(t, h) == 0);
    int16_t ser[1024];
    size_t len = 0;
    CHECK(huf_tree_serialize(t, ser, &len) == 0 && len == 21);
    huf_tree_t *t2 = NULL;
    CHECK(huf_tree_init(&t2) == 0);
    CHECK(huf_tree_deserialize(t2, ser, len) == 0);
    uint8_t code[512];
    size_t clen = sizeof(code);
    CHECK(huf_node_to_string(t->leaves['a'], code, &clen) == 0 && clen == 2);
    CHECK(huf_tree_free(&t) == 0 && huf_tree_free(&t2) == 0 && huf_histogram_free(&h) == 0);

    int x = 0;
    CHECK(with_macros(&x, 3) == 0 && x == 3 && with_macros(NULL, 3) == HUF_ERROR_INVALID_ARGUMENT &&
          with_macros(&x, 11) == HUF_ERROR_INVALID_ARGUMENT);
    CHECK(!strcmp(huf_error_string(HUF_ERROR_BTREE_CORRUPTED), huf_error_string(6)));
    printf("host sanitizer driver ok\n");
    return 0;
}
