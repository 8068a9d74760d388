/* TEST INFRASTRUCTURE: stands in for the CUDA shim when the host C layer is linked alone for the
 * ASan/UBSan lane (tests/test_host_sanitizers.py).  Every device entry point fails the way the
 * real shim does without a GPU; the unit-level objects under test never reach them. */
#include <huffman/b200.h>

huf_error_t huf_b200_ctx_create(huf_b200_ctx_t **ctx, int device) { (void)device; if (ctx) *ctx = 0; return HUF_ERROR_FATAL; }
huf_error_t huf_b200_ctx_destroy(huf_b200_ctx_t **ctx) { if (ctx) *ctx = 0; return HUF_ERROR_SUCCESS; }
int huf_b200_device_count(void) { return 0; }
uint64_t huf_b200_encode_bound(uint64_t length, uint64_t blocksize) { (void)blocksize; return length * 2 + 4096; }
huf_error_t huf_b200_encode_host(huf_b200_ctx_t *c, const huf_b200_source_t *s, uint64_t l, uint64_t b,
                                 const huf_b200_sink_t *d, uint64_t *n)
{ (void)c; (void)s; (void)l; (void)b; (void)d; (void)n; return HUF_ERROR_FATAL; }
huf_error_t huf_b200_decode_host(huf_b200_ctx_t *c, const huf_b200_source_t *s, uint64_t l,
                                 const huf_b200_sink_t *d, uint64_t *n)
{ (void)c; (void)s; (void)l; (void)d; (void)n; return HUF_ERROR_FATAL; }
huf_error_t huf_b200_encode_host_multi(huf_b200_ctx_t *const *c, int n, const void *h, uint64_t l, uint64_t b,
                                       const huf_b200_sink_t *d, uint64_t *s)
{ (void)c; (void)n; (void)h; (void)l; (void)b; (void)d; (void)s; return HUF_ERROR_FATAL; }
huf_error_t huf_b200_decode_host_multi(huf_b200_ctx_t *const *c, int n, const void *h, uint64_t a, uint64_t l,
                                       const huf_b200_sink_t *d, uint64_t *u)
{ (void)c; (void)n; (void)h; (void)a; (void)l; (void)d; (void)u; return HUF_ERROR_FATAL; }
huf_error_t huf_b200_host_register(void *p, uint64_t n) { (void)p; (void)n; return HUF_ERROR_FATAL; }
huf_error_t huf_b200_host_unregister(void *p) { (void)p; return HUF_ERROR_SUCCESS; }
