/*
 * huffman.h — public C API of the B200-native libhuffman block codec.
 *
 * Source-compatible with ybubnov/libhuffman 1.0.3: same type names, field order and
 * sizes (LP64), same function names, argument meaning and huf_error_t values, so that
 * code written against the reference's <huffman.h> / <huffman/...h> recompiles unchanged
 * and the cffi-based `huffmanfile` package binds to this library unchanged.
 *
 * Only huf_encode()/huf_decode() run on the GPU (sm_100a kernels behind the C-ABI shim
 * declared in <huffman/b200.h>); there is no CPU fallback for them.  The small
 * histogram/tree/symbol/bufio objects below are host-side API objects kept for link
 * compatibility with the reference's unit tests and the cffi cdef; the codec does not
 * route through them.
 *
 * Every declaration cites the reference declaration it replaces as
 * [ref: file:line] relative to the reference tree.
 *
 * The text between `#define CFFI_...` / `#undef CFFI_...` fences is valid cffi cdef input
 * (no preprocessor lines inside), mirroring how the reference's setup_ffi.py:8-23 scrapes
 * its headers.
 */
#ifndef HUFFMAN_B200_PUBLIC_API_H
#define HUFFMAN_B200_PUBLIC_API_H

#include <stdint.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Buffer size constants [ref: include/huffman/common.h:7-17]. */
#define HUF_1KIB_BUFFER   1024
#define HUF_64KIB_BUFFER  65536
#define HUF_128KIB_BUFFER 131072
#define HUF_256KIB_BUFFER 262144
#define HUF_512KIB_BUFFER 524288
#define HUF_1MIB_BUFFER   1048576

/* Tree constants [ref: include/huffman/tree.h:8-18]. */
#define HUF_ASCII_COUNT   256   /* alphabet size */
#define HUF_BTREE_LEN     1024  /* max int16 elements the decoder accepts per tree */
#define HUF_HISTOGRAM_LEN 512   /* 256 leaves + 256 internal-node weight slots */
#define HUF_LEAF_NODE     -1    /* "no child" marker in the serialised tree */

#define CFFI_huffman_b200_api

/* ---- errors [ref: include/huffman/errors.h:6-31] ---------------------------------------- */

typedef enum {
    HUF_ERROR_SUCCESS,            /* 0 */
    HUF_ERROR_MEMORY_ALLOCATION,  /* 1: host or device allocation failed */
    HUF_ERROR_INVALID_ARGUMENT,   /* 2: NULL where an object is required */
    HUF_ERROR_READ_WRITE,         /* 3: stream error or short read */
    HUF_ERROR_FATAL,              /* 4: unrecoverable (CUDA runtime/driver failure) */
    HUF_ERROR_BTREE_OVERFLOW,     /* 5: serialised tree length < 0 or > HUF_BTREE_LEN */
    HUF_ERROR_BTREE_CORRUPTED,    /* 6: bit stream walks off the tree */
} huf_error_t;

const char* huf_error_string(huf_error_t error);

/* ---- streams [ref: include/huffman/io.h:11-32] ------------------------------------------ */

typedef struct __huf_read_writer {
    void *stream;
    huf_error_t (*write)(void *stream, const void *buf, size_t count);
    /* On return *count holds the bytes delivered; fewer than asked (or 0) at end of data. */
    huf_error_t (*read)(void *stream, void *buf, size_t *count);
} huf_read_writer_t;

/* Growable in-memory stream.  *buf is calloc'd here, may be replaced on growth, and is
 * owned by the caller after huf_memclose (free() it). */
huf_error_t huf_memopen(huf_read_writer_t **self, void **buf, size_t capacity);
huf_error_t huf_memlen(const huf_read_writer_t *self, size_t *len);
huf_error_t huf_memcap(const huf_read_writer_t *self, size_t *cap);
huf_error_t huf_memrewind(huf_read_writer_t *self);
huf_error_t huf_memclose(huf_read_writer_t **self);

/* File-descriptor stream over read(2)/write(2); the descriptor stays owned by the caller. */
huf_error_t huf_fdopen(huf_read_writer_t **self, int fd);
huf_error_t huf_fdclose(huf_read_writer_t **self);

/* ---- codec configuration [ref: include/huffman/config.h:10-45] -------------------------- */

typedef struct __huf_encoder_config {
    uint64_t length;            /* encode: bytes to pull; decode: compressed bytes to consume */
    uint64_t blocksize;         /* encode: block size, 0 => one block of `length`; decode: ignored */
    size_t reader_buffer_size;  /* bufio capacity hints; 0 = unbuffered */
    size_t writer_buffer_size;
    huf_read_writer_t *reader;
    huf_read_writer_t *writer;
} huf_config_t;

huf_error_t huf_config_init(huf_config_t **self);
huf_error_t huf_config_free(huf_config_t **self);

/* ---- codec entry points: THE GPU HOT PATH ------------------------------------------------
 * [ref: include/huffman/encoder.h:10-26, include/huffman/decoder.h:10-26] */

typedef struct __huf_encoder huf_encoder_t;
typedef struct __huf_decoder huf_decoder_t;

huf_error_t huf_encoder_init(huf_encoder_t **self, const huf_config_t *config);
huf_error_t huf_encoder_free(huf_encoder_t **self);
huf_error_t huf_encode(const huf_config_t *config);

huf_error_t huf_decoder_init(huf_decoder_t **self, const huf_config_t *config);
huf_error_t huf_decoder_free(huf_decoder_t **self);
huf_error_t huf_decode(const huf_config_t *config);

/* ---- buffered and bit-level I/O objects [ref: include/huffman/bufio.h:11-97] ------------ */

typedef struct __huf_bufio_read_writer {
    uint8_t *bytes;
    size_t offset;
    size_t capacity;
    size_t length;
    uint64_t have_been_processed;
    huf_read_writer_t *read_writer;
} huf_bufio_read_writer_t;

typedef struct __huf_bit_read_writer {
    uint8_t bits;
    uint8_t offset;
} huf_bit_read_writer_t;

void huf_bit_write(huf_bit_read_writer_t *self, uint8_t bit);
void huf_bit_read_writer_reset(huf_bit_read_writer_t *self);

huf_error_t huf_bufio_read_writer_init(huf_bufio_read_writer_t **self, huf_read_writer_t *read_writer, size_t size);
huf_error_t huf_bufio_read_writer_free(huf_bufio_read_writer_t **self);
huf_error_t huf_bufio_read_writer_flush(huf_bufio_read_writer_t *self);
huf_error_t huf_bufio_write(huf_bufio_read_writer_t *self, const void *buf, size_t size);
huf_error_t huf_bufio_read(huf_bufio_read_writer_t *self, void *buf, size_t size);
huf_error_t huf_bufio_read_uint8(huf_bufio_read_writer_t *self, uint8_t *byte);
huf_error_t huf_bufio_write_uint8(huf_bufio_read_writer_t *self, uint8_t byte);

/* ---- histogram object [ref: include/huffman/histogram.h:10-49] -------------------------- */

typedef struct __huf_histogram {
    uint64_t *frequencies;
    size_t iota;    /* element width in bytes (1..8) */
    size_t length;  /* number of counters */
    size_t start;   /* smallest element seen, (size_t)-1 when empty */
} huf_histogram_t;

huf_error_t huf_histogram_init(huf_histogram_t **self, size_t iota, size_t length);
huf_error_t huf_histogram_free(huf_histogram_t **self);
huf_error_t huf_histogram_reset(huf_histogram_t *self);
huf_error_t huf_histogram_populate(huf_histogram_t *self, void *buf, size_t len);

/* ---- allocation helper [ref: include/huffman/malloc.h:10-11] ---------------------------- */

huf_error_t huf_malloc(void** ptr, size_t size, size_t num);

/* ---- symbol table object [ref: include/huffman/symbol.h:10-79] -------------------------- */

typedef struct __huf_symbol_mapping_element {
    size_t length;
    uint8_t *coding;  /* '0'/'1' characters, leaf -> root order, NUL terminated */
} huf_symbol_mapping_element_t;

typedef struct __huf_symbol_mapping {
    size_t length;
    huf_symbol_mapping_element_t **symbols;
} huf_symbol_mapping_t;

huf_error_t huf_symbol_mapping_element_init(huf_symbol_mapping_element_t **self, const uint8_t *coding, size_t length);
huf_error_t huf_symbol_mapping_element_free(huf_symbol_mapping_element_t **self);
huf_error_t huf_symbol_mapping_init(huf_symbol_mapping_t **self, size_t length);
huf_error_t huf_symbol_mapping_free(huf_symbol_mapping_t **self);
huf_error_t huf_symbol_mapping_insert(huf_symbol_mapping_t *self, size_t position, huf_symbol_mapping_element_t *element);
huf_error_t huf_symbol_mapping_get(huf_symbol_mapping_t *self, size_t position, huf_symbol_mapping_element_t **element);
huf_error_t huf_symbol_mapping_reset(huf_symbol_mapping_t *self);

/* ---- tree object [ref: include/huffman/tree.h:24-82] ------------------------------------ */

typedef struct __huf_node {
    int16_t index;  /* byte value for leaves, 256.. for merge nodes */
    struct __huf_node *parent;
    struct __huf_node *left;
    struct __huf_node *right;
} huf_node_t;

typedef struct __huf_tree {
    huf_node_t **leaves;  /* HUF_HISTOGRAM_LEN slots */
    huf_node_t *root;
} huf_tree_t;

huf_error_t huf_node_to_string(const huf_node_t *self, uint8_t *buf, size_t *len);
huf_error_t huf_tree_init(huf_tree_t **self);
huf_error_t huf_tree_free(huf_tree_t **self);
huf_error_t huf_tree_reset(huf_tree_t *self);
huf_error_t huf_tree_deserialize(huf_tree_t *self, const int16_t *buf, size_t len);
huf_error_t huf_tree_serialize(huf_tree_t *self, int16_t *buf, size_t *len);
huf_error_t huf_tree_from_histogram(huf_tree_t *self, huf_histogram_t *histogram);

#undef CFFI_huffman_b200_api

#ifdef __cplusplus
}
#endif

#endif /* HUFFMAN_B200_PUBLIC_API_H */
