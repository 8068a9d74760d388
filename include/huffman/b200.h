/*
 * <huffman/b200.h> — the C-ABI shim between the host C layer and the sm_100a CUDA kernels.
 *
 * This is the drop-in boundary for the hot path: plain pointers and sizes, huf_error_t
 * status, no CUDA or torch types in any signature (a cudaStream_t is passed as void*).
 * huf_encode()/huf_decode() in the host layer call exactly these entry points; a reference
 * maintainer replacing the body of the reference's block loops would bind the same symbols
 * (see INTEGRATION.md for the stub).
 *
 * What each entry point replaces in the reference (paths relative to the reference tree):
 *   huf_b200_encode_*   the per-block body of huf_encode, src/encoder.c:288-374
 *                       = huf_histogram_populate (src/histogram.c:73-103)
 *                       + huf_tree_from_histogram (src/tree.c:292-427)
 *                       + __huf_create_char_coding (src/encoder.c:40-81)
 *                       + huf_tree_serialize (src/tree.c:233-289)
 *                       + header emit (src/encoder.c:325-342)
 *                       + __huf_encode_block (src/encoder.c:85-131)
 *   huf_b200_decode_*   the per-block body of huf_decode, src/decoder.c:218-276
 *                       = header parse + bound check (src/decoder.c:220-239)
 *                       + huf_tree_deserialize (src/tree.c:138-227)
 *                       + __huf_decode_block (src/decoder.c:34-96)
 *
 * All `d_*` pointers are DEVICE pointers, 16-byte aligned (cudaMalloc gives 256).
 * Functions ending in _async only enqueue work on `stream`; the matching _finish waits for
 * it and returns the status/size.  One context serves one call at a time.
 */
#ifndef HUFFMAN_B200_SHIM_H
#define HUFFMAN_B200_SHIM_H

#include "../huffman.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct huf_b200_ctx huf_b200_ctx_t;

/* `stream` arguments take a cudaStream_t handle.  NULL is CUDA's default stream (the value
 * torch.cuda.current_stream().cuda_stream has unless a side stream is active);
 * HUF_B200_STREAM_PRIVATE selects the context's own non-blocking stream (what the host layer
 * uses so that huf_encode/huf_decode never serialise with an application's other streams). */
#define HUF_B200_STREAM_PRIVATE ((void *)(intptr_t)-1)

/* Options for huf_b200_ctx_set_option. */
enum {
    /* 0 (default): tree_len > HUF_BTREE_LEN is HUF_ERROR_BTREE_OVERFLOW exactly like the
     * reference decoder (src/decoder.c:237-239).  1: additionally accept the 1025-element
     * tree the reference encoder emits for blocks that contain all 256 byte values
     * (src/encoder.c:270, src/tree.c:264).  Env var HUF_B200_ACCEPT_1025=1 sets the default
     * for contexts created afterwards (this is how huf_decode() is switched). */
    HUF_B200_OPT_ACCEPT_1025 = 1,
    /* 1: bracket every kernel launch of the following *_async calls with CUDA events on the
     * launching stream; read them with huf_b200_kernel_times after *_finish.  Large encode calls
     * then run their passes one after the other so that the times add up; value 2 keeps the
     * pass pipeline on and reports every launch as "name@start_ms duration_ms". */
    HUF_B200_OPT_KERNEL_TIMING = 2,
    /* 1: launch the instance of the fast decode kernel that does not rely on its lookup table
     * being 8 KB aligned in the shared window (what the library falls back to by itself when
     * its probe finds the table elsewhere).  For tests. */
    HUF_B200_OPT_FORCE_LUT_ADD = 3,
    /* 1: large encode calls run their passes one after the other on the caller's stream instead
     * of as a pipeline over the context's side streams (env HUF_B200_NO_OVERLAP=1 sets the
     * default).  Results are identical; for measurements and tests. */
    HUF_B200_OPT_NO_OVERLAP = 4,
};

/* Create a context on CUDA device `device` (< 0: the current device).  Fails with
 * HUF_ERROR_FATAL when no usable sm_100 device/driver is present: there is no CPU path. */
huf_error_t huf_b200_ctx_create(huf_b200_ctx_t **ctx, int device);
huf_error_t huf_b200_ctx_destroy(huf_b200_ctx_t **ctx);
huf_error_t huf_b200_ctx_set_option(huf_b200_ctx_t *ctx, int option, int64_t value);

/* Upper bound of the encoded size of `length` bytes at `blocksize` (0 => one block). */
uint64_t huf_b200_encode_bound(uint64_t length, uint64_t blocksize);

/* Number of blocks huf_encode produces for (length, blocksize). */
uint64_t huf_b200_block_count(uint64_t length, uint64_t blocksize);

/* Encode d_in[0..length) into d_out (capacity out_capacity bytes, >= encode_bound).
 * Kernels: K1 segment histogram, K2 per-block code build, block-offset scan, K3 pack. */
huf_error_t huf_b200_encode_async(huf_b200_ctx_t *ctx, const void *d_in, uint64_t length,
                                  uint64_t blocksize, void *d_out, uint64_t out_capacity,
                                  void *stream);
/* Wait; *out_len = bytes written to d_out. */
huf_error_t huf_b200_encode_finish(huf_b200_ctx_t *ctx, uint64_t *out_len);

/* After encode_finish: device pointer to u64[nblocks + 1] byte offsets of every block in
 * d_out (exclusive scan of the compressed block sizes; last entry = total).  This is the
 * array the host concatenation of per-GPU slabs uses; it never alters the stream bytes. */
huf_error_t huf_b200_encode_block_offsets(huf_b200_ctx_t *ctx, const uint64_t **d_offsets,
                                          uint64_t *nblocks);

/* Decode.  d_in holds `avail` readable compressed bytes; blocks are started while the
 * consumed byte count is < `length` (the reference's loop condition, src/decoder.c:218).
 * Kernels: K4 header-candidate scan + chain/offset scan, K5 table decode. */
huf_error_t huf_b200_decode_async(huf_b200_ctx_t *ctx, const void *d_in, uint64_t avail,
                                  uint64_t length, void *d_out, uint64_t out_capacity,
                                  void *stream);
/* Optional block index for the NEXT huf_b200_decode_async call on this context: d_offsets is a
 * device array of `nblocks` ascending byte offsets of block headers in d_in (entry 0 = 0), e.g.
 * what huf_b200_encode_block_offsets returned for this very stream.  With it the decoder skips
 * the speculative header scan (the stream format has no index, src/encoder.c:325-342).  The
 * index never changes results: every block is still proven by the chain check, and a wrong
 * index only costs a rescan.  The array must stay valid until decode_finish. */
huf_error_t huf_b200_decode_hint_offsets(huf_b200_ctx_t *ctx, const uint64_t *d_offsets,
                                         uint64_t nblocks);
/* Wait; *out_len = decoded bytes valid in d_out, *consumed = compressed bytes consumed.
 * Returns the reference's error code for the first failing block (earlier blocks' output is
 * valid), HUF_ERROR_READ_WRITE when a block runs past `avail`. */
huf_error_t huf_b200_decode_finish(huf_b200_ctx_t *ctx, uint64_t *out_len,
                                   uint64_t *consumed);

/* Plan only: total decoded size and block count of the stream (runs K4, synchronises).
 * Used by the host layer to size the output buffer.  `status` receives the error the
 * decode would stop with (sizes then cover the blocks before it). */
huf_error_t huf_b200_decode_plan(huf_b200_ctx_t *ctx, const void *d_in, uint64_t avail,
                                 uint64_t length, uint64_t *out_len, uint64_t *nblocks,
                                 void *stream);

/* Like huf_b200_decode_async, but the pass starts at byte offset `first` of d_in, which the
 * caller has proven to be a block start (the end of the previous block); output begins at
 * d_out[0].  This is how the host lane decodes a stream span by span. */
huf_error_t huf_b200_decode_async_at(huf_b200_ctx_t *ctx, const void *d_in, uint64_t avail,
                                     uint64_t length, uint64_t first, void *d_out,
                                     uint64_t out_capacity, void *stream);

/* Byte-range decode, the unit of a multi-GPU decode of ONE stream (the
 * reference walks the stream serially, src/decoder.c:218-276): decode every block that STARTS in
 * [start, stop) of d_in.  `start` need not be a block start -- the first block is the first
 * header candidate at or behind it, unless start_is_block says that `start` is one (the
 * stream start, or a chain end the caller has proven) --
 * and the last block may reach past `stop` (the caller supplies that overlap in `avail`).
 * _finish returns where the first block was found (*first, ~0 when no block starts in the
 * range), where the chain of decoded blocks ends (*end) and the decoded bytes (in d_out from
 * offset 0).  Neighbouring ranges are stitched by the caller: range g is consistent when
 * end(g-1) == first(g); the decoded slabs concatenate in range order. */
huf_error_t huf_b200_decode_range_async(huf_b200_ctx_t *ctx, const void *d_in, uint64_t avail,
                                        uint64_t start, uint64_t stop, int start_is_block,
                                        void *d_out, uint64_t out_capacity, void *stream);
huf_error_t huf_b200_decode_range_finish(huf_b200_ctx_t *ctx, uint64_t *first, uint64_t *end,
                                         uint64_t *out_len);

/* ---- host-buffer lanes ----------------------------------------------------------------------
 * What huf_encode / huf_decode run below their stream objects (reference src/io.c:9-226,
 * src/bufio.c:149-287): the bytes come from a source and go to a sink in HOST memory, and the
 * library overlaps source -> pinned -> HBM, the kernels, and HBM -> pinned -> sink over spans of
 * whole blocks (three CUDA streams, a pool of copy threads).  A source is either contiguous
 * memory (`data`/`size`, any host memory) or a pull callback that fills the pinned buffer it is
 * handed; a sink either lends memory (`reserve` returns where the next `count` bytes go, `commit`
 * accounts for them) or takes pushes. */
typedef struct huf_b200_source {
    const void *data;   /* contiguous readable bytes, or NULL: use pull */
    uint64_t size;      /* bytes readable at data */
    /* Fill dst with up to `want` of the next bytes; *got < want means end of data. */
    huf_error_t (*pull)(void *arg, void *dst, uint64_t want, uint64_t *got);
    void *arg;
} huf_b200_source_t;

typedef struct huf_b200_sink {
    /* Make room for `count` more bytes and return where they go in *dst (NULL: use push). */
    huf_error_t (*reserve)(void *arg, uint64_t count, void **dst);
    huf_error_t (*commit)(void *arg, uint64_t count);
    huf_error_t (*push)(void *arg, const void *src, uint64_t count);
    void *arg;
    /* Optional, lending sinks only: where the next byte goes and how many bytes fit there
     * without the buffer moving.  When that range is page-locked (huf_b200_host_register) the
     * lanes copy device -> sink directly into it and only call commit. */
    huf_error_t (*room)(void *arg, void **dst, uint64_t *avail);
} huf_b200_sink_t;

/* huf_encode below the streams: `length` bytes of `src` in blocks of `blocksize` (0 => one
 * block).  A source that ends early gives HUF_ERROR_READ_WRITE after the complete blocks were
 * delivered (src/encoder.c:296-299).  *consumed = source bytes taken. */
huf_error_t huf_b200_encode_host(huf_b200_ctx_t *ctx, const huf_b200_source_t *src, uint64_t length,
                                 uint64_t blocksize, const huf_b200_sink_t *dst, uint64_t *consumed);

/* huf_decode below the streams: blocks are started while consumed < `length`
 * (src/decoder.c:218); a contiguous source lends all of its bytes (the last block may reach
 * past `length`), a pull source is asked for `length` bytes and for more only when a block
 * needs them.  Output decoded before a failing block is still delivered.  *consumed =
 * compressed bytes of the whole blocks decoded. */
huf_error_t huf_b200_decode_host(huf_b200_ctx_t *ctx, const huf_b200_source_t *src, uint64_t length,
                                 const huf_b200_sink_t *dst, uint64_t *consumed);

/* Several GPUs in one call.  ctxs[0..ndev) are contexts on different devices.
 * Encode: device g takes the contiguous block range [g*B/ndev, (g+1)*B/ndev) of the input
 * (blocks are independent, src/encoder.c:345,360-373), the slabs are placed in `dst` at the
 * exclusive scan of their sizes: the result is byte-identical to the one-device stream.
 * Decode: device g scans the byte range [g*C/ndev, (g+1)*C/ndev) of the ONE stream (plus an
 * overlap for its last block) and decodes the blocks that start in it; the host validates the
 * chain across the ranges (end of range g-1 == first block of range g) and concatenates the
 * decoded slabs; a seam that does not validate falls back to the one-device lane, so results
 * never depend on the split.  No device-to-device traffic, no collective.  `dst` must be a
 * lending sink (reserve/commit).  slab_sizes (optional) receives the ndev slab sizes. */
huf_error_t huf_b200_encode_host_multi(huf_b200_ctx_t *const *ctxs, int ndev, const void *h_in,
                                       uint64_t length, uint64_t blocksize,
                                       const huf_b200_sink_t *dst, uint64_t *slab_sizes);
huf_error_t huf_b200_decode_host_multi(huf_b200_ctx_t *const *ctxs, int ndev, const void *h_in,
                                       uint64_t avail, uint64_t length, const huf_b200_sink_t *dst,
                                       uint64_t *consumed);
/* Plan of a byte-range decode (see huf_b200_decode_range_async): decoded size of the blocks
 * whose headers the scan finds in [start, stop). */
huf_error_t huf_b200_decode_range_plan(huf_b200_ctx_t *ctx, const void *d_in, uint64_t avail,
                                       uint64_t start, uint64_t stop, int start_is_block,
                                       uint64_t *out_len, uint64_t *nblocks, void *stream);

/* Counters for benches/tests: kernels launched by the last *_async call. */
uint64_t huf_b200_last_launch_count(const huf_b200_ctx_t *ctx);
/* After decode_finish: candidate blocks of the last pass that the fast decode lane declined
 * and the general lane decoded (foreign tree shapes, code words beyond the table reach,
 * corrupt or truncated blocks, speculative header candidates that are no blocks). */
uint64_t huf_b200_last_slow_blocks(const huf_b200_ctx_t *ctx);

/* With HUF_B200_OPT_KERNEL_TIMING on: writes one "kernel_name milliseconds\n" line per launch
 * of the last call into buf (NUL terminated). */
huf_error_t huf_b200_kernel_times(huf_b200_ctx_t *ctx, char *buf, uint64_t buflen);

/* Raw device memory helpers so that non-CUDA hosts (C, ctypes, cgo) can stage buffers
 * without linking the CUDA runtime themselves. */
huf_error_t huf_b200_dev_alloc(void **d_ptr, uint64_t bytes);
huf_error_t huf_b200_dev_free(void *d_ptr);
huf_error_t huf_b200_copy_h2d(void *d_dst, const void *h_src, uint64_t bytes);
huf_error_t huf_b200_copy_d2h(void *h_dst, const void *d_src, uint64_t bytes);
/* Page-lock / release caller memory for direct DMA (cudaHostRegister).  The host lanes use a
 * contiguous source or a lending sink in place -- no bounce copy through their own pinned
 * buffers -- when its bytes are page-locked.  Registering costs ~0.3 s per GiB (and releasing
 * about as much), so it only pays for buffers that live through many calls: huf_memopen streams
 * are registered by the library from their HUF_B200_PIN_AFTER-th codec call on (default 4;
 * 0 = never) and released when they grow or are closed. */
huf_error_t huf_b200_host_register(void *ptr, uint64_t bytes);
huf_error_t huf_b200_host_unregister(void *ptr);
/* Counter for benches/tests: spans and results the host lanes copied straight from / into
 * page-locked caller memory since the library was loaded. */
uint64_t huf_b200_direct_copy_count(void);
/* Number of visible CUDA devices (0 when the driver/GPU is missing). */
int huf_b200_device_count(void);

#ifdef __cplusplus
}
#endif

#endif /* HUFFMAN_B200_SHIM_H */
