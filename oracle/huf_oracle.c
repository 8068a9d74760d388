/*
 * huf_oracle.c — CPU restatement of libhuffman's block codec.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the B200 path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product library
 * (libhuffman_b200.so) never links, loads or calls anything in this directory.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle.py) against
 *   - the reference's own golden vectors (test/encode_test.c:35, test/decode_test.c:32-74,
 *     test/tree_test.c:25-31) and the survey's known-answer vectors, committed under
 *     tests/golden/, and
 *   - the unmodified reference compiled from /root/reference/src by oracle/Makefile into
 *     oracle/_ref/libhuffman_ref.so, on randomized inputs (ties, 1..256 symbols, multi-block).
 *
 * What is restated (reference file:line, relative to /root/reference):
 *   block loop / header layout          src/encoder.c:288-374, src/decoder.c:218-276
 *   byte histogram                      src/histogram.c:73-103
 *   min-pair merge + tie-break + unary  src/tree.c:292-427
 *   leaf->root path == code word        src/tree.c:12-47, src/encoder.c:40-81,106-108
 *   pre-order tree (de)serialisation    src/tree.c:138-289
 *   MSB-first bit packing, zero pad     src/bufio.c:18-32, src/encoder.c:85-131
 *   bit-walk decode + error codes       src/decoder.c:34-96,237-239
 *
 * The code below is written from the behavioural spec, not from the C text:
 * flat index arrays instead of pointer-linked nodes, no streams, buffer in / buffer out.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* huf_error_t values, include/huffman/errors.h:6-27 of the reference. */
enum {
    ORC_OK = 0,
    ORC_ENOMEM = 1,
    ORC_EINVAL = 2,
    ORC_EIO = 3,
    ORC_EFATAL = 4,
    ORC_EOVERFLOW = 5,
    ORC_ECORRUPT = 6,
};

#define NSYM 256
#define NSLOT 512
#define NIL (-1)

typedef struct {
    int left[NSLOT];   /* child indices, NIL when absent */
    int right[NSLOT];
    int parent[NSLOT];
    int root;          /* index of the unary root */
    int nsym;          /* distinct symbols */
} orc_tree_t;

/*
 * Merge rule (src/tree.c:318-414): among live slots j < next with w[j] != 0 order by
 * (weight ascending, index DESCENDING); first -> left child, second -> right child of a new
 * slot `next`; when only one live slot is left it is wrapped by a unary node which is the root.
 */
static void
orc_build_tree(const uint64_t freq[NSYM], orc_tree_t *t)
{
    uint64_t w[NSLOT];
    int next = NSYM;

    for (int i = 0; i < NSLOT; i++) {
        w[i] = i < NSYM ? freq[i] : 0;
        t->left[i] = t->right[i] = t->parent[i] = NIL;
    }
    t->nsym = 0;
    for (int i = 0; i < NSYM; i++) {
        t->nsym += freq[i] != 0;
    }
    t->root = NIL;
    if (t->nsym == 0) {
        return;
    }

    for (;;) {
        int a = NIL, b = NIL;
        /* a = best, b = second best under (weight asc, index desc). */
        for (int j = 0; j < next; j++) {
            if (!w[j]) {
                continue;
            }
            if (a == NIL || w[j] <= w[a]) {
                b = a;
                a = j;
            } else if (b == NIL || w[j] <= w[b]) {
                b = j;
            }
        }
        t->left[next] = a;
        t->parent[a] = next;
        w[next] = w[a];
        w[a] = 0;
        if (b != NIL) {
            t->right[next] = b;
            t->parent[b] = next;
            w[next] += w[b];
            w[b] = 0;
        }
        next++;
        if (b == NIL) {
            t->root = next - 1;
            return;
        }
    }
}

/* Pre-order dump: index, left subtree, right subtree; absent child -> -1 (src/tree.c:233-270). */
static int
orc_serialize(const orc_tree_t *t, int16_t *out)
{
    int stack[2 * NSLOT + 4];
    int sp = 0, n = 0;

    stack[sp++] = t->root;
    while (sp) {
        int v = stack[--sp];
        if (v == NIL) {
            out[n++] = -1;
            continue;
        }
        out[n++] = (int16_t)v;
        stack[sp++] = t->right[v];
        stack[sp++] = t->left[v];
    }
    return n;
}

/* Bits of symbol s, root -> leaf, one byte per bit; returns the length. */
static int
orc_code_bits(const orc_tree_t *t, int s, uint8_t *bits)
{
    uint8_t rev[NSLOT];
    int n = 0;

    for (int v = s; t->parent[v] != NIL; v = t->parent[v]) {
        rev[n++] = t->right[t->parent[v]] == v;
    }
    for (int i = 0; i < n; i++) {
        bits[i] = rev[n - 1 - i];
    }
    return n;
}

/* ---- exported helpers (ctypes-friendly, plain pointers) -------------------------------- */

/*
 * Code book of one block: freq[256] -> len[256] (0 = absent), codes as '0'/'1' strings in a
 * 256 x 512 char matrix (NUL terminated), serialised tree and its element count.
 */
int
huf_oracle_codebook(const uint64_t *freq, uint16_t *len, char *codes, int16_t *tree,
                    int *tree_len)
{
    orc_tree_t t;
    uint8_t bits[NSLOT];

    orc_build_tree(freq, &t);
    memset(len, 0, NSYM * sizeof(*len));
    memset(codes, 0, NSYM * NSLOT);
    *tree_len = 0;
    if (t.root == NIL) {
        return ORC_OK;
    }
    for (int s = 0; s < NSYM; s++) {
        if (!freq[s]) {
            continue;
        }
        int n = orc_code_bits(&t, s, bits);
        len[s] = (uint16_t)n;
        for (int i = 0; i < n; i++) {
            codes[s * NSLOT + i] = (char)('0' + bits[i]);
        }
    }
    *tree_len = orc_serialize(&t, tree);
    return ORC_OK;
}

/* Worst case stream size for `length` bytes at `blocksize` (0 => one block). */
uint64_t
huf_oracle_encode_bound(uint64_t length, uint64_t blocksize)
{
    if (!length) {
        return 0;
    }
    if (!blocksize) {
        blocksize = length;
    }
    uint64_t nblocks = (length + blocksize - 1) / blocksize;
    /* header 10 + 2*1025, payload < 10 bits per symbol for 256 symbols; be generous. */
    return nblocks * (10 + 2 * 1025 + 8) + length * 2 + (length / 8);
}

typedef struct {
    uint8_t *p;
    uint64_t cap, n;
    uint32_t acc;
    int fill; /* bits already in acc (0..7) */
} orc_bitw_t;

static inline int
orc_put_byte(orc_bitw_t *w, uint8_t b)
{
    if (w->n >= w->cap) {
        return 0;
    }
    w->p[w->n++] = b;
    return 1;
}

/*
 * huf_encode restated: stream of [u64 orig_len | i16 tree_len | i16 tree[] | payload].
 * Q1: a 256-symbol block serialises 1025 elements; the reference emits them all
 * and the 1025th is -1, which the pre-order dump produces naturally.
 */
int
huf_oracle_encode(const uint8_t *in, uint64_t length, uint64_t blocksize, uint8_t *out,
                  uint64_t out_cap, uint64_t *out_len)
{
    orc_bitw_t w = { out, out_cap, 0, 0, 0 };
    orc_tree_t t;
    int16_t tree[2 * NSLOT + 4];
    uint8_t bits[NSLOT];
    /* per symbol: length and the bits, MSB-first packed into up to 8 u64 words. */
    static const int WORDS = NSLOT / 64;
    uint64_t code[NSYM][NSLOT / 64];
    int clen[NSYM];

    if (!in && length) {
        return ORC_EINVAL;
    }
    if (!blocksize) {
        blocksize = length;
    }
    for (uint64_t off = 0; off < length; off += blocksize) {
        uint64_t need = length - off < blocksize ? length - off : blocksize;
        const uint8_t *src = in + off;
        uint64_t freq[NSYM] = { 0 };

        for (uint64_t i = 0; i < need; i++) {
            freq[src[i]]++;
        }
        orc_build_tree(freq, &t);
        for (int s = 0; s < NSYM; s++) {
            clen[s] = 0;
            if (!freq[s]) {
                continue;
            }
            clen[s] = orc_code_bits(&t, s, bits);
            memset(code[s], 0, sizeof(code[s]));
            for (int i = 0; i < clen[s]; i++) {
                if (bits[i]) {
                    code[s][i / 64] |= 1ull << (63 - i % 64);
                }
            }
        }
        int tl = orc_serialize(&t, tree);
        int16_t tl16 = (int16_t)tl;

        /* header */
        if (w.n + 10 + 2ull * tl > w.cap) {
            return ORC_EIO;
        }
        memcpy(w.p + w.n, &need, 8);
        memcpy(w.p + w.n + 8, &tl16, 2);
        memcpy(w.p + w.n + 10, tree, 2ull * tl);
        w.n += 10 + 2ull * tl;

        /* payload: MSB first, byte flushed when full, zero padded at block end. */
        w.acc = 0;
        w.fill = 0;
        for (uint64_t i = 0; i < need; i++) {
            int s = src[i], n = clen[s];
            for (int k = 0; k < n; k++) {
                uint32_t bit = (uint32_t)((code[s][k / 64] >> (63 - k % 64)) & 1);
                w.acc = (w.acc << 1) | bit;
                if (++w.fill == 8) {
                    if (!orc_put_byte(&w, (uint8_t)w.acc)) {
                        return ORC_EIO;
                    }
                    w.acc = 0;
                    w.fill = 0;
                }
            }
        }
        if (w.fill) {
            if (!orc_put_byte(&w, (uint8_t)(w.acc << (8 - w.fill)))) {
                return ORC_EIO;
            }
        }
        (void)WORDS;
    }
    *out_len = w.n;
    return ORC_OK;
}

/*
 * Decoder grammar (src/tree.c:138-208): T := -1 | v T T, read pre-order from at most n
 * elements; when elements run out the remaining children are absent; trailing elements are
 * ignored.  Any int16 may label a node; a node without children is a leaf and emits
 * (uint8_t)label (src/decoder.c:78).
 */
typedef struct {
    int16_t label[1026];
    int left[1026], right[1026];
    int n;
} orc_dtree_t;

static int
orc_parse(const int16_t *buf, int n, int *pos, orc_dtree_t *t)
{
    if (*pos >= n) {
        return NIL;
    }
    int16_t v = buf[(*pos)++];
    if (v == -1) {
        return NIL;
    }
    int id = t->n++;
    t->label[id] = v;
    t->left[id] = orc_parse(buf, n, pos, t);
    t->right[id] = orc_parse(buf, n, pos, t);
    return id;
}

/*
 * huf_decode restated.  `avail` = bytes the reader can deliver, `length` = config.length
 * (compressed bytes to consume; checked only between blocks, src/decoder.c:218).
 * `accept_1025` = 0 is reference behaviour (tree_len > 1024 -> BTREE_OVERFLOW, Q2);
 * 1 additionally accepts the 1025-element tree the reference encoder emits for 256-symbol
 * blocks (Q1).  Deliberate deviation Q4: an absent root with orig_len > 0 returns
 * BTREE_CORRUPTED where the reference dereferences NULL.
 */
int
huf_oracle_decode(const uint8_t *in, uint64_t avail, uint64_t length, uint8_t *out,
                  uint64_t out_cap, uint64_t *out_len, uint64_t *consumed, int accept_1025)
{
    uint64_t pos = 0, produced = 0;
    int rc = ORC_OK;
    orc_dtree_t *t = malloc(sizeof(*t));
    int16_t *tree = malloc(sizeof(int16_t) * 1026);

    if (!t || !tree) {
        free(t);
        free(tree);
        return ORC_ENOMEM;
    }
    while (length > pos) {
        uint64_t orig_len;
        int16_t tl;

        if (avail - pos < 8) {
            rc = ORC_EIO;
            break;
        }
        memcpy(&orig_len, in + pos, 8);
        pos += 8;
        if (avail - pos < 2) {
            rc = ORC_EIO;
            break;
        }
        memcpy(&tl, in + pos, 2);
        pos += 2;
        if (tl < 0 || tl > (accept_1025 ? 1025 : 1024)) {
            rc = ORC_EOVERFLOW;
            break;
        }
        if (avail - pos < 2ull * (uint64_t)tl) {
            rc = ORC_EIO;
            break;
        }
        memcpy(tree, in + pos, 2ull * (uint64_t)tl);
        pos += 2ull * (uint64_t)tl;

        int p = 0;
        t->n = 0;
        int root = orc_parse(tree, tl, &p, t);
        int node = root;
        uint64_t restored = 0;

        while (restored < orig_len && rc == ORC_OK) {
            if (pos >= avail) {
                rc = ORC_EIO;
                break;
            }
            uint8_t byte = in[pos++];
            for (int k = 7; k >= 0; k--) {
                if (node == NIL) { /* Q4: absent root */
                    rc = ORC_ECORRUPT;
                    break;
                }
                node = (byte >> k) & 1 ? t->right[node] : t->left[node];
                if (node == NIL) {
                    rc = ORC_ECORRUPT;
                    break;
                }
                if (t->left[node] != NIL || t->right[node] != NIL) {
                    continue;
                }
                if (produced >= out_cap) {
                    rc = ORC_EIO;
                    break;
                }
                out[produced++] = (uint8_t)t->label[node];
                restored++;
                node = root;
                if (restored >= orig_len) {
                    break;
                }
            }
        }
        if (rc != ORC_OK) {
            break;
        }
    }
    *out_len = produced;
    if (consumed) {
        *consumed = pos;
    }
    free(t);
    free(tree);
    return rc;
}
