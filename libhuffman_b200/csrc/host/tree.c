/*
 * tree.c — the pointer-linked huf_tree_t API object [ref: src/tree.c:12-427].
 * Host-side object for API/link compatibility (test/tree_test.c dereferences root/leaves);
 * the codec's tree lives in index arrays inside kernel K2 / K4.
 *
 * The merge rule is SURVEY.md §5.2's restatement: repeatedly take the two live weights
 * that come first under (weight ascending, slot index descending); first becomes the left
 * child, second the right child of a new slot 256, 257, ...; the last survivor is wrapped in a
 * node that has only a left child, and that node is the root.
 */
#include <string.h>

#include "internal.h"

#define SLOTS HUF_HISTOGRAM_LEN

huf_error_t
huf_node_to_string(const huf_node_t *self, uint8_t *buf, size_t *len)
{
    HUF_REQUIRE(buf);
    HUF_REQUIRE(len);

    size_t n = 0;
    /* climb to the root, writing '0' for "I am a left child", '1' otherwise */
    for (const huf_node_t *v = self; v && v->parent && n < *len; v = v->parent) {
        buf[n++] = v->parent->left == v ? '0' : '1';
    }
    *len = n;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_tree_init(huf_tree_t **self)
{
    HUF_REQUIRE(self);
    HUF_TRY(huf_malloc((void **)self, sizeof(**self), 1));
    return huf_malloc((void **)&(*self)->leaves, sizeof(huf_node_t *), SLOTS);
}

static void
release_subtree(huf_node_t *v)
{
    /* explicit stack: depth can reach the node count for degenerate trees */
    huf_node_t *stack[2 * HUF_BTREE_LEN + 8];
    int sp = 0;

    if (v) {
        stack[sp++] = v;
    }
    while (sp) {
        huf_node_t *n = stack[--sp];
        if (n->left) {
            stack[sp++] = n->left;
        }
        if (n->right) {
            stack[sp++] = n->right;
        }
        free(n);
    }
}

huf_error_t
huf_tree_reset(huf_tree_t *self)
{
    HUF_REQUIRE(self);
    release_subtree(self->root);
    self->root = NULL;
    memset(self->leaves, 0, sizeof(huf_node_t *) * SLOTS);
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_tree_free(huf_tree_t **self)
{
    HUF_REQUIRE(self);
    if (*self) {
        release_subtree((*self)->root);
        free((*self)->leaves);
        free(*self);
    }
    *self = NULL;
    return HUF_ERROR_SUCCESS;
}

/* Grammar T := -1 | v T T over at most `len` elements; missing elements = absent children;
 * trailing elements are ignored. */
static huf_error_t
parse(huf_node_t **out, const int16_t *buf, size_t len, size_t *pos, int depth)
{
    if (*pos >= len || depth > 2 * HUF_BTREE_LEN) {
        return HUF_ERROR_SUCCESS;
    }
    int16_t label = buf[(*pos)++];
    if (label == HUF_LEAF_NODE) {
        return HUF_ERROR_SUCCESS;
    }
    huf_node_t *v = NULL;
    HUF_TRY(huf_malloc((void **)&v, sizeof(*v), 1));
    *out = v;
    v->index = label;
    HUF_TRY(parse(&v->left, buf, len, pos, depth + 1));
    if (v->left) {
        v->left->parent = v;
    }
    HUF_TRY(parse(&v->right, buf, len, pos, depth + 1));
    if (v->right) {
        v->right->parent = v;
    }
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_tree_deserialize(huf_tree_t *self, const int16_t *buf, size_t len)
{
    HUF_REQUIRE(self);
    HUF_REQUIRE(buf);
    size_t pos = 0;
    return parse(&self->root, buf, len, &pos, 0);
}

huf_error_t
huf_tree_serialize(huf_tree_t *self, int16_t *buf, size_t *len)
{
    HUF_REQUIRE(self);
    HUF_REQUIRE(buf);
    HUF_REQUIRE(len);

    const huf_node_t *stack[2 * HUF_BTREE_LEN + 8];
    int sp = 0;
    size_t n = 0;

    stack[sp++] = self->root;
    while (sp) {
        const huf_node_t *v = stack[--sp];
        if (!v) {
            buf[n++] = HUF_LEAF_NODE;
            continue;
        }
        buf[n++] = v->index;
        stack[sp++] = v->right;
        stack[sp++] = v->left;
    }
    *len = n;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_tree_from_histogram(huf_tree_t *self, huf_histogram_t *histogram)
{
    HUF_REQUIRE(self);
    HUF_REQUIRE(histogram);

    uint64_t *w = histogram->frequencies; /* consumed: merged slots are zeroed */
    size_t limit = histogram->length < SLOTS ? histogram->length : SLOTS;
    int kid[SLOTS][2];
    int made = HUF_ASCII_COUNT;
    int root = -1;

    while ((size_t)made < limit) {
        int first = -1, second = -1;
        for (int j = 0; j < made; j++) {
            if (!w[j]) {
                continue;
            }
            if (first < 0 || w[j] <= w[first]) {
                second = first;
                first = j;
            } else if (second < 0 || w[j] <= w[second]) {
                second = j;
            }
        }
        if (first < 0) {
            break; /* empty histogram: no tree */
        }
        kid[made][0] = first;
        kid[made][1] = second;
        w[made] = w[first] + (second >= 0 ? w[second] : 0);
        w[first] = 0;
        if (second >= 0) {
            w[second] = 0;
        }
        root = made++;
        if (second < 0) {
            break; /* the survivor has just been wrapped: unary root */
        }
    }
    if (root < 0) {
        return HUF_ERROR_SUCCESS;
    }

    /* materialise the pointer structure the public struct promises */
    huf_node_t *node[SLOTS] = { 0 };
    for (int v = HUF_ASCII_COUNT; v <= root; v++) {
        for (int side = 0; side < 2; side++) {
            int c = kid[v][side];
            if (c >= 0 && !node[c]) {
                HUF_TRY(huf_malloc((void **)&node[c], sizeof(huf_node_t), 1));
                node[c]->index = (int16_t)c;
            }
        }
        if (!node[v]) {
            HUF_TRY(huf_malloc((void **)&node[v], sizeof(huf_node_t), 1));
            node[v]->index = (int16_t)v;
        }
        node[v]->left = kid[v][0] >= 0 ? node[kid[v][0]] : NULL;
        node[v]->right = kid[v][1] >= 0 ? node[kid[v][1]] : NULL;
        if (node[v]->left) {
            node[v]->left->parent = node[v];
        }
        if (node[v]->right) {
            node[v]->right->parent = node[v];
        }
    }
    for (int s = 0; s < HUF_ASCII_COUNT; s++) {
        self->leaves[s] = node[s];
    }
    self->root = node[root];
    return HUF_ERROR_SUCCESS;
}
