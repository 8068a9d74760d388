// enc_build.cuh — K2 of the encode path split in three kernels (blocks of at most 4 MiB; larger
// blocks keep the fused k_build<uint64_t> of enc_kernels.cuh):
//
//   K2a k_build_sort   one WARP per block: block histogram = sum of its segment histograms,
//                      keys (weight << 9 | 511 - symbol) sorted ascending -> blk_keys.
//   K2b k_build_merge  one LANE per block: the min-pair merge of huf_tree_from_histogram
//                      (reference src/tree.c:292-427) is a strictly serial chain of up to 256
//                      steps, so 32 blocks advance in lockstep per warp with every per-block
//                      array transposed in shared memory ([index][lane]: conflict-free whatever
//                      index a lane touches).  Exact tie-break: two-queue merge in which equal
//                      weight runs of merge nodes are consumed newest first and a merge node
//                      beats a leaf of equal weight (later index wins, src/tree.c:341,347).
//                      Emits {left, right, leaves below} per merge node -> blk_nodes.
//   K2c k_build_codes  one WARP per block: parents, code words and pre-order positions by
//                      climbing to the root (__huf_create_char_coding, src/encoder.c:40-81;
//                      huf_tree_serialize, src/tree.c:233-289), code table, segment bit
//                      offsets, block size.
#pragma once

#include "enc_kernels.cuh"

namespace hufb200 {

constexpr int kMergeDyn = 256 * 32 * 4 + 256 * 32;  // keys / merge nodes (u32) and run lengths (u8) per lane

// ------------------------------------------------------------------------------------------
// K2a
// ------------------------------------------------------------------------------------------

// Compare-exchange network of the bitonic sort, element e = 8 * lane + i held in register i of
// `lane`: partners at distance j < 8 are registers of the same lane, partners at distance j >= 8
// the same register of lane ^ (j >> 3), reached by one shuffle.  An element keeps the smaller key
// of its pair when it is the lower one of an ascending pair or the upper one of a descending pair.
__device__ __forceinline__ void sort256_in_registers(uint32_t (&key)[8], int lane)
{
#pragma unroll
    for (uint32_t k = 2; k <= 256; k <<= 1) {
#pragma unroll
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            if (j >= 8) {
                const bool keep_min = ((lane & (j >> 3)) == 0) == ((lane & (k >> 3)) == 0);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const uint32_t other = __shfl_xor_sync(kFull, key[i], (int)(j >> 3));
                    key[i] = keep_min ? min(key[i], other) : max(key[i], other);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if ((i & j) == 0) {  // (i, i | j) is a pair; its direction: bit k of the element index
                        const bool up = k >= 8 ? (lane & (k >> 3)) == 0 : (i & k) == 0;
                        const uint32_t lo = min(key[i], key[i | j]), hi = max(key[i], key[i | j]);
                        key[i] = up ? lo : hi;
                        key[i | j] = up ? hi : lo;
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kBuildWarps * 32) k_build_sort(EncArgs a)
{
    const int lane = lane_id();
    const uint64_t bl = (uint64_t)blockIdx.x * kBuildWarps + warp_in_cta();  // pass-local block
    if (bl >= a.npass) return;
    const uint64_t b = a.blk0 + bl;
    const uint32_t kMax = ~0u;

    const uint64_t blen = blk_len_of(a, b);
    const uint32_t nseg_b = (uint32_t)((blen + a.seg - 1) / a.seg);
    const uint64_t g0 = bl * a.nspb;

    // block histogram: lane owns symbols 8*lane .. 8*lane+7
    uint32_t cnt[8];
#pragma unroll
    for (int i = 0; i < 8; i++) cnt[i] = 0;
    for (uint32_t k = 0; k < nseg_b; k++) {
        uint4 v = reinterpret_cast<const uint4 *>(a.seg_hist + (g0 + k) * 256)[lane];
        cnt[0] += v.x & 0xffffu; cnt[1] += v.x >> 16;
        cnt[2] += v.y & 0xffffu; cnt[3] += v.y >> 16;
        cnt[4] += v.z & 0xffffu; cnt[5] += v.z >> 16;
        cnt[6] += v.w & 0xffffu; cnt[7] += v.w >> 16;
    }
    // keys (absent symbols sort to the end), sorted without leaving the registers: no shared
    // memory, no barriers, 2 instructions per element and exchange step
    uint32_t key[8];
    uint32_t present = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t s = lane * 8 + i;
        key[i] = cnt[i] ? make_key<uint32_t>(cnt[i], s) : kMax;
        present += cnt[i] != 0;
    }
    const uint32_t n = warp_sum(present);  // distinct symbols, >= 1
    sort256_in_registers(key, lane);
    uint4 *dst = reinterpret_cast<uint4 *>(a.blk_keys + bl * 256) + 2 * lane;
    dst[0] = make_uint4(key[0], key[1], key[2], key[3]);
    dst[1] = make_uint4(key[4], key[5], key[6], key[7]);
    if (lane == 0) a.blk_meta[bl * 4 + 3] = n;
}

// ------------------------------------------------------------------------------------------
// K2b
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(32) k_build_merge(EncArgs a)
{
#ifdef HUF_EMU
    uint32_t *arr = reinterpret_cast<uint32_t *>(hufemu::dyn_smem());
#else
    extern __shared__ __align__(16) uint32_t arr[];
#endif
    // One array of 256 words per lane, transposed (arr[j * 32 + lane]: the bank is the lane's
    // whatever j each lane is at), serves both queues of the two-queue merge: the sorted leaf
    // keys sit in it from the start, and merge node 256 + j parks its (weight << 8 | leaves
    // below - 1) in word j -- free by then: when node j is made, 2 (j + 1) items have been
    // consumed, at most j of them nodes, so the leaf cursor is already past j + 1.
    // runlen[j * 32 + lane]: nodes of equal weight from node j on, kept at the first of a run.
    const int lane = lane_id();
    const uint64_t bl = (uint64_t)blockIdx.x * 32 + lane;  // pass-local block
    const bool live = bl < a.npass;
    const uint32_t kMax = ~0u;
    uint32_t *mine = arr + lane;
    uint8_t *runlen = reinterpret_cast<uint8_t *>(arr + 256 * 32) + lane;
#define SLOT(j) mine[(j) * 32]
#define RUNLEN(j) runlen[(j) * 32]
    // all 256 keys of my block (absent symbols sort to the end as kMax): sixty-four 16-byte loads
    // in flight together, instead of a load per pick on the chain
    if (live) {
        const uint4 *keys = reinterpret_cast<const uint4 *>(a.blk_keys + bl * 256);
#pragma unroll 8
        for (int i = 0; i < 64; i++) {
            const uint4 v = keys[i];
            SLOT(4 * i) = v.x;
            SLOT(4 * i + 1) = v.y;
            SLOT(4 * i + 2) = v.z;
            SLOT(4 * i + 3) = v.w;
        }
    }
    __syncwarp();
    if (!live) return;
    uint2 *out = reinterpret_cast<uint2 *>(a.blk_nodes) + bl * 256;

    // Leaves are consumed in key order.  Merge nodes are created with non-decreasing weight, so
    // they form runs of equal weight in creation order; the live ones are [head, top) -- the
    // front run, consumed newest first, its weight in `runw` -- plus [nxt, made).  A run is
    // complete before its first node is consumed (what is made afterwards is heavier), so its
    // length can be kept at its first node and opening a run is two loads, not a scan.
    uint32_t li = 0, head = 0, top = 0, nxt = 0, made = 0, runw = 0;
    uint32_t run0 = 0, run_n = 0, last_w = kMax;  // the newest run: first node, length, weight
    uint32_t k0 = SLOT(0);                          // key of the next leaf (kMax: none left)
    for (;;) {
        uint32_t pick0 = 0, pick1 = kNone16, w0 = 0, w1 = 0, nl = 0;
        int got = 0;
#pragma unroll
        for (int s = 0; s < 2; s++) {
            if (top == head && nxt < made) {  // front run used up: open the next one
                head = nxt;
                runw = SLOT(nxt) >> 8;
                top = nxt = nxt + RUNLEN(nxt);
            }
            const bool has_i = top > head;
            if (!has_i && k0 == kMax) break;  // nothing left (second pick only)
            // One straight-line pick: the newest node of the front run against the next leaf
            // (key = weight << 9 | 511 - index: a node beats a leaf of equal weight, a newer node
            // an older one); both candidates' data are at hand, the winner is selected.
            const uint32_t ikey = has_i ? ((runw << 9) | (256u - top)) : kMax;
            const bool node = ikey < k0;
            const uint32_t nv = SLOT(has_i ? top - 1u : 0u);
            const uint32_t pick = node ? 255u + top : 511u - (k0 & 511u);
            const uint32_t pw = node ? runw : (k0 >> 9);
            const uint32_t pl = node ? (nv & 0xffu) + 1u : 1u;
            top -= node ? 1u : 0u;
            const bool spent = node && top == head;
            head = spent ? nxt : head;
            top = spent ? nxt : top;
            if (!node) {
                li++;
                k0 = li < 256u ? SLOT(li) : kMax;
            }
            if (s == 0) {
                pick0 = pick;
                w0 = pw;
            } else {
                pick1 = pick;
                w1 = pw;
            }
            nl += pl;
            got++;
        }
        const uint32_t weight = w0 + (got == 2 ? w1 : 0u);
        SLOT(made) = (weight << 8) | (nl - 1u);
        out[made] = make_uint2(pick0 | (pick1 << 16), nl);
        // the new node starts a run or extends the newest one
        const bool same = weight == last_w;
        run0 = same ? run0 : made;
        run_n = same ? run_n + 1u : 1u;
        last_w = weight;
        RUNLEN(run0) = (uint8_t)run_n;
        // an open front run that is still untouched and is the newest run grows with it
        if (top > head && top == nxt && nxt == made && weight == runw) {
            top++;
            nxt++;
        }
        made++;
        if (got < 2) break;
    }
#undef SLOT
#undef RUNLEN
}

// ------------------------------------------------------------------------------------------
// K2c
// ------------------------------------------------------------------------------------------

// Code word, length and pre-order position of a node are sums over the edges of its path from the
// root.  Every merge node keeps a link "I hang `len` edges below merge node `anc`, reached by the
// path bits `bits`, `pos` elements behind it in the serialised tree"; a round replaces the link by
// its composition with the link of `anc` (pointer jumping), so after ceil(log2(depth)) rounds every
// link ends at the root -- a few hundred independent instructions per lane where climbing from
// every node to the root was a chain of four dependent shared-memory loads per level and node.
struct CodesSmem {
    uint64_t bits[256];      // path bits from `anc` down to merge node 256 + j (the low `len` ones)
    uint32_t link[256];      // anc | len << 16
    uint16_t pos[256];       // serialised elements between `anc` and the node
    uint16_t isz[256];       // serialised size (elements) of the subtree of merge node 256 + j
    uint16_t leaf_up[256];   // symbol s: parent merge node | right child << 15; kNone16: absent
    uint16_t leaf_pos[256];  // serialised elements between the parent and the leaf
    uint8_t len[256];
    __align__(16) int16_t tree[kTreeStride];  // the serialised tree, assembled here and copied out in lines
};

__global__ void __launch_bounds__(kBuildWarps * 32) k_build_codes(EncArgs a)
{
    __shared__ CodesSmem sm_all[kBuildWarps];
    const int lane = lane_id();
    const uint64_t bl = (uint64_t)blockIdx.x * kBuildWarps + warp_in_cta();
    if (bl >= a.npass) return;
    const uint64_t b = a.blk0 + bl;
    CodesSmem &sm = sm_all[warp_in_cta()];
    const uint64_t blen = blk_len_of(a, b);
    const uint32_t nseg_b = (uint32_t)((blen + a.seg - 1) / a.seg);
    const uint64_t g0 = bl * a.nspb;
    const uint32_t n = a.blk_meta[bl * 4 + 3];  // distinct symbols = merge nodes made
    const uint32_t root = n - 1;                // the one-child root is the last merge node

    const uint2 *nodes = reinterpret_cast<const uint2 *>(a.blk_nodes) + bl * 256;
    uint2 mine[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t j = lane + 32 * i;
        sm.leaf_up[j] = kNone16;
        sm.len[j] = 0;
        mine[i] = j < n ? nodes[j] : make_uint2(0, 0);
        // a subtree with L leaves below a two-child node serialises to 4L - 1 elements; the
        // one-child root adds itself and its absent right child
        sm.isz[j] = (uint16_t)((mine[i].x >> 16) == kNone16 ? 4 * mine[i].y + 1 : 4 * mine[i].y - 1);
    }
    __syncwarp();
    // every merge node hands its children their first link: one edge, to itself
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t j = lane + 32 * i;
        if (j >= n) continue;
        const uint32_t l = mine[i].x & 0xffffu, r = mine[i].x >> 16;
        const uint32_t behind_left = 1u + (l < 256u ? 3u : sm.isz[l - 256u]);
        if (l < 256u) {
            sm.leaf_up[l] = (uint16_t)j;
            sm.leaf_pos[l] = 1;
        } else {
            sm.link[l - 256u] = j | (1u << 16);
            sm.bits[l - 256u] = 0;
            sm.pos[l - 256u] = 1;
        }
        if (r == kNone16) {
        } else if (r < 256u) {
            sm.leaf_up[r] = (uint16_t)(j | 0x8000u);
            sm.leaf_pos[r] = (uint16_t)behind_left;
        } else {
            sm.link[r - 256u] = j | (1u << 16);
            sm.bits[r - 256u] = 1;
            sm.pos[r - 256u] = (uint16_t)behind_left;
        }
        if (j == root) {
            sm.link[j] = j;  // zero edges below itself: composing with it changes nothing
            sm.bits[j] = 0;
            sm.pos[j] = 0;
        }
    }
    __syncwarp();
    // pointer jumping; links are read by everybody, then written by their owners
    for (int round = 0; round < 9; round++) {
        uint32_t up_link[8], up_pos[8];
        uint64_t up_bits[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t j = lane + 32 * i;
            const uint32_t anc = j < n ? sm.link[j] & 0xffffu : 0u;
            up_link[i] = sm.link[anc];
            up_bits[i] = sm.bits[anc];
            up_pos[i] = sm.pos[anc];
        }
        __syncwarp();
        bool moving = false;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t j = lane + 32 * i;
            if (j >= n) continue;
            const uint32_t len = sm.link[j] >> 16;
            sm.bits[j] |= len < 64u ? up_bits[i] << len : 0ull;
            sm.pos[j] = (uint16_t)(sm.pos[j] + up_pos[i]);
            sm.link[j] = (up_link[i] & 0xffffu) | ((len + (up_link[i] >> 16)) << 16);
            moving |= (up_link[i] & 0xffffu) != root;
        }
        __syncwarp();
        if (!__any_sync(kFull, moving)) break;
    }

    // (element writes scattered over global memory cost a sector each: the tree is put together
    // in shared memory and leaves as 16-byte lines)
    const uint32_t tree_len = sm.isz[root];
    int16_t *tree = sm.tree;
    uint32_t *tab32 = a.blk_table + bl * 512;
    uint32_t my_max = 0;
    uint64_t code_of[8];
    uint32_t len_of[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t s = lane + 32 * i;
        uint64_t code = 0;
        uint32_t len = 0;
        const uint32_t up = sm.leaf_up[s];
        if (up != kNone16) {
            const uint32_t pj = up & 0x7fffu;
            len = (sm.link[pj] >> 16) + 1u;
            code = (sm.bits[pj] << 1) | (up >> 15);
            const uint32_t pos = (uint32_t)sm.pos[pj] + sm.leaf_pos[s];
            tree[pos] = (int16_t)s;
            tree[pos + 1] = -1;
            tree[pos + 2] = -1;
            sm.len[s] = (uint8_t)len;
        }
        code_of[i] = code;
        len_of[i] = len;
        my_max = max(my_max, len);
        if (s < n) tree[sm.pos[s]] = (int16_t)(256u + s);  // merge node 256 + s
    }
    if (lane == 0) tree[tree_len - 1] = -1;  // absent right child of the one-child root
    __syncwarp();
    {
        uint4 *dst = reinterpret_cast<uint4 *>(a.blk_tree + bl * kTreeStride);
        const uint4 *src = reinterpret_cast<const uint4 *>(sm.tree);
        for (uint32_t k = lane; 8 * k < tree_len; k += 32) dst[k] = src[k];
    }

    const uint32_t max_len = warp_max(my_max);
    const uint32_t fmt = max_len <= 26 ? 0u : 1u;
    if (max_len > 56 && lane == 0) {
        atomicMax(&a.status[0], (uint32_t)kErrFatal);
        a.status[1] = max_len;
    }
    if ((max_len > 16 || n == 1) && lane == 0) atomicAdd(&a.status[2], 1u);  // k_pack_wide has work
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t s = lane + 32 * i;
        const uint32_t len = len_of[i];
        if (fmt == 0) {
            tab32[s] = len ? ((uint32_t)code_of[i] << (32 - len)) | len : 0u;
        } else {
            const uint64_t e = len ? (code_of[i] << (64 - len)) | len : 0ull;
            reinterpret_cast<uint64_t *>(tab32)[s] = e;
        }
    }
    __syncwarp();

    // payload bit offset of every segment = running dot(segment histogram, code length)
    uint32_t l8[8];
#pragma unroll
    for (int i = 0; i < 8; i++) l8[i] = sm.len[lane * 8 + i];
    uint64_t run = 0;
    for (uint32_t k = 0; k < nseg_b; k++) {
        uint4 v = reinterpret_cast<const uint4 *>(a.seg_hist + (g0 + k) * 256)[lane];
        uint32_t d = (v.x & 0xffffu) * l8[0] + (v.x >> 16) * l8[1] +
                     (v.y & 0xffffu) * l8[2] + (v.y >> 16) * l8[3] +
                     (v.z & 0xffffu) * l8[4] + (v.z >> 16) * l8[5] +
                     (v.w & 0xffffu) * l8[6] + (v.w >> 16) * l8[7];
        d = warp_sum(d);
        if (lane == 0) a.seg_bitoff[g0 + k] = run;
        run += d;
    }
    if (lane == 0) {
        a.blk_bits[bl] = run;
        a.blk_size[b] = (uint64_t)kHdrFixed + 2ull * tree_len + ((run + 7) >> 3);
        uint32_t *m = a.blk_meta + bl * 4;
        m[0] = tree_len;
        m[1] = max_len;
        m[2] = fmt;
        m[3] = n;
    }
}

}  // namespace hufb200
