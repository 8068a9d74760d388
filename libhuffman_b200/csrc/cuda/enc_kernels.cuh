// enc_kernels.cuh — sm_100a kernels of the encode path.
//
//   K1  k_seg_hist      one warp per input segment: byte histogram in warp-private shared
//                       memory counters fed by 16-byte streaming loads; u16[256] per segment.
//                       Replaces huf_histogram_populate (reference src/histogram.c:73-103).
//   K2  k_build<W>      one warp per block: block histogram = sum of its segment histograms,
//                       exact min-pair merge with the reference's tie-break, code words,
//                       serialised tree, per-segment payload bit offsets, block byte size.
//                       Replaces huf_tree_from_histogram (src/tree.c:292-427),
//                       __huf_create_char_coding (src/encoder.c:40-81) + huf_node_to_string
//                       (src/tree.c:12-47) and huf_tree_serialize (src/tree.c:233-289).
//                       Used with 64-bit keys for blocks above 4 MiB only; ordinary blocks go
//                       through the three kernels of enc_build.cuh (sort / lane-per-block merge
//                       / codes).
//       k_scan_sizes    exclusive scan of block byte sizes -> block offsets in the stream.
//   K3  (enc_pack.cuh)  k_pack and k_pack_wide; this file keeps what they share: the owned byte
//                       range of a segment, the per-symbol accumulator of the ragged ends, the
//                       segmented carry scan, and the block header (emit_block_header).
//
// Data layout in HBM (all sizes for nblocks blocks, nspb segments per block):
//   seg_hist   u16 [nblocks*nspb][256]    written by K1, read by K2
//   seg_bitoff u64 [nblocks*nspb]         payload bit offset of each segment inside its block
//   blk_bits   u64 [nblocks]              payload bits per block
//   blk_size   u64 [nblocks]              10 + 2*tree_len + ceil(bits/8)
//   blk_off    u64 [nblocks+1]            exclusive scan of blk_size
//   blk_table  u32 [nblocks][512]         fmt0: 256 x u32 (code<<(32-len) | len), len <= 26
//                                         fmt1: 256 x u64 (code<<(64-len) | len), len <= 56
//   blk_tree   i16 [nblocks][kTreeStride] pre-order tree, -1 = absent child
//   blk_meta   u32 [nblocks][4]           {tree_len, max_len, fmt, nsym}
//   blk_keys   u32 [nblocks][256]         sorted merge keys, K2a -> K2b
//   blk_nodes  u32 [nblocks][256][2]      merge nodes, K2b -> K2c
#pragma once

#include "common.cuh"

namespace hufb200 {

struct EncArgs {
    const uint8_t *in;
    uint64_t length;
    uint64_t blocksize;
    uint64_t nblocks;   // blocks in the whole call
    uint64_t blk0;      // first block of this pass
    uint64_t npass;     // blocks in this pass (workspace arrays are indexed pass-locally)
    uint32_t seg;    // segment size in bytes (multiple of 16, <= 16384)
    uint32_t nspb;   // segments per full block
    uint8_t *out;
    uint64_t out_cap;
    // workspace
    uint16_t *seg_hist;
    uint64_t *seg_bitoff;
    uint64_t *blk_bits;
    uint64_t *blk_size;
    uint64_t *blk_off;
    uint32_t *blk_table;
    int16_t *blk_tree;
    uint32_t *blk_meta;
    uint32_t *blk_keys;   // [npass][256] sorted merge keys (enc_build.cuh)
    uint32_t *blk_nodes;  // [npass][256][2] merge nodes: left | right << 16, leaves below
    uint32_t *status;  // [0] error code, [1] detail
};

constexpr int kEncWarps = 8;      // warps per CTA in K1/K3
constexpr int kBuildWarps = 4;    // warps per CTA in K2
constexpr uint32_t kNone16 = 0xffffu;

__device__ __forceinline__ uint64_t blk_len_of(const EncArgs &a, uint64_t b)
{
    uint64_t start = b * a.blocksize;
    uint64_t left = a.length - start;
    return left < a.blocksize ? left : a.blocksize;
}

// Block header (reference src/encoder.c:325-342): u64 orig_len | i16 tree_len | i16 tree[tree_len],
// written by one warp.  The serialised tree is up to 2 KB: the 16-byte lines of the output that hold
// nothing but tree elements are assembled from aligned words of the workspace copy (shifted to the
// header's byte alignment, which is arbitrary) and stored whole; only the bytes around them -- the
// ten fixed ones, the ragged head and tail of the tree -- go byte by byte.
__device__ __forceinline__ void emit_block_header(const EncArgs &a, uint64_t bl, uint64_t blen, uint32_t tree_len,
                                                  uint64_t boff, int lane)
{
    const int16_t *tree = a.blk_tree + bl * kTreeStride;
    const uint32_t *tw = reinterpret_cast<const uint32_t *>(tree);
    const uint32_t hlen = kHdrFixed + 2 * tree_len;
    uint8_t *dst = a.out + boff;
    // header bytes [i0, i1): whole output lines made of tree bytes only
    uint32_t i0 = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15)) & 15u;
    if (i0 < (uint32_t)kHdrFixed) i0 += 16;
    if (i0 > hlen) i0 = hlen;
    const uint32_t i1 = i0 + ((hlen - i0) & ~15u);
    for (uint32_t i = i0 + 16 * lane; i < i1; i += 512) {
        const uint32_t t0 = i - kHdrFixed;  // byte offset into the serialised tree
        const uint32_t *w = tw + (t0 >> 2);
        const uint32_t sh = (t0 & 3) * 8;
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
        *reinterpret_cast<uint4 *>(dst + i) = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh),
                                                         __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
    }
    const uint32_t around = i0 + (hlen - i1);
    for (uint32_t e = lane; e < around; e += 32) {
        const uint32_t i = e < i0 ? e : e - i0 + i1;
        uint32_t v;
        if (i < 8) {
            v = (uint32_t)(blen >> (8 * i));
        } else if (i < 10) {
            v = tree_len >> (8 * (i - 8));
        } else {
            v = (uint32_t)(uint16_t)tree[(i - 10) >> 1] >> (8 * (i & 1));
        }
        dst[i] = (uint8_t)v;
    }
}

// ------------------------------------------------------------------------------------------
// K1: per-segment byte histogram.
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ void hist_word(uint32_t *h, uint32_t w)
{
    atomicAdd(&h[w & 0xffu], 1u);
    atomicAdd(&h[(w >> 8) & 0xffu], 1u);
    atomicAdd(&h[(w >> 16) & 0xffu], 1u);
    atomicAdd(&h[w >> 24], 1u);
}

__device__ __forceinline__ void hist_u4(uint32_t *h, const uint4 &v)
{
    hist_word(h, v.x);
    hist_word(h, v.y);
    hist_word(h, v.z);
    hist_word(h, v.w);
}

// kHistSets counter sets per warp (lane & (kHistSets-1) picks one), skewed by 8 banks.  Measured
// on B200 with Zipf(1.1) input: 1 set 0.33 ms / GiB, 4 sets 0.42 ms (the skew moves the hot
// symbols onto each other's banks), so one set it is.  Counting the segment's most frequent
// byte value in a register (SIMD compare + popc, its atomics predicated off) is slower too
// (0.49 ms): the extra compare work costs more than the serialised atomics it removes.  Round 2
// tried two sets picked by lane parity with the second at index s ^ 16 (the copy of a hot low
// symbol lands on the bank of a symbol sixteen ranks colder): 0.385 ms on Zipf, 0.46 ms on
// uniform bytes -- the extra XOR per atomic and the second set cost more than the conflicts.
constexpr int kHistSets = 1;
constexpr int kHistStride = 256 + 8;

__global__ void __launch_bounds__(kEncWarps * 32) k_seg_hist(EncArgs a)
{
    __shared__ uint32_t sh[kEncWarps][kHistSets * kHistStride];
    const int lane = lane_id();
    const int w = warp_in_cta();
    const uint64_t g = (uint64_t)blockIdx.x * kEncWarps + w;  // pass-local segment index
    if (g >= a.npass * a.nspb) return;

    uint32_t *hw = sh[w];
    for (int i = lane; i < kHistSets * kHistStride; i += 32) hw[i] = 0;
    __syncwarp();
    uint32_t *h = hw + (lane & (kHistSets - 1)) * kHistStride;

    const uint64_t b = a.blk0 + g / a.nspb;
    const uint32_t k = (uint32_t)(g % a.nspb);
    const uint64_t blen = blk_len_of(a, b);
    const uint64_t soff = (uint64_t)k * a.seg;
    uint32_t slen = 0;
    if (soff < blen) slen = (uint32_t)((blen - soff) < a.seg ? (blen - soff) : a.seg);
    const uint8_t *p = a.in + b * a.blocksize + soff;

    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const uint4 *p4 = reinterpret_cast<const uint4 *>(p);
        const uint32_t n16 = slen >> 4;
        uint32_t i = lane;
        // 4 independent 16-byte loads in flight per lane
        for (; i + 96 < n16; i += 128) {
            uint4 v0 = ld_stream_u4(p4 + i);
            uint4 v1 = ld_stream_u4(p4 + i + 32);
            uint4 v2 = ld_stream_u4(p4 + i + 64);
            uint4 v3 = ld_stream_u4(p4 + i + 96);
            hist_u4(h, v0);
            hist_u4(h, v1);
            hist_u4(h, v2);
            hist_u4(h, v3);
        }
        for (; i < n16; i += 32) hist_u4(h, ld_stream_u4(p4 + i));
        for (uint32_t j = (n16 << 4) + lane; j < slen; j += 32) atomicAdd(&h[p[j]], 1u);
    } else {
        for (uint32_t j = lane; j < slen; j += 32) atomicAdd(&h[p[j]], 1u);
    }
    __syncwarp();

    // lane l owns bins 8l .. 8l+7 -> one 16-byte store per lane, 512 contiguous bytes per warp
    uint32_t c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c[i] = 0;
#pragma unroll
        for (int q = 0; q < kHistSets; q++) c[i] += hw[q * kHistStride + lane * 8 + i];
    }
    uint4 o;
    o.x = c[0] | (c[1] << 16);
    o.y = c[2] | (c[3] << 16);
    o.z = c[4] | (c[5] << 16);
    o.w = c[6] | (c[7] << 16);
    reinterpret_cast<uint4 *>(a.seg_hist + g * 256)[lane] = o;
}

// ------------------------------------------------------------------------------------------
// K2: per-block code build.
// ------------------------------------------------------------------------------------------

template <typename W>
struct BuildSmem {
    W key[256];          // block histogram, then leaf keys sorted ascending
    W iw[256];           // weight of merge node 256 + j
    uint16_t isz[256];   // serialised size (elements) of the subtree of merge node 256 + j
    uint16_t lch[256];   // children of merge node 256 + j
    uint16_t rch[256];
    uint16_t par[512];   // parent of every node
    uint16_t rend[256];  // for the first merge node of an equal-weight run: one past its last
    uint8_t len[256];    // code length per symbol (0 = absent)
};

// Key orders nodes by (weight ascending, node index descending): key = weight << 9 | (511 - idx).
template <typename W>
__device__ __forceinline__ W make_key(W weight, uint32_t idx)
{
    return (weight << 9) | (W)(511u - idx);
}

template <typename W>
__global__ void __launch_bounds__(kBuildWarps * 32) k_build(EncArgs a)
{
    __shared__ BuildSmem<W> sm_all[kBuildWarps];
    const int lane = lane_id();
    const uint64_t bl = (uint64_t)blockIdx.x * kBuildWarps + warp_in_cta();  // pass-local block
    if (bl >= a.npass) return;
    const uint64_t b = a.blk0 + bl;
    BuildSmem<W> &sm = sm_all[warp_in_cta()];
    const W kMax = ~(W)0;

    const uint64_t blen = blk_len_of(a, b);
    const uint32_t nseg_b = (uint32_t)((blen + a.seg - 1) / a.seg);
    const uint64_t g0 = bl * a.nspb;

    // (1) block histogram: lane owns symbols 8*lane .. 8*lane+7
    W cnt[8];
#pragma unroll
    for (int i = 0; i < 8; i++) cnt[i] = 0;
    for (uint32_t k = 0; k < nseg_b; k++) {
        uint4 v = reinterpret_cast<const uint4 *>(a.seg_hist + (g0 + k) * 256)[lane];
        cnt[0] += v.x & 0xffffu; cnt[1] += v.x >> 16;
        cnt[2] += v.y & 0xffffu; cnt[3] += v.y >> 16;
        cnt[4] += v.z & 0xffffu; cnt[5] += v.z >> 16;
        cnt[6] += v.w & 0xffffu; cnt[7] += v.w >> 16;
    }
    uint32_t present = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t s = lane * 8 + i;
        sm.key[s] = cnt[i] ? make_key<W>(cnt[i], s) : kMax;
        present += cnt[i] != 0;
        sm.len[s] = 0;
    }
    const uint32_t n = warp_sum(present);  // distinct symbols, >= 1
    for (int i = lane; i < 512; i += 32) sm.par[i] = kNone16;
    __syncwarp();

    // (2) bitonic sort of the 256 leaf keys (absent symbols sort to the end)
    for (uint32_t k = 2; k <= 256; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (uint32_t t = lane; t < 128; t += 32) {
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const uint32_t p = i | j;
                const bool up = (i & k) == 0;
                const W x = sm.key[i], y = sm.key[p];
                if ((x > y) == up) {
                    sm.key[i] = y;
                    sm.key[p] = x;
                }
            }
            __syncwarp();
        }
    }

    // (3) two-queue merge, serial in lane 0 with all queue state in registers.  Leaves are
    // consumed in key order from sm.key (two keys prefetched).  Merge nodes are created with
    // non-decreasing weight, so they form runs of equal weight in creation order; the live
    // ones are [head, top) -- the front run, consumed newest first, its weight in `runw` --
    // plus [nxt, made).  Run extents are recorded when nodes are created (sm.rend), so no
    // scanning is needed when a run is opened.
    if (lane == 0) {
        uint32_t li = 0, head = 0, top = 0, nxt = 0, made = 0, run_first = 0;
        W runw = 0, last_w = 0;
        W k0 = sm.key[0];                 // n >= 1
        W k1 = n > 1 ? sm.key[1] : kMax;
        for (;;) {
            uint32_t pick0 = 0, pick1 = kNone16;
            W w0 = 0, w1 = 0;
            int got = 0;
#pragma unroll
            for (int s = 0; s < 2; s++) {
                if (top == head && nxt < made) {  // front run used up: open the next one
                    head = nxt;
                    runw = sm.iw[nxt];
                    top = nxt = sm.rend[nxt];
                }
                const bool has_i = top > head;
                const W ikey = has_i ? make_key<W>(runw, 255u + top) : kMax;
                if (ikey == kMax && k0 == kMax) break;  // nothing left (second pick only)
                uint32_t pick;
                W pw;
                if (ikey < k0) {
                    top--;
                    pick = 256u + top;
                    pw = runw;
                    if (top == head) head = top = nxt;
                } else {
                    pick = 511u - (uint32_t)(k0 & 511u);
                    pw = k0 >> 9;
                    li++;
                    k0 = k1;
                    k1 = li + 1 < n ? sm.key[li + 1] : kMax;
                }
                if (s == 0) {
                    pick0 = pick;
                    w0 = pw;
                } else {
                    pick1 = pick;
                    w1 = pw;
                }
                got++;
            }
            const uint32_t me = 256u + made;
            uint32_t size = 1 + (pick0 < 256u ? 3u : sm.isz[pick0 - 256u]);
            sm.lch[made] = (uint16_t)pick0;
            sm.par[pick0] = (uint16_t)me;
            W weight = w0;
            if (got == 2) {
                sm.par[pick1] = (uint16_t)me;
                size += pick1 < 256u ? 3u : sm.isz[pick1 - 256u];
                weight += w1;
            } else {
                size += 1;  // the absent right child of the unary root
            }
            sm.rch[made] = (uint16_t)pick1;
            sm.iw[made] = weight;
            sm.isz[made] = (uint16_t)size;
            // run bookkeeping: extend the newest run or start a new one
            if (made == 0 || weight != last_w) run_first = made;
            sm.rend[run_first] = (uint16_t)(made + 1);
            last_w = weight;
            // an open front run that is still untouched and is the newest run grows with it
            if (top > head && top == nxt && nxt == made && weight == runw) {
                top++;
                nxt++;
            }
            made++;
            if (got < 2) break;
        }
    }
    __syncwarp();

    // (4) climb from every node to the root: code word + pre-order position.
    const uint32_t root = 255u + n;  // n merge nodes were made
    const uint32_t tree_len = sm.isz[n - 1];
    int16_t *tree = a.blk_tree + bl * kTreeStride;
    uint32_t *tab32 = a.blk_table + bl * 512;
    uint32_t my_max = 0;

    // leaves first (need code + length), lane handles symbols lane, lane+32, ...
    uint64_t code_of[8];
    uint32_t len_of[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t s = lane + 32 * i;
        uint64_t code = 0;
        uint32_t len = 0, pos = 0;
        if (sm.par[s] != kNone16) {
            uint32_t x = s;
            while (x != root) {
                const uint32_t p = sm.par[x];
                const uint32_t pj = p - 256u;
                const bool is_right = sm.rch[pj] == x;
                if (is_right) {
                    const uint32_t l = sm.lch[pj];
                    code |= 1ull << len;
                    pos += l < 256u ? 3u : sm.isz[l - 256u];
                }
                pos += 1;
                len++;
                x = p;
            }
            tree[pos] = (int16_t)s;
            tree[pos + 1] = -1;
            tree[pos + 2] = -1;
            sm.len[s] = (uint8_t)len;
        }
        code_of[i] = code;
        len_of[i] = len;
        my_max = max(my_max, len);
    }
    // merge nodes: position only
    for (uint32_t v = 256u + lane; v <= root; v += 32) {
        uint32_t pos = 0, x = v;
        while (x != root) {
            const uint32_t p = sm.par[x];
            const uint32_t pj = p - 256u;
            if (sm.rch[pj] == x) {
                const uint32_t l = sm.lch[pj];
                pos += l < 256u ? 3u : sm.isz[l - 256u];
            }
            pos += 1;
            x = p;
        }
        tree[pos] = (int16_t)v;
    }
    if (lane == 0) tree[tree_len - 1] = -1;  // absent right child of the unary root

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

    // (5) payload bit offset of every segment = running dot(segment histogram, code length)
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

// ------------------------------------------------------------------------------------------
// Block-offset scan: blk_off[i] = sum_{j<i} blk_size[j], blk_off[n] = total.  One CTA.
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(kScanThreads) k_scan_sizes(const uint64_t *__restrict__ in,
                                                             uint64_t *__restrict__ out,
                                                             uint64_t n, uint64_t cap,
                                                             uint32_t *status)
{
    __shared__ uint64_t warp_tot[kScanThreads / 32 + 1];
    // out[0] holds the running total of the previous passes (0 for the first pass)
    const uint64_t base = out[0];
    __syncthreads();
    const uint64_t total = cta_excl_scan(in, out, n, base, warp_tot);
    if (threadIdx.x == 0) {
        out[n] = total;
        if (total > cap) atomicMax(&status[0], (uint32_t)kErrNoMem);
    }
}

// ------------------------------------------------------------------------------------------
// K3: bit packing.
// ------------------------------------------------------------------------------------------

constexpr int kStageWords = 912;  // 512 symbols * 56 bits / 32 + the kept line + slack

struct PackSmem {
    uint32_t table[kEncWarps][512];       // per-warp copy of the block's code table
    uint32_t stage[kEncWarps][kStageWords];
};

// Owned output byte range of one warp and the store helpers that respect it.
struct OutRange {
    uint8_t *out;
    uint64_t b0, b1;          // owned bytes [b0, b1)
    uint64_t full_lo, full_hi;  // word indices fully inside the owned range
};

__device__ __forceinline__ void store_word(const OutRange &r, uint64_t widx, uint32_t be)
{
    if (widx >= r.full_lo && widx < r.full_hi) {
        reinterpret_cast<uint32_t *>(r.out)[widx] = bswap32(be);
    } else {
        const uint64_t base = widx << 2;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint64_t addr = base + j;
            if (addr >= r.b0 && addr < r.b1) r.out[addr] = (uint8_t)(be >> (24 - 8 * j));
        }
    }
}

// Insert the top `l` bits of `t` (left aligned, other bits zero) at bit `nb` of the hi:lo
// window; a finished 32-bit word goes to the staging buffer with a plain store (every word
// is finished by exactly one lane; the bits earlier lanes left in a lane's first word are
// OR-ed in afterwards, see pack_flush_carry).  The body is small enough to be predicated.
struct BitAcc {
    uint32_t hi, lo, nb, widx;
};

__device__ __forceinline__ void acc_put(BitAcc &s, uint32_t *stage, uint32_t t, uint32_t l)
{
    s.hi |= t >> s.nb;
    s.lo |= __funnelshift_r(0u, t, s.nb);
    s.nb += l;
    if (s.nb >= 32) {
        stage[s.widx] = s.hi;
        s.hi = s.lo;
        s.lo = 0;
        s.nb -= 32;
        s.widx++;
    }
}

// After every lane ran its symbols through acc_put: hand the unfinished tail of each lane
// to the lane that finishes that word.  A segmented OR-scan over the lanes (segments start
// at lanes that finished at least one word) yields, for every lane, the bits its
// predecessors left in its first word; `carry` is the same thing across iterations.
// Returns the warp's new carry (the unfinished word after lane 31).
__device__ __forceinline__ uint32_t pack_flush_carry(const BitAcc &acc, uint32_t first_widx,
                                                    uint32_t *stage, uint32_t carry)
{
    const int lane = lane_id();
    const bool done_one = acc.widx != first_widx;   // this lane finished >= 1 word
    const uint32_t heads = __ballot_sync(kFull, done_one);
    uint32_t v = acc.hi;  // own unfinished bits (zero when nb == 0)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t pv = __shfl_up_sync(kFull, v, d);
        // take the predecessor's run only if no segment starts in lanes (lane-d, lane]
        if (lane >= d && ((heads >> (lane - d + 1)) & ((1u << d) - 1u)) == 0) v |= pv;
    }
    uint32_t in = __shfl_up_sync(kFull, v, 1);
    const uint32_t below = heads & ((1u << lane) - 1u);  // segment starts among earlier lanes
    if (lane == 0) in = 0;
    if (below == 0) in |= carry;
    if (done_one) {
        if (in) stage[first_widx] |= in;  // own word: written by this lane above
    }
    // lanes that finished nothing pass (in | own bits) on: that is what the scan computed
    uint32_t out = __shfl_sync(kFull, v, 31);
    if (heads == 0) out |= carry;
    return out;
}

// (k_pack_wide, the general lane of K3, lives in enc_pack.cuh behind the fast lane whose
// accumulator it shares)

}  // namespace hufb200
