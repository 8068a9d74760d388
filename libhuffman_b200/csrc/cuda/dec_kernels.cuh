// dec_kernels.cuh — sm_100a kernels of the decode path.
//
// The stream has no block index (no compressed-length field, reference src/encoder.c:325-342),
// so block starts are found speculatively and then proven:
//
//   K4a k_find<EMIT>   header-candidate scan over every byte offset (signature: tree_len =
//                      4n+1, tree[0] = 255+n, plausible orig_len), two passes: count per chunk,
//                      then ordered emit.  Offset `first` is always a candidate (true start).
//   K4b k_gather/scan  orig_len of each candidate -> exclusive scan -> output offsets.
//   K4d k_tree         (dec_fast.cuh) lane-per-candidate tree walk for encoder-shaped trees.
//   K5  k_decode       (dec_fast.cuh) the fast lane: chunked warm-up / verify decode of clean
//                      blocks, symbols written once.
//   K5s k_decode_slow  the general lane, for every block the fast lane declines: header parse +
//                      tree_len bound check (src/decoder.c:220-239), tree -> node arrays +
//                      12-bit lookup table in shared memory with the whole acceptance grammar
//                      (replaces huf_tree_deserialize, src/tree.c:138-227), self-synchronising
//                      speculative sub-block decode with a sync-point fix-up loop and a
//                      symbol-count scan (replaces __huf_decode_block, src/decoder.c:34-96),
//                      the reference's error codes.
//   K4c k_verify       chain validation: end(j) must equal candidate(j+1); first violation or
//                      error ends the proven chain.  The host restarts after it if needed, so
//                      the result is exact; speculation only affects speed.
#pragma once

#include "common.cuh"

namespace hufb200 {

constexpr int kFindWarps = 8;
constexpr uint32_t kFindChunk = 32768;  // bytes scanned per warp
constexpr int kDecThreads = 512;
constexpr int kLutBits = 12;
constexpr int kLutSize = 1 << kLutBits;
constexpr int kMaxNodes = 1026;

// LUT entry (u16): bit15 = LONG (low 11 bits: node reached after kLutBits bits),
// bit14 = DEAD (low 4 bits: 1-based depth of the bit that walks into an absent child),
// else  [11:8] code length 1..12, [7:0] symbol.
constexpr uint16_t kLutLong = 0x8000;
constexpr uint16_t kLutDead = 0x4000;

struct DecArgs {
    const uint8_t *in;
    uint64_t avail;      // readable bytes
    uint64_t length;     // blocks may start while offset < length
    uint64_t first;      // offset this pass begins at: a proven block start (first_proven), or
                         // just the lower bound of the byte range whose blocks are wanted
    uint64_t out_base;   // output bytes already produced before `first`
    uint32_t first_proven;
    uint8_t *out;
    uint64_t out_cap;
    uint32_t accept_1025;
    uint32_t count_only; // plan mode: do not write output
    uint32_t stage_cap;  // bytes of dynamic shared memory usable as output staging
    uint32_t mul14;      // 1 << 14, as a run-time value (dec_fast.cuh bulk_idx)
    // workspace
    uint32_t *chunk_cnt;   // [nchunks]
    uint32_t *slots;       // [nchunks][kFindSlots] chunk-relative candidate offsets (sparse mode)
    uint64_t *chunk_off;   // [nchunks + 1]
    uint64_t *cand;        // [max_cand]  candidate byte offsets, ascending
    uint64_t *olen;        // [max_cand]  orig_len per candidate
    uint64_t *out_off;     // [max_cand + 1]
    uint64_t *end_off;     // [max_cand]  byte offset just past the block
    uint32_t *blk_status;  // [max_cand]
    uint32_t *meta;        // [max_cand]  fast-lane eligibility + table facts (dec_fast.cuh)
    uint32_t *terms;       // [term_slots][kTermStride]  ordered table terminals per candidate
    uint64_t term_slots;   // candidates that own a terminal slot (the rest take the slow lane)
    uint64_t max_cand;
    uint64_t nchunks;
    uint64_t *result;      // [0] ncand [1] proven blocks [2] status [3] consumed [4] out bytes
                           // [5] chain complete flag [6] candidates found (may exceed max_cand)
                           // [7] largest orig_len among the candidates
                           // [8] sparse-mode slot overflow (rerun with the two-pass scan)
                           // [9] largest block extent (header + payload bytes) seen
                           // [10] blocks the fast lane left to k_decode_slow
                           // [11] work counter of k_decode (next candidate to hand out)
                           // [12] offset of the first candidate of the pass (~0: none)
};

// The header scan walks 16-byte aligned chunks from the aligned offset at or in front of `first`
// (so that its vector loads stay aligned whatever `first` is); offsets in front of `first` are
// never candidates.
__device__ __host__ __forceinline__ uint64_t find_base(uint64_t first) { return first & ~uint64_t(15); }

// ------------------------------------------------------------------------------------------
// byte-granular, bounds-checked readers
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t rd_u16(const uint8_t *p) { return p[0] | ((uint32_t)p[1] << 8); }

__device__ __forceinline__ uint64_t rd_u64(const uint8_t *p)
{
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

// Full signature test at byte offset `off` (reference encoder output only; foreign streams
// with other tree shapes are still decoded through the chain restart).
__device__ __forceinline__ bool header_plausible(const DecArgs &a, uint64_t off)
{
    if (off + 12 > a.avail) return false;
    const uint8_t *p = a.in + off;
    const uint32_t tl = rd_u16(p + 8);
    const uint32_t t0 = rd_u16(p + 10);
    if ((tl & 3u) != 1u || tl < 5u || tl > 1025u) return false;
    if (t0 != 255u + (tl >> 2)) return false;
    const uint64_t pay0 = off + kHdrFixed + 2ull * tl;
    if (pay0 > a.avail) return false;
    const uint64_t ol = rd_u64(p);
    if (ol == 0) return false;
    // every symbol costs at least one bit
    if (ol > 8ull * (a.avail - pay0)) return false;
    return true;
}

// ------------------------------------------------------------------------------------------
// K4a: candidate scan.  Each warp owns kFindChunk consecutive byte offsets.
// Pre-filter: byte at offset+11 (high byte of tree[0] = 255+n, n in 1..256) must be 0x01.
// ------------------------------------------------------------------------------------------

// MODE 0: count per chunk.  MODE 1: ordered emit into cand[] (after the chunk scan).
// MODE 2: single pass for sparse streams: count and park up to kFindSlots candidates per chunk
// in a slot array (k_compact moves them into cand[]); denser chunks raise result[8] and the
// host reruns with the exact two-pass scheme (MODE 0 + MODE 1).
constexpr uint32_t kFindSlots = 8;

// Borrow-trick prefilter over the 16 offsets of one lane: bytes o+9 .. o+26 are the words d2, d3
// (the lane's own bytes 8..15) and n0, n1, n2 (bytes 16..27, the next lane's first twelve).  A
// zero byte in z marks an offset whose byte o+11 is 0x01 (high byte of tree[0] = 255 + n) and
// whose byte o+9 is below 8 (high byte of tree_len <= 0x04); the subtraction flags every zero
// byte (and, harmlessly, sometimes the byte above one).  1 in 8000 random offsets survives.
__device__ __forceinline__ uint32_t find_prefilter(uint32_t d2, uint32_t d3, uint32_t n0, uint32_t n1, uint32_t n2)
{
    const uint32_t d[5] = {d2, d3, n0, n1, n2};
    uint32_t any = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t w11 = __funnelshift_r(d[j], d[j + 1], 24);
        const uint32_t w9 = __funnelshift_r(d[j], d[j + 1], 8);
        const uint32_t z = (w11 ^ 0x01010101u) | (w9 & 0xf8f8f8f8u);
        any |= (z - 0x01010101u) & ~z;
    }
    return any & 0x80808080u;
}

// Exact candidate mask of a lane whose prefilter fired: the per-offset byte tests, then the full
// signature of every survivor.
__device__ __forceinline__ uint32_t find_exact(const DecArgs &a, uint64_t o0, uint64_t lim, uint32_t d2, uint32_t d3,
                                               uint32_t n0, uint32_t n1, uint32_t n2)
{
    const uint32_t d[5] = {d2, d3, n0, n1, n2};
    uint32_t pre = 0, mask = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t w11 = __funnelshift_r(d[j], d[j + 1], 24);
        const uint32_t w9 = __funnelshift_r(d[j], d[j + 1], 8);
        const uint32_t eq = __vcmpeq4(w11, 0x01010101u) & __vcmpleu4(w9, 0x04040404u);
        pre |= ((eq & 1u) | ((eq >> 7) & 2u) | ((eq >> 14) & 4u) | ((eq >> 21) & 8u)) << (4 * j);
    }
    while (pre) {
        const int i = __ffs(pre) - 1;
        pre &= pre - 1;
        if (o0 + i >= a.first && o0 + i < lim && header_plausible(a, o0 + i)) mask |= 1u << i;
    }
    return mask;
}

// (HUF_FIND_MINCTA: build knob of the occupancy experiment -- CTAs per SM the register allocation must allow)
template <int MODE>
#ifdef HUF_FIND_MINCTA
__global__ void __launch_bounds__(kFindWarps * 32, HUF_FIND_MINCTA) k_find(DecArgs a)
#else
__global__ void __launch_bounds__(kFindWarps * 32) k_find(DecArgs a)
#endif
{
    constexpr bool EMIT = MODE != 0;
    const int lane = lane_id();
    const uint64_t chunk = (uint64_t)blockIdx.x * kFindWarps + warp_in_cta();
    if (chunk >= a.nchunks) return;
    const uint64_t lim = a.length < a.avail ? a.length : a.avail;  // starts must be < lim
    const uint64_t c0 = find_base(a.first) + chunk * kFindChunk;
    if (MODE == 1 && a.chunk_cnt[chunk] == 0) return;

    uint64_t wr = MODE == 1 ? a.chunk_off[chunk] : 0;
    uint32_t total = 0;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(a.in) & 15) == 0;

    const uint64_t in_aligned = a.avail & ~uint64_t(15);  // bytes readable with 16-byte loads
    // the 16-byte group that holds a proven start (none: an offset no group has)
    const uint64_t proven16 = a.first_proven && a.first < lim ? find_base(a.first) : ~uint64_t(0);
    const int nl = (lane + 1) & 31;

    // ordered emit of one row's candidates: lanes in order, offsets in order inside a lane
    auto emit = [&](uint32_t mask, uint64_t o0) {
        const uint32_t n = __popc(mask);
        total += n;
        if (EMIT && __any_sync(kFull, mask != 0)) {
            const uint32_t incl = warp_incl_scan(n);
            uint64_t at = wr + incl - n;
            uint32_t m = mask;
            while (m) {
                const int i = __ffs(m) - 1;
                m &= m - 1;
                if (MODE == 1) {
                    if (at < a.max_cand) a.cand[at] = o0 + i;
                } else if (at < kFindSlots) {
                    a.slots[chunk * kFindSlots + at] = (uint32_t)(o0 + i - c0);
                }
                at++;
            }
            wr += __shfl_sync(kFull, incl, 31);
        }
    };

    // Interior chunk: every offset of it lies behind `first` and in front of `lim`, the bytes its
    // last lane looks ahead into are readable with vector loads, and no proven start sits in it.
    // Nothing has to be tested per row then: the loop streams 512 bytes per row through the
    // prefilter (the row behind is in flight while one is tested) and only a row with a survivor
    // -- one in 16 of a compressed payload -- builds exact masks.
    const uint64_t lim_lean = lim < in_aligned ? lim : in_aligned;
    if (vec_ok && c0 >= a.first && c0 + kFindChunk + 544 <= lim_lean && (proven16 < c0 || proven16 >= c0 + kFindChunk)) {
        const uint4 *p = reinterpret_cast<const uint4 *>(a.in + c0) + lane;
        // (three rows in flight per lane behind the one being tested: 2 KB per warp, which is what
        // it takes to keep HBM busy at this occupancy)
        uint4 cur = ld_stream_u4(p), nxt = ld_stream_u4(p + 32), nx2 = ld_stream_u4(p + 64), nx3 = ld_stream_u4(p + 96);
#pragma unroll 4
        for (uint32_t r = 0; r < kFindChunk / 512; r++) {
            // (the row behind the chunk's last one is readable and feeds lane 31's look-ahead there)
            const uint4 nn = r + 4 <= kFindChunk / 512 ? ld_stream_u4(p + (r + 4) * 32) : make_uint4(0, 0, 0, 0);
            const uint32_t n0 = __shfl_sync(kFull, lane == 0 ? nxt.x : cur.x, nl);
            const uint32_t n1 = __shfl_sync(kFull, lane == 0 ? nxt.y : cur.y, nl);
            const uint32_t n2 = __shfl_sync(kFull, lane == 0 ? nxt.z : cur.z, nl);
            const uint32_t hit = find_prefilter(cur.z, cur.w, n0, n1, n2);
            if (__any_sync(kFull, hit != 0)) {
                const uint64_t o0 = c0 + (uint64_t)r * 512 + lane * 16;
                emit(hit ? find_exact(a, o0, lim, cur.z, cur.w, n0, n1, n2) : 0u, o0);
            }
            cur = nxt;
            nxt = nx2;
            nx2 = nx3;
            nx3 = nn;
        }
    } else {
    for (uint32_t it0 = 0; it0 < kFindChunk / 512; it0 += 4) {
        if (c0 + (uint64_t)it0 * 512 >= lim) break;  // warp-uniform: nothing left in this chunk
        // four independent 16-byte loads in flight per lane; lane 0 also fetches the 16 bytes
        // behind the fourth row (what lane 31 needs to look past its own bytes there)
        uint4 v[5];
#pragma unroll
        for (int u = 0; u < 5; u++) {
            const uint64_t o0 = c0 + (uint64_t)(it0 + u) * 512 + lane * 16;
            v[u] = make_uint4(0, 0, 0, 0);
            if ((u < 4 || lane == 0) && vec_ok && o0 + 16 <= in_aligned) v[u] = ld_stream_u4(a.in + o0);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint64_t o0 = c0 + (uint64_t)(it0 + u) * 512 + lane * 16;  // offsets o0 .. o0+15
            uint32_t mask = 0;                                               // bit i: o0+i is a candidate
            // the 16 bytes that follow come from the next lane; lane 31's are lane 0's of the next
            // row (already in registers): lane 0 offers that row to the rotating shuffle
            const uint32_t n0 = __shfl_sync(kFull, lane == 0 ? v[u + 1].x : v[u].x, nl);
            const uint32_t n1 = __shfl_sync(kFull, lane == 0 ? v[u + 1].y : v[u].y, nl);
            const uint32_t n2 = __shfl_sync(kFull, lane == 0 ? v[u + 1].z : v[u].z, nl);
            if (o0 < lim) {
                if (vec_ok && o0 + 32 <= in_aligned) {
                    if (find_prefilter(v[u].z, v[u].w, n0, n1, n2))
                        mask = find_exact(a, o0, lim, v[u].z, v[u].w, n0, n1, n2);
                } else {
                    for (int i = 0; i < 16; i++) {
                        const uint64_t o = o0 + i;
                        if (o >= a.first && o < lim && o + 12 <= a.avail && a.in[o + 11] == 1 && header_plausible(a, o))
                            mask |= 1u << i;
                    }
                }
                if (o0 == proven16) mask |= 1u << (uint32_t)(a.first & 15);  // a proven start is always block 0
            }
            emit(mask, o0);
        }
    }
    }
    if (MODE != 1) {
        total = warp_sum(total);
        if (lane == 0) {
            a.chunk_cnt[chunk] = total;
            if (MODE == 2 && total > kFindSlots) a.result[8] = 1;
        }
    }
}

// MODE 2 follow-up: slots -> cand[], in stream order.
__global__ void k_compact(DecArgs a)
{
    const uint64_t chunk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (chunk >= a.nchunks) return;
    const uint32_t n = min(a.chunk_cnt[chunk], kFindSlots);
    const uint64_t at = a.chunk_off[chunk];
    const uint64_t c0 = find_base(a.first) + chunk * kFindChunk;
    for (uint32_t i = 0; i < n; i++) {
        if (at + i < a.max_cand) a.cand[at + i] = c0 + a.slots[chunk * kFindSlots + i];
    }
}

// Candidate list from a caller-supplied block index (the offset array an encoder returns,
// huf_b200_encode_block_offsets) instead of the header scan.  The entries are only hints: the
// chain validation still proves every block, and a wrong index just costs a restart with the
// scan.  Offsets must ascend; those at or behind the consumable range are dropped.
__global__ void k_hint(DecArgs a, const uint64_t *__restrict__ hint, uint64_t n)
{
    const uint64_t lim = a.length < a.avail ? a.length : a.avail;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && i < a.max_cand) {
        const uint64_t off = i == 0 ? a.first : hint[i];  // the proven start is always block 0 (hints are only used with one)
        a.cand[i] = off < lim ? off : lim;
    }
    if (i == 0) {
        // count = entries in front of `lim` (binary search over the ascending array)
        uint64_t lo = 0, hi = n < a.max_cand ? n : a.max_cand;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if ((mid == 0 ? a.first : hint[mid]) < lim) lo = mid + 1; else hi = mid;
        }
        a.result[0] = lo ? lo : 1;
        a.result[6] = lo ? lo : 1;
    }
}

// chunk counts -> exclusive offsets, total candidate count -> result[0].  One CTA.
__global__ void __launch_bounds__(kScanThreads) k_scan_chunks(DecArgs a)
{
    __shared__ uint64_t warp_tot[kScanThreads / 32 + 1];
    const uint64_t n = a.nchunks;
    const uint64_t run = cta_excl_scan(a.chunk_cnt, a.chunk_off, n, 0, warp_tot);
    if (threadIdx.x == 0) {
        a.chunk_off[n] = run;
        a.result[0] = run < a.max_cand ? run : a.max_cand;
        a.result[6] = run;  // > max_cand: workspace too small, the host re-sizes and reruns
    }
}

// K4b: orig_len of every candidate (bounded so that the scan cannot overflow).
__global__ void k_gather(DecArgs a)
{
    const uint64_t n = a.result[0];
    // the two maxima are reduced per warp before they touch the shared counters (one atomic
    // per candidate on the same address serialises in L2: 25 us for 16 K candidates)
    unsigned long long max_ol = 0, max_ext = 0;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n;
         j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t off = a.cand[j];
        uint64_t ol = 0;
        if (off + 8 <= a.avail) ol = rd_u64(a.in + off);
        // a block that cannot complete inside the readable bytes produces no counted output
        const uint64_t room = a.avail > off ? a.avail - off : 0;
        if (ol > 8ull * room) ol = 0;
        a.olen[j] = ol;
        if (ol > max_ol) max_ol = ol;
        // extent of this block's header + payload as the candidate list sees it: sizes the
        // shared-memory payload staging of k_decode on the next call
        const uint64_t next = j + 1 < n ? a.cand[j + 1] : a.avail;
        if (next > off && next - off > max_ext) max_ext = next - off;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        const unsigned long long o1 = __shfl_xor_sync(kFull, max_ol, d), o2 = __shfl_xor_sync(kFull, max_ext, d);
        if (o1 > max_ol) max_ol = o1;
        if (o2 > max_ext) max_ext = o2;
    }
    if (lane_id() == 0) {
        if (max_ol) atomicMax(reinterpret_cast<unsigned long long *>(&a.result[7]), max_ol);
        if (max_ext) atomicMax(reinterpret_cast<unsigned long long *>(&a.result[9]), max_ext);
    }
}

// exclusive scan of olen[0..ncand) -> out_off, with ncand read from device memory.
__global__ void __launch_bounds__(kScanThreads) k_scan_olen(DecArgs a)
{
    __shared__ uint64_t warp_tot[kScanThreads / 32 + 1];
    const uint64_t n = a.result[0];
    const uint64_t run = cta_excl_scan(a.olen, a.out_off, n, a.out_base, warp_tot);
    if (threadIdx.x == 0) a.out_off[n] = run;
}

// ------------------------------------------------------------------------------------------
// K5: block decode.
// ------------------------------------------------------------------------------------------

constexpr int kMaxElems = 1040;   // serialised tree elements (<= 1025) padded

struct DecSmem {
    int16_t elems[kMaxElems];    // serialised tree as read from the stream; node id == element index
    int16_t open[kMaxElems];     // open child slots before element i is consumed (1 for i == 0)
    int16_t open_min[kMaxElems / 32 + 2];  // minimum of open[] per aligned run of 32 elements
    int16_t lch[kMaxElems];      // child node ids, -1 = absent
    int16_t rch[kMaxElems];
    int16_t rraw[kMaxElems];     // element index that fills the right slot (-1: elements ran out)
    uint16_t pre[kMaxElems];     // code prefix (depth bits) of nodes at depth <= kLutBits
    uint8_t lvl[2][kMaxElems];   // [d & 1][i] == d: node i sits at depth d (0xff: deeper than the table or
                                 // unused).  Two planes: round d reads plane d & 1 and marks the children in
                                 // the other one, so no thread reads a byte another one writes in that round
    uint16_t lut[kLutSize + 2];  // [kLutSize] = sentinel: first bit walks off a one-child root
    uint32_t sub_end[kDecThreads + 1];
    uint32_t warp_tot[kDecThreads / 32];
    uint32_t n_term;
    uint32_t n_wide;
    uint32_t n_eff;              // elements that belong to the tree
    uint32_t skip;               // 1: the root has only a left child, the table starts below it
    int32_t root;
    uint32_t hdr_status;
    uint32_t tree_len;
    uint64_t orig_len;
    uint32_t err_pos;     // smallest payload bit position at which the true chain failed
    uint32_t end_bit;
    uint64_t total_syms;
};

// Terminals of the depth-limited tree (leaf, absent child, or inner node at table depth) are
// collected in the dynamic shared memory area, which is free until the decode phases start.
struct Terminal {
    uint16_t start;   // first LUT index covered
    uint16_t entry;
    uint16_t depth;
    uint16_t pad;
};

// Stateless MSB-first bit access to one block's payload, addressed by payload-relative bit
// position.  SMEM = true: the payload was staged into shared memory as big-endian 32-bit
// words (bit 0 of word 0 is `bias` bits in front of payload bit 0), so a 32-bit window costs
// two LDS and one funnel shift and carries no per-thread state: every lane of a warp runs
// the same instructions.  SMEM = false reads the bytes from global memory, bounds checked
// (used when the payload does not fit and for the serial tail walk).
template <bool SMEM>
struct Bits {
    const uint32_t *sw;   // staged words
    uint32_t bias;        // bits between staged bit 0 and payload bit 0
    uint32_t last_word;   // highest index that may be read (words beyond read as the last one)
    uint32_t lshift;      // 32 - kLutBits - skip: table index = window >> lshift, capped at kLutSize
    const uint8_t *in;    // whole stream (global)
    uint64_t avail;
    uint64_t pay0;

    // 32 stream bits starting at payload bit `pos`.  CLAMP = false is for callers that keep
    // pos inside the staged extent by construction.
    template <bool CLAMP = true>
    __device__ __forceinline__ uint32_t window(uint32_t pos) const
    {
        if (SMEM) {
            const uint32_t p = pos + bias;
            uint32_t byte = (p >> 3) & ~3u;
            if (CLAMP) byte = min(byte, last_word << 2);
            const uint32_t *w = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(sw) + byte);
            return __funnelshift_l(w[1], w[0], p & 31);
        }
        const uint64_t byte = pay0 + (pos >> 3);
        uint64_t v = 0;
#pragma unroll
        for (int q = 0; q < 5; q++) {
            const uint64_t at = byte + q;
            v = (v << 8) | (at < avail ? in[at] : 0u);
        }
        return (uint32_t)(v >> (8 - (pos & 7)));
    }
};

// One decode step at payload bit position `pos`.  Returns the symbol (>= 0) and advances, or
// -1 when the walk dies: then `pos` advances by one bit (any deterministic rule works for a
// speculative start; on a proven start the caller records the error) and *dead_at is the bit
// whose consumption walks into the absent child.
template <bool SMEM, bool CLAMP = true>
__device__ __forceinline__ int decode_one(const DecSmem &sm, const Bits<SMEM> &bits, uint32_t &pos,
                                          uint32_t *dead_at)
{
    // a root with only a left child (every tree the reference encoder emits) costs one bit
    // per code word that carries no information: the table is indexed behind it, and a set
    // first bit lands on the sentinel entry
    const uint16_t e = sm.lut[min(bits.template window<CLAMP>(pos) >> bits.lshift, (uint32_t)kLutSize)];
    if (!(e & (kLutLong | kLutDead))) {
        pos += e >> 8;
        return e & 0xff;
    }
    if (e & kLutDead) {
        *dead_at = pos + (e & 0xf) - 1;
        pos += 1;
        return -1;
    }
    // long code: continue bit by bit from the node reached after kLutBits bits
    int node = e & 0x7ff;
    uint32_t p = pos + (32 - bits.lshift);
    for (;;) {
        const int bit = bits.window(p) >> 31;
        const int nx = bit ? sm.rch[node] : sm.lch[node];
        if (nx < 0) {
            *dead_at = p;
            pos += 1;  // same rule as a table miss: resume one bit after the failed start
            return -1;
        }
        p++;
        node = nx;
        if (sm.lch[node] < 0 && sm.rch[node] < 0) {
            // keep the caller inside the staged extent even on a runaway (corrupt) walk
            pos = SMEM ? min(p, (bits.last_word << 5) - bits.bias - 64) : p;
            return (uint8_t)sm.elems[node];
        }
    }
}

// Count symbols from `pos` until the position reaches `limit` (inside the staged extent).
template <bool SMEM>
__device__ __forceinline__ uint32_t count_span(const DecSmem &sm, const Bits<SMEM> &bits,
                                               uint32_t pos, uint32_t limit, uint32_t *end)
{
    uint32_t n = 0, d;
    while (pos < limit) n += decode_one<SMEM, false>(sm, bits, pos, &d) >= 0;
    *end = pos;
    return n;
}

// Decode `need` symbols from the proven position `pos` into dst (nullptr: nowhere).  Output
// leaves in 16-byte stores once dst is aligned.  A dead walk is an error of the stream: its
// first bit is returned in *dead (the step then counts as a symbol so that the walk stays
// bounded; the block is reported as failed anyway).  Returns the end position.
template <bool SMEM>
__device__ __forceinline__ uint32_t emit_span(const DecSmem &sm, const Bits<SMEM> &bits,
                                              uint32_t pos, uint32_t need, uint8_t *dst,
                                              uint32_t *dead)
{
    uint32_t n = 0, first_dead = 0xffffffffu;
    auto one = [&]() -> uint32_t {
        uint32_t d = 0;
        const int sy = decode_one<SMEM, false>(sm, bits, pos, &d);
        if (sy < 0 && first_dead == 0xffffffffu) first_dead = d;
        return (uint32_t)sy & 0xffu;
    };
    if (!dst) {
        while (n < need) {
            one();
            n++;
        }
    } else {
        const uint32_t head = min(need, (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15));
        while (n < head) dst[n++] = (uint8_t)one();
        while (need - n >= 16) {
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                w[q] = one();
                w[q] |= one() << 8;
                w[q] |= one() << 16;
                w[q] |= one() << 24;
            }
            *reinterpret_cast<uint4 *>(dst + n) = make_uint4(w[0], w[1], w[2], w[3]);
            n += 16;
        }
        while (n < need) dst[n++] = (uint8_t)one();
    }
    *dead = first_dead;
    return pos;
}

// The decode phases of one block, for either payload source.
template <bool SMEM>
__device__ __forceinline__ void decode_phases(DecSmem &sm, const Bits<SMEM> &bits, const DecArgs &a,
                                              uint64_t j, uint64_t orig_len, bool doomed,
                                              uint32_t cover_bits, uint32_t sub)
{
    const int tid = threadIdx.x;
    // sub-blocks tile the guessed extent only: what lies behind it is the next block's header,
    // which this block's table would chew through one dead bit at a time
    const uint32_t my_lo = (uint32_t)min((uint64_t)tid * sub, (uint64_t)cover_bits);
    const uint32_t my_hi = (uint32_t)min((uint64_t)(tid + 1) * sub, (uint64_t)cover_bits);
    // phase 1: speculative count of every sub-block from its nominal start
    uint32_t start = my_lo, end = my_lo, cnt = 0;
    if (my_lo < my_hi) cnt = count_span(sm, bits, start, my_hi, &end);
    sm.sub_end[tid] = end;
    __syncthreads();

    // phase 2: sync-point fix-up.  Thread t's true start is where thread t-1 ended.  When that
    // differs from the start it used, it walks the new and the old trajectory in lockstep only
    // until they meet (Huffman codes re-synchronise within a few symbols); from there on the
    // old count and end stay valid.  Thread t is final after at most t rounds, in practice two.
    for (int round = 0; round < kDecThreads; round++) {
        const uint32_t want = tid == 0 ? 0u : sm.sub_end[tid - 1];
        const bool redo = tid > 0 && want != start;
        __syncthreads();
        if (redo) {
            if (want >= my_hi) {
                cnt = 0;
                end = want;
            } else {
                const bool has_old = start < my_hi;
                uint32_t pa = want, pb = start, ca = 0, cb = 0, d;
                // advance whichever trajectory is behind until both stand on the same bit
                while (pa < my_hi && !(has_old && pa == pb)) {
                    if (!has_old || pa < pb || pb >= my_hi) {
                        ca += decode_one<SMEM, false>(sm, bits, pa, &d) >= 0;
                    } else {
                        cb += decode_one<SMEM, false>(sm, bits, pb, &d) >= 0;
                    }
                }
                if (has_old && pa == pb && pa < my_hi) {
                    cnt = ca + (cnt - cb);  // merged: the rest of the old walk is reused
                } else {
                    cnt = ca;               // ran to the boundary on its own
                    end = pa;
                }
            }
            start = want;
            sm.sub_end[tid] = end;
        }
        if (!__syncthreads_or(redo)) break;
    }

    // phase 3: symbol-count scan -> output index of every sub-block
    const uint32_t incl = warp_incl_scan(cnt);
    if ((tid & 31) == 31) sm.warp_tot[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        const uint32_t t = tid < kDecThreads / 32 ? sm.warp_tot[tid] : 0;
        const uint32_t ti = warp_incl_scan(t);
        if (tid < kDecThreads / 32) sm.warp_tot[tid] = ti - t;
    }
    __syncthreads();
    const uint64_t before = (uint64_t)sm.warp_tot[tid >> 5] + incl - cnt;  // symbols before mine
    if (tid == kDecThreads - 1) sm.total_syms = before + cnt;

    const uint64_t out0 = a.out_off[j];
    const bool can_write = !a.count_only && !doomed && out0 + orig_len <= a.out_cap;

    // phase 4: every start is proven now: decode for real, straight into the output.  A dead
    // walk met on the way is an error of the stream, not of the speculation, so sub-blocks in
    // front of the block's last symbol are walked to their very end.
    if (before < orig_len && start < end) {
        const bool finisher = before + cnt >= orig_len;  // the block completes in here
        const uint32_t need = finisher ? (uint32_t)(orig_len - before) : cnt;
        uint8_t *dst = can_write ? a.out + out0 + before : nullptr;
        uint32_t dead;
        uint32_t pos = emit_span(sm, bits, start, need, dst, &dead);
        if (finisher) {
            sm.end_bit = pos;
        } else {
            uint32_t d = 0xffffffffu;
            while (pos < end) {  // only dead bits can be left in here
                uint32_t dd = 0;
                if (decode_one(sm, bits, pos, &dd) < 0 && d == 0xffffffffu) d = dd;
            }
            if (dead == 0xffffffffu) dead = d;
        }
        if (dead != 0xffffffffu) atomicMin(&sm.err_pos, dead);
    }
    __syncthreads();
}

// The general lane: every block the fast lane (dec_fast.cuh) marked kRedoStatus -- foreign tree
// shapes, deep trees, corrupt or truncated blocks -- with the full acceptance grammar and the
// reference's error codes.
constexpr uint32_t kRedoStatus = 0x80;

__global__ void __launch_bounds__(kDecThreads) k_decode_slow(DecArgs a)
{
#ifdef HUF_EMU
    uint8_t *stage = hufemu::dyn_smem();
#else
    extern __shared__ __align__(16) uint8_t stage[];
#endif
    __shared__ DecSmem sm;
    const int tid = threadIdx.x;
    const uint64_t ncand = a.result[0];
    if (a.result[10] == 0) return;  // nothing was left over

    for (uint64_t j = blockIdx.x; j < ncand; j += gridDim.x) {
        if (a.blk_status[j] != kRedoStatus) continue;
        __syncthreads();
        const uint64_t off = a.cand[j];
        const uint64_t next_cand = (j + 1 < ncand) ? a.cand[j + 1] : a.avail;

        // ---- header (src/decoder.c:220-239)
        if (tid == 0) {
            uint32_t st = kOk;
            uint64_t ol = 0;
            uint32_t tl = 0;
            if (off + 8 > a.avail) {
                st = kErrIO;
            } else {
                ol = rd_u64(a.in + off);
                if (off + 10 > a.avail) {
                    st = kErrIO;
                } else {
                    tl = rd_u16(a.in + off + 8);
                    if (tl >= 0x8000u || tl > (a.accept_1025 ? 1025u : 1024u)) {
                        st = kErrOverflow;
                    } else if (off + 10 + 2ull * tl > a.avail) {
                        st = kErrIO;
                    }
                }
            }
            sm.hdr_status = st;
            sm.orig_len = ol;
            sm.tree_len = tl;
            sm.err_pos = 0xffffffffu;
            sm.end_bit = 0;
            sm.n_term = 0;
            sm.n_wide = 0;
            sm.n_eff = tl;
            sm.root = -1;
            sm.total_syms = 0;
        }
        __syncthreads();
        if (sm.hdr_status != kOk) {
            if (tid == 0) {
                a.blk_status[j] = sm.hdr_status;
                a.end_off[j] = off;
            }
            continue;
        }
        const uint32_t tl = sm.tree_len;
        const uint64_t orig_len = sm.orig_len;
        const uint64_t pay0 = off + kHdrFixed + 2ull * tl;

        // ---- tree -> node arrays, in parallel (grammar of src/tree.c:138-208: T := -1 | v T T,
        // missing elements = absent children, trailing elements ignored).
        // open[i] = child slots still open before element i is consumed: a node consumes one
        // and opens two, an absent marker consumes one.  The tree ends where open hits 0.
        constexpr int kPer = (kMaxElems + kDecThreads - 1) / kDecThreads;  // 5 elements per thread
        {
            int v[kPer];
            int sum = 0;
#pragma unroll
            for (int q = 0; q < kPer; q++) {
                const uint32_t i = tid * kPer + q;
                int e = -1;
                if (i < tl) {
                    e = (int16_t)rd_u16(a.in + off + kHdrFixed + 2ull * i);
                    sm.elems[i] = (int16_t)e;
                }
                v[q] = i < tl ? (e != -1 ? 1 : -1) : 0;
                sum += v[q];
            }
            int incl = warp_incl_scan(sum);
            if ((tid & 31) == 31) sm.warp_tot[tid >> 5] = (uint32_t)incl;
            __syncthreads();
            if (tid < 32) {
                const int t = tid < kDecThreads / 32 ? (int)sm.warp_tot[tid] : 0;
                const int ti = warp_incl_scan(t);
                if (tid < kDecThreads / 32) sm.warp_tot[tid] = (uint32_t)(ti - t);
            }
            __syncthreads();
            int run = 1 + (int)sm.warp_tot[tid >> 5] + incl - sum;
#pragma unroll
            for (int q = 0; q < kPer; q++) {
                const uint32_t i = tid * kPer + q;
                if (i < tl) {
                    sm.open[i] = (int16_t)run;
                    if (run == 0) atomicMin(&sm.n_eff, i);
                    sm.lvl[0][i] = 0xff;
                    sm.lvl[1][i] = 0xff;
                }
                run += v[q];
            }
        }
        for (uint32_t i = tid; i < kLutSize; i += kDecThreads) sm.lut[i] = kLutDead | 1;
        __syncthreads();
        const uint32_t n_eff = sm.n_eff;

        // minimum of open[] over every aligned run of 32 elements: the right-child search
        // below skips whole runs that cannot contain its target
        for (uint32_t blk = tid >> 5; blk * 32 < n_eff; blk += kDecThreads / 32) {
            const uint32_t i = blk * 32 + (tid & 31);
            int v = i < n_eff ? (int)sm.open[i] : 0x7fff;
#pragma unroll
            for (int dd = 16; dd; dd >>= 1) v = min(v, __shfl_xor_sync(kFull, v, dd));
            if ((tid & 31) == 0) sm.open_min[blk] = (int16_t)v;
        }
        __syncthreads();

        // children: left slot is filled by the next element; the right slot by the first later
        // element that sees the same number of open slots as this node did.
        for (uint32_t i = tid; i < n_eff; i += kDecThreads) {
            int l = -1, r = -1, rr = -1;
            if (sm.elems[i] != -1) {
                if (i + 1 < n_eff && sm.elems[i + 1] != -1) l = (int)(i + 1);
                const int want = sm.open[i];
                uint32_t k = i + 2;
                // finish the current run of 32, then hop over runs whose minimum is too high
                while (k < n_eff && (k & 31) && sm.open[k] > want) k++;
                if (k < n_eff && (k & 31) == 0 && sm.open[k] > want) {
                    while (k < n_eff && sm.open_min[k >> 5] > want) k += 32;
                    while (k < n_eff && sm.open[k] > want) k++;
                }
                if (k < n_eff && i + 1 < n_eff) {
                    rr = (int)k;
                    if (sm.elems[k] != -1) r = (int)k;
                }
            }
            sm.lch[i] = (int16_t)l;
            sm.rch[i] = (int16_t)r;
            sm.rraw[i] = (int16_t)rr;
        }
        __syncthreads();
        if (tid == 0) {
            sm.skip = 0;
            sm.lut[kLutSize] = kLutDead | 1;
            if (n_eff > 0 && sm.elems[0] != -1) {
                sm.root = 0;
                if (sm.lch[0] >= 0 && sm.rch[0] < 0) {
                    sm.skip = 1;              // table root = the only child, one bit down
                    sm.lvl[1][sm.lch[0]] = 1;
                    sm.pre[sm.lch[0]] = 0;
                } else {
                    sm.lvl[0][0] = 0;
                    sm.pre[0] = 0;
                }
            }
        }
        __syncthreads();
        const uint32_t skip = sm.skip;

        // depth-limited expansion, one level per round: leaves and absent children become
        // table terminals, inner nodes at the table depth become long-code continuations.
        Terminal *term = reinterpret_cast<Terminal *>(stage);
        uint16_t *wide = reinterpret_cast<uint16_t *>(stage + sizeof(Terminal) * (2 * kMaxElems + 8));
        for (uint32_t d = skip; d <= (uint32_t)kLutBits + skip; d++) {
            const uint32_t de = d - skip;  // depth below the table root
            for (uint32_t i = tid; i < n_eff; i += kDecThreads) {
                if (sm.lvl[d & 1][i] != d) continue;
                const uint32_t p = sm.pre[i];
                const int l = sm.lch[i], r = sm.rch[i];
                const bool leaf = l < 0 && r < 0;
                uint32_t n_new = 0;
                Terminal t[2];
                if (leaf) {
                    if (d >= 1) {
                        t[0].start = (uint16_t)(p << (kLutBits - de));
                        t[0].entry = (uint16_t)((d << 8) | (uint8_t)sm.elems[i]);
                        t[0].depth = (uint16_t)de;
                        n_new = 1;
                    }  // a root without children keeps the all-dead table
                } else if (de == (uint32_t)kLutBits) {
                    t[0].start = (uint16_t)p;
                    t[0].entry = (uint16_t)(kLutLong | i);
                    t[0].depth = (uint16_t)de;
                    n_new = 1;
                } else {
                    const int kids[2] = {l, r};
#pragma unroll
                    for (int side = 0; side < 2; side++) {
                        const uint32_t cp = (p << 1) | (uint32_t)side;
                        if (kids[side] >= 0) {
                            sm.lvl[(d + 1) & 1][kids[side]] = (uint8_t)(d + 1);
                            sm.pre[kids[side]] = (uint16_t)cp;
                        } else {
                            // consuming this bit walks into an absent child
                            t[n_new].start = (uint16_t)(cp << (kLutBits - de - 1));
                            t[n_new].entry = (uint16_t)(kLutDead | (d + 1));
                            t[n_new].depth = (uint16_t)(de + 1);
                            n_new++;
                        }
                    }
                }
                if (n_new) {
                    const uint32_t at = atomicAdd(&sm.n_term, n_new);
                    for (uint32_t q = 0; q < n_new; q++) {
                        term[at + q] = t[q];
                        if (t[q].depth < (uint32_t)kLutBits - 6) wide[atomicAdd(&sm.n_wide, 1u)] = (uint16_t)(at + q);
                    }
                }
            }
            __syncthreads();
        }

        // ---- fill the lookup table from the terminal list (ranges are disjoint)
        {
            const uint32_t nt = sm.n_term;
            for (uint32_t t = tid; t < nt; t += kDecThreads) {
                const Terminal tm = term[t];
                const uint32_t span = 1u << (kLutBits - tm.depth);
                if (span <= 64) {
                    for (uint32_t i = 0; i < span; i++) sm.lut[tm.start + i] = tm.entry;
                }
            }
            // wide ranges (depth <= 5, at most 63 of them): all threads cooperate
            const uint32_t nw = sm.n_wide;
            for (uint32_t w = 0; w < nw; w++) {
                const Terminal tm = term[wide[w]];
                const uint32_t span = 1u << (kLutBits - tm.depth);
                for (uint32_t i = tid; i < span; i += kDecThreads) sm.lut[tm.start + i] = tm.entry;
            }
        }
        __syncthreads();

        // ---- trivial / impossible blocks
        if (orig_len == 0) {
            if (tid == 0) {
                a.blk_status[j] = kOk;
                a.end_off[j] = pay0;
            }
            continue;
        }
        if (sm.root < 0) {  // Q4: absent root with symbols to produce
            if (tid == 0) {
                a.blk_status[j] = kErrCorrupt;
                a.end_off[j] = pay0;
            }
            continue;
        }
        // Bit positions inside a block are 32-bit.  The readable room behind the header is
        // clipped accordingly; a block that really needs more than ~512 MiB of payload is
        // refused (HUF_ERROR_FATAL) instead of being mis-decoded.
        const uint64_t pay_room_full = 8ull * (a.avail - pay0);
        const bool room_clipped = pay_room_full > 0xffffff00ull;
        const uint64_t pay_room_bits = room_clipped ? 0xffffff00ull : pay_room_full;
        const bool doomed = orig_len > pay_room_full;  // cannot finish: error is EOF or dead walk
        if (!doomed && orig_len >= 0xfffffff0ull) {
            if (tid == 0) {
                a.blk_status[j] = kErrFatal;
                a.end_off[j] = pay0;
            }
            continue;
        }

        // ---- sub-block split over the guessed payload extent
        uint64_t guess_bytes = next_cand > pay0 ? next_cand - pay0 : 0;
        if (guess_bytes > a.avail - pay0) guess_bytes = a.avail - pay0;
        const uint32_t guess_bits = (uint32_t)min(8 * guess_bytes, pay_room_bits);
        const uint32_t room_bits = (uint32_t)pay_room_bits;
        uint32_t sub = (guess_bits + kDecThreads - 1) / kDecThreads;
        sub = (sub + 31u) & ~31u;
        if (sub < 64) sub = 64;

        // ---- stage the payload into shared memory as big-endian words (coalesced 16-byte
        // loads, bytes past `avail` read as zero), then run the decode phases on it
        const uint64_t base16 = pay0 & ~uint64_t(15);
        const uint32_t cover_bits = guess_bits;                          // bits the threads cover
        const uint64_t want_bytes = (pay0 - base16) + (((uint64_t)cover_bits + 7) >> 3) + 64;  // + slack for unclamped windows
        const uint64_t want_chunks = (want_bytes + 15) >> 4;
        const bool staged = want_chunks * 16 + 16 <= a.stage_cap &&
                            (reinterpret_cast<uintptr_t>(a.in) & 15) == 0;
        if (staged) {
            uint32_t *sw = reinterpret_cast<uint32_t *>(stage);
            for (uint32_t c = tid; c < (uint32_t)want_chunks; c += kDecThreads) {
                const uint64_t byte = base16 + 16ull * c;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (byte + 16 <= a.avail) {
                    v = ld_stream_u4(a.in + byte);
                } else if (byte < a.avail) {
                    uint32_t w[4] = {0, 0, 0, 0};
                    for (int q = 0; q < 16; q++) {
                        if (byte + q < a.avail) w[q >> 2] |= (uint32_t)a.in[byte + q] << (8 * (q & 3));
                    }
                    v = make_uint4(w[0], w[1], w[2], w[3]);
                }
                reinterpret_cast<uint4 *>(sw)[c] =
                    make_uint4(bswap32(v.x), bswap32(v.y), bswap32(v.z), bswap32(v.w));
            }
            if (tid == 0) {  // one spare word so that window() may read index + 1
                sw[want_chunks * 4] = 0;
            }
            __syncthreads();
            Bits<true> bits;
            bits.sw = sw;
            bits.bias = (uint32_t)(8 * (pay0 - base16));
            bits.last_word = (uint32_t)(want_chunks * 4 - 1);
            bits.lshift = 32 - kLutBits - skip;
            bits.in = a.in;
            bits.avail = a.avail;
            bits.pay0 = pay0;
            decode_phases<true>(sm, bits, a, j, orig_len, doomed, cover_bits, sub);
        } else {
            Bits<false> bits;
            bits.sw = nullptr;
            bits.bias = 0;
            bits.last_word = 0;
            bits.lshift = 32 - kLutBits - skip;
            bits.in = a.in;
            bits.avail = a.avail;
            bits.pay0 = pay0;
            decode_phases<false>(sm, bits, a, j, orig_len, doomed, cover_bits, sub);
        }

        // The guessed extent was short (the next candidate was a false positive inside this
        // payload) or the block cannot complete: the chain runs past the last sub-block, whose
        // end is proven by now, and one thread walks on serially from global memory.  Rare.
        if (tid == 0 && sm.total_syms < orig_len) {
            const uint32_t from = sm.sub_end[kDecThreads - 1];
            const uint64_t have = sm.total_syms;
            const uint64_t rest = orig_len - have;
            const uint64_t out0 = a.out_off[j];
            const bool can_write = !a.count_only && !doomed && out0 + orig_len <= a.out_cap;
            Bits<false> bits;
            bits.sw = nullptr;
            bits.bias = 0;
            bits.last_word = 0;
            bits.lshift = 32 - kLutBits - skip;
            bits.in = a.in;
            bits.avail = a.avail;
            bits.pay0 = pay0;
            uint32_t pos = from, got = 0, dead = 0xffffffffu;
            uint8_t *dst = can_write ? a.out + out0 + have : nullptr;
            while (pos < room_bits && got < rest) {
                uint32_t d = 0;
                const int sy = decode_one(sm, bits, pos, &d);
                if (sy >= 0) {
                    if (dst) dst[got] = (uint8_t)sy;
                    got++;
                } else if (dead == 0xffffffffu) {
                    dead = d;
                }
            }
            if (dead != 0xffffffffu) atomicMin(&sm.err_pos, dead);
            if (got == rest) {
                sm.end_bit = pos;
            } else {
                atomicMin(&sm.err_pos, room_bits);  // ran out of readable bytes
            }
        }
        __syncthreads();

        // A dead walk at a bit the reader could still deliver is BTREE_CORRUPTED; running out
        // of readable bytes first is READ_WRITE (src/decoder.c:53,69-71).
        uint32_t status = kOk;
        if (sm.err_pos != 0xffffffffu) {
            status = sm.err_pos >= room_bits ? (room_clipped ? kErrFatal : kErrIO) : kErrCorrupt;
        } else if (sm.end_bit > room_bits) {
            // the finishing code word extends past the readable bytes
            status = room_clipped ? kErrFatal : kErrIO;
        }
        if (tid == 0) {
            a.blk_status[j] = status;
            a.end_off[j] = pay0 + (((uint64_t)sm.end_bit + 7) >> 3);
        }
    }
}

// ------------------------------------------------------------------------------------------
// K4c: chain validation.  One CTA.
// result[1] = number of proven blocks, [2] = status of the first failing block (or OK),
// [3] = byte offset reached, [4] = output bytes produced by the proven blocks,
// [5] = 1 when the chain reached `length` (or an error) and no restart is needed.
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(kScanThreads) k_verify(DecArgs a)
{
    __shared__ unsigned long long first_bad;
    const uint64_t n = a.result[0];
    if (threadIdx.x == 0) first_bad = n;
    __syncthreads();
    // (four candidates per thread and round, all their loads in flight together: the kernel is
    // one CTA and a chain of dependent memory round trips otherwise)
    for (uint64_t j0 = threadIdx.x; j0 < n; j0 += 4 * kScanThreads) {
        uint32_t st[4];
        uint64_t eo[4], nx[4], oo[4], ol[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint64_t j = j0 + (uint64_t)q * kScanThreads;
            const bool in = j < n;
            st[q] = in ? a.blk_status[j] : (uint32_t)kOk;
            eo[q] = in ? a.end_off[j] : 0;
            nx[q] = in && j + 1 < n ? a.cand[j + 1] : eo[q];
            oo[q] = in ? a.out_off[j] : 0;
            ol[q] = in ? a.olen[j] : 0;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint64_t j = j0 + (uint64_t)q * kScanThreads;
            bool bad = st[q] != kOk;
            if (!bad && j + 1 < n && eo[q] != nx[q]) bad = true;
            // a block whose output does not fit is not proven either
            if (!bad && !a.count_only && oo[q] + ol[q] > a.out_cap) bad = true;
            if (j < n && bad) atomicMin(&first_bad, (unsigned long long)j);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint64_t fb = first_bad;
        uint64_t proven = fb, status = kOk, reached, produced, done = 0;
        if (fb < n && a.blk_status[fb] != kOk) {
            status = a.blk_status[fb];  // blocks before fb are valid, fb fails: stop here
            done = 1;
            reached = fb ? a.end_off[fb - 1] : a.first;
            produced = a.out_off[fb];
        } else if (fb < n && !a.count_only && a.out_off[fb] + a.olen[fb] > a.out_cap) {
            status = kErrNoMem;  // output capacity exhausted (host grows the buffer and resumes)
            done = 1;
            reached = fb ? a.end_off[fb - 1] : a.first;
            produced = a.out_off[fb];
        } else if (fb < n) {
            // block fb decoded fine but does not end where the next candidate starts:
            // fb itself is proven, restart after it
            proven = fb + 1;
            reached = a.end_off[fb];
            produced = a.out_off[fb + 1];
        } else {
            reached = n ? a.end_off[n - 1] : a.first;
            produced = a.out_off[n];
        }
        if (!a.first_proven && n == 0) done = 1;  // no block starts in the byte range: nothing to do
        a.result[12] = n ? a.cand[0] : ~0ull;
        if (!done && reached >= a.length) done = 1;
        if (!done && reached >= a.avail) {
            // `length` asks for another block but the readable bytes end here: the reference
            // fails on the read of the next header (src/decoder.c:220-229, bufio short read)
            status = kErrIO;
            done = 1;
        }
        a.result[1] = proven;
        a.result[2] = status;
        a.result[3] = reached;
        a.result[4] = produced;
        a.result[5] = done;
    }
}

}  // namespace hufb200
