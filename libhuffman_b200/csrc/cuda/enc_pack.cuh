// enc_pack.cuh — K3, the bit packing.
//
// k_pack (fast lane): blocks whose longest code word is <= 16 bits (every block of ordinary data).
// One warp per segment, 16 symbols per lane and iteration.  Code words are looked up with one
// 4-byte shared-memory load (left-aligned code in the top half, length below: one bank word per
// entry), two neighbours are concatenated in registers (<= 32 bits), the lane's bit offset comes
// from a warp scan of the lengths.  Every lane then pushes its eight pairs through a one-word
// accumulator keyed by its running bit position and stores each word it COMPLETES with a plain
// shared-memory store: sixteen code words of at least two bits fill at least one word, so every
// word of the window is completed by exactly one lane, and what a lane leaves unfinished behind
// its last word boundary travels to its right neighbour through one shuffle and is OR-ed into
// that lane's first word (lane 31's remainder is the carry into the next iteration).  No atomics,
// no zeroing of the window.  Whole 16-byte lines leave as coalesced big-endian stores through a
// running pointer; only the first and last bytes of a segment, which share a word with a
// neighbouring segment, are written byte-wise.
//
// k_pack_wide (general lane, at the end of this file): blocks with a code word of 17 ... 56 bits
// and one-symbol blocks.  Whole rows of 512 symbols are packed the same way with one put per code
// word (<= 26 bits, 32-bit table entries) or two (<= 56 bits, 64-bit entries); ragged ends,
// unaligned buffers and one-symbol blocks take its per-symbol loop with a segmented carry scan.
//
// Replaces __huf_encode_block + huf_bit_write (reference src/encoder.c:85-131,
// src/bufio.c:18-32) and the header writes (src/encoder.c:325-342).
#pragma once

#include <type_traits>

#include "enc_kernels.cuh"

namespace hufb200 {

constexpr uint32_t kPackFastMaxLen = 16;
constexpr int kPackStageWords = 288;  // 16 symbols x 16 bits x 32 lanes = 256 words + kept line, padded

struct PackFastSmem {
    uint32_t table[kEncWarps][256];            // code << (32 - len) in the top 16 bits | len
    __align__(16) uint32_t stage[kEncWarps][kPackStageWords];
};

// A lane's bit accumulator: `cur` is the word under construction, `pos` the lane's running bit
// position (only its low five bits -- the bits already in cur -- and the carry into bit 5 are
// looked at, so whatever sits above them rides along: the sum words carry code bits up there),
// `at` the shared address cur goes to once complete.  A put appends up to 32 bits (t, left
// aligned; the sum word s carries their number in its low six bits) and stores at most one word:
// a word boundary was crossed exactly when bit 5 of the position flipped.  What does not fit
// becomes the new word under construction -- and is zero when nothing was stored, so one select
// replaces the two-register hand-over -- and the byte address advances by itself.
struct PackAcc {
    uint32_t cur, pos;
    saddr_t at;
};

__device__ __forceinline__ void pack_put(PackAcc &s, uint32_t t, uint32_t sum)
{
    const uint32_t x = __funnelshift_r(t, 0u, s.pos);   // t >> (pos & 31)
    const uint32_t y = __funnelshift_r(0u, t, s.pos);   // t << (32 - (pos & 31)); 0 for an empty cur
    const uint32_t n = s.pos + sum;
    const uint32_t w = s.cur | x;
#ifdef HUF_EMU
    const bool full = ((n ^ s.pos) & 32u) != 0;
    if (full) sts_u32(s.at, w);
    s.cur = full ? y : w;
    s.at += full ? 4u : 0u;
#else
    // (spelled out: one test, a predicated store, a select and a predicated add -- the kernel is
    // bound by the integer pipe, and the compiler's own version spends three more operations on
    // the test and the address)
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 f;\n\t"
        "xor.b32 f, %3, %5;\n\t"
        "and.b32 f, f, 32;\n\t"
        "setp.ne.u32 p, f, 0;\n\t"
        "@p st.shared.u32 [%1], %2;\n\t"
        "selp.u32 %0, %4, %2, p;\n\t"
        "@p add.u32 %1, %1, 4;\n\t"
        "}"
        : "=r"(s.cur), "+r"(s.at)
        : "r"(w), "r"(n), "r"(y), "r"(s.pos)
        : "memory");
#endif
    s.pos = n;
}

// (HUF_PACK_MINCTA: build knob of the occupancy experiment -- CTAs per SM the register allocation must allow)
#ifdef HUF_PACK_MINCTA
__global__ void __launch_bounds__(kEncWarps * 32, HUF_PACK_MINCTA)
#else
__global__ void __launch_bounds__(kEncWarps * 32)
#endif
k_pack(EncArgs a)
{
    __shared__ PackFastSmem sm;
    const int lane = lane_id();
    const int w = warp_in_cta();
    const uint64_t g = (uint64_t)blockIdx.x * kEncWarps + w;  // pass-local segment index
    if (g >= a.npass * a.nspb) return;
    if (a.status[0] != kOk) return;

    const uint64_t bl = g / a.nspb;
    const uint32_t *meta = a.blk_meta + bl * 4;
    if (meta[1] > kPackFastMaxLen || meta[3] == 1) return;  // k_pack_wide takes this block
    const uint64_t b = a.blk0 + bl;
    const uint32_t k = (uint32_t)(g % a.nspb);
    const uint64_t blen = blk_len_of(a, b);
    const uint64_t soff = (uint64_t)k * a.seg;
    if (soff >= blen) return;
    const uint32_t slen = (uint32_t)((blen - soff) < a.seg ? (blen - soff) : a.seg);
    const uint32_t nseg_b = (uint32_t)((blen + a.seg - 1) / a.seg);
    const uint8_t *blk_in = a.in + b * a.blocksize;
    const uint8_t *p = blk_in + soff;

    const uint32_t tree_len = meta[0];
    const uint64_t boff = a.blk_off[b];
    const uint64_t pay0 = boff + kHdrFixed + 2ull * tree_len;  // first payload byte
    const uint64_t bits_total = a.blk_bits[bl];
    const uint64_t o = a.seg_bitoff[g];
    const bool last_seg = (k + 1 == nseg_b);
    const uint64_t o_end = last_seg ? bits_total : a.seg_bitoff[g + 1];

    // block header: written by the warp that owns segment 0
    if (k == 0) emit_block_header(a, bl, blen, tree_len, boff, lane);

    // per-warp copy of the code table (codes are at most 16 bits here: the low half holds the length)
    uint32_t *tab = sm.table[w];
    {
        const uint32_t *src = a.blk_table + bl * 512;
        for (int i = lane; i < 256; i += 32) tab[i] = src[i];
    }
    uint32_t *stage = sm.stage[w];
    const saddr_t stage_s = smem_addr(stage);
    const saddr_t tab_s = saddr_pin(smem_addr(tab));  // (pinned: otherwise rebuilt from the thread index per iteration)
    __syncwarp();

    OutRange r;
    r.out = a.out;
    r.b0 = pay0 + (o >> 3);
    r.b1 = last_seg ? pay0 + ((bits_total + 7) >> 3) : pay0 + (o_end >> 3);
    r.full_lo = (r.b0 + 3) >> 2;
    r.full_hi = r.b1 >> 2;

    // Global bit cursor, kept relative to a 16-byte line of the output so that finished lines
    // leave as 128-bit stores: stage[0] is the first word of the line the cursor is in, q the
    // bits of the window in front of the cursor.  Words of that line in front of the cursor's
    // word belong to other segments (never written from here: OutRange), the bits in front of the
    // cursor inside its word are the carry.  The first byte of the segment may begin with the last
    // bits of the previous segment's final code words: rebuild them so this warp owns the whole
    // byte.
    const uint64_t gbit = (pay0 << 3) + o;
    uint32_t q = (uint32_t)(gbit & 127);    // bits in front of the cursor inside its line
    // The output word index of stage[0] (a multiple of 4) is kept relative to the owned words
    // [full_lo, full_hi), in 32 bits: `lead` words until the window starts inside them (at most
    // 3, never positive again once the first line has left), `room` words from the window to
    // their end; `line` is where the lane's line of the window goes.
    const uint64_t wbase0 = (gbit >> 7) << 2;
    int32_t lead = (int32_t)(r.full_lo - wbase0);
    int32_t room = (int32_t)(r.full_hi - wbase0);
    uint4 *line = reinterpret_cast<uint4 *>(r.out) + (wbase0 >> 2) + lane;
    uint32_t carry = 0;                     // unfinished word in front of the cursor (left aligned)
    {
        const uint32_t rb = (uint32_t)(o & 7);
        if (rb) {
            uint32_t val = 0;
            if (lane == 0) {
                uint32_t got = 0;
                uint64_t idx = soff;
                while (got < rb) {
                    idx--;
                    const uint32_t e = tab[blk_in[idx]];
                    const uint32_t l = e & 31u;
                    val |= ((e & 0xffff0000u) >> (32 - l)) << got;
                    got += l;
                }
                val &= (1u << rb) - 1u;
            }
            val = __shfl_sync(kFull, val, 0);
            carry = val << (32 - (q & 31));  // rb != 0 implies q & 31 != 0
        }
    }
    if (lane < 4) stage[lane] = 0;  // (words of the first line in front of the cursor: not ours, never output)
    __syncwarp();

    const bool aligned = (reinterpret_cast<uintptr_t>(p) & 15) == 0;
    uint4 pre = make_uint4(0, 0, 0, 0);  // software pipeline: the lane's next 16 input bytes
    if (aligned && slen >= 512) pre = ld_stream_u4(p + lane * 16);
    // One iteration = 16 symbols per lane.  FULL: every lane has 16 symbols and the input is
    // 16-byte aligned (all iterations but the last of a segment).
    auto iteration = [&](auto full_tag, uint32_t base) {
        constexpr bool FULL = decltype(full_tag)::value;
        const uint32_t my0 = base + lane * 16;
        uint32_t sym[4] = {0, 0, 0, 0};
        uint32_t nvalid = 16;
        if (FULL) {
            // this iteration's bytes were requested one iteration ago; request the next ones now
            sym[0] = pre.x; sym[1] = pre.y; sym[2] = pre.z; sym[3] = pre.w;
            if (base + 1024 <= slen) pre = ld_stream_u4(p + my0 + 512);
        } else {
            nvalid = my0 < slen ? min(16u, slen - my0) : 0u;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if ((uint32_t)j < nvalid) sym[j >> 2] |= (uint32_t)p[my0 + j] << (8 * (j & 3));
            }
        }
        // ---- look up, concatenate neighbours (<= 32 bits), sum the lengths.  The length sits in
        // the low five bits of an entry: the funnel shifter takes it from there (wrap mode), and
        // the sum of two entries still carries the sum of their lengths in its low six bits.
        // (entries are code << 16 | length with bits [15:5] clear: sums of entries carry the sum
        // of the lengths in their low half whatever the codes add up to)
        uint32_t t[8], sp[8], sum_all = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t s0 = __byte_perm(sym[j >> 1], 0, 0x4440 + 2 * (j & 1));
            const uint32_t s1 = __byte_perm(sym[j >> 1], 0, 0x4441 + 2 * (j & 1));
            uint32_t e0 = lds_u32(tab_s + 4 * s0), e1 = lds_u32(tab_s + 4 * s1);
            if (!FULL) {
                if ((uint32_t)(2 * j) >= nvalid) e0 = 0;
                if ((uint32_t)(2 * j + 1) >= nvalid) e1 = 0;
            }
            t[j] = (e0 & 0xffff0000u) | __funnelshift_r(e1 & 0xffff0000u, 0u, e0);
            sp[j] = e0 + e1;
            sum_all += sp[j];
        }
        const uint32_t total_l = sum_all & 0xffffu;
        const uint32_t incl = warp_incl_scan(total_l);
        const uint32_t total = __shfl_sync(kFull, incl, 31);
        const uint32_t start = q + incl - total_l;
        const uint32_t first_widx = start >> 5;
        if (FULL) {
            PackAcc acc;
            acc.cur = 0;
            acc.pos = start;
            acc.at = stage_s + 4 * first_widx;
#pragma unroll
            for (int j = 0; j < 8; j++) pack_put(acc, t[j], sp[j]);
            // every lane completed at least one word: my first word still lacks the bits my left
            // neighbour left behind its last word boundary (lane 0: the carry of the iteration
            // before); lane 31's leftover is the next carry
            uint32_t in = __shfl_up_sync(kFull, acc.cur, 1);
            if (lane == 0) in = carry;
            carry = __shfl_sync(kFull, acc.cur, 31);
            stage[first_widx] |= in;
        } else {
            BitAcc acc;
            acc.hi = acc.lo = 0;
            acc.nb = start & 31;
            acc.widx = first_widx;
#pragma unroll
            for (int j = 0; j < 8; j++) acc_put(acc, stage, t[j], sp[j] & 63u);
            carry = pack_flush_carry(acc, first_widx, stage, carry);
        }
        __syncwarp();

        // ---- finished 16-byte lines leave coalesced; the unfinished line stays in front.  An
        // iteration completes at most (127 + 32 * 16 * 16) / 128 = 64 lines: two per lane.
        const uint32_t nlines = (q + total) >> 7;
        if (lead <= 0 && (int32_t)(4 * nlines) <= room) {
            // (every iteration but a segment's first and last: all lines lie inside the owned words)
            if ((uint32_t)lane < nlines) {
                const uint4 v = reinterpret_cast<const uint4 *>(stage)[lane];
                line[0] = make_uint4(bswap32(v.x), bswap32(v.y), bswap32(v.z), bswap32(v.w));
            }
            if ((uint32_t)lane + 32 < nlines) {
                const uint4 v = reinterpret_cast<const uint4 *>(stage)[lane + 32];
                line[32] = make_uint4(bswap32(v.x), bswap32(v.y), bswap32(v.z), bswap32(v.w));
            }
        } else {
            const uint64_t wbase = r.full_hi - (uint64_t)(int64_t)room;
            for (uint32_t L = lane; L < nlines; L += 32) {
                const uint64_t w0 = wbase + 4 * L;
                const uint4 v = reinterpret_cast<const uint4 *>(stage)[L];
                if (w0 >= r.full_lo && w0 + 4 <= r.full_hi) {
                    reinterpret_cast<uint4 *>(r.out)[w0 >> 2] =
                        make_uint4(bswap32(v.x), bswap32(v.y), bswap32(v.z), bswap32(v.w));
                } else {
                    store_word(r, w0, v.x);
                    store_word(r, w0 + 1, v.y);
                    store_word(r, w0 + 2, v.z);
                    store_word(r, w0 + 3, v.w);
                }
            }
        }
        // the words of the unfinished line that are complete move to the front of the window (its
        // later words are stored by the lanes that complete them in the next iteration)
        const uint32_t keep = lane < 4 ? stage[4 * nlines + lane] : 0u;
        q = (q + total) & 127;
        line += nlines;
        lead -= (int32_t)(4 * nlines);
        room -= (int32_t)(4 * nlines);
        __syncwarp();
        if (lane < 4) stage[lane] = keep;
        __syncwarp();
    };
    const bool out_ok = (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
    uint32_t base = 0;
    if (aligned && out_ok) {
        for (; base + 512 <= slen; base += 512) iteration(std::true_type{}, base);
    }
    for (; base < slen; base += 512) iteration(std::false_type{}, base);

    // what is left of the last line: finished words and the trailing partial word (the carry), of
    // which only the owned bytes are written
    if (lane == 0 && (q & 31)) stage[q >> 5] = carry;
    __syncwarp();
    if (lane < 4 && 32 * (uint32_t)lane < q) store_word(r, r.full_hi - (uint64_t)(int64_t)room + lane, stage[lane]);
}

// General lane of K3: blocks whose longest code word exceeds kPackWideMinLen - 1 bits, and blocks
// of a single symbol (the fast lane in enc_pack.cuh takes all others); status[2] counts them.
constexpr uint32_t kPackWideMinLen = 17;

__global__ void __launch_bounds__(kEncWarps * 32) k_pack_wide(EncArgs a)
{
    __shared__ PackSmem sm;
    const int lane = lane_id();
    const int w = warp_in_cta();
    const uint64_t g = (uint64_t)blockIdx.x * kEncWarps + w;  // pass-local segment index
    if (g >= a.npass * a.nspb) return;
    if (a.status[0] != kOk) return;
    if (a.status[2] == 0) return;  // no deep block in this call

    const uint64_t bl = g / a.nspb;
    // (blocks of one symbol have a 1-bit code: sixteen symbols of a lane do not fill a word, which
    // the fast lane's hand-over between neighbouring lanes relies on)
    if (a.blk_meta[bl * 4 + 1] < kPackWideMinLen && a.blk_meta[bl * 4 + 3] != 1) return;
    const uint64_t b = a.blk0 + bl;
    const uint32_t k = (uint32_t)(g % a.nspb);
    const uint64_t blen = blk_len_of(a, b);
    const uint64_t soff = (uint64_t)k * a.seg;
    if (soff >= blen) return;
    const uint32_t slen = (uint32_t)((blen - soff) < a.seg ? (blen - soff) : a.seg);
    const uint32_t nseg_b = (uint32_t)((blen + a.seg - 1) / a.seg);
    const uint8_t *blk_in = a.in + b * a.blocksize;
    const uint8_t *p = blk_in + soff;

    const uint32_t *meta = a.blk_meta + bl * 4;
    const uint32_t tree_len = meta[0];
    const uint32_t fmt = meta[2];
    const uint64_t boff = a.blk_off[b];
    const uint64_t pay0 = boff + kHdrFixed + 2ull * tree_len;  // first payload byte
    const uint64_t bits_total = a.blk_bits[bl];
    const uint64_t o = a.seg_bitoff[g];
    const bool last_seg = (k + 1 == nseg_b);
    const uint64_t o_end = last_seg ? bits_total : a.seg_bitoff[g + 1];

    // block header: written by the warp that owns segment 0
    if (k == 0) emit_block_header(a, bl, blen, tree_len, boff, lane);

    // per-warp copy of the code table
    uint32_t *tab = sm.table[w];
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.blk_table + bl * 512);
        uint4 *dst = reinterpret_cast<uint4 *>(tab);
        const int n16 = fmt == 0 ? 64 : 128;
        for (int i = lane; i < n16; i += 32) dst[i] = src[i];
    }
    __syncwarp();

    OutRange r;
    r.out = a.out;
    r.b0 = pay0 + (o >> 3);
    r.b1 = last_seg ? pay0 + ((bits_total + 7) >> 3) : pay0 + (o_end >> 3);
    r.full_lo = (r.b0 + 3) >> 2;
    r.full_hi = r.b1 >> 2;

    // Global bit cursor.  The first byte of the segment may begin with the last bits of the
    // previous segment's final code words: rebuild them so this warp owns the whole byte.
    uint64_t gbit = (pay0 << 3) + o;
    uint32_t q = (uint32_t)(gbit & 31);
    uint64_t wbase = gbit >> 5;
    uint32_t carry = 0;
    {
        const uint32_t rb = (uint32_t)(o & 7);
        if (rb) {
            uint32_t val = 0;
            if (lane == 0) {
                uint32_t got = 0;
                uint64_t idx = soff;
                while (got < rb) {
                    idx--;
                    const uint32_t s = blk_in[idx];
                    uint32_t c, l;
                    if (fmt == 0) {
                        const uint32_t e = tab[s];
                        l = e & 31u;
                        c = (e & ~31u) >> (32 - l);
                    } else {
                        const uint64_t e = reinterpret_cast<const uint64_t *>(tab)[s];
                        l = (uint32_t)(e & 0xffu);
                        c = (uint32_t)((e & ~0xffull) >> (64 - l));  // low bits suffice
                    }
                    val |= c << got;
                    got += l;
                }
                val &= (1u << rb) - 1u;
            }
            val = __shfl_sync(kFull, val, 0);
            carry = val << (32 - q);
        }
    }

    uint32_t *stage = sm.stage[w];
    const bool aligned = (reinterpret_cast<uintptr_t>(p) & 15) == 0;
    const int per_lane = fmt == 0 ? 16 : 4;              // symbols per lane per iteration
    const uint32_t step = 32u * per_lane;
    uint32_t base = 0;

    // Whole rows of 512 symbols go the way of the fast lane (enc_pack.cuh: 16 symbols per lane,
    // one-word accumulator keyed by the running bit position, completed words stored plainly,
    // leftovers handed to the right neighbour, 16-byte line stores) with one put per code word
    // of up to 26 bits, two per code word of up to 56: every code word of such a block has at
    // least two bits, so every lane completes a word per row.  The ragged end of a segment,
    // unaligned buffers and one-symbol blocks (1-bit code words) take the general loop below.
    if (aligned && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 && meta[3] != 1 && slen >= 512) {
        uint32_t ql = (uint32_t)(gbit & 127);  // bits in front of the cursor inside its 16-byte line
        const uint64_t wbase0 = (gbit >> 7) << 2;
        int32_t lead = (int32_t)(r.full_lo - wbase0);
        int32_t room = (int32_t)(r.full_hi - wbase0);
        uint4 *line = reinterpret_cast<uint4 *>(r.out) + (wbase0 >> 2) + lane;
        const saddr_t stage_s = smem_addr(stage);
        const saddr_t tab_s = saddr_pin(smem_addr(tab));
        if (lane < 4) stage[lane] = 0;  // (words of the first line in front of the cursor: not ours, never output)
        __syncwarp();
        uint4 pre = ld_stream_u4(p + lane * 16);
        auto row = [&](auto wide_tag) {
            constexpr bool WIDE = decltype(wide_tag)::value;  // 64-bit table entries
            const uint32_t sym[4] = {pre.x, pre.y, pre.z, pre.w};
            if (base + 1024 <= slen) pre = ld_stream_u4(p + base + 512 + lane * 16);
            uint32_t hi[16], lo[WIDE ? 16 : 1], total_l = 0;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const uint32_t s = __byte_perm(sym[j >> 2], 0, 0x4440 + (j & 3));
                if constexpr (WIDE) {
                    lds_u32x2(tab_s + 8 * s, lo[j], hi[j]);
                    total_l += lo[j] & 0xffu;
                } else {
                    hi[j] = lds_u32(tab_s + 4 * s);
                    total_l += hi[j] & 31u;
                }
            }
            const uint32_t incl = warp_incl_scan(total_l);
            const uint32_t total = __shfl_sync(kFull, incl, 31);
            const uint32_t start = ql + incl - total_l;
            const uint32_t first_widx = start >> 5;
            PackAcc acc;
            acc.cur = 0;
            acc.pos = start;
            acc.at = stage_s + 4 * first_widx;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if constexpr (WIDE) {
                    const uint32_t l = lo[j] & 0xffu;
                    const uint32_t l1 = min(l, 32u);
                    pack_put(acc, hi[j], l1);
                    pack_put(acc, lo[j] & ~0xffu, l - l1);
                } else {
                    pack_put(acc, hi[j] & ~31u, hi[j] & 31u);
                }
            }
            uint32_t in = __shfl_up_sync(kFull, acc.cur, 1);
            if (lane == 0) in = carry;
            carry = __shfl_sync(kFull, acc.cur, 31);
            stage[first_widx] |= in;
            __syncwarp();
            const uint32_t nlines = (ql + total) >> 7;
            if (lead <= 0 && (int32_t)(4 * nlines) <= room) {
                for (uint32_t L = lane; L < nlines; L += 32) {
                    const uint4 v = reinterpret_cast<const uint4 *>(stage)[L];
                    line[L - lane] = make_uint4(bswap32(v.x), bswap32(v.y), bswap32(v.z), bswap32(v.w));
                }
            } else {
                const uint64_t wline = r.full_hi - (uint64_t)(int64_t)room;
                for (uint32_t L = lane; L < nlines; L += 32) {
                    const uint4 v = reinterpret_cast<const uint4 *>(stage)[L];
                    store_word(r, wline + 4 * L, v.x);
                    store_word(r, wline + 4 * L + 1, v.y);
                    store_word(r, wline + 4 * L + 2, v.z);
                    store_word(r, wline + 4 * L + 3, v.w);
                }
            }
            const uint32_t keep = lane < 4 ? stage[4 * nlines + lane] : 0u;
            ql = (ql + total) & 127;
            line += nlines;
            lead -= (int32_t)(4 * nlines);
            room -= (int32_t)(4 * nlines);
            __syncwarp();
            if (lane < 4) stage[lane] = keep;
            __syncwarp();
        };
        if (fmt == 0) {
            for (; base + 512 <= slen; base += 512) row(std::false_type{});
        } else {
            for (; base + 512 <= slen; base += 512) row(std::true_type{});
        }
        // back to the word cursor of the general loop: the complete words of the unfinished line leave now
        const uint64_t wline = r.full_hi - (uint64_t)(int64_t)room;
        if ((uint32_t)lane < (ql >> 5)) store_word(r, wline + lane, stage[lane]);
        __syncwarp();
        q = ql & 31;
        wbase = wline + (ql >> 5);
    }

    for (; base < slen; base += step) {
        // ---- load this lane's symbols
        const uint32_t my0 = base + lane * per_lane;
        uint32_t sym[4] = {0, 0, 0, 0};  // 16 bytes, little endian in words
        uint32_t nvalid = 0;
        if (my0 < slen) nvalid = min((uint32_t)per_lane, slen - my0);
        if (per_lane == 16 && aligned && nvalid == 16) {
            const uint4 v = ld_stream_u4(p + my0);
            sym[0] = v.x; sym[1] = v.y; sym[2] = v.z; sym[3] = v.w;
        } else {
            // ragged tail / unaligned input: byte loads, static register indices
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if ((uint32_t)j < nvalid) sym[j >> 2] |= (uint32_t)p[my0 + j] << (8 * (j & 3));
            }
        }

        uint32_t total_l = 0;
        BitAcc acc;
        if (fmt == 0) {
            // ---- look up code words, sum the lengths
            uint32_t e[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const uint32_t s = (sym[j >> 2] >> (8 * (j & 3))) & 0xffu;
                e[j] = (uint32_t)j < nvalid ? tab[s] : 0u;
                total_l += e[j] & 31u;
            }
            const uint32_t incl = warp_incl_scan(total_l);
            const uint32_t total = __shfl_sync(kFull, incl, 31);
            const uint32_t start = q + incl - total_l;
            acc.hi = acc.lo = 0;
            acc.nb = start & 31;
            acc.widx = start >> 5;
            const uint32_t first_widx = acc.widx;
#pragma unroll
            for (int j = 0; j < 16; j++) acc_put(acc, stage, e[j] & ~31u, e[j] & 31u);
            carry = pack_flush_carry(acc, first_widx, stage, carry);
            __syncwarp();

            // ---- copy finished words out, coalesced
            const uint32_t nfull = (q + total) >> 5;
            for (uint32_t i = lane; i < nfull; i += 32) store_word(r, wbase + i, stage[i]);
            q = (q + total) & 31;
            wbase += nfull;
            __syncwarp();
        } else {
            uint64_t e[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t s = (sym[0] >> (8 * j)) & 0xffu;
                e[j] = (uint32_t)j < nvalid ? reinterpret_cast<const uint64_t *>(tab)[s] : 0ull;
                total_l += (uint32_t)(e[j] & 0xffu);
            }
            const uint32_t incl = warp_incl_scan(total_l);
            const uint32_t total = __shfl_sync(kFull, incl, 31);
            const uint32_t start = q + incl - total_l;
            acc.hi = acc.lo = 0;
            acc.nb = start & 31;
            acc.widx = start >> 5;
            const uint32_t first_widx = acc.widx;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t l = (uint32_t)(e[j] & 0xffu);
                const uint64_t t = e[j] & ~0xffull;
                const uint32_t l1 = min(l, 32u);
                acc_put(acc, stage, (uint32_t)(t >> 32), l1);
                acc_put(acc, stage, (uint32_t)t, l - l1);
            }
            carry = pack_flush_carry(acc, first_widx, stage, carry);
            __syncwarp();

            const uint32_t nfull = (q + total) >> 5;
            for (uint32_t i = lane; i < nfull; i += 32) store_word(r, wbase + i, stage[i]);
            q = (q + total) & 31;
            wbase += nfull;
            __syncwarp();
        }
    }
    // trailing partial word: only its owned bytes are written
    if (q && lane == 0) store_word(r, wbase, carry);
}

}  // namespace hufb200
