// dec_fast.cuh — the fast lane of the decode path (sm_100a).
//
//   K4d k_tree     one LANE per candidate block (32 headers staged in shared memory per warp
//                  with coalesced loads): serial pre-order walk of the serialised tree (grammar
//                  of reference src/tree.c:138-208) -- the slot being filled is known by its left
//                  aligned path bits and depth, a terminal moves on by adding 2^(32 - depth) and
//                  the carry of that addition is the pop -- producing the ordered
//                  list of lookup-table terminals (leaf / absent child / long-code prefix) and
//                  the shortest code length; leaves below the table reach (13 bits) become
//                  sorted (code, length, symbol) records that the decoder searches.  Only the
//                  shape every reference encoder emits is eligible: a root with a left child
//                  only (src/tree.c:410-413), codes of at most 32 bits.  Anything else --
//                  foreign tree shapes, broken headers, zero-length blocks -- is left to
//                  k_decode_slow, which implements the whole acceptance grammar.
//   K5  k_decode   one 128-thread CTA per eligible candidate (handed out dynamically, four
//                  CTAs per SM), the block's payload processed in chunks:
//                  (0) chunk staged in shared memory as big-endian words (coalesced 16-byte
//                      loads, all of a thread's requests in flight together and issued before
//                      the previous chunk is copied out)
//                  (1) every thread warms up on the bits in front of its sub-block (Huffman
//                      codes self-synchronise within ~20 code words), then decodes its sub-block
//                      and WRITES the symbols into a private shared-memory region (4 symbols per
//                      32-bit store): no symbol is decoded twice.  The walk keeps three staged
//                      words in registers (every word is loaded once), looks four code words up
//                      per iteration through a 12-bit table behind the root bit whose 8 KB
//                      aligned base is OR-ed into the index, and stores into regions that are
//                      interleaved by thread (bank = thread: no conflicts).  Sub-blocks are sized
//                      for the block's average code length; a region that runs full makes the
//                      CTA repeat the chunk with the length that is safe for the shortest code
//                  (2) verification: a thread's first code word must start where its
//                      predecessor's last one ended.  A thread that did not synchronise decodes
//                      its sub-block again from the proven position; rounds repeat until nothing
//                      moves (codes that never re-synchronise, e.g. p(k) = 2^-(k+1), degrade to a
//                      sequential ripple over the threads of a chunk: slow, never wrong)
//                  (3) symbol-count scan -> output position of every region
//                  (4) regions compacted into a linear shared buffer (aliasing the stage) and
//                      copied out with coalesced 16-byte stores.
//                  Replaces __huf_decode_block (reference src/decoder.c:34-96) for clean blocks.
//                  A block that shows anything irregular (dead walk on the proven trajectory,
//                  payload running out) is handed to k_decode_slow untouched, so error codes and
//                  partial-output behaviour stay those of the general lane.
#pragma once

#include <type_traits>

#include "dec_kernels.cuh"

namespace hufb200 {

constexpr uint32_t kRedo = kRedoStatus;   // blk_status value: block left to k_decode_slow
constexpr uint32_t kNone = 0xffffffffu;

// per-candidate scratch slot: table terminals ((table start << 16) | table entry, ascending
// start) grow from the front, long-code records (left-aligned code, length << 8 | symbol;
// ascending code) from the back
constexpr int kTermStride = 800;
constexpr int kLongMax = 256;
constexpr int kLongBits = 32;             // longest code word the fast lane handles
constexpr int kTreeReach = kLutBits + 1;  // code bits resolved by the table incl. the root bit

// Fast-lane table entries (u16): leaf = symbol << 8 | length (1..13); the two special kinds
// carry a length field of 1 so that a walk which only needs to make progress (warm-up) can
// add bits [3:0] blindly.
constexpr uint32_t kFastDead = 0x41;   // walks into an absent child
constexpr uint32_t kFastLong = 0x81;   // code word longer than the table reach
constexpr uint32_t kFastFlags = 0xc0u;

// two meta words per candidate: [0] bit 0 = eligible for the fast lane, [15:8] shortest code
// length, [31:16] number of terminals; [1] number of long-code records
constexpr uint32_t kMetaFast = 1u;

constexpr int kFT = 128;                  // threads per CTA.  256 threads with sub-blocks half as long (kRegWords 25,
                                          // kMaxSubWords 13) measure within 2-4 % either way: Zipf and Fibonacci
                                          // shapes prefer 128, uniform / text / geometric prefer 256
#ifndef HUF_DEC_REGWORDS
#define HUF_DEC_REGWORDS 51
#endif
#ifndef HUF_DEC_MINCTA
#define HUF_DEC_MINCTA 4
#endif
constexpr int kRegWords = HUF_DEC_REGWORDS;  // words per thread region (+ one row behind them that
                                          // takes the stores of a region that has run full)
constexpr int kRegCap = 4 * kRegWords;    // symbols a region can hold
constexpr int kRegRow = kFT * 4;          // regions are interleaved: word c of thread t sits at
                                          // row c, column t -- the bank is the thread's, so region
                                          // stores and loads never conflict whatever c each lane is at
constexpr int kMaxSubWords = 29;          // payload words per thread per chunk (odd)
constexpr int kFastStage = kFT * kMaxSubWords * 4 + 64;  // staged payload bytes (+ start skew, slack)
constexpr int kFastOutWin = kFT * kMaxSubWords * 4;      // compaction window (multiple of 16)
// Dynamic shared memory of k_decode (the kernel has no static shared memory, so the block
// starts at shared-window address 0x400 behind the 1 KB the system reserves):
//   [0, kFastStage)            staged payload / compaction window
//   [kFastLutOff, +8 KB)       lookup table, at window address 0x4000: 8 KB aligned, so the
//                              table index is OR-ed into the base
//   [kFastRegOff, ...)         symbol regions, one per thread
//   [kFastTailOff, ...)        FastTail: long-code records, re-speculation classes
//   [kFastSmallOff, kFastDyn)  FastSmall: per-thread hand-over words and CTA scalars
// kFastDyn + 1 KB stays below a quarter of the 228 KB an SM offers: four CTAs per SM.
constexpr int kFastLutOff = 15360;
constexpr int kFastLutAlign = 2 * kLutSize;               // 8 KB
constexpr int kFastRegOff = kFastLutOff + 2 * kLutSize;
constexpr int kFastTailOff = kFastRegOff + (kRegWords + 1) * kRegRow;

constexpr int kRespecMin = 2;             // threads still failing after one repair round: a ripple, re-speculate
constexpr int kRespecStarts = 16;         // warm-up start offsets tried for the alternate trajectory
constexpr int kTreeWin = 520;             // serialised tree elements staged per lane at a time
constexpr int kTreeHdrStride = 1052;      // staged bytes per candidate (263 words, odd)
constexpr int kTreeDyn = 32 * kTreeHdrStride;

// ------------------------------------------------------------------------------------------
// K4d: tree walk, one lane per candidate, 32 candidates per CTA.
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(32) k_tree(DecArgs a)
{
#ifdef HUF_EMU
    uint8_t *hdr = hufemu::dyn_smem();
#else
    extern __shared__ __align__(16) uint8_t hdr[];
#endif
    const int lane = lane_id();
    const uint64_t ncand = a.result[0];
    const bool in_ok = (reinterpret_cast<uintptr_t>(a.in) & 3) == 0;

    for (uint64_t g0 = (uint64_t)blockIdx.x * 32; g0 < ncand; g0 += (uint64_t)gridDim.x * 32) {
        __syncwarp();
        // ---- every lane reads the fixed header fields of its own candidate ...
        uint64_t my_off = 0, my_ol = 0;
        uint32_t my_tl = 0;
        bool my_ok = false;
        {
            const uint64_t j = g0 + lane;
            bool ok = in_ok && j < ncand && j < a.term_slots;
            if (ok) {
                my_off = a.cand[j];
                ok = my_off + kHdrFixed + 4 <= a.avail;
            }
            if (ok) {
                my_ol = rd_u64(a.in + my_off);
                my_tl = rd_u16(a.in + my_off + 8);
                ok = my_ol != 0 && my_tl >= 5 && my_tl <= (a.accept_1025 ? 1025u : 1024u) &&
                     my_off + kHdrFixed + 2ull * my_tl <= a.avail;
            }
            my_ok = ok;
        }
        // ---- walk state of my tree (kept across window reloads)
        const uint64_t j = g0 + lane;
        uint32_t meta = 0, nlong = 0;
        bool active = j < ncand && my_ok && my_ol;  // still walking
        bool ok = active;
        const uint32_t tl = my_tl;
        uint32_t *slot = a.terms + j * kTermStride;
        // The walk fills slots in pre-order; the slot being filled is known by its left-aligned
        // path bits K and its depth D alone.  A terminal (leaf or absent child) at depth D covers
        // 2^-D of the code space, so the next open slot starts at K + 2^(32 - D) -- the carry of
        // that addition runs through the ancestors that were entered to the right and stops at
        // the deepest one that still has its right slot open, which is exactly where a pre-order
        // walk continues -- and its depth is the position of the lowest set bit of the sum.  An
        // inner node just goes one level down to the left (K unchanged).  The tree is complete
        // when the sum wraps to zero.  (Eligible trees are at most kLongBits = 32 deep.)
        uint32_t i = 1, D = 1;        // slot being filled: element index, depth ...
        uint32_t K = 0;               // ... and path bits, left aligned
        uint32_t nterm = 0, min_len = kTreeReach + 1;
        uint32_t wb = 0;              // first element of my staged window
        bool first = true;

        while (__any_sync(kFull, active)) {
            // ---- the warp stages a window of kTreeWin elements for every walking lane: whole
            // aligned words copied verbatim, asynchronously, all in flight together
            for (int c = 0; c < 32; c++) {
                if (!__shfl_sync(kFull, (int)active, c)) continue;
                const uint64_t b0 = __shfl_sync(kFull, my_off, c) + kHdrFixed + 2ull * __shfl_sync(kFull, wb, c);
                const uint32_t left = __shfl_sync(kFull, tl, c) - __shfl_sync(kFull, wb, c);
                const uint32_t cnt = left < (uint32_t)kTreeWin ? left : (uint32_t)kTreeWin;
                const uint64_t w0 = b0 >> 2, w1 = (b0 + 2ull * cnt + 3) >> 2;  // word range
                uint32_t *dst = reinterpret_cast<uint32_t *>(hdr + c * kTreeHdrStride);
                for (uint64_t k = w0 + lane; k < w1; k += 32) {
                    if (4 * k + 4 <= a.avail) {
                        cp_async4(dst + (k - w0), reinterpret_cast<const uint32_t *>(a.in) + k);
                    } else {
                        uint32_t v = 0;
                        for (int q = 0; q < 4; q++) {
                            if (4 * k + q < a.avail) v |= (uint32_t)a.in[4 * k + q] << (8 * q);
                        }
                        dst[k - w0] = v;
                    }
                }
            }
            cp_async_wait_all();
            __syncwarp();

            if (active) {
                uint32_t *hw = reinterpret_cast<uint32_t *>(hdr + lane * kTreeHdrStride);
                const uint32_t skew = (uint32_t)((my_off + kHdrFixed + 2ull * wb) & 3);
                // The staged words keep the stream's byte skew.  A header at an odd address is
                // moved down by one byte once, so that every element is an aligned 16-bit word
                // and the walk reads it with one sign-extending load.
                if (skew & 1) {
                    uint32_t lo = hw[0];
#pragma unroll 8
                    for (int w = 0; w < kTreeHdrStride / 4 - 1; w++) {
                        const uint32_t hi = hw[w + 1];
                        hw[w] = __funnelshift_r(lo, hi, 8);
                        lo = hi;
                    }
                    hw[kTreeHdrStride / 4 - 1] = lo >> 8;
                }
                int16_t *el = reinterpret_cast<int16_t *>(hw) + ((skew & 2) >> 1);  // el[k]: element wb + k
                // elements at and behind tree_len read as absent children (src/tree.c:154-160):
                // three of them are written behind the last element, and reads are clamped there
                const uint32_t pad = tl - wb;
                if (pad <= (uint32_t)kTreeWin) el[pad] = el[pad + 1] = el[pad + 2] = -1;
                if (first) {
                    ok = el[0] != -1 && el[1] != -1;  // a root with something below its left edge
                    first = false;
                }
                while (ok) {
                    // elements i .. i+2 must lie in the window (or past the end of the tree)
                    const uint32_t rel = i - wb;
                    if (rel + 3 > (uint32_t)kTreeWin && wb + (uint32_t)kTreeWin < tl) {
                        wb = i;  // slide the window; the warp reloads it
                        break;
                    }
                    const uint32_t at = min(rel, pad);
                    const int e = el[at], e1 = el[at + 1], e2 = el[at + 2];
                    // One straight-line body for the three kinds of slot content (leaf: v, absent,
                    // absent / inner node / absent child): the lanes of a warp sit on different
                    // kinds all the time, so the kinds are selected, not branched on; only the
                    // exits (error, tree complete, window slide) leave the loop.
                    const bool is_abs = e == -1;
                    const bool childless = (e1 & e2) == -1;
                    const bool is_leaf = !is_abs && childless;
                    const bool is_inner = !is_abs && !childless;
                    // the right slot of the root must stay empty (table sits behind the root bit);
                    // leaves below an inner node at depth 32 would need more than 32 bits
                    if ((K == 0x80000000u && !is_abs) || (is_inner && D >= (uint32_t)kLongBits)) {
                        ok = false;
                        break;
                    }
                    const bool in_reach = D <= (uint32_t)kTreeReach;
                    // table terminals: leaves and absent children inside the reach, inner nodes
                    // exactly at the table depth (long codes); leaves below become records
                    const bool emit_term = in_reach && (!is_inner || D == (uint32_t)kTreeReach);
                    const bool emit_long = is_leaf && !in_reach;
                    // (whether the slot can hold everything is decided once, at the end: terminals
                    // past its capacity pile up on its last word meanwhile; the records cannot
                    // leave it -- at most (1025 + 32) / 3 leaves -- and what the two lists
                    // overwrite of each other is never read, such a tree is not eligible)
                    const uint32_t sym = (uint32_t)(e & 0xff);
                    if (emit_term) {
                        const uint32_t entry = is_leaf ? ((sym << 8) | D) : (is_abs ? kFastDead : kFastLong);
                        slot[min(nterm, (uint32_t)kTermStride - 1u)] = ((K >> (32 - kTreeReach)) << 16) | entry;
                        nterm++;
                        if (is_leaf) min_len = min(min_len, D);
                    }
                    if (emit_long) {
                        slot[kTermStride - 2 * (nlong + 1)] = K;
                        slot[kTermStride - 2 * (nlong + 1) + 1] = (D << 8) | sym;
                        nlong++;
                    }
                    i += is_leaf ? 3u : 1u;
                    const uint32_t next = K + (0x80000000u >> (D - 1u));  // behind a terminal at depth D
                    if (!is_inner && next == 0) {  // every slot filled: the tree is complete
                        // (one word of the slot stays free, and the records are limited)
                        if (nterm + 2 * nlong < (uint32_t)kTermStride && nlong <= (uint32_t)kLongMax)
                            meta = kMetaFast | (min_len << 8) | (nterm << 16);
                        ok = false;
                        break;
                    }
                    D = is_inner ? D + 1u : 33u - (uint32_t)__ffs((int)next);
                    K = is_inner ? K : next;
                }
                if (!ok) active = false;
            }
        }
        if (j < ncand) {
            a.meta[2 * j] = meta;
            a.meta[2 * j + 1] = nlong;
        }
    }
}

// ------------------------------------------------------------------------------------------
// K5: fast block decode.
// ------------------------------------------------------------------------------------------

struct FastSmall {
    uint32_t sub_end[kFT];
    uint32_t warp_tot[kFT / 32];
    uint32_t nlong;
    uint32_t redo;
    uint32_t fin_found;
    uint32_t fin_end;
    uint32_t total;
    uint32_t ovf;                  // a region ran full: the chunk is repeated with safe sub-blocks
    unsigned long long next_j;
    unsigned long long mbar;       // mbarrier of the bulk copy that stages the payload
};

struct FastTail {
    uint32_t long_code[kLongMax];  // left-aligned code words longer than the table reach, ascending
    uint32_t c_start[2][kFT];      // re-speculation: first code word of the primary / alternate
    uint32_t c_end[2][kFT];        // trajectory of every sub-block and where it leaves it
    uint32_t pred[kFT];            // predicted true start of every sub-block
    uint16_t long_ent[kLongMax];   // length << 8 | symbol
};

constexpr int kFastSmallOff = kFastTailOff + (int)sizeof(FastTail);
constexpr int kFastDyn = kFastSmallOff + (int)sizeof(FastSmall);
static_assert(kFastStage <= kFastLutOff, "the stage must fit in front of the table");
static_assert(kFastDyn + 1024 <= 233472 / HUF_DEC_MINCTA, "CTAs per SM");
static_assert(sizeof(FastTail) % 8 == 0, "FastSmall holds 64-bit words");
static_assert(kFastStage % 16 == 0 && kFastRegOff % 16 == 0 && kFastTailOff % 8 == 0, "alignment");

// Byte offset of the table entry selected by the twelve bits BEHIND the root bit of the left
// aligned window (every code word starts with the 0 bit of the one-child root, src/tree.c:410-413).
// The root bit itself is checked separately: a walk on a set root bit is dead.
__device__ __forceinline__ uint32_t fast_idx(uint32_t win) { return (win >> (32 - kTreeReach - 1)) & (2u * kLutSize - 2u); }

// Shared-window address of the table entry at byte offset `off`.  kLutOr: the table base is 8 KB
// aligned, so the offset is OR-ed in (one three-input logic instruction forms the address out of
// the shifted window, the mask and the base); otherwise it is added.  Which one applies depends
// on where the dynamic shared memory of the kernel starts in the shared window: the host probes
// that once per context (k_smem_base) and launches the matching instance.
template <bool kLutOr>
__device__ __forceinline__ saddr_t lut_at(saddr_t lut_s, uint32_t off)
{
    return kLutOr ? saddr_or(lut_s, off) : lut_s + off;
}

// The three staged words from the one that holds bit `pos` on, kept in registers while a walk
// moves through its sub-block: every staged word is loaded once per walk (the loads of 32
// lanes at unrelated positions hit random banks, so each one costs several wavefronts).
struct BitWin {
    uint32_t w0, w1, w2;
};

// (the stage holds the stream bytes as they are; a word becomes MSB-first bits by a byte swap)
__device__ __forceinline__ void win_load(BitWin &b, saddr_t sw_s, uint32_t pos)
{
    lds_u32x3(sw_s + ((pos >> 5) << 2), b.w0, b.w1, b.w2);
    b.w0 = bswap32(b.w0);
    b.w1 = bswap32(b.w1);
    b.w2 = bswap32(b.w2);
}

// The walk moved from `pos` to `np` (at most two words further).  Branch-free: predicated
// register moves and predicated loads of the one or two new words.
__device__ __forceinline__ void win_advance(BitWin &b, saddr_t sw_s, uint32_t pos, uint32_t np)
{
    const uint32_t adv = (np >> 5) - (pos >> 5);
    const saddr_t at = sw_s + ((np >> 5) << 2);
#ifdef HUF_EMU
    if (adv == 1) {
        b.w0 = b.w1;
        b.w1 = b.w2;
        b.w2 = bswap32(lds_u32(at + 8));
    } else if (adv == 2) {
        b.w0 = b.w2;
        b.w1 = bswap32(lds_u32(at + 4));
        b.w2 = bswap32(lds_u32(at + 8));
    }
#else
    asm volatile(
        "{\n\t"
        ".reg .pred p1, p2;\n\t"
        "setp.ne.u32 p1, %4, 0;\n\t"
        "setp.gt.u32 p2, %4, 1;\n\t"
        "@p1 mov.u32 %0, %1;\n\t"
        "@p1 mov.u32 %1, %2;\n\t"
        "@p2 mov.u32 %0, %2;\n\t"
        "@p2 ld.shared.u32 %1, [%3+4];\n\t"
        "@p1 ld.shared.u32 %2, [%3+8];\n\t"
        "@p2 prmt.b32 %1, %1, 0, 0x0123;\n\t"
        "@p1 prmt.b32 %2, %2, 0, 0x0123;\n\t"
        "}"
        : "+r"(b.w0), "+r"(b.w1), "+r"(b.w2)
        : "r"(at), "r"(adv));
#endif
}

// The window moves on by `adv` (0, 1 or 2) words; `at` is the shared address of its new first
// word.  Branch-free: predicated register moves and predicated loads of the one or two new words.
__device__ __forceinline__ void win_step(BitWin &b, saddr_t at, uint32_t adv)
{
#ifdef HUF_EMU
    if (adv == 1) {
        b.w0 = b.w1;
        b.w1 = b.w2;
        b.w2 = bswap32(lds_u32(at + 8));
    } else if (adv == 2) {
        b.w0 = b.w2;
        b.w1 = bswap32(lds_u32(at + 4));
        b.w2 = bswap32(lds_u32(at + 8));
    }
#else
    asm volatile(
        "{\n\t"
        ".reg .pred p1, p2;\n\t"
        "setp.ne.u32 p1, %4, 0;\n\t"
        "setp.gt.u32 p2, %4, 1;\n\t"
        "@p1 mov.u32 %0, %1;\n\t"
        "@p1 mov.u32 %1, %2;\n\t"
        "@p2 mov.u32 %0, %2;\n\t"
        "@p2 ld.shared.u32 %1, [%3+4];\n\t"
        "@p1 ld.shared.u32 %2, [%3+8];\n\t"
        "@p2 prmt.b32 %1, %1, 0, 0x0123;\n\t"
        "@p1 prmt.b32 %2, %2, 0, 0x0123;\n\t"
        "}"
        : "+r"(b.w0), "+r"(b.w1), "+r"(b.w2)
        : "r"(at), "r"(adv));
#endif
}

// a * b + c on the multiply-add pipe (the compiler would pick the integer ALU's shift-add)
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c)
{
#ifdef HUF_EMU
    return a * b + c;
#else
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#endif
}

// Shared address of staged word `idx` (tests/emu: addresses are 64-bit host pointers there, which
// a 32-bit multiply-add would cut off -- harmless in the main thread of a non-PIE interpreter,
// whose heap lies below 4 GB, fatal in the worker threads of the multi-device host lanes).
__device__ __forceinline__ saddr_t saddr_word(uint32_t idx, saddr_t base)
{
#ifdef HUF_EMU
    return base + (saddr_t)idx * 4u;
#else
    return mad_u32(idx, 4u, base);
#endif
}

// Table offset for the bulk walk.  The walk is bound by the integer ALU pipe (shifts, logic,
// compares: one warp instruction every other cycle) while the multiply-add pipe idles, so the
// right shift by 18 is done as the high half of a multiplication by 2^14; the factor comes
// from the kernel arguments because a literal would be turned back into a shift.
__device__ __forceinline__ uint32_t bulk_idx(uint32_t win, uint32_t mul14)
{
#if defined(HUF_EMU) || defined(HUF_NO_MULHI_IDX)
    (void)mul14;
    return fast_idx(win);
#else
    return __umulhi(win, mul14) & (2u * kLutSize - 2u);
#endif
}

// Length field for the blind walk: the entry, or 1 when the window starts on a set root bit
// (two instructions: arithmetic shift + one three-input logic op).
__device__ __forceinline__ uint32_t blind_len(uint32_t e, uint32_t h)
{
#ifdef HUF_EMU
    const uint32_t m = (uint32_t)((int32_t)h >> 31);
    return (e & ~m) | (m & 1u);
#else
    uint32_t m, l;
    asm("shr.s32 %0, %1, 31;" : "=r"(m) : "r"(h));
    asm("lop3.b32 %0, %1, %2, 1, 0xb8;" : "=r"(l) : "r"(e), "r"(m));  // (e & ~m) | (m & 1)
    return l;
#endif
}

// Four consecutive table entries from `pos` on: the three window words are shifted into a
// 64-bit left-aligned bit buffer that is then shifted by every code length (4 x 13 bits fit; the
// shifter takes the low five bits of an entry, i.e. its length, 1 for the special kinds).
// Returns the position behind the fourth.  h0..h3 are the four windows: their top bits tell
// whether a walk started on a set root bit; the entries and the returned position are only
// meaningful up to the first such walk (it is dead and advances one bit).
template <bool kLutOr>
__device__ __forceinline__ uint32_t fast_look4(const BitWin &b, saddr_t lut_s, uint32_t pos,
                                               uint32_t &e0, uint32_t &e1, uint32_t &e2, uint32_t &e3,
                                               uint32_t &h0, uint32_t &h1, uint32_t &h2, uint32_t &h3)
{
    const uint32_t w0 = b.w0, w1 = b.w1, w2 = b.w2;
    h0 = __funnelshift_l(w1, w0, pos);                 // stream bits pos .. pos+31
    uint32_t lo = __funnelshift_l(w2, w1, pos);        // stream bits pos+32 .. pos+63
    e0 = lds_u16(lut_at<kLutOr>(lut_s, fast_idx(h0)));
    h1 = __funnelshift_l(lo, h0, e0);
    lo = __funnelshift_l(0u, lo, e0);
    e1 = lds_u16(lut_at<kLutOr>(lut_s, fast_idx(h1)));
    h2 = __funnelshift_l(lo, h1, e1);
    lo = __funnelshift_l(0u, lo, e1);
    e2 = lds_u16(lut_at<kLutOr>(lut_s, fast_idx(h2)));
    h3 = __funnelshift_l(lo, h2, e2);
    e3 = lds_u16(lut_at<kLutOr>(lut_s, fast_idx(h3)));
    // lengths sit in bits [3:0], bits [5:4] are clear in every kind of entry: no carries
    return pos + ((e0 + e1 + e2 + e3) & 0x3fu);
}

// One exact decode step at staged bit position `pos`.  Returns the next position in the low
// word; bit 63 is set, with the symbol in bits [39:32], when a code word starts at `pos` (table
// hit or long-code record).  A dead walk (set root bit, absent child) advances one bit: any
// deterministic rule serves a speculative start -- this one re-synchronises quickly on skewed
// codes -- and a dead step on the proven trajectory sends the block to the general lane.
constexpr uint64_t kStepOk = 1ull << 63;
template <bool kLutOr>
__device__ __noinline__ uint64_t fast_step(saddr_t sw_s, saddr_t lut_s, const FastTail *ft, uint32_t n,
                                           uint32_t pos)
{
    uint32_t w0, w1;
    lds_u32x2(sw_s + ((pos >> 5) << 2), w0, w1);
    const uint32_t win = __funnelshift_l(bswap32(w1), bswap32(w0), pos);
    const uint32_t e = lds_u16(lut_at<kLutOr>(lut_s, fast_idx(win)));
    const uint32_t root = win >> 31;
    if (!((e & kFastFlags) | root)) {
        return kStepOk | ((uint64_t)(e >> 8) << 32) | (pos + (e & 0xfu));
    }
    if (!root && (e & 0x80u) && n) {
        // last record with code <= window (prefix-free codes: the only possible match), found
        // with a fixed number of halving steps
        uint32_t lo = 0;
#pragma unroll
        for (uint32_t step = kLongMax / 2; step; step >>= 1) {
            const uint32_t mid = lo + step;
            if (mid < n && ft->long_code[mid] <= win) lo = mid;
        }
        const uint32_t ent = ft->long_ent[lo];
        const uint32_t len = ent >> 8;
        if (((win ^ ft->long_code[lo]) >> (32 - len)) == 0) {
            return kStepOk | ((uint64_t)(ent & 0xffu) << 32) | (pos + len);
        }
    }
    return pos + 1;
}

// CTA barrier behind divergent per-thread loops: the warp is explicitly reconverged first
// (the aligned barrier instruction requires all 32 lanes to arrive together).
__device__ __forceinline__ void cta_sync()
{
#ifndef HUF_EMU
    asm volatile("bar.warp.sync 0xffffffff;" ::: "memory");
#else
    __syncwarp();
#endif
    __syncthreads();
}

// Copy symbols [s0, s0 + n) of an interleaved region (word j of the region at regw[j * kFT])
// to dst (shared memory, arbitrary alignment): up to three bytes to align the destination, then
// whole words assembled with one funnel shift each, then up to three bytes.
__device__ __forceinline__ void region_copy(uint8_t *dst, const uint32_t *regw, uint32_t s0, uint32_t n)
{
    // four symbols from symbol i on: two region words and one funnel shift
    auto four_at = [&](uint32_t i) -> uint32_t {
        const uint32_t *p = regw + (i >> 2) * kFT;
        return __funnelshift_r(p[0], p[kFT], (i & 3) * 8);
    };
    const uint32_t head = min(n, (uint32_t)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3));
    if (head) {
        const uint32_t hw = four_at(s0);
        dst[0] = (uint8_t)hw;
        if (head > 1) dst[1] = (uint8_t)(hw >> 8);
        if (head > 2) dst[2] = (uint8_t)(hw >> 16);
    }
    const uint32_t sb = s0 + head;
    const uint32_t sh = (sb & 3) * 8;
    const uint32_t *__restrict__ sp = regw + (sb >> 2) * kFT;
    uint32_t *__restrict__ d = reinterpret_cast<uint32_t *>(dst + head);
    const uint32_t nw = (n - head) >> 2;
    uint32_t prev = sp[0];
    uint32_t k = 0;
    for (; k + 4 <= nw; k += 4) {
        const uint32_t a1 = sp[(k + 1) * kFT], a2 = sp[(k + 2) * kFT], a3 = sp[(k + 3) * kFT],
                       a4 = sp[(k + 4) * kFT];
        d[k] = __funnelshift_r(prev, a1, sh);
        d[k + 1] = __funnelshift_r(a1, a2, sh);
        d[k + 2] = __funnelshift_r(a2, a3, sh);
        d[k + 3] = __funnelshift_r(a3, a4, sh);
        prev = a4;
    }
    for (; k < nw; k++) {
        const uint32_t a1 = sp[(k + 1) * kFT];
        d[k] = __funnelshift_r(prev, a1, sh);
        prev = a1;
    }
    const uint32_t done = head + 4 * nw;
    if (done < n) {
        const uint32_t tw = four_at(s0 + done);
        dst[done] = (uint8_t)tw;
        if (done + 1 < n) dst[done + 1] = (uint8_t)(tw >> 8);
        if (done + 2 < n) dst[done + 2] = (uint8_t)(tw >> 16);
    }
}

// Phase timing of thread 0 of every CTA (debug builds with -DHUF_PHASE_PROF only): cycles per
// phase are summed over all CTAs into g_fast_prof, scripts/phase_prof.py prints them.
#ifdef HUF_PHASE_PROF
__device__ unsigned long long g_fast_prof[16];
#define HUF_PROF(k)                                \
    if (tid == 0) {                                \
        const unsigned long long now_ = clock64(); \
        pacc[k] += now_ - pt;                      \
        pt = now_;                                 \
    }
#define HUF_PROF_CNT(k) \
    if (tid == 0) pacc[k]++;
#else
#define HUF_PROF(k)
#define HUF_PROF_CNT(k)
#endif

template <bool kLutOr>
__device__ __forceinline__ void decode_fast_body(DecArgs a, uint8_t *dyn)
{
    FastSmall &sm = *reinterpret_cast<FastSmall *>(dyn + kFastSmallOff);
    FastTail &ft = *reinterpret_cast<FastTail *>(dyn + kFastTailOff);
    const int tid = threadIdx.x;
#ifdef HUF_PHASE_PROF
    unsigned long long pacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    unsigned long long pt = clock64();
#endif
    const uint64_t ncand = a.result[0];
    uint32_t *sw = reinterpret_cast<uint32_t *>(dyn);  // staged payload, big-endian words
    uint16_t *lut = reinterpret_cast<uint16_t *>(dyn + kFastLutOff);
    uint8_t *regions = dyn + kFastRegOff;
    const uint32_t *regw = reinterpret_cast<const uint32_t *>(regions) + tid;  // my region: regw[c * kFT]
    const saddr_t sw_s = saddr_pin(smem_addr(sw)), lut_s = saddr_pin(smem_addr(lut));
    const saddr_t reg_s = saddr_pin(smem_addr(regw));  // first word of my region
    uint32_t tma_phase = 0;    // parity of the staging barrier's next phase
    bool tma_pending = false;  // a bulk copy into the stage buffer is in flight ...
    uint64_t tma_base = 0;     // ... from this stream offset
    if (tid == 0) mbar_init(&sm.mbar, 1);
#ifndef HUF_EMU
    // (the host launches this instance only after probing the shared window; a mismatch is a
    // loud failure, never a silent mis-decode)
    if (kLutOr && (lut_s & (uint32_t)(kFastLutAlign - 1))) __trap();
#endif
    const bool in_ok = (reinterpret_cast<uintptr_t>(a.in) & 15) == 0;
    const uint32_t mul14 = a.mul14;  // 1 << (32 - 18), see bulk_idx

    for (;;) {
        // blocks are handed out dynamically (result[11] is the work counter): block costs differ
        cta_sync();
        if (tid == 0) sm.next_j = atomicAdd(reinterpret_cast<unsigned long long *>(&a.result[11]), 1ull);
        cta_sync();
        const uint64_t j = sm.next_j;
        if (j >= ncand) break;
        const uint32_t meta = a.meta[2 * j];
        const uint32_t nlong = a.meta[2 * j + 1];
        const uint64_t off = a.cand[j];
        uint64_t orig_len = 0;
        uint32_t tl = 0;
        bool fast = (meta & kMetaFast) && in_ok;
        if (fast) {
            orig_len = rd_u64(a.in + off);
            tl = rd_u16(a.in + off + 8);
        }
        const uint64_t pay0 = off + kHdrFixed + 2ull * tl;
        if (fast && orig_len > 8ull * (a.avail - pay0)) fast = false;  // cannot complete: error lane
        if (!fast) {
            if (tid == 0) {
                a.blk_status[j] = kRedo;
                atomicAdd(reinterpret_cast<unsigned long long *>(&a.result[10]), 1ull);
            }
            continue;
        }
        const uint32_t min_len = (meta >> 8) & 0xffu;
        const uint32_t nterm = meta >> 16;
        // the block this CTA decodes next: request its header, terminals and first chunk into L2
        if (j + gridDim.x < ncand) {
            const uint64_t noff = a.cand[j + gridDim.x] + 128ull * (uint32_t)tid;
            if (noff < a.avail) prefetch_l2(a.in + noff);
            if (tid < 26) prefetch_l2(a.terms + (j + gridDim.x) * kTermStride + 32 * tid);
        }

        // Staging.  The payload window of a chunk (kFastStage bytes from the 16-byte aligned
        // `base16`) is brought in by ONE bulk copy of the TMA unit, requested by thread 0 as soon
        // as the stage buffer is free and the window's start is known -- at the start of a block
        // and behind the copy-out of every chunk -- so it flies during the table build and the
        // chunk planning; everybody waits on its mbarrier at the top of the chunk.  Only a
        // window that reaches past the readable bytes (the last chunks of a stream) is loaded
        // with ordinary 16-byte loads, zero-filled behind `avail`.
        auto stage_request = [&](uint64_t from16) {
            tma_pending = from16 + (uint64_t)kFastStage <= a.avail;
            tma_base = from16;
            if (tma_pending && tid == 0) {
                fence_async_proxy();  // the compaction window was written through the generic proxy
                tma_load_1d(sw, a.in + from16, (uint32_t)kFastStage, &sm.mbar);
            }
#ifdef HUF_EMU
            cta_sync();  // the emulator copies at request time: this stands in for the mbarrier wait
#endif
        };
        auto stage_wait = [&]() {
            mbar_wait(&sm.mbar, tma_phase);
            tma_phase ^= 1u;
            tma_pending = false;
        };
        stage_request(pay0 & ~uint64_t(15));  // the first chunk's window lands during the table build

        // ---- lookup table from the ordered terminal list: every thread fills 16 entries
        {
            uint32_t *tb = reinterpret_cast<uint32_t *>(regions);
            const uint32_t *src = a.terms + j * kTermStride;
            for (uint32_t k = tid; k < nterm; k += kFT) tb[k] = src[k];
            for (uint32_t r = tid; r < nlong; r += kFT) {
                ft.long_code[r] = src[kTermStride - 2 * (r + 1)];
                ft.long_ent[r] = (uint16_t)src[kTermStride - 2 * (r + 1) + 1];
            }
            if (tid == 0) {
                sm.redo = 0;
                sm.ovf = 0;
                sm.nlong = nlong;
            }
            cta_sync();
            constexpr int kPer = kLutSize / kFT;  // 16
            const uint32_t idx0 = (uint32_t)tid * kPer;
            uint32_t lo = 0, hi = nterm;  // last terminal with start <= idx0 (terminal 0 starts at 0)
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if ((tb[mid] >> 16) <= idx0) lo = mid; else hi = mid;
            }
            uint32_t k = lo;
            uint32_t cur = tb[k] & 0xffffu;
            uint32_t nxt = k + 1 < nterm ? tb[k + 1] >> 16 : kNone;
            uint32_t w[kPer / 2];
#pragma unroll
            for (int q = 0; q < kPer; q++) {
                if (idx0 + q == nxt) {
                    k++;
                    cur = tb[k] & 0xffffu;
                    nxt = k + 1 < nterm ? tb[k + 1] >> 16 : kNone;
                }
                if (q & 1) w[q >> 1] |= cur << 16; else w[q >> 1] = cur;
            }
#pragma unroll
            for (int q = 0; q < kPer / 8; q++) {
                reinterpret_cast<uint4 *>(lut)[tid * (kPer / 8) + q] =
                    make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
            }
        }
        cta_sync();

        HUF_PROF(0);
        HUF_PROF_CNT(9);
        // ---- chunk loop
        const uint64_t next_cand = (j + 1 < ncand) ? a.cand[j + 1] : a.avail;
        const uint64_t room_end = 8ull * a.avail;
        uint64_t guess_end = next_cand < a.avail ? next_cand : a.avail;
        guess_end = 8ull * (guess_end > pay0 ? guess_end : pay0);
        bool use_guess = guess_end > 8ull * pay0;
        uint64_t next_bit = 8ull * pay0;  // absolute stream bit of the next code word (proven)
        uint64_t produced = 0, end_bit = 0;
        const uint64_t out0 = a.out_off[j];
        const bool can_write = !a.count_only && out0 + orig_len <= a.out_cap;
        // Sub-block length.  Safe: even a run of the shortest code word cannot overfill a region.
        // Speculative: sized for the block's AVERAGE code length (regions two thirds full; three
        // quarters for short average codes, where it measures 7 % faster -- on the Zipf shape it
        // measures slower: 1.79 against 1.74 ms), which
        // is far longer for ordinary data; a region that does run full makes the CTA repeat the
        // chunk with the safe length and keep it for the rest of the block.
        uint32_t safe_cap_w = min((uint32_t)kMaxSubWords, (kRegCap * min_len) / 32u);
        if (!(safe_cap_w & 1)) safe_cap_w--;  // kRegCap / 32 = 7, so never below 7 (odd: the staged
                                              // words of the threads of a warp start in different banks)
        uint32_t sub_cap_w = safe_cap_w;
        // (sizes are heuristics: single-precision arithmetic is exact enough and much cheaper
        // than 64-bit integer division)
        uint32_t warm = 192;
        if (use_guess) {
            const float avg_bits = (float)(guess_end - 8ull * pay0) / (float)orig_len;  // per code word
            #ifdef HUF_DEC_FILL
            const float w = (float)kRegCap * HUF_DEC_FILL * avg_bits * (1.0f / 32.0f);
#else
            const float w = (float)(avg_bits < 5.0f ? kRegCap * 3 / 4 : kRegCap * 2 / 3) * avg_bits * (1.0f / 32.0f);
#endif
            uint32_t spec = w < (float)kMaxSubWords ? (uint32_t)w : (uint32_t)kMaxSubWords;
            if (!(spec & 1)) spec--;
            if (spec > safe_cap_w && spec <= (uint32_t)kMaxSubWords) sub_cap_w = spec;
            // warm-up distance: ~24 average code words.  A Zipf(1.1) code leaves 0.13 % of the
            // walks unsynchronised after 20 code words and halves that every 2.5 more; with 256
            // sub-blocks per chunk a failed walk (one repair round for the whole CTA) costs more
            // than the four extra code words of everybody
            const float wb = 24.0f * avg_bits;
            warm = wb < 64.0f ? 64u : (wb > 384.0f ? 384u : (uint32_t)wb);
        }
        float inv_chunk_cap = 1.0f / (float)((uint32_t)kFT * 32u * sub_cap_w);
        uint32_t status = kOk;
        for (;;) {
            HUF_PROF_CNT(7);
            const uint64_t limit = use_guess ? guess_end : room_end;
            if (next_bit >= limit) {
                if (use_guess) {
                    use_guess = false;  // the next candidate sat inside this payload: go on
                    continue;
                }
                status = kRedo;  // out of readable bits before orig_len symbols: error lane
                break;
            }
            const uint64_t base16 = (next_bit >> 3) & ~uint64_t(15);
            const uint32_t rel_start = (uint32_t)(next_bit - 8ull * base16);
            const uint32_t A = rel_start & ~31u;
            // split what is left evenly over as few chunks as possible (32-bit arithmetic: a
            // span beyond 4 Gbit only needs the right order of magnitude)
            const uint64_t span64 = limit - (8ull * base16 + A);
            const uint32_t span = span64 > 0xf0000000ull ? 0xf0000000u : (uint32_t)span64;
            // (any odd length between 7 and sub_cap_w words is valid: the divisions only have
            // to be about right, so they are single-precision multiplications)
            const float fspan = (float)span;
            const uint32_t nch = (uint32_t)(fspan * inv_chunk_cap) + 1u;
            uint32_t subw = (uint32_t)(fspan / (float)(nch * (uint32_t)kFT * 32u)) + 1u;
            subw |= 1u;
            if (subw < 7) subw = 7;
            if (subw > sub_cap_w) subw = sub_cap_w;
            const uint32_t sub = subw * 32u;
            const uint64_t lim_rel = limit - 8ull * base16;
            const uint64_t cov64 = (uint64_t)A + (uint64_t)kFT * sub;
            const uint32_t cover = (uint32_t)(cov64 < lim_rel ? cov64 : lim_rel);
            // threads whose sub-block starts inside the covered bits take part in this chunk
            const uint32_t nact = cov64 <= lim_rel ? (uint32_t)kFT : (cover - A + sub - 1) / sub;

            // (0) stage the chunk: 16-byte loads, bytes past `avail` read as zero; the lines of
            // the chunk behind it are requested into L2 meanwhile
            {
                const uint64_t nxt = base16 + ((cover + 7) >> 3) + 128ull * (uint32_t)tid;
                if (nxt < a.avail && 128u * (uint32_t)tid < (kFT * sub >> 3) + 256u) prefetch_l2(a.in + nxt);
                if (tma_pending && tma_base == base16) {
                    stage_wait();  // every thread sees the bytes once the barrier's phase is over
                } else {
                    if (tma_pending) stage_wait();  // (a window nobody wants: let it land first)
                    const uint32_t n16 = ((cover + 7) >> 3) / 16 + 2;
                    for (uint32_t c = tid; c < n16; c += kFT) {
                        const uint64_t byte = base16 + 16ull * c;
                        uint4 v = make_uint4(0, 0, 0, 0);
                        if (byte + 16 <= a.avail) {
                            v = ld_stream_u4(a.in + byte);
                        } else if (byte < a.avail) {
                            uint32_t w[4] = {0, 0, 0, 0};
                            for (int r = 0; r < 16; r++) {
                                if (byte + r < a.avail) w[r >> 2] |= (uint32_t)a.in[byte + r] << (8 * (r & 3));
                            }
                            v = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                        reinterpret_cast<uint4 *>(sw)[c] = v;
                    }
                    cta_sync();
                }
            }
            HUF_PROF(1);

            // (1) warm-up in front of my sub-block, then decode it into my region
            const uint32_t my_lo = tid == 0 ? rel_start : min(A + (uint32_t)tid * sub, cover);
            const uint32_t my_hi = min(A + (uint32_t)(tid + 1) * sub, cover);
            uint32_t pos = my_lo, cnt = 0, last_dead = kNone;
            // blind walk (every entry advances by its length field, dead entries and walks on a
            // set root bit by one bit, long code words by their exact length) from `from` to the
            // first position >= limit
            auto blind_walk = [&](auto long_tag, uint32_t from, uint32_t limit) -> uint32_t {
                constexpr bool HAS_LONG = decltype(long_tag)::value;  // the block has long-code records
                uint32_t p = from;
                if (p >= limit) return p;
                BitWin b;
                win_load(b, sw_s, p);
                for (;;) {
                    // like fast_look4, with the length of a walk on a set root bit forced to 1
                    // (mask = all ones under a set root bit, computed beside the table load)
                    const uint32_t h0 = __funnelshift_l(b.w1, b.w0, p);
                    uint32_t lo = __funnelshift_l(b.w2, b.w1, p);
                    const uint32_t l0 = blind_len(lds_u16(lut_at<kLutOr>(lut_s, fast_idx(h0))), h0);
                    const uint32_t h1 = __funnelshift_l(lo, h0, l0);
                    lo = __funnelshift_l(0u, lo, l0);
                    const uint32_t l1 = blind_len(lds_u16(lut_at<kLutOr>(lut_s, fast_idx(h1))), h1);
                    const uint32_t h2 = __funnelshift_l(lo, h1, l1);
                    lo = __funnelshift_l(0u, lo, l1);
                    const uint32_t l2 = blind_len(lds_u16(lut_at<kLutOr>(lut_s, fast_idx(h2))), h2);
                    const uint32_t h3 = __funnelshift_l(lo, h2, l2);
                    const uint32_t l3 = blind_len(lds_u16(lut_at<kLutOr>(lut_s, fast_idx(h3))), h3);
                    const uint32_t np = p + ((l0 + l1 + l2 + l3) & 0x3fu);
                    if (HAS_LONG && ((l0 | l1 | l2 | l3) & 0x80u)) {
                        // a code word longer than the table reach among the four: this group is
                        // walked with exact steps (a blind single-bit step would leave the
                        // trajectory of the exact walk and cost a repair round later)
                        for (int k = 0; k < 4 && p < limit; k++) p = (uint32_t)fast_step<kLutOr>(sw_s, lut_s, &ft, sm.nlong, p);
                        if (p >= limit) return p;
                        win_load(b, sw_s, p);
                        continue;
                    }
                    if (np < limit) {  // four steps at a time while the fifth starts in front of the limit
                        win_advance(b, sw_s, p, np);
                        p = np;
                        continue;
                    }
                    p += l0 & 0x1fu;
                    if (p < limit) p += l1 & 0x1fu;
                    if (p < limit) p += l2 & 0x1fu;
                    if (p < limit) p += l3 & 0x1fu;
                    return p;
                }
            };
            // (blocks without long code words -- nearly all -- run the loop without that test)
            auto blind_to = [&](uint32_t from, uint32_t limit) -> uint32_t {
                return sm.nlong ? blind_walk(std::true_type{}, from, limit) : blind_walk(std::false_type{}, from, limit);
            };
            // start `warm` bits early (or at the proven chunk start when that is closer)
            const uint32_t warm_from = my_lo > rel_start + warm ? my_lo - warm : rel_start;
            if (tid > 0 && my_lo < my_hi) pos = blind_to(warm_from, my_lo);
            // (1b) decode my sub-block from `start` into my region, (2) verify: my first code
            // word must begin where my predecessor's last one ended.  A thread that was not
            // synchronised decodes its sub-block again from the proven position (codes that
            // synchronise badly make that frequent; the loop is the same fast one), and since
            // its end may move, the check repeats until nothing changes.
            uint32_t start = pos;  // first code word of mine (speculative unless tid == 0)
            uint32_t end = pos;
            uint32_t chk_pos = pos, chk_cnt = 0;  // start of my last look-up group, symbols before it
            bool chk_dead = false;                // a dead step in front of it
            bool walk = true;
            for (int round = 0; round <= kFT; round++) {
                if (walk) {
                    pos = start;
                    last_dead = kNone;
                    chk_pos = start;
                    chk_cnt = 0;
                    chk_dead = false;
                    // Symbol number n of my region lives in byte n & 3 of word n >> 2.  Only whole
                    // words are stored: up to three pending symbols wait in the top bytes of
                    // `acc` (earliest lowest), so a group of four always leaves as one 32-bit
                    // store even after single steps have left the count unaligned.
                    saddr_t wp = reg_s;            // word the next store goes to
                    // a full region keeps storing into the row behind it (no branch in the loop);
                    // reaching that row is what reports the overflow afterwards
                    const saddr_t wp_end = reg_s + (saddr_t)kRegWords * kRegRow;
                    uint32_t acc = 0, npend = 0;   // pending symbols
                    uint32_t sh = 32;              // 32 - 8 * npend
                    auto put = [&](uint32_t sy) {
                        acc = (acc >> 8) | (sy << 24);
                        npend++;
                        if (npend == 4) {
                            sts_u32(wp, acc);
                            wp = min(wp + (saddr_t)kRegRow, wp_end);
                            npend = 0;
                        }
                        sh = 32 - 8 * npend;
                    };
                    BitWin b;
#ifndef HUF_NO_BULK_WALK
                    // Bulk phase: groups of four look-ups that lie inside my sub-block whatever
                    // they decode to (a group consumes at most 4 x 13 bits), with the checks
                    // for special entries and set root bits deferred to the end of the phase:
                    // the entries and windows of all groups are OR-ed together, and if anything
                    // irregular shows up the phase is thrown away and the careful loop below
                    // decodes the sub-block from its start (so what is kept is exactly what that
                    // loop would have produced).  Per group this saves the check, a branch pair
                    // and the pending-byte bookkeeping; the third look-up takes its window from
                    // the group's first 64 bits with the summed lengths (<= 26 bits), which saves
                    // two funnel shifts.  Blocks with long-code records never qualify.
                    if (!sm.nlong && pos + 4u * (uint32_t)kTreeReach <= my_hi) {
                        const uint32_t bulk_last = my_hi - 4u * (uint32_t)kTreeReach;  // last start of a group
                        uint32_t p = pos, fe = 0, fh = 0;
                        uint32_t pw = p >> 5;   // staged word that holds bit p (w0 of the window)
                        saddr_t wq = wp;
                        win_load(b, sw_s, p);
                        do {
                            const uint32_t h0 = __funnelshift_l(b.w1, b.w0, p);
                            const uint32_t lo = __funnelshift_l(b.w2, b.w1, p);
                            const uint32_t e0 = lds_u16(lut_at<kLutOr>(lut_s, bulk_idx(h0, mul14)));
                            const uint32_t h1 = __funnelshift_l(lo, h0, e0);
                            const uint32_t e1 = lds_u16(lut_at<kLutOr>(lut_s, bulk_idx(h1, mul14)));
                            const uint32_t s2 = e0 + e1;                 // low five bits: l0 + l1 <= 26
                            const uint32_t h2 = __funnelshift_l(lo, h0, s2);
                            const uint32_t lo2 = __funnelshift_l(0u, lo, s2);
                            const uint32_t e2 = lds_u16(lut_at<kLutOr>(lut_s, bulk_idx(h2, mul14)));
                            const uint32_t h3 = __funnelshift_l(lo2, h2, e2);
                            const uint32_t e3 = lds_u16(lut_at<kLutOr>(lut_s, bulk_idx(h3, mul14)));
                            const uint32_t np = p + ((s2 + e2 + e3) & 0x3fu);
                            fe |= e0 | e1 | e2 | e3;
                            fh |= h0 | h1 | h2 | h3;
                            const uint32_t lo2s = __byte_perm(e0, e1, 0x0051);
                            const uint32_t hi2s = __byte_perm(e2, e3, 0x0051);
                            sts_u32(wq, __byte_perm(lo2s, hi2s, 0x5410));
                            wq = min(wq + (saddr_t)kRegRow, wp_end);
                            // window advance by 0, 1 or 2 words (the word index travels with the
                            // loop, the address is a multiply-add: both off the integer ALU pipe)
                            const uint32_t nw = np >> 5;
                            win_step(b, saddr_word(nw, sw_s), nw - pw);
                            pw = nw;
                            p = np;
                        } while (p <= bulk_last);
                        if (!((fe & kFastFlags) | (fh >> 31))) {
                            pos = p;   // all plain table hits: keep them
                            wp = wq;
                            HUF_PROF_CNT(12);
                        } else {
                            HUF_PROF_CNT(13);
                        }
                    }
#endif
                    HUF_PROF_CNT(14);
                    if (pos < my_hi) win_load(b, sw_s, pos);
                    while (pos < my_hi) {
                        uint32_t e0, e1, e2, e3, h0, h1, h2, h3;
                        const uint32_t np = fast_look4<kLutOr>(b, lut_s, pos, e0, e1, e2, e3, h0, h1, h2, h3);
                        if (!(((e0 | e1 | e2 | e3) & kFastFlags) | ((h0 | h1 | h2 | h3) >> 31))) {
                            // four plain table hits
                            const uint32_t lo2 = __byte_perm(e0, e1, 0x0051);
                            const uint32_t hi2 = __byte_perm(e2, e3, 0x0051);
                            const uint32_t four = __byte_perm(lo2, hi2, 0x5410);
                            if (np <= my_hi) {
                                // ... that all start inside my sub-block: one 32-bit store
                                win_advance(b, sw_s, pos, np);
                                pos = np;
                                sts_u32(wp, __funnelshift_rc(acc, four, sh));  // pending bytes below, new ones above
                                acc = four;                                    // its top bytes are the new pending ones
                                wp = min(wp + (saddr_t)kRegRow, wp_end);
                                continue;
                            }
                            // the sub-block ends among them: the first is mine, the others may be
                            // (where this last group begins is remembered: should the block end
                            // in it, the search for its exact end bit starts here)
                            chk_pos = pos;
                            chk_cnt = (uint32_t)((wp - reg_s) / (uint32_t)kRegRow) * 4u + npend;
                            chk_dead = last_dead != kNone;
                            put(four & 0xffu);
                            pos += e0 & 0xfu;
                            if (pos < my_hi) {
                                put((four >> 8) & 0xffu);
                                pos += e1 & 0xfu;
                            }
                            if (pos < my_hi) {
                                put((four >> 16) & 0xffu);
                                pos += e2 & 0xfu;
                            }
                            if (pos < my_hi) {
                                put(four >> 24);
                                pos += e3 & 0xfu;
                            }
                            break;
                        }
                        // irregular (special entry or dead root among the four): one exact step
                        {
                            const uint32_t at = pos;
                            const uint64_t r = fast_step<kLutOr>(sw_s, lut_s, &ft, sm.nlong, pos);
                            pos = (uint32_t)r;
                            if (r & kStepOk) {
                                put((uint32_t)(r >> 32) & 0xffu);
                            } else {
                                last_dead = at;
                            }
                        }
                        if (pos < my_hi) win_load(b, sw_s, pos);
                    }
                    if (npend) sts_u32(wp, acc >> sh);
                    if (wp == wp_end) sm.ovf = 1;
                    cnt = (uint32_t)((wp - reg_s) / (uint32_t)kRegRow) * 4u + npend;
                    end = pos;
                    sm.sub_end[tid] = end;
                }
                if (round == 0) HUF_PROF(2);
                cta_sync();
                if (sm.ovf) break;  // (uniform: nobody walks again before the next barrier)
                const uint32_t want = (tid == 0 || (uint32_t)tid >= nact) ? start : sm.sub_end[tid - 1];
                walk = want != start;
                const uint32_t had = start;  // what my region was decoded from
                start = want;
                __syncwarp();
                const int nfail = __syncthreads_count(walk);
                if (nfail) HUF_PROF_CNT(8);
                if (nfail == 0) break;
                if (round == 1 && nfail >= kRespecMin) {
                    // Re-speculation.  Isolated failures are gone after one repair round; threads
                    // that still fail do so because a repaired predecessor ended elsewhere: the
                    // code synchronises badly (e.g. p(k) = 2^-(k+1) keeps a shifted parse shifted
                    // for ever) and repairs would ripple through the chunk one round each.  Every
                    // thread looks for a second trajectory class (blind walks from 16 adjacent
                    // start bits reach every trajectory that crosses the warm-up zone), follows
                    // it through its sub-block, and one thread chains the classes from the proven
                    // chunk start.  The result only PREDICTS starts: the verification rounds
                    // below still decide, so a wrong prediction costs time, never correctness.
                    uint32_t alt = kNone, alt_end = kNone;
                    if (tid > 0 && my_lo < my_hi) {
                        for (uint32_t k = 1; k < (uint32_t)kRespecStarts && alt == kNone; k++) {
                            if (warm_from + k >= my_lo) break;
                            const uint32_t b = blind_to(warm_from + k, my_lo);
                            if (b != had) alt = b;
                        }
                        if (alt != kNone) alt_end = blind_to(alt, my_hi);
                    }
                    ft.c_start[0][tid] = had;
                    ft.c_end[0][tid] = end;
                    ft.c_start[1][tid] = alt;
                    ft.c_end[1][tid] = alt_end;
                    cta_sync();
                    if (tid == 0) {
                        uint32_t cur = ft.c_end[0][0];  // thread 0 decoded from the proven start
                        ft.pred[0] = ft.c_start[0][0];
                        for (uint32_t t = 1; t < nact; t++) {
                            ft.pred[t] = cur;
                            if (cur == ft.c_start[0][t]) {
                                cur = ft.c_end[0][t];
                            } else if (cur == ft.c_start[1][t]) {
                                cur = ft.c_end[1][t];
                            } else {
                                cur = ft.c_end[0][t];  // unknown class: guess, verification repairs
                            }
                        }
                    }
                    cta_sync();
                    if ((uint32_t)tid < nact) {
                        start = ft.pred[tid];
                        walk = start != had;
                    }
                }
            }
            if (sm.ovf) {
                // a region ran full under the speculative sub-block length: nothing of this
                // chunk has left shared memory yet, so it is simply decoded again
                cta_sync();
                if (tid == 0) sm.ovf = 0;
                sub_cap_w = safe_cap_w;
                inv_chunk_cap = 1.0f / (float)((uint32_t)kFT * 32u * sub_cap_w);
                HUF_PROF(3);
                continue;
            }
            HUF_PROF(3);
            // (3) symbol-count scan
            const uint32_t incl = warp_incl_scan(cnt);
            if ((tid & 31) == 31) sm.warp_tot[tid >> 5] = incl;
            if (tid == 0) sm.fin_found = 0;
            cta_sync();
            // (a handful of warp totals: every thread adds up the ones in front of its warp
            // itself, which saves a second barrier round)
            uint32_t before = incl - cnt, total = 0;
#pragma unroll
            for (int wq = 0; wq < kFT / 32; wq++) {
                const uint32_t t = sm.warp_tot[wq];
                if (wq < (tid >> 5)) before += t;
                total += t;
            }
            const uint32_t last_end = sm.sub_end[nact - 1];  // (final since the verification; read
                                                             // here, behind a barrier, so that the
                                                             // chunk needs none at its end)
            const uint64_t remaining = orig_len - produced;
            const bool in_blk = (uint64_t)before < remaining && cnt > 0;
            const bool fin = in_blk && (uint64_t)before + cnt >= remaining;
            const uint32_t ncopy = in_blk ? (fin ? (uint32_t)(remaining - before) : cnt) : 0;
            if (in_blk && !fin && last_dead != kNone) sm.redo = 1;  // dead walk on the proven chain
            if (fin) {
                // the block ends inside my sub-block: find the bit behind its last code word
                // (a dead step ends the search: the block goes to the general lane anyway, and
                // the symbol count of a dead walk does not bound where this one would stop)
                // The search starts at the last look-up group of my walk when the block ends in
                // it (the usual case: the block ends where the next header was found), else at
                // my first code word; whole groups of four are taken with the table loop.
                uint32_t p = start, q = 0, dead = 0;
                if (chk_cnt <= ncopy) {
                    p = chk_pos;
                    q = chk_cnt;
                    dead = chk_dead ? 1u : 0u;
                }
                while (q < ncopy && !dead) {
                    if (q + 4 <= ncopy) {
                        BitWin b;
                        win_load(b, sw_s, p);
                        uint32_t e0, e1, e2, e3, h0, h1, h2, h3;
                        const uint32_t np = fast_look4<kLutOr>(b, lut_s, p, e0, e1, e2, e3, h0, h1, h2, h3);
                        if (!(((e0 | e1 | e2 | e3) & kFastFlags) | ((h0 | h1 | h2 | h3) >> 31))) {
                            p = np;
                            q += 4;
                            continue;
                        }
                    }
                    const uint64_t r = fast_step<kLutOr>(sw_s, lut_s, &ft, sm.nlong, p);
                    p = (uint32_t)r;
                    if (r & kStepOk) q++; else dead = 1;
                }
                if (dead) sm.redo = 1;
                sm.fin_end = p;
                sm.fin_found = 1;
            }
            cta_sync();
            HUF_PROF(4);
            if (sm.redo) {
                status = kRedo;
                break;
            }
            const bool fin_found = sm.fin_found != 0;
            const uint32_t total_copy = (uint64_t)total < remaining ? total : (uint32_t)remaining;
            HUF_PROF(10);

            // (4) compaction into the (now free) stage buffer, coalesced copy-out
            if (can_write && total_copy) {
                uint8_t *dst0 = a.out + out0 + produced;  // first output byte of this chunk
                const uint32_t m = (uint32_t)(reinterpret_cast<uintptr_t>(dst0) & 15);
                uint8_t *obuf = dyn;
                const uint32_t y0 = m + before, y1 = y0 + ncopy;  // my bytes in aligned position space
                const uint32_t yend = m + total_copy;
                for (uint32_t wb = 0; wb < yend; wb += kFastOutWin) {
                    const uint32_t we = wb + kFastOutWin;
                    const uint32_t c0 = max(y0, wb), c1 = min(y1, we);
                    if (c0 < c1) region_copy(obuf + (c0 - wb), regw, c0 - y0, c1 - c0);
                    HUF_PROF(11);
                    cta_sync();
                    HUF_PROF(5);
                    const uint32_t vend = min(yend, we) - wb;          // valid bytes end (window relative)
                    const uint32_t vbeg = wb == 0 ? m : 0;             // valid bytes begin
                    uint8_t *gbase = dst0 - m + wb;                    // 16-byte aligned
                    const uint32_t nlines = (vend + 15) >> 4;
                    // whole lines: 16-byte stores; the first and the last line may be partial:
                    // their bytes go one per thread (threads 0..15 and 16..31)
                    const uint32_t l_lo = (vbeg + 15) >> 4, l_hi = vend >> 4;
                    for (uint32_t L = l_lo + tid; L < l_hi; L += kFT) {
                        *reinterpret_cast<uint4 *>(gbase + 16 * L) = *reinterpret_cast<const uint4 *>(obuf + 16 * L);
                    }
                    if (tid < 32) {
                        // line 0 when it starts late, line l_hi when it ends early (they may coincide)
                        const uint32_t q = (tid < 16 ? 0u : 16u * l_hi) + ((uint32_t)tid & 15u);
                        const bool first_part = tid < 16 && l_lo > 0;
                        const bool last_part = tid >= 16 && l_hi < nlines && (l_hi > 0 || l_lo == 0);
                        if ((first_part || last_part) && q >= vbeg && q < vend) gbase[q] = obuf[q];
                    }
                    cta_sync();
                }
            }
            HUF_PROF(6);
            produced += total_copy;
            if (fin_found) {
                end_bit = 8ull * base16 + sm.fin_end;
                break;
            }
            next_bit = 8ull * base16 + last_end;
            // (the copy-out ended with a barrier: nobody reads the stage buffer any more; a
            // chunk that wrote nothing has not passed one since the region copies)
            if (!(can_write && total_copy)) cta_sync();
            stage_request((next_bit >> 3) & ~uint64_t(15));
        }

        if (tma_pending) stage_wait();  // (left early: the stage buffer must be quiet for the next block)
        if (status == kOk && end_bit > room_end) status = kRedo;  // last code word leaves the readable bytes
        if (tid == 0) {
            a.blk_status[j] = status;
            a.end_off[j] = (end_bit + 7) >> 3;
            if (status == kRedo) atomicAdd(reinterpret_cast<unsigned long long *>(&a.result[10]), 1ull);
        }
    }
#ifdef HUF_PHASE_PROF
    if (tid == 0) {
        for (int k = 0; k < 16; k++) atomicAdd(&g_fast_prof[k], pacc[k]);
    }
#endif
}

#ifdef HUF_EMU
#define HUF_DYN_SMEM(name) uint8_t *name = hufemu::dyn_smem()
#else
#define HUF_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#endif

// The instance every launch uses today: dynamic shared memory starts at shared-window address
// 0x400 (no static shared memory, 1 KB reserved by the system), which puts the table at 0x4000.
__global__ void __launch_bounds__(kFT, HUF_DEC_MINCTA) k_decode(DecArgs a)
{
    HUF_DYN_SMEM(dyn);
    decode_fast_body<true>(a, dyn);
}

// Same kernel with the table offset ADDED to its base: launched when the probe finds the table
// at an address that is not 8 KB aligned (a driver or toolkit that reserves a different amount
// of shared memory), one more integer instruction per code word.
__global__ void __launch_bounds__(kFT, HUF_DEC_MINCTA) k_decode_unaligned(DecArgs a)
{
    HUF_DYN_SMEM(dyn);
    decode_fast_body<false>(a, dyn);
}

// Where does dynamic shared memory of a kernel without static shared memory start?
__global__ void k_smem_base(uint32_t *out)
{
    HUF_DYN_SMEM(dyn);
    if (threadIdx.x == 0) out[0] = (uint32_t)smem_addr(dyn);
}

}  // namespace hufb200
