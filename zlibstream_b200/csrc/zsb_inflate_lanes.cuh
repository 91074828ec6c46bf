// zsb_inflate_lanes.cuh -- inflate of full-flush segments, ONE LANE PER SEGMENT (32 segments per warp).
//
// Replaces the same reference code as zsb_inflate.cu: Inflate.Decompress (Inflate.cs:103-357), InflateBlocks.Process
// (InflateBlocks.cs:130-713), InfCodes.Inflate_fast / Process (InfCodes.cs:106-817), InfTree.Huft_build
// (InfTree.cs:125-365, RFC 1951 semantics -- SURVEY Q15).
//
// Why: with a warp per segment the symbol loop keeps one lane of 32 busy and the kernel is issue-bound (~50 warp
// instructions per symbol, 33 GB/s).  Here every lane runs the whole decoder state machine on its own segment, so one
// issued instruction serves up to 32 segments.  What makes that workable:
//   * per-lane state is small: a 9-bit literal/length root table and a 6-bit distance root table of u16 entries in
//     shared memory (1.4 KiB per lane, 4 warps per SM); longer codes take a canonical-code search over <= 6 (9) packed
//     limits, their symbols kept in a short per-lane list;
//   * the state machine does a bounded amount of work per iteration in every state (a few symbols, a few code lengths,
//     a few 8-byte pieces of a stored block), so a lane parsing a block header never stalls the other 31 for long;
//   * the only warp-cooperative step is the table build: the lanes that need one are served in turn by all 32 lanes
//     (counting sort of the code lengths, root-table fill), ~1-2 k cycles per block instead of ~10 k for one lane alone;
//   * output is assembled in a 64-bit register and stored as aligned 8-byte words; a match is copied 8 bytes at a time
//     from the segment's own earlier output (L1/L2), a run (distance < 15) is widened to a multiple of its period first.
//
// Everything except the cooperative build's warp plumbing is __host__ __device__: tests/ build a host emulation of the
// state machine from this very header (tools/inflate_lanes_host.cu) to check the logic without a GPU.  That emulation is
// test infrastructure; the product path is the kernel in zsb_inflate_lanes.cu.
#pragma once

#include <string.h>

#include "zsb_kernels.cuh"

namespace zsb {
// ZSB_LN_NS: the translation unit zsb_inflate_lanes_small.cu builds a second instance of everything below with smaller
// tables (more resident warps, for inputs with far more segments than resident lanes) in its own namespace.
#ifndef ZSB_LN_NS
#define ZSB_LN_NS lanes
#endif
namespace ZSB_LN_NS {

#define ZL_HD __host__ __device__ __forceinline__

#ifndef ZSB_LN_ROOT_L
#define ZSB_LN_ROOT_L 8
#endif
constexpr int ROOT_L = ZSB_LN_ROOT_L;   // literal/length root table bits
#ifndef ZSB_LN_ROOT_D
#define ZSB_LN_ROOT_D 6
#endif
constexpr int ROOT_D = ZSB_LN_ROOT_D;   // distance root table bits
#ifndef ZSB_LN_LIT_STREAK
#define ZSB_LN_LIT_STREAK 4
#endif
constexpr int LIT_STREAK = ZSB_LN_LIT_STREAK;   // literals taken in one step (<= 8: they travel in one 64-bit word)
constexpr int ROOT_BL = 7;              // bit-length code: always complete within the root (MAX_BL_BITS)
#ifndef ZSB_LN_LONG_L
#define ZSB_LN_LONG_L 96
#endif
constexpr int LONG_L = ZSB_LN_LONG_L;            // long-code literal/length symbols kept in shared memory (the rest: global)
constexpr int LENS_STRIDE = 320;        // per-lane global scratch: code lengths of the block header (bytes)
constexpr int GLONG_STRIDE = 288;       // per-lane global scratch: all long-code literal/length symbols (u16)
constexpr int SCRATCH_STRIDE = LENS_STRIDE + 2 * GLONG_STRIDE;   // bytes per lane
static_assert(ROOT_L >= ROOT_BL && ROOT_L <= 12, "the bit-length table is built in the literal/length root table");

struct LaneTables {                       // per lane, shared memory
    uint16_t root_l[1 << ROOT_L];         // (symbol << 4) | code length; 0 = longer than the root or unused
    uint16_t root_d[1 << ROOT_D];
    uint32_t meta_l[16 - ROOT_L];         // [0] = limit(ROOT);  [i] = limit(ROOT+i) | long-list offset << 16
    uint32_t meta_d[16 - ROOT_D];         //   limit(l) = left-justified (15-bit) end of the codes of length <= l
    uint16_t long_l[LONG_L];
    uint8_t long_d[32];
};

struct WarpScratch {                      // per CTA, shared memory: the cooperative table build (one warp at a time, see `lock`)
    unsigned long long loc[32];           // per helper: count of each literal/length code length, 4 bits each
    uint16_t pre[16][32];                 // [length][helper]: symbols of that length owned by earlier helpers
    uint8_t pre_d[16][32];
    uint8_t dlen[32];                     // code length of the small alphabet's symbol `helper`
    uint16_t cnt_l[16], cnt_d[16], first_l[16], first_d[16], off_l[16], off_d[16];
    int rc_l, rc_d;
    int lock, pad_;                       // 1 while a warp of the CTA runs a build
};

struct HelperRegs {                       // what a helper lane carries from phase 0 to phase 3
    uint8_t ll[9];
    uint8_t dl;
};

enum LaneStateId { S_BLOCK = 0, S_SYM, S_CLEN, S_STORED, S_BUILD_BL, S_BUILD_LD, S_FINISH, S_EXIT };

struct Lane {
    const uint8_t* in;     // first compressed byte of the segment
    uint32_t pos;          // next byte to load, relative to `in`
    uint32_t end;          // compressed size of the segment
    uint32_t limit;        // bytes that may be loaded (to the end of the whole compressed buffer)
    uint64_t acc;          // bit accumulator, LSB first; bits above nbits are zero
    int nbits;
    uint32_t pre;          // the input word at `pos`, loaded one refill ahead (the L1 left beside 224 KiB of tables is
    int pre_ok;            //   too small to keep input lines, so an on-demand word costs an L2 round trip)
    uint8_t* ob8;          // 8-byte aligned base of the output: segment start minus sh
    uint32_t sh;           // misalignment of the segment's first output byte (0 for cudaMalloc'ed buffers)
    uint32_t v;            // sh + bytes produced; wbuf holds the bytes [v & ~7, v)
    uint32_t vcap;         // sh + room
    uint64_t wbuf;
    uint64_t wprev;        // the word completed last: with wbuf, the most recent 8..15 output bytes without a memory access
    int st, err;
    uint32_t seg;
    int nl, nd, ci, prevlen, fixed;
    uint32_t stored_rem;
    uint32_t cp_rem;       // bytes of the current symbol not yet appended (incl. the pending unit)
    uint32_t pend_n;       // bytes of the pending unit, 0 = none
    int pend_kind;         // UnitKind
    uint64_t pend_val;     // U_LIT: one to LIT_STREAK bytes
    uint32_t cp_D;         // U_MEM: distance
    uint64_t plo, phi;     // U_PAT: 16 bytes of the periodic sequence
    uint32_t ph, pd;       //        phase of the next unit, period
    int count;             // count-only pass (segment index building): nothing is loaded from or stored to `out`
    uint32_t found;        // count-only: offset just past the full-flush marker that ended the piece, 0 = none yet
    uint32_t carry;        // Q14: leading bytes that repeat the previous segment's last byte (Rle streams only)
    uint32_t carry_ok;     // this segment is not the stream's first: it may open with such a run
    uint32_t final_seen;   // a block with BFINAL has been entered: the segment (a one-piece stream) ends with it
    uint64_t bl_lens;      // 19 x 3 bits
};

// ---- primitives ------------------------------------------------------------------------------------------------------
ZL_HD uint32_t ld_in_u32(const uint8_t* p)   // aligned, read-only input
{
#ifdef __CUDA_ARCH__
    return __ldg(reinterpret_cast<const uint32_t*>(p));
#else
    uint32_t v;
    memcpy(&v, p, 4);
    return v;
#endif
}
ZL_HD uint64_t ld_in_u64(const uint8_t* p)
{
#ifdef __CUDA_ARCH__
    return __ldg(reinterpret_cast<const unsigned long long*>(p));
#else
    uint64_t v;
    memcpy(&v, p, 8);
    return v;
#endif
}
ZL_HD uint8_t ld_in_u8(const uint8_t* p)
{
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
// the segment's own earlier output: plain (coherent) loads -- a thread always sees its own stores
ZL_HD uint64_t ld_out_u64(const uint8_t* p)
{
#ifdef __CUDA_ARCH__
    return *reinterpret_cast<const unsigned long long*>(p);
#else
    uint64_t v;
    memcpy(&v, p, 8);
    return v;
#endif
}
ZL_HD void st_out_u64(uint8_t* p, uint64_t v)
{
#ifdef __CUDA_ARCH__
    *reinterpret_cast<unsigned long long*>(p) = v;
#else
    memcpy(p, &v, 8);
#endif
}
ZL_HD void st_out_u8(uint8_t* p, uint8_t v)
{
#ifdef __CUDA_ARCH__
    *p = v;
#else
    *p = v;
#endif
}
ZL_HD uint8_t ld_scratch_u8(const uint8_t* p)   // written by another lane of the same warp: read at L2
{
#ifdef __CUDA_ARCH__
    return __ldcg(p);
#else
    return *p;
#endif
}
ZL_HD uint16_t ld_scratch_u16(const uint16_t* p)
{
#ifdef __CUDA_ARCH__
    return __ldcg(p);
#else
    return *p;
#endif
}
ZL_HD uint32_t brev32(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return __brev(x);
#else
    x = (x >> 16) | (x << 16);
    x = ((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8);
    x = ((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2);
    x = ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1);
    return x;
#endif
}

// ---- bit reader ------------------------------------------------------------------------------------------------------
// After refill() nbits >= 33 unless the input is exhausted.  Bits are dropped without checking; running out shows up as
// nbits < 0 at the end of the step (missing bits read as zero in between).
ZL_HD void refill(Lane& L)
{
    if (L.nbits > 32) return;
    const uint8_t* p = L.in + L.pos;
    if (((uintptr_t)p & 3u) == 0 && L.pos + 4 <= L.limit) {
        const uint32_t w = L.pre_ok ? L.pre : ld_in_u32(p);
        L.acc |= (uint64_t)w << L.nbits;
        L.pos += 4;
        L.nbits += 32;
        L.pre_ok = L.pos + 4 <= L.limit;
        if (L.pre_ok) L.pre = ld_in_u32(p + 4);   // not needed before the next refill: its latency hides behind ~4 symbols
        return;
    }
    L.pre_ok = 0;
    // start of a segment, after a stored block, last bytes of the buffer
    while (L.nbits <= 56 && L.pos < L.limit) {
        L.acc |= (uint64_t)ld_in_u8(L.in + L.pos) << L.nbits;
        L.pos++;
        L.nbits += 8;
        if (L.nbits > 32 && ((uintptr_t)(L.in + L.pos) & 3u) == 0) break;
    }
}
ZL_HD uint32_t peek(const Lane& L, int k) { return (uint32_t)L.acc & ((1u << k) - 1u); }   // k <= 16
ZL_HD void drop(Lane& L, int k)
{
    L.acc >>= k;
    L.nbits -= k;
}
ZL_HD uint32_t byte_pos(const Lane& L) { return L.pos - (uint32_t)(L.nbits >> 3); }

// ---- output ----------------------------------------------------------------------------------------------------------
ZL_HD void store_word(Lane& L, uint32_t vw, uint64_t w)
{
    if (L.count) return;
    if (vw >= L.sh) st_out_u64(L.ob8 + vw, w);
    else
        for (uint32_t k = L.sh - vw; k < 8; k++) st_out_u8(L.ob8 + vw + k, (uint8_t)(w >> (8 * k)));   // first word of an unaligned segment
}
ZL_HD void append_byte(Lane& L, uint32_t b)
{
    const uint32_t a = L.v & 7u;
    L.wbuf |= (uint64_t)b << (8 * a);
    L.v++;
    if (a == 7) {
        store_word(L, L.v - 8, L.wbuf);
        L.wprev = L.wbuf;
        L.wbuf = 0;
    }
}
// n = 1..8 bytes, val has nothing above them
ZL_HD void append_chunk(Lane& L, uint64_t val, uint32_t n)
{
    const uint32_t a = L.v & 7u;
    L.wbuf |= val << (8 * a);
    if (a + n >= 8) {
        store_word(L, L.v & ~7u, L.wbuf);
        L.wprev = L.wbuf;
        L.wbuf = a ? val >> (8 * (8 - a)) : 0;
    }
    L.v += n;
}
// makes the bytes of the partial word visible in memory (they stay in wbuf as well)
ZL_HD void flush_partial(Lane& L)
{
    if (L.count) return;
    const uint32_t a = L.v & 7u, vw = L.v & ~7u;
    for (uint32_t k = 0; k < a; k++)
        if (vw + k >= L.sh) st_out_u8(L.ob8 + vw + k, (uint8_t)(L.wbuf >> (8 * k)));
}
// n = 1..8 bytes of the segment's output at sp; every one of them is in memory (sp + n <= v & ~7, or flushed)
ZL_HD uint64_t load_out8(const Lane& L, uint32_t sp, uint32_t n)
{
    const uint32_t q = sp & ~7u, b = sp & 7u;
    uint64_t val = ld_out_u64(L.ob8 + q) >> (8 * b);
    if (b + n > 8) val |= ld_out_u64(L.ob8 + q + 8) << (8 * (8 - b));
    if (n < 8) val &= (1ull << (8 * n)) - 1ull;
    return val;
}

// ---- output units ----------------------------------------------------------------------------------------------------
// Every symbol is emitted as "units" of at most 8 bytes, one unit per iteration of the lane (so a long match never makes
// the other 31 lanes wait): a literal is a unit of one byte; a match is cut into 8-byte units that are
//   U_MEM   read from the segment's earlier output (distance >= 9; below 15 the word under assembly is flushed first),
//   U_PAT   generated from a 16-byte periodic extension of the last `d` bytes held in registers (distance <= 8: runs),
//   U_ZERO  placeholders of a Q14 carried run.
enum UnitKind { U_LIT = 0, U_MEM, U_PAT, U_ZERO };

// Sets up U_PAT for a match of distance d <= 8 at the current position (InfCodes.cs:243-283 copies byte by byte).
ZL_HD void pattern_setup(Lane& L, uint32_t d)
{
    L.ph = 0;
    L.pd = d;
    if (L.count) return;
    // the last d <= 8 bytes are in registers: `a` bytes in the word under assembly, the eight before them in wprev
    // (a distance never reaches before the segment, so what is needed of wprev has been written)
    const uint32_t a = L.v & 7u, sft = 8u + a - d;   // bytes to drop from the 8 + a recent ones
    uint64_t lo = sft < 8 ? (L.wprev >> (8 * sft)) | (sft ? L.wbuf << (8 * (8 - sft)) : 0) : L.wbuf >> (8 * (sft - 8));
    if (d < 8) lo &= (1ull << (8 * d)) - 1ull;
    uint64_t hi = 0;
    for (uint32_t f = d; f < 16; f *= 2) {   // (lo, hi) |= (lo, hi) << 8f
        if (f < 8) {
            hi |= (hi << (8 * f)) | (lo >> (64 - 8 * f));
            lo |= lo << (8 * f);
        } else hi |= lo << (8 * (f - 8));
    }
    L.plo = lo;
    L.phi = hi;
    L.ph = 0;
    L.pd = d;
}

// ---- steps -----------------------------------------------------------------------------------------------------------
ZL_HD void lane_fail(Lane& L, int e)
{
    L.err = e;
    L.st = S_FINISH;
}

ZL_HD int bl_order(int i)   // InflateBlocks.cs:21 `border`
{
    // 16 17 18 0 8 7 9 6 10 5 11 4 | 12 3 13 2 14 1 15, five bits each
    const unsigned long long lo = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 | 7ull << 25 | 9ull << 30 | 6ull << 35 |
                                  10ull << 40 | 5ull << 45 | 11ull << 50 | 4ull << 55;
    const unsigned long long hi = 12ull | 3ull << 5 | 13ull << 10 | 2ull << 15 | 14ull << 20 | 1ull << 25 | 15ull << 30;
    return (int)((i < 12 ? lo >> (5 * i) : hi >> (5 * (i - 12))) & 31u);
}

// Block header (InflateBlocks.cs:190-250, 366-470)
ZL_HD void block_step(Lane& L)
{
    if (L.final_seen && !L.count) {   // the final block of a one-piece stream (Inflate.cs:286-300): nothing follows but padding
        L.st = S_FINISH;
        return;
    }
    const uint32_t used = byte_pos(L);
    if (L.nbits < 0 || used > L.end) return lane_fail(L, IE_TRUNCATED);   // ran past the segment's bytes
    if (used == L.end) {                                                  // clean end, after the byte-aligned full-flush marker
        if (L.count) return lane_fail(L, IE_TRUNCATED);                   // a piece ends AT a marker, not at the end of the data
        L.st = S_FINISH;
        return;
    }
    refill(L);
    if (L.nbits < 3) return lane_fail(L, IE_TRUNCATED);
    const uint32_t hdr = peek(L, 3);
    drop(L, 3);
    const uint32_t type = hdr >> 1;
    L.final_seen = hdr & 1u;          // segments of a full-flush stream never hold one: the stream's `03 00` lies outside them
    if (type == 0) {
        drop(L, L.nbits & 7);
        refill(L);
        if (L.nbits < 32) return lane_fail(L, IE_TRUNCATED);
        const uint32_t l = peek(L, 16);
        drop(L, 16);
        const uint32_t nl = peek(L, 16);
        drop(L, 16);
        if (((~nl) & 0xFFFFu) != l) return lane_fail(L, IE_STORED_LEN);
        const uint32_t src = byte_pos(L);
        if (src + l > L.end) return lane_fail(L, IE_TRUNCATED);
        if (L.v + l > L.vcap) return lane_fail(L, IE_OUTPUT_OVERFLOW);
        L.pos = src;
        L.pre_ok = 0;
        L.acc = 0;
        L.nbits = 0;
        L.stored_rem = l;
        L.st = l ? S_STORED : S_BLOCK;
        if (L.count && l == 0) {   // the marker that ends every full flush
            L.found = src;
            L.st = S_FINISH;
        }
    } else if (type == 1) {
        if (peek(L, 7) == 0) {   // the empty static block of every flush: END_BLOCK is seven zero bits
            drop(L, 7);
            if (L.nbits < 0) return lane_fail(L, IE_TRUNCATED);
        } else {
            L.fixed = 1;
            L.nl = 288;
            L.nd = 30;
            L.st = S_BUILD_LD;
        }
    } else if (type == 2) {
        refill(L);
        const uint32_t v = peek(L, 14);
        drop(L, 14);
        if ((v & 0x1f) > 29 || ((v >> 5) & 0x1f) > 29) return lane_fail(L, IE_TOO_MANY_SYMS);
        L.nl = 257 + (int)(v & 0x1f);
        L.nd = 1 + (int)((v >> 5) & 0x1f);
        const int nb = 4 + (int)(v >> 10);
        uint64_t bl = 0;
        for (int i = 0; i < nb; i++) {
            refill(L);
            bl |= (uint64_t)peek(L, 3) << (3 * bl_order(i));
            drop(L, 3);
        }
        if (L.nbits < 0) return lane_fail(L, IE_TRUNCATED);
        L.bl_lens = bl;
        L.fixed = 0;
        L.st = S_BUILD_BL;
    } else return lane_fail(L, IE_BLOCK_TYPE);
}

// Stored block body (InflateBlocks.cs:330-364): up to 32 bytes per step
ZL_HD void stored_step(Lane& L)
{
    for (int it = 0; it < 4 && L.stored_rem; it++) {
        const uint32_t n = L.stored_rem < 8 ? L.stored_rem : 8;
        const uint8_t* p = L.in + L.pos;
        const uint32_t b = (uint32_t)((uintptr_t)p & 7u);
        const uint8_t* q = p - b;
        uint64_t val = 0;
        if (!L.count) {
            val = ld_in_u64(q) >> (8 * b);
            if (b + n > 8) val |= ld_in_u64(q + 8) << (8 * (8 - b));
            if (n < 8) val &= (1ull << (8 * n)) - 1ull;
        }
        append_chunk(L, val, n);
        L.pos += n;
        L.stored_rem -= n;
    }
    if (!L.stored_rem) L.st = S_BLOCK;
}

// One code of the dynamic header's code-length sequence (InflateBlocks.cs:472-569), decoded through the bit-length
// table that sits in the literal/length root table until the real one is built.
ZL_HD void clen_step(Lane& L, const LaneTables& T, uint8_t* glens)
{
    refill(L);
    const uint32_t e = T.root_l[L.acc & ((1u << ROOT_BL) - 1u)];
    if (!e) return lane_fail(L, L.nbits <= 0 ? IE_TRUNCATED : IE_BAD_REPEAT);
    drop(L, (int)(e & 15u));
    const int sym = (int)(e >> 4), total = L.nl + L.nd;
    if (sym < 16) {
        glens[L.ci++] = (uint8_t)sym;
        L.prevlen = sym;
    } else {
        const int eb = sym == 18 ? 7 : sym - 14, rb = sym == 18 ? 11 : 3;
        int rep = rb + (int)peek(L, eb);
        drop(L, eb);
        if (L.ci + rep > total || (sym == 16 && L.ci < 1)) return lane_fail(L, IE_BAD_REPEAT);
        const uint8_t val = sym == 16 ? (uint8_t)L.prevlen : (uint8_t)0;
        if (sym != 16) L.prevlen = 0;
        while (rep--) glens[L.ci++] = val;
    }
    if (L.nbits < 0) return lane_fail(L, IE_TRUNCATED);
    if (L.ci == total) L.st = S_BUILD_LD;
}

// Canonical search for a code longer than the root table.  Returns (symbol index in the long list << 4) | length, or 0.
template <int ROOT>
ZL_HD uint32_t long_code(const uint32_t* meta, uint64_t acc)
{
    const uint32_t c = brev32((uint32_t)acc) >> 17;   // the next 15 bits, first bit on top
    uint32_t prev = meta[0] & 0xffffu;
#pragma unroll
    for (int i = 1; i <= 15 - ROOT; i++) {
        const uint32_t m = meta[i], lim = m & 0xffffu;
        if (c < lim) return (((m >> 16) + ((c - prev) >> (15 - ROOT - i))) << 4) | (uint32_t)(ROOT + i);
        prev = lim;
    }
    return 0;
}

// One iteration of the symbol loop (InfCodes.cs:106-369), software-pipelined so that the L2 round trip of a match copy
// hides behind the decoding of the NEXT symbol:
//   (a) issue the loads of the pending unit (its source address is known since the previous iteration),
//   (b) if that unit completes its symbol, decode the next literal/length (+ distance) code -- bit work and shared-memory
//       lookups only, nothing that depends on the output,
//   (c) append the pending unit,
//   (d) turn the decoded symbol into the next pending unit (checks against the output position happen here).
// Nothing is pending across a block boundary: END_BLOCK leaves pend_n == 0.
ZL_HD void sym_step(Lane& L, const LaneTables& T, const uint16_t* glong)
{
    // (a)
    uint64_t w0 = 0, w1 = 0;
    const uint32_t pn = L.pend_n;
    if (pn && L.pend_kind == U_MEM && !L.count) {
        if (L.cp_D < 15) flush_partial(L);
        const uint32_t sp = L.v - L.cp_D, q = sp & ~7u, b = sp & 7u;
        w0 = ld_out_u64(L.ob8 + q);
        if (b + pn > 8) w1 = ld_out_u64(L.ob8 + q + 8);
    }
    // (b)
    const bool dec = L.cp_rem == pn;
    int sym = 0;
    uint64_t lits = 0;
    uint32_t nlit = 0;
    uint32_t len = 0, d = 0;
    int derr = IE_OK;
    if (dec) {
        refill(L);
        uint32_t e = T.root_l[L.acc & ((1u << ROOT_L) - 1u)];
        if (e) {
            drop(L, (int)(e & 15u));
            sym = (int)(e >> 4);
        } else {
            e = long_code<ROOT_L>(T.meta_l, L.acc);
            if (!e) derr = L.nbits <= 0 ? IE_TRUNCATED : IE_BAD_LCODE;
            else {
                drop(L, (int)(e & 15u));
                const uint32_t idx = e >> 4;
                sym = idx < (uint32_t)LONG_L ? T.long_l[idx] : ld_scratch_u16(glong + idx);
            }
        }
        if (sym < 256 && derr == IE_OK) {
            // more literals in the same step while the next code is one too (root-table codes only; most of a text
            // stream's symbols are literals): up to four bytes per step for a dozen instructions each
            lits = (uint64_t)sym;
            nlit = 1;
#pragma unroll
            for (int x = 1; x < LIT_STREAK; x++) {
                if (x == 4) refill(L);
                const uint32_t e2 = T.root_l[L.acc & ((1u << ROOT_L) - 1u)];
                if (!(e2 >= 16u && e2 < (256u << 4) && (int)(e2 & 15u) <= L.nbits)) break;
                drop(L, (int)(e2 & 15u));
                lits |= (uint64_t)(e2 >> 4) << (8 * x);
                nlit = (uint32_t)x + 1;
            }
        }
        if (sym > 256 && derr == IE_OK) {
            const int ls = sym - 257;
            if (ls >= 29) derr = IE_BAD_LCODE;
            else {
                {   // Trees.cs:85-115 length code -> base + extra bits
                    const int eb = (ls < 8 || ls == 28) ? 0 : (ls - 4) >> 2;
                    const uint32_t base = ls < 8 ? 3u + (uint32_t)ls : ls == 28 ? 258u : 3u + ((4u + ((uint32_t)ls & 3u)) << eb);
                    len = base + peek(L, eb);
                    drop(L, eb);
                }
                refill(L);
                e = T.root_d[L.acc & ((1u << ROOT_D) - 1u)];
                int ds = 0;
                if (e) {
                    drop(L, (int)(e & 15u));
                    ds = (int)(e >> 4);
                } else {
                    e = long_code<ROOT_D>(T.meta_d, L.acc);
                    if (!e) derr = L.nbits <= 0 ? IE_TRUNCATED : IE_BAD_DCODE;
                    else {
                        drop(L, (int)(e & 15u));
                        ds = T.long_d[(e >> 4) & 31u];
                    }
                }
                if (ds >= 30 && derr == IE_OK) derr = IE_BAD_DCODE;
                {   // Trees.cs:75-98
                    const int eb = ds < 4 ? 0 : (ds >> 1) - 1;
                    const uint32_t base = ds < 4 ? 1u + (uint32_t)ds : 1u + ((2u + ((uint32_t)ds & 1u)) << (eb & 15));
                    d = base + peek(L, eb & 15);
                    drop(L, eb & 15);
                }
            }
        }
        if (derr == IE_OK && L.nbits < 0 && sym != 256) derr = IE_TRUNCATED;   // END_BLOCK: block_step checks
    }
    // (c)
    if (pn) {
        uint64_t val;
        const int kind = L.pend_kind;
        if (kind == U_LIT) val = L.pend_val;
        else if (kind == U_MEM) {
            const uint32_t b = (L.v - L.cp_D) & 7u;
            val = w0 >> (8 * b);
            if (b + pn > 8) val |= w1 << (8 * (8 - b));
        } else if (kind == U_PAT) {
            const uint32_t ph = L.ph;
            val = ph ? (L.plo >> (8 * ph)) | (L.phi << (8 * (8 - ph))) : L.plo;
            uint32_t nx = ph + (8u % L.pd);   // the next unit starts 8 bytes further into the period
            if (nx >= L.pd) nx -= L.pd;
            L.ph = nx;
        } else val = 0;
        if (pn < 8) val &= (1ull << (8 * pn)) - 1ull;
        append_chunk(L, val, pn);
        L.cp_rem -= pn;
        // a run needs no loads: up to three more units of it in the same iteration (a 258-byte run in 9 steps, not 33)
        if (kind == U_PAT || kind == U_ZERO) {
#pragma unroll 1
            for (int x = 0; x < 3 && L.cp_rem >= 16; x++) {
                uint64_t v2 = 0;
                if (kind == U_PAT) {
                    const uint32_t ph = L.ph;
                    v2 = ph ? (L.plo >> (8 * ph)) | (L.phi << (8 * (8 - ph))) : L.plo;
                    uint32_t nx = ph + (8u % L.pd);
                    if (nx >= L.pd) nx -= L.pd;
                    L.ph = nx;
                }
                append_chunk(L, v2, 8);
                L.cp_rem -= 8;
            }
        }
    }
    // (d)
    if (!dec) {
        L.pend_n = L.cp_rem < 8 ? L.cp_rem : 8;
        return;
    }
    L.pend_n = 0;
    if (derr != IE_OK) return lane_fail(L, derr);
    if (sym < 256) {
        if (L.v + nlit > L.vcap) return lane_fail(L, IE_OUTPUT_OVERFLOW);
        L.pend_kind = U_LIT;
        L.pend_val = lits;
        L.pend_n = nlit;
        L.cp_rem = nlit;
        return;
    }
    if (sym == 256) {
        L.st = S_BLOCK;
        return;
    }
    if (L.v + len > L.vcap) return lane_fail(L, IE_OUTPUT_OVERFLOW);
    const uint32_t o = L.v - L.sh;
    if (d > o || o - d < L.carry) {
        // Q14 (Deflate.Rle.cs:47-50): an Rle segment may open with distance-1 runs of the previous segment's last byte.
        // Those bytes are placeholders here; inflate_carry_fixup_kernel fills them in once every segment is decoded.
        if (!(d == 1 && o == L.carry && L.carry_ok)) return lane_fail(L, IE_DIST_TOO_FAR);
        L.carry += len;
        L.pend_kind = U_ZERO;
    } else if (d <= 8) {
        pattern_setup(L, d);
        L.pend_kind = U_PAT;
    } else {
        L.cp_D = d;
        L.pend_kind = U_MEM;
    }
    L.cp_rem = len;
    L.pend_n = len < 8 ? len : 8;
}

// Start of a segment: a full-flush segment begins byte aligned with an empty history (Deflate.cs:596-605).
ZL_HD void begin_segment(Lane& L, const InflateParams& P, uint32_t seg)
{
    L.seg = seg;
    L.carry_ok = (seg > 0 || P.continuation) ? 1u : 0u;
    L.final_seen = 0;
    L.err = IE_OK;
    L.st = S_BLOCK;
    L.pos = 0;
    L.pre = 0;
    L.pre_ok = 0;
    L.acc = 0;
    L.nbits = 0;
    L.end = 0;
    L.limit = 0;
    L.wbuf = 0;
    L.wprev = 0;
    L.fixed = 0;
    L.nl = L.nd = L.ci = L.prevlen = 0;
    L.stored_rem = 0;
    L.cp_rem = L.pend_n = 0;
    L.pend_kind = U_LIT;
    L.pend_val = L.cp_D = L.ph = 0;
    L.pd = 1;
    L.plo = L.phi = 0;
    L.carry = 0;
    L.bl_lens = 0;
    const uint64_t c_begin = seg == 0 ? P.first_offset : P.seg_end[seg - 1];
    const uint64_t c_end = P.count_only ? P.count_body_end : P.seg_end[seg];
    L.count = P.count_only;
    L.found = 0;
    const uint64_t o_begin = P.count_only ? 0 : (uint64_t)seg * P.seg_size;
    uint64_t room = o_begin < P.out_cap ? P.out_cap - o_begin : 0;
    if (room > P.seg_size || P.count_only) room = P.seg_size;
    uint8_t* ob = P.out + o_begin;
    L.sh = (uint32_t)((uintptr_t)ob & 7u);
    L.ob8 = ob - L.sh;
    L.v = L.sh;
    L.vcap = L.sh + (uint32_t)room;
    L.in = P.in + (c_begin <= P.n ? c_begin : 0);
    if (c_end > P.n || c_begin > c_end || c_end - c_begin > 0xF0000000ull) {
        L.err = IE_TRUNCATED;
        L.st = S_FINISH;
        return;
    }
    L.end = (uint32_t)(c_end - c_begin);
    const uint64_t left = P.n - c_begin;
    L.limit = left > 0xFFFFFF00ull ? 0xFFFFFF00u : (uint32_t)left;
}

// End of a segment: write back what is still in the register, record length and status.
ZL_HD void finish_segment(Lane& L, const InflateParams& P)
{
    flush_partial(L);
    const uint32_t o = L.v - L.sh;
    // every segment but the last holds exactly seg_size bytes (that is what places the output)
    if (L.err == IE_OK && !L.count && L.seg + 1 < P.nseg && o != P.seg_size) L.err = IE_SEG_SIZE;
    P.seg_out_len[L.seg] = o;
    P.seg_status[L.seg] = L.err;
    if (P.seg_found_end) P.seg_found_end[L.seg] = L.found ? (uint64_t)(L.in - P.in) + L.found : 0;
    if (P.seg_carry) P.seg_carry[L.seg] = L.carry;
    if (L.carry && P.carry_flag) *P.carry_flag = 1;
}

// Q14 fix-up: the byte the carried prefix of segment k repeats = the last byte before it that is not itself carried.
ZL_HD bool carry_fill_value(const InflateParams& P, uint32_t k, uint8_t& val)
{
    for (uint32_t j = k; j > 0;) {
        j--;
        const uint32_t len = P.seg_out_len[j];
        if (len > P.seg_carry[j]) {
            val = P.out[(uint64_t)j * P.seg_size + len - 1];
            return true;
        }
    }
    // every earlier segment of this call is a carried run from end to end: the byte before the call (known to the caller
    // of a later shard, zsb_inflate_fill_carry), or nothing yet
    if (P.prev_byte >= 0) {
        val = (uint8_t)P.prev_byte;
        return true;
    }
    return false;
}

#ifndef ZSB_LN_SYMS_PER_STEP
#define ZSB_LN_SYMS_PER_STEP 4
#endif
constexpr int SYMS_PER_STEP = ZSB_LN_SYMS_PER_STEP;

// One bounded unit of work in whatever state the lane is in (S_BUILD_* and S_FINISH are the caller's).
ZL_HD void lane_step(Lane& L, const LaneTables& T, uint8_t* glens, const uint16_t* glong)
{
    if (L.st == S_SYM) {
#pragma unroll 1
        for (int r = 0; r < SYMS_PER_STEP && L.st == S_SYM; r++) sym_step(L, T, glong);
    } else if (L.st == S_CLEN) {
#pragma unroll 1
        for (int r = 0; r < SYMS_PER_STEP && L.st == S_CLEN; r++) clen_step(L, T, glens);
    } else if (L.st == S_BLOCK) block_step(L);
    else if (L.st == S_STORED) stored_step(L);
}

// ---- cooperative table build: four phases, every helper h = 0..31 runs each phase, a warp barrier between phases -------
// mode BL: the 19-symbol bit-length code (lengths in `bl`, 3 bits each) -> 7-bit table in T.root_l.
// mode LD: literal/length code of nl symbols (9 consecutive symbols per helper) and distance code of nd symbols (one per
//          helper), lengths from glens (or the fixed code).
ZL_HD int fixed_llen(int n) { return n <= 143 ? 8 : n <= 255 ? 9 : n <= 279 ? 7 : 8; }   // InfTree.cs:16-96

ZL_HD void build_p0(int h, WarpScratch& sc, LaneTables& T, HelperRegs& R, bool ld, uint64_t bl, const uint8_t* glens, int nl,
                    int nd, int fixed)
{
    if (ld) {
        uint32_t* zl = reinterpret_cast<uint32_t*>(T.root_l) + h * ((1 << ROOT_L) / 64);
        for (int i = 0; i < (1 << ROOT_L) / 64; i++) zl[i] = 0;
        reinterpret_cast<uint32_t*>(T.root_d)[h] = 0;   // 64 entries = 32 words
        unsigned long long c = 0;
        for (int k = 0; k < 9; k++) {
            const int s = h * 9 + k;
            int l = 0;
            if (s < nl) l = fixed ? fixed_llen(s) : ld_scratch_u8(glens + s);
            R.ll[k] = (uint8_t)l;
            if (l) c += 1ull << (4 * l);
        }
        sc.loc[h] = c;
        int dl = 0;
        if (h < nd) dl = fixed ? 5 : ld_scratch_u8(glens + nl + h);
        R.dl = (uint8_t)dl;
        sc.dlen[h] = (uint8_t)dl;
    } else {
        uint32_t* zl = reinterpret_cast<uint32_t*>(T.root_l) + h * 2;   // 128 entries = 64 words
        zl[0] = 0;
        zl[1] = 0;
        const int dl = h < 19 ? (int)((bl >> (3 * h)) & 7u) : 0;
        R.dl = (uint8_t)dl;
        sc.dlen[h] = (uint8_t)dl;
    }
}

ZL_HD void build_p1(int h, WarpScratch& sc, bool ld)
{
    if (h < 16) {
        if (ld && h) {
            uint32_t run = 0;
            for (int j = 0; j < 32; j++) {
                sc.pre[h][j] = (uint16_t)run;
                run += (uint32_t)(sc.loc[j] >> (4 * h)) & 15u;
            }
            sc.cnt_l[h] = (uint16_t)run;
        }
    } else {
        const int l = h - 16;
        uint32_t run = 0;
        for (int j = 0; j < 32; j++) {
            sc.pre_d[l][j] = (uint8_t)run;
            run += sc.dlen[j] == l ? 1u : 0u;
        }
        sc.cnt_d[l] = (uint16_t)run;
    }
}

// First codes, long-list offsets and the packed limits of one alphabet.  0 ok, 1 no codes, -1 oversubscribed,
// -2 incomplete (InfTree.cs huft_build rules: a single one-bit code is the only incomplete set accepted).
ZL_HD int build_meta(const uint16_t* cnt, int root, uint16_t* first, uint16_t* off, uint32_t* meta)
{
    int left = 1, maxlen = 0, total = 0;
    uint32_t code = 0, longs = 0;
    int rc = 0;
    for (int l = 1; l <= 15; l++) {
        const int c = cnt[l];
        code = (code + (l > 1 ? cnt[l - 1] : 0)) << 1;
        first[l] = (uint16_t)code;
        left = (left << 1) - c;
        if (left < 0 && rc == 0) rc = -1;
        if (c) maxlen = l;
        total += c;
        if (l >= root) {
            const uint32_t lim = ((code + (uint32_t)c) << (15 - l)) & 0xffffu;
            off[l] = (uint16_t)longs;
            if (meta) meta[l - root] = lim | (longs << 16);
            if (l > root) longs += (uint32_t)c;
        }
    }
    if (total == 0) return 1;
    if (rc == 0 && left > 0 && maxlen != 1) rc = -2;
    return rc;
}

ZL_HD void build_p2(int h, WarpScratch& sc, LaneTables& T, bool ld)
{
    if (h == 0 && ld) sc.rc_l = build_meta(sc.cnt_l, ROOT_L, sc.first_l, sc.off_l, T.meta_l);
    if (h == 1) sc.rc_d = ld ? build_meta(sc.cnt_d, ROOT_D, sc.first_d, sc.off_d, T.meta_d)
                             : build_meta(sc.cnt_d, ROOT_BL, sc.first_d, sc.off_d, nullptr);
}

ZL_HD void fill_root(uint16_t* table, int root, uint32_t code, int l, int sym)
{
    const uint32_t r = brev32(code) >> (32 - l);
    const uint16_t e = (uint16_t)((sym << 4) | l);
    for (uint32_t k = r; k < (1u << root); k += 1u << l) table[k] = e;
}

ZL_HD void build_p3(int h, const WarpScratch& sc, LaneTables& T, const HelperRegs& R, bool ld, uint16_t* glong, int fixed)
{
    if (ld) {
        const bool l_ok = sc.rc_l == 0 || fixed, d_ok = sc.rc_d == 0 || sc.rc_d == -2 || fixed;
        if (l_ok) {
            unsigned long long seen = 0;
            for (int k = 0; k < 9; k++) {
                const int l = R.ll[k];
                if (!l) continue;
                const uint32_t rank = sc.pre[l][h] + ((uint32_t)(seen >> (4 * l)) & 15u);
                seen += 1ull << (4 * l);
                const int s = h * 9 + k;
                if (l <= ROOT_L) fill_root(T.root_l, ROOT_L, sc.first_l[l] + rank, l, s);
                else {
                    const uint32_t idx = sc.off_l[l] + rank;
                    if (idx < (uint32_t)LONG_L) T.long_l[idx] = (uint16_t)s;
                    if (idx < (uint32_t)GLONG_STRIDE) glong[idx] = (uint16_t)s;
                }
            }
        }
        const int l = R.dl;
        if (d_ok && l) {
            const uint32_t rank = sc.pre_d[l][h];
            if (l <= ROOT_D) fill_root(T.root_d, ROOT_D, sc.first_d[l] + rank, l, h);
            else T.long_d[(sc.off_d[l] + rank) & 31u] = (uint8_t)h;
        }
    } else {
        const int l = R.dl;
        if (sc.rc_d == 0 && l) fill_root(T.root_l, ROOT_BL, sc.first_d[l] + sc.pre_d[l][h], l, h);
    }
}

// What the lane that asked for the build does with the outcome (InflateBlocks.cs:440-470, 571-626; InfTree.cs:377-427)
ZL_HD void after_build(Lane& L, bool ld, int rc_l, int rc_d)
{
    if (!ld) {
        if (rc_d == -1) return lane_fail(L, IE_OVERSUB_BL);
        if (rc_d != 0) return lane_fail(L, IE_INCOMPLETE_BL);
        L.ci = 0;
        L.prevlen = 0;
        L.st = S_CLEN;
        return;
    }
    if (!L.fixed) {
        if (rc_l == -1) return lane_fail(L, IE_OVERSUB_L);
        if (rc_l != 0) return lane_fail(L, IE_INCOMPLETE_L);
        if (rc_d == -1) return lane_fail(L, IE_OVERSUB_D);
        if (rc_d == -2) return lane_fail(L, IE_INCOMPLETE_D);
        if (rc_d == 1 && L.nl > 257) return lane_fail(L, IE_EMPTY_D);
    }
    L.st = S_SYM;
}

}  // namespace ZSB_LN_NS
}  // namespace zsb
