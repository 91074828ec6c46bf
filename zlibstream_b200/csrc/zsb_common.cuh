// zsb_common.cuh -- shared constants, records and device helpers of libzsb200 (sm_100a only).
//
// Vocabulary follows the reference (src/ZlibStream/Deflate.cs, Trees.cs): window, head, prev,
// strStart, lookahead, lit/dist symbols, blocks of at most 16383 symbols, full-flush segments.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace zsb {

// Deflate.cs:17-78, Trees.cs:15-70
constexpr int MIN_MATCH = 3;
constexpr int MAX_MATCH = 258;
constexpr int MIN_LOOKAHEAD = MAX_MATCH + MIN_MATCH + 1;  // 262
constexpr int W_SIZE = 32768;                             // windowBits 15 (ZlibStream.cs:14-15)
constexpr int W_MASK = W_SIZE - 1;
constexpr int WINDOW_SIZE = 2 * W_SIZE;
constexpr int MAX_DIST = W_SIZE - MIN_LOOKAHEAD;          // 32506
constexpr int HASH_SIZE = 32768;                          // memLevel 8 -> hashBits 15 (Deflate.cs:294-296)
constexpr int HASH_MASK = HASH_SIZE - 1;
constexpr int LIT_BUFSIZE = 16384;                        // Deflate.cs:298
constexpr int BLOCK_SYMS = LIT_BUFSIZE - 1;               // a block is flushed at exactly 16383 symbols (Q9)
constexpr int MAX_BITS = 15;
constexpr int MAX_BL_BITS = 7;
constexpr int LITERALS = 256;
constexpr int END_BLOCK = 256;
constexpr int LENGTH_CODES = 29;
constexpr int L_CODES = LITERALS + 1 + LENGTH_CODES;  // 286
constexpr int D_CODES = 30;
constexpr int BL_CODES = 19;
constexpr int HEAP_SIZE = 2 * L_CODES + 1;  // 573
constexpr int CHUNK_OUT = 512;              // ZlibOutputStream.cs:53

enum { FN_STORED = 0, FN_FAST = 1, FN_SLOW = 2, FN_RLE = 3 };
enum { BT_STORED = 0, BT_STATIC = 1, BT_DYNAMIC = 2 };

// One deflate block as the parser leaves it and the Huffman stage completes it.
struct BlockRec {
    uint32_t sym_begin;   // index of the block's first symbol inside the segment's symbol array
    uint32_t sym_count;   // lastLit at flush (<= 16383)
    uint32_t in_begin;    // segment-relative offset of the first input byte the block covers
    uint32_t in_len;      // strStart - blockStart at flush ("stored_len")
    uint32_t buf_valid;   // blockStart >= 0 in window coordinates (Trees.cs:608 `buf != -1`)
    // filled by the Huffman stage
    uint32_t type;        // BT_*
    uint32_t hdr_bits;    // dynamic blocks: bits of the tree description after the 3 type bits
    uint32_t body_bits;   // code + extra bits of all symbols plus the END_BLOCK code
    uint64_t bit_offset;  // where the 3 type bits go, relative to the segment's first output bit
};

struct LevelCfg {
    int good, lazy, nice, chain, fn;
};

// ---- small device helpers ---------------------------------------------------------------------

__device__ __forceinline__ uint32_t lds_u32_unaligned(const uint8_t* base, int idx)
{
    // 4-byte little-endian load at any byte index of a shared array (array must be padded by 8 bytes).
    const uint32_t* p = reinterpret_cast<const uint32_t*>(base + (idx & ~3));
    uint32_t lo = p[0], hi = p[1];
    return __funnelshift_r(lo, hi, (idx & 3) * 8);
}

__device__ __forceinline__ int length_code_of(int lc)  // Trees.cs:166-183 (lc = length - 3)
{
    if (lc < 8) return lc;
    if (lc == 255) return 28;
    int n = 31 - __clz(lc);
    return 4 * n - 4 + ((lc >> (n - 2)) & 3);
}
__device__ __forceinline__ int length_extra_bits(int code)  // Trees.cs:85-89
{
    return (code < 8 || code == 28) ? 0 : (code - 4) >> 2;
}
__device__ __forceinline__ int length_base(int code)  // Trees.cs:111-115
{
    if (code < 8) return code;
    if (code == 28) return 0;  // the reference's table holds 0 for code 28 (no extra bits are sent)
    int e = (code - 4) >> 2;
    return ((4 + (code & 3)) << e);
}
__device__ __forceinline__ int dist_code_of(int d)  // Trees.cs:131-161,193-201 (d = distance - 1)
{
    if (d < 4) return d;
    int n = 31 - __clz(d);
    return 2 * n + ((d >> (n - 1)) & 1);
}
__device__ __forceinline__ int dist_extra_bits(int code) { return code < 4 ? 0 : (code >> 1) - 1; }  // Trees.cs:94-98
__device__ __forceinline__ int dist_base(int code)                                                    // Trees.cs:75-80
{
    if (code < 4) return code;
    int e = (code >> 1) - 1;
    return (2 + (code & 1)) << e;
}
__device__ __forceinline__ int static_llen(int n) { return n <= 143 ? 8 : n <= 255 ? 9 : n <= 279 ? 7 : 8; }  // Trees.Static.cs:57-82

__device__ __forceinline__ uint32_t bit_reverse(uint32_t code, int len) { return __brev(code) >> (32 - len); }  // Trees.cs:269-281

__device__ __forceinline__ uint32_t static_lcode(int n)  // canonical code of the fixed literal/length tree, bit-reversed
{
    uint32_t c;
    int len;
    if (n <= 143) { c = 0x30 + n; len = 8; }
    else if (n <= 255) { c = 0x190 + (n - 144); len = 9; }
    else if (n <= 279) { c = n - 256; len = 7; }
    else { c = 0xC0 + (n - 280); len = 8; }
    return bit_reverse(c, len);
}

constexpr uint32_t ADLER_BASE = 65521u;  // Adler32.cs:27

}  // namespace zsb
