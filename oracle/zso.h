/*
 * zso.h -- CPU ORACLE for the SixLabors.ZlibStream DEFLATE hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may build,
 * link or call anything in oracle/.  The product (libzsb200.so, zlibstream_b200/) never does.
 *
 * It is a plain-C restatement (not a copy) of the reference's managed algorithm:
 *   Deflate.cs, Deflate.Fast.cs, Deflate.Slow.cs, Deflate.Rle.cs, Deflate.Stored.cs,
 *   Deflate.Intrinsics.cs:295-307 (hash), Trees.cs, Trees.Static.cs,
 *   ZlibStream.cs:197-222 (ReadBuffer), ZlibOutputStream.cs:125-168,213-256 (512-byte loop),
 *   Adler32.cs:270-326.
 * Inflate follows RFC 1950/1951 (the reference's own inflater mis-builds dynamic
 * tables, InfTree.cs:312,341) with the reference's error strings (Inflate.cs,
 * InflateBlocks.cs, InfTree.cs, InfCodes.cs).
 *
 * Parity pin: the 36 compressed sizes published in the reference's benchmarks.md
 * (tests/test_oracle_golden.py) + round trips through system zlib.
 */
#ifndef ZSO_H
#define ZSO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* CompressionState.cs:9-55 */
enum {
    ZSO_VERSION_ERROR = -6,
    ZSO_BUF_ERROR = -5,
    ZSO_MEM_ERROR = -4,
    ZSO_DATA_ERROR = -3,
    ZSO_STREAM_ERROR = -2,
    ZSO_ERRNO = -1,
    ZSO_OK = 0,
    ZSO_STREAM_END = 1,
    ZSO_NEED_DICT = 2
};

/* FlushMode.cs:9-35 */
enum { ZSO_NO_FLUSH = 0, ZSO_PARTIAL_FLUSH = 1, ZSO_SYNC_FLUSH = 2, ZSO_FULL_FLUSH = 3, ZSO_FINISH = 4 };

/* CompressionStrategy.cs:9-37 */
enum { ZSO_DEFAULT_STRATEGY = 0, ZSO_FILTERED = 1, ZSO_HUFFMAN_ONLY = 2, ZSO_RLE = 3, ZSO_FIXED = 4 };

/* Deflate.Intrinsics.cs:295-307: SSE4.2 CRC32C (>= netcoreapp3.1 on x86) or multiplicative */
enum { ZSO_HASH_CRC32C = 0, ZSO_HASH_MUL = 1 };

/* ---- ZlibOutputStream-equivalent compressor -------------------------------------- */

typedef struct zso_deflater zso_deflater;

/* Optional trace of what the compressor emitted; used by tests to diff GPU stages. */
typedef struct zso_block_trace {
    int type;            /* 0 stored, 1 static, 2 dynamic */
    int last;            /* BFINAL */
    int n_symbols;       /* lastLit at flush */
    int stored_len;      /* input bytes covered */
    uint64_t bit_start;  /* absolute bit offset in the output stream where the 3 header bits go */
} zso_block_trace;

typedef void (*zso_block_cb)(void* user, const zso_block_trace* blk,
                             const uint16_t* dist, const uint8_t* lc /* n_symbols each */);

/* level -1..9 (CompressionLevel.cs), strategy 0..4, flush_mode = Options.FlushMode used on every Write. */
zso_deflater* zso_deflater_new(int level, int strategy, int flush_mode, int hash_variant);
void zso_deflater_free(zso_deflater* d);
void zso_deflater_set_trace(zso_deflater* d, zso_block_cb cb, void* user);

/* == ZlibOutputStream.Write(buf, 0, n) (ZlibOutputStream.cs:125-168). Returns ZSO_OK or an error code. */
int zso_deflater_write(zso_deflater* d, const uint8_t* buf, size_t n);
/* == ZlibOutputStream.Dispose -> Finish (ZlibOutputStream.cs:213-256). */
int zso_deflater_finish(zso_deflater* d);

const uint8_t* zso_deflater_output(const zso_deflater* d, size_t* len);
uint64_t zso_deflater_total_in(const zso_deflater* d);
uint64_t zso_deflater_total_out(const zso_deflater* d);
uint32_t zso_deflater_adler(const zso_deflater* d);
const char* zso_deflater_message(const zso_deflater* d);

/* Work counters */
typedef struct zso_counters {
    uint64_t inserts, searches, chain_steps, compare_bytes, literals, matches, match_bytes, blocks;
} zso_counters;
void zso_deflater_counters(const zso_deflater* d, zso_counters* out);

/*
 * One stream, one Write per segment with FlushMode.FullFlush, then Finish.
 * seg_end[k] = TotalOut after Write k (exact compressed offset of the end of segment k,
 * header included).  out may be NULL to only measure.  Returns ZSO_OK, or ZSO_BUF_ERROR if out_cap
 * is too small (out_len is still set to the needed size).
 */
int zso_deflate_segments(const uint8_t* in, size_t n, size_t seg_size, int level, int strategy,
                         int hash_variant, uint8_t* out, size_t out_cap, size_t* out_len,
                         uint64_t* seg_end /* ceil(n/seg_size) entries or NULL */, uint32_t* adler);

/* Single stream, single Write, NoFlush, Finish: the reference's benchmark call pattern. */
int zso_deflate_oneshot(const uint8_t* in, size_t n, int level, int strategy, int hash_variant,
                        uint8_t* out, size_t out_cap, size_t* out_len, uint32_t* adler);

/* Worst-case output bound for zso_deflate_segments / oneshot. */
size_t zso_deflate_bound(size_t n, size_t seg_size);

/* ---- Inflate (RFC 1950/1951) ------------------------------------------------------ */

/* Decodes one zlib stream. msg (optional) receives a static string using the reference's wording. */
int zso_inflate(const uint8_t* in, size_t n, uint8_t* out, size_t out_cap, size_t* out_len,
                size_t* in_used, uint32_t* adler, const char** msg);

/* Decodes raw deflate blocks from `in` until a BFINAL block or until `stop_at_out` bytes have been
 * produced and the decoder sits at a block boundary.  hist/hist_len is an optional preset window. */
int zso_inflate_raw(const uint8_t* in, size_t n, uint8_t* out, size_t out_cap, size_t* out_len,
                    size_t* in_used, const char** msg);

/* ---- Adler-32 ----------------------------------------------------------------------- */

uint32_t zso_adler32(uint32_t seed, const uint8_t* p, size_t n);                 /* Adler32.cs:270-326 */
uint32_t zso_adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2);           /* zlib formula   */

/* ---- timing helpers for bench.py's CPU legs (zso_bench.c): native pinned threads, no Python in the timed region -------- */
int zso_bench_deflate_mt(const uint8_t* in, size_t n, size_t seg_size, int level, int strategy, int hash_variant, int threads,
                         int reps, double* seconds, uint64_t* out_bytes);
int zso_bench_inflate_mt(const uint8_t* comp, const uint64_t* seg_ends, size_t nseg, size_t first_offset, size_t seg_size,
                         int which /* 1 system zlib, 2 oracle */, int threads, int reps, double* seconds, uint64_t* decoded);

/* ---- hash (exposed for table generation checks) ------------------------------------ */
uint32_t zso_crc32c_u32(uint32_t crc, uint32_t v);                               /* Sse42.Crc32  */

#ifdef __cplusplus
}
#endif
#endif /* ZSO_H */
