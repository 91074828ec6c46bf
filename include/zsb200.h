/*
 * zsb200.h -- C ABI of libzsb200: B200-native (sm_100a) segmented DEFLATE / INFLATE / Adler-32
 * behind the SixLabors.ZlibStream API.
 *
 * The reference has no FFI seam of its own; the natural one is its internal z_stream class
 * `ZLibStream` (src/ZlibStream/ZlibStream.cs), driven by `ZlibOutputStream.WriteCore/Finish`
 * (ZlibOutputStream.cs:125-168,213-256) and `ZlibInputStream.ReadCore/Finish`
 * (ZlibInputStream.cs:133-186,230-266).  Every entry point below names the reference member it
 * replaces.  A C# host binds these with [DllImport("zsb200")]; the Python host
 * mirror in zlibstream_b200/ binds them with ctypes.
 *
 * Conventions:
 *   - return values are the reference's CompressionState codes (CompressionState.cs:9-55);
 *     anything other than ZSB_OK / ZSB_STREAM_END is an error and zsb_last_error() holds the text
 *     the reference would put in ZLibStream.Message ("incorrect data check", ...);
 *   - the caller owns input and output memory (the reference pins them for the duration of the call,
 *     ZlibOutputStream.cs:132-134); the context owns all device scratch;
 *   - a context is not thread-safe; distinct contexts are independent;
 *   - pointers are host pointers unless the matching ZSB_*_DEVICE flag is given;
 *   - there is NO CPU fallback: every call needs a CUDA device and fails with ZSB_ERRNO otherwise.
 */
#ifndef ZSB200_H
#define ZSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define ZSB_API __declspec(dllexport)
#else
#define ZSB_API __attribute__((visibility("default")))
#endif

/* CompressionState.cs:9-55 */
#define ZSB_VERSION_ERROR (-6)
#define ZSB_BUF_ERROR (-5)
#define ZSB_MEM_ERROR (-4)
#define ZSB_DATA_ERROR (-3)
#define ZSB_STREAM_ERROR (-2)
#define ZSB_ERRNO (-1)
#define ZSB_OK 0
#define ZSB_STREAM_END 1
#define ZSB_NEED_DICT 2

/* CompressionStrategy.cs:9-37 */
#define ZSB_DEFAULT_STRATEGY 0
#define ZSB_FILTERED 1
#define ZSB_HUFFMAN_ONLY 2
#define ZSB_RLE 3
#define ZSB_FIXED 4

/* Deflate.Intrinsics.cs:295-307: which UpdateHash the reference build would use */
#define ZSB_HASH_CRC32C 0 /* >= netcoreapp3.1 with SSE4.2 (the published numbers) */
#define ZSB_HASH_MUL 1    /* netstandard / net472 / non-x86 */

/* flags */
#define ZSB_IN_DEVICE 0x1u      /* `in` is a device pointer */
#define ZSB_OUT_DEVICE 0x2u     /* `out` is a device pointer */
#define ZSB_NO_HEADER 0x4u      /* deflate: do not emit / inflate: do not expect the 2-byte zlib header   */
#define ZSB_NO_TRAILER 0x8u     /* deflate: do not emit / inflate: do not expect `03 00` + Adler-32        */
#define ZSB_OFFSETS_DEVICE 0x10u /* seg_end_offsets is a device pointer */
#define ZSB_SINGLE_STREAM 0x40u  /* the stream is ONE piece: written with FlushMode.NoFlush (one Write + Dispose), final block + Adler-32 */
#define ZSB_SYNC_FLUSH 0x80u     /* deflate: every segment ends with FlushMode.SyncFlush, not FullFlush (the hash table is kept:
                                    one CTA compresses the segments in order; the stream inflates as ONE piece)               */
#define ZSB_PARTIAL_FLUSH 0x100u /* deflate: every segment ends with FlushMode.PartialFlush (Tr_align; like ZSB_SYNC_FLUSH, and the
                                    segments are not byte aligned: seg_end_offsets count the whole bytes written so far)      */
#define ZSB_CONTINUATION 0x20u   /* inflate: `in` is a later shard of a stream (its first segment may open with a Q14 run) */

typedef struct zsb_ctx zsb_ctx;

/* Replaces `new ZLibStream(options)` (ZlibStream.cs:18-29) + the pooled buffers of
 * Deflate.Buffers.cs:91-198.  device = CUDA ordinal. */
ZSB_API int zsb_ctx_create(int device, zsb_ctx** ctx);
/* Replaces ZLibStream.Dispose (ZlibStream.cs:225-245). */
ZSB_API void zsb_ctx_destroy(zsb_ctx* ctx);
/* Replaces ZLibStream.Message (ZlibStream.cs:64-67). Never NULL. */
ZSB_API const char* zsb_last_error(const zsb_ctx* ctx);
/* Work is enqueued on this cudaStream_t (NULL = the context's own stream). */
ZSB_API int zsb_ctx_set_stream(zsb_ctx* ctx, void* cuda_stream);
ZSB_API void* zsb_ctx_get_stream(const zsb_ctx* ctx);
/* Number of kernel launches issued by this context so far (bench.py's gpu_launches). */
ZSB_API uint64_t zsb_ctx_launch_count(const zsb_ctx* ctx);
/* Device time (ms, CUDA events on the context's stream) of the LZ77 / Huffman / inflate kernels of
 * the last call; any pointer may be NULL. */
ZSB_API int zsb_ctx_last_kernel_ms(const zsb_ctx* ctx, float* lz77_ms, float* huffman_ms, float* inflate_ms);
/* The same with the two level-1 LZ77 kernels apart: out[0..4] = lz77 stage, huffman stage, inflate, lz77_lists_kernel,
 * lz77_resolve_kernel (the last two are 0 when the call did not take the two-kernel path). */
ZSB_API int zsb_ctx_last_kernel_ms_detail(const zsb_ctx* ctx, float* out, int n);

/* Upper bound of what zsb_deflate_segments can write for n input bytes. */
ZSB_API size_t zsb_deflate_bound(size_t n, size_t seg_size);

/*
 * Replaces the loop `foreach segment: ZlibOutputStream.Write(segment)` with
 * Options.FlushMode = FullFlush (ZlibOutputStream.cs:125-168 -> ZLibStream.Deflate ->
 * Deflate.Compress, Deflate.cs:436-636) followed, unless ZSB_NO_TRAILER, by Dispose/Finish
 * (ZlibOutputStream.cs:213-256).  Output is bit-identical to the reference's for the same
 * level / strategy / hash variant / segment boundaries.
 *
 *   in, n               input bytes; every segment is seg_size bytes except possibly the last
 *   seg_size            multiple of 32768 and >= 65536 (the window refill schedule of Deflate.cs:967-1019
 *                       is only segment-independent for those)
 *   level               -1..9 (CompressionLevel.cs); strategy 0..4; hash_variant ZSB_HASH_*
 *   first_segment_index index of in's first segment inside the whole stream: only stream segment 0 starts
 *                       at window index 0 and has the header counted in the 512-byte chunk arithmetic (Q6, Q10)
 *   halo                with first_segment_index > 0: the 32768 bytes that precede `in` in the stream (same
 *                       memory space as `in`), or NULL; only Rle (Q14) and a short last segment read it
 *   out, out_cap        receives [header] segments... [03 00 adler32]; a DEVICE out pointer must be 4-byte aligned
 *                       (ZSB_STREAM_ERROR otherwise); input pointers and inflate's pointers may have any alignment
 *   out_len             bytes written
 *   seg_end_offsets     optional, one per segment: offset in `out` of the end of the segment
 *                       (== ZlibOutputStream.TotalOut after that Write, minus the bytes this call did not emit)
 *   adler               Adler-32 (seed 1) of `in`; combine across shards with zsb_adler32_combine
 */
ZSB_API int zsb_deflate_segments(zsb_ctx* ctx, const void* in, size_t n, size_t seg_size, int level, int strategy,
                                 int hash_variant, uint64_t first_segment_index, const void* halo, unsigned flags,
                                 void* out, size_t out_cap, size_t* out_len, uint64_t* seg_end_offsets,
                                 uint32_t* adler);

/*
 * Replaces the loop over ZlibInputStream.Read (ZlibInputStream.cs:133-186 -> ZLibStream.Inflate ->
 * Inflate.Decompress, Inflate.cs:103-357) for a full-flush-segmented stream.
 *
 *   seg_end_offsets     nseg offsets into `in` of each segment's end (as produced by zsb_deflate_segments);
 *                       NULL = build the index on the device by scanning for the full-flush marker
 *   seg_size            uncompressed size of every segment but the last (needed to place output)
 *   out_len             bytes produced; adler = Adler-32 of the output, checked against the trailer unless
 *                       ZSB_NO_TRAILER ("incorrect data check")
 */
ZSB_API int zsb_inflate_segments(zsb_ctx* ctx, const void* in, size_t n, const uint64_t* seg_end_offsets,
                                 size_t nseg, size_t seg_size, unsigned flags, void* out, size_t out_cap,
                                 size_t* out_len, uint32_t* adler);

/*
 * Q14 (Deflate.Rle.cs:47-50): an Rle segment may open with a distance-1 run of the byte BEFORE it.  Inside one call the
 * library fills such runs from the previous segment; for the first segment of a call (a later shard of a stream) that byte
 * is not in `in`: this returns how many leading output bytes of the last zsb_inflate_segments call were left unset, for the
 * caller to fill with the last byte of the previous shard.
 *
 * zsb_inflate_fill_carry fills those bytes (and those of following segments that are one carried run from end to end) in the
 * DEVICE output buffer of the last zsb_inflate_segments call with prev_byte and returns the Adler-32 of the completed output.
 * A first segment may only open with such a run when the call was made with ZSB_CONTINUATION ("invalid distance too far back"
 * otherwise, as at the start of a stream).
 */
ZSB_API int zsb_ctx_last_inflate_carry(const zsb_ctx* ctx, uint32_t* first_segment_carry);
ZSB_API int zsb_inflate_fill_carry(zsb_ctx* ctx, uint8_t prev_byte, uint32_t* adler);

/*
 * Segment index of a full-flush-segmented stream that comes without a side-car, e.g. one written by the reference itself
 * (SURVEY.md section 8f-1).  The reference's own resynchronisation primitive is InflateSync (Inflate.cs:386-448), which
 * looks for the same 00 00 FF FF marker one flush point at a time.  Candidates are validated by decoding (chance
 * occurrences of the pattern inside stored or Huffman data are merged away) and the repeated markers of Q10 are kept with
 * the segment they follow.
 *
 *   seg_end_offsets, cap   receives the end offset of every segment (may be NULL to query the count only)
 *   nseg                   number of segments found
 *   uncompressed_size      total decoded size (every segment but the last is seg_size bytes)
 *   flags                  ZSB_IN_DEVICE, ZSB_OFFSETS_DEVICE, ZSB_NO_HEADER, ZSB_NO_TRAILER
 */
ZSB_API int zsb_build_segment_index(zsb_ctx* ctx, const void* in, size_t n, size_t seg_size, unsigned flags,
                                    uint64_t* seg_end_offsets, size_t cap, size_t* nseg, uint64_t* uncompressed_size);

/*
 * PNG front end: the caller the reference is written for is ImageSharp's PNG encoder
 * (/root/reference/readme.md:22), which filters scanlines on the CPU and streams them through ZlibOutputStream.  Here the
 * pixels are filtered on the device (PNG specification section 9; filter 5 = the minimum-sum-of-absolute-differences choice per
 * scanline of section 12.8), compressed as full-flush segments by zsb_deflate_segments (so the IDAT payload is bit-identical to
 * what the reference writes for the same filtered bytes), check-summed (CRC-32 of every IDAT chunk) and framed: the result is a
 * complete PNG file (signature, IHDR, IDAT..., IEND).  8 bits per sample; color_type 0 gray, 2 RGB, 4 gray + alpha, 6 RGBA.
 *
 *   seg_size     full-flush segment size of the IDAT stream (0 = 65536)
 *   idat_bytes   zlib bytes per IDAT chunk, a multiple of 65536 (0 = one chunk if the stream fits, else 1 GiB chunks)
 *   flags        ZSB_IN_DEVICE / ZSB_OUT_DEVICE
 */
ZSB_API size_t zsb_png_bound(uint32_t width, uint32_t height, int color_type, size_t seg_size, size_t idat_bytes);
ZSB_API int zsb_png_encode(zsb_ctx* ctx, const void* pixels, uint32_t width, uint32_t height, int color_type, int filter,
                           int level, int strategy, size_t seg_size, size_t idat_bytes, unsigned flags, void* out,
                           size_t out_cap, size_t* out_len);

/* Replaces Adler32.Calculate(adler, span) (Adler32.cs:51-78). flags: ZSB_IN_DEVICE. */
ZSB_API int zsb_adler32(zsb_ctx* ctx, uint32_t seed, const void* p, size_t n, unsigned flags, uint32_t* result);
/* adler32_combine; absent from the reference. Pure host function. */
ZSB_API uint32_t zsb_adler32_combine(uint32_t adler1, uint32_t adler2, uint64_t len2);

/* Library / build identification: "zsb200 <version> sm_100a". */
ZSB_API const char* zsb_version(void);

/* Diagnostics for tests: raw copy of an intermediate device buffer of the last zsb_deflate_segments batch
 * (0 symbols, 1 block records, 2 block counts, 3 segment layouts, 4 code tables, 5 per-segment Adler sums). */
ZSB_API int zsb_debug_fetch(zsb_ctx* ctx, int what, uint64_t byte_offset, void* dst, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* ZSB200_H */
