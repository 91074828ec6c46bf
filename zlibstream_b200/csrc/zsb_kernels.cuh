// zsb_kernels.cuh -- kernel parameter blocks and host-side launchers shared by the .cu files.
#pragma once

#include "zsb_common.cuh"

namespace zsb {

// ---- LZ77 (zsb_lz77.cu) ------------------------------------------------------------------------
struct Lz77Params {
    const uint8_t* in;        // first byte of the first segment handled by this launch
    size_t n;                 // bytes from `in` to the end of the input
    uint32_t seg_size;
    uint32_t nseg;
    uint64_t first_seg_index; // stream index of segment 0 of this launch
    const uint8_t* halo;      // the 32768 bytes preceding `in` in the stream, or nullptr
    int strategy, hash_variant;
    int good, lazy, nice, chain, fn;
    uint32_t* sym;            // [nseg][sym_stride]  dist | lc << 16
    size_t sym_stride;
    BlockRec* blocks;         // [nseg][max_blocks]
    uint32_t max_blocks;
    uint32_t* nblocks;        // [nseg]
    uint2* seg_adler;         // [nseg] (s1, s2) of each segment, seed 1
    unsigned int* work_counter;
    const uint16_t* crc_tab;  // 4 * 256
    unsigned long long* stats; // optional diagnostics counters (ZSB_LZ_STATS=1), else nullptr
    // level 1, two-kernel path (zsb_lz77_lists.cu + zsb_lz77_resolve.cu): 32-byte candidate records, [nseg][seg_size] x 2 uint4;
    // nullptr = not available (the batch parser of zsb_lz77.cu runs instead).  work_counter then points to three counters.
    uint4* lists = nullptr;
    cudaEvent_t ev_mid = nullptr;   // optional: recorded between the two kernels
    int chained = 0;                // FlushMode.SyncFlush: segments share the hash table and the window (one CTA, in order)
};
constexpr int LZ_LIST_K = 10;          // candidates per record
bool lz77_lists_applies(const Lz77Params& P);
size_t lz77_lists_bytes_per_segment(uint32_t seg_size);
cudaError_t launch_lz77_lists(const Lz77Params& P, int num_sms, cudaStream_t stream);     // kernel A: records + Adler sums
cudaError_t launch_lz77_resolve(const Lz77Params& P, int num_sms, cudaStream_t stream);   // kernel B: symbols + block records
size_t lz77_smem_bytes();
cudaError_t launch_lz77(const Lz77Params& P, int num_sms, cudaStream_t stream);
// HuffmanOnly and Rle need no hash chains: zsb_symbolize.cu produces their symbols, blocks and Adler sums in parallel.
bool symbolize_applies(const Lz77Params& P);
cudaError_t launch_symbolize(const Lz77Params& P, int num_sms, cudaStream_t stream);
// Level 0 (DeflateStored, Deflate.Stored.cs:24-84): block records only, no symbols, no Adler sums (the caller checksums the input).
cudaError_t launch_stored_blocks(const Lz77Params& P, cudaStream_t stream);

// ---- Huffman stage (zsb_huffman.cu) --------------------------------------------------------------
constexpr int kHdrWords = 80;   // dynamic tree description: <= 14 + 57 + 316 * 7 bits
constexpr int kTailWords = 8;   // per-segment tail: empty static blocks and full-flush markers (Q10)

struct BlockCodesDev {
    uint32_t lcode[L_CODES];   // code | len << 16, code already bit-reversed
    uint32_t dcode[D_CODES];
    uint32_t hdr[kHdrWords];   // LSB-first bit string sent after the 3 block-type bits of a dynamic block
};

struct SegLayout {
    uint64_t out_offset;   // byte offset of the segment's first output byte inside this call's output
    uint32_t out_bytes;    // compressed size of the segment
    uint32_t tail_nbits;
    uint64_t tail_bit_offset;       // relative to the segment's first output bit
    uint32_t tail[kTailWords];      // LSB-first bits
    uint32_t bit_shift;             // FlushMode.PartialFlush: the segment starts this many bits into its first byte (else 0)
    uint32_t pad_;
};

struct HuffParams {
    const uint8_t* in;
    size_t n;
    uint32_t seg_size;
    uint32_t nseg;
    uint64_t first_seg_index;
    int strategy;
    const uint32_t* sym;
    size_t sym_stride;
    BlockRec* blocks;
    uint32_t max_blocks;
    const uint32_t* nblocks;
    BlockCodesDev* codes;     // [nseg][max_blocks]
    SegLayout* layout;        // [nseg]
    uint64_t out_base;        // bytes already in `out` before the first segment of this launch (header, earlier batches)
    uint64_t* total_out;      // device scalar: out offset after the last segment of this launch
    uint64_t* seg_end;        // optional [nseg]: end offset of each segment inside `out`
    uint8_t* out;
    size_t out_cap;
    int* status;              // device int: set to ZSB_BUF_ERROR if out_cap is too small
    int partial_flush = 0;    // FlushMode.PartialFlush: Tr_align instead of the stored marker; segments are not byte aligned
    uint64_t* total_bits = nullptr;   // partial_flush: device scalar, the stream's bit position after the last segment
    int level0 = 0;           // level 0: every block is stored (Trees.cs:586-594 with level == 0), also the empty ones of Q10
    int finish_mode = 0;      // 1: one Write with FlushMode.NoFlush, then Finish: the last block carries BFINAL, no flush marker
};
cudaError_t launch_block_trees(const HuffParams& P, cudaStream_t stream);
cudaError_t launch_layout(const HuffParams& P, cudaStream_t stream);
cudaError_t launch_pack(const HuffParams& P, cudaStream_t stream);

// ---- Adler-32 (zsb_adler32.cu) -------------------------------------------------------------------
// Folds per-segment (s1, s2) pairs (seed 1 each) into one checksum of the concatenation.
cudaError_t launch_adler_fold(const uint2* seg_adler, uint32_t nseg, uint32_t seg_size, size_t n, uint32_t seed,
                              uint32_t* result, cudaStream_t stream);
// Stand-alone checksum: per-chunk partial sums + fold.
cudaError_t launch_adler32(const uint8_t* p, size_t n, uint32_t seed, uint2* scratch, uint32_t scratch_len,
                           uint32_t* result, cudaStream_t stream);
constexpr uint32_t kAdlerChunk = 1u << 16;

// ---- Inflate (zsb_inflate.cu) ---------------------------------------------------------------------
struct InflateParams {
    const uint8_t* in;
    size_t n;                  // compressed bytes
    const uint64_t* seg_end;   // [nseg] end offsets of each segment inside `in`
    uint64_t first_offset;     // offset of the first segment's first byte (2 after a zlib header)
    uint32_t nseg;
    uint32_t seg_size;
    uint8_t* out;
    size_t out_cap;            // total uncompressed capacity
    uint32_t* seg_out_len;     // [nseg] bytes produced per segment
    uint2* seg_adler;          // [nseg]
    int* seg_status;           // [nseg] 0 ok, else error id
    unsigned int* work_counter;
    uint32_t* seg_carry = nullptr;   // [nseg] Q14: leading bytes that repeat the previous segment's last byte (lane kernel)
    unsigned int* carry_flag = nullptr;   // set to 1 if any seg_carry is non-zero
    // lane kernel, segment index building: count_only decodes lengths and status only; piece k starts at
    // seg_end[k-1] (first_offset for k = 0) and runs to the first empty stored block (the full-flush marker,
    // Deflate.cs:710-712), whose end offset goes to seg_found_end[k]; it may read up to count_body_end and produce at
    // most seg_size bytes
    int continuation = 0;      // `in` does not start at the stream's first segment: segment 0 of the call may open with a Q14 run
    int prev_byte = -1;        // carry fix-up: the byte before the call's first segment, once the caller knows it (else -1)
    int count_only = 0;
    uint64_t count_body_end = 0;
    uint64_t* seg_found_end = nullptr;
};
cudaError_t launch_inflate(const InflateParams& P, int num_sms, cudaStream_t stream);
// Lane-per-segment kernel (zsb_inflate_lanes.cu): the default.  scratch: inflate_lanes_scratch_bytes(num_sms) bytes.
size_t inflate_lanes_scratch_bytes(int num_sms);
cudaError_t launch_inflate_lanes(const InflateParams& P, int num_sms, uint8_t* scratch, size_t scratch_bytes, cudaStream_t stream);
// Same kernel with smaller per-lane tables (zsb_inflate_lanes_small.cu): more resident warps, for very many segments.
cudaError_t launch_inflate_lanes_small(const InflateParams& P, int num_sms, uint8_t* scratch, size_t scratch_bytes, cudaStream_t stream);
// Q14 fix-up after the lane kernel: fills the carried leading bytes of Rle segments.
cudaError_t launch_inflate_carry_fixup(const InflateParams& P, cudaStream_t stream);
// Lists the end offset of every 00 00 FF FF inside in[begin, end) (unordered; *ncand may exceed cand_cap).
cudaError_t launch_index_markers(const uint8_t* in, size_t begin, size_t end, uint64_t* cand, uint32_t cand_cap, uint32_t* ncand,
                                 cudaStream_t stream);

// ---- PNG front end (zsb_png.cu) ---------------------------------------------------------------------------------------
struct PngFilterParams {
    const uint8_t* px;     // [height][rowbytes] raw samples, 8 bits each
    uint8_t* out;          // [height][1 + rowbytes]: filter type byte + filtered bytes
    uint32_t rowbytes, height, bpp;
    int filter;            // 0 None, 1 Sub, 2 Up, 3 Average, 4 Paeth, 5 adaptive (per row)
};
cudaError_t launch_png_filter(const PngFilterParams& P, cudaStream_t stream);
constexpr uint32_t kCrcChunk = 1u << 16;
struct CrcMults { uint32_t level[8], chunk, x8; };
uint32_t crc32_host(const uint8_t* p, size_t n);
uint32_t crc32_combine_host(uint32_t crc1, uint32_t crc2, uint64_t len2);
// CRC-32 of every kCrcChunk bytes of p (the last chunk may be shorter)
cudaError_t launch_crc32_chunks(const uint8_t* p, size_t n, uint32_t* chunk_crc, cudaStream_t stream);
// crc("IDAT" || bytes of IDAT k) for IDAT chunks of idat_bytes (a multiple of kCrcChunk) stream bytes each
cudaError_t launch_png_idat_crc(const uint32_t* chunk_crc, size_t n, size_t idat_bytes, uint32_t nidat, uint32_t* out, cudaStream_t stream);

enum InflateErr {
    IE_OK = 0, IE_BLOCK_TYPE, IE_STORED_LEN, IE_TOO_MANY_SYMS, IE_BAD_REPEAT, IE_OVERSUB_BL, IE_INCOMPLETE_BL,
    IE_OVERSUB_L, IE_INCOMPLETE_L, IE_OVERSUB_D, IE_INCOMPLETE_D, IE_EMPTY_D, IE_BAD_LCODE, IE_BAD_DCODE,
    IE_DIST_TOO_FAR, IE_TRUNCATED, IE_OUTPUT_OVERFLOW, IE_SEG_SIZE
};

}  // namespace zsb
