// zsb_inflate.cu -- inflate of full-flush-segmented deflate streams: one group of GW lanes per segment.
//
// Replaces Inflate.Decompress (Inflate.cs:103-357, block loop), InflateBlocks.Process
// (InflateBlocks.cs:130-713), InfCodes.Inflate_fast / Process (InfCodes.cs:106-817) and
// InfTree.Huft_build (InfTree.cs:125-365; RFC 1951 semantics, not the reference's broken table copy,
// SURVEY Q15).  A full-flush segment starts byte aligned with an empty history (Deflate.cs:596-605), so
// segments decode independently; the per-warp state is two lookup tables in shared memory, output goes
// straight to HBM and back-references are served from L2.
#include "zsb_kernels.cuh"

namespace zsb {

namespace {

// One GROUP of GW lanes per segment (GW = 8: four segments per warp).  With a whole warp per segment the kernel is
// issue-bound -- the serial symbol decode keeps one lane of 32 busy, ~50 warp instructions per symbol, 40 warps per SM
// (32 GB/s measured); narrower groups make every issued instruction serve several segments.
#ifndef ZSB_INF_GW
#define ZSB_INF_GW 32
#endif
constexpr int GW = ZSB_INF_GW;
constexpr int INF_THREADS = 64;
constexpr int INF_GROUPS = INF_THREADS / GW;
constexpr int PEND_PER_LANE = GW >= 32 ? 1 : 4;   // bytes of an in-flight match copy each lane carries
constexpr int PEND_MAX = PEND_PER_LANE * GW;
constexpr int FAST_L = 10;   // root table bits, literal/length
constexpr int FAST_D = 8;    // root table bits, distance

struct WarpTables {
    uint16_t fast_l[1 << FAST_L];   // (sym << 4) | len, 0 = longer than FAST_L bits
    uint16_t fast_d[1 << FAST_D];
    uint16_t sorted_l[288];
    uint16_t sorted_d[32];
    uint16_t count_l[16], first_l[16], offset_l[16];
    uint16_t count_d[16], first_d[16], offset_d[16];
    uint16_t code_tmp[320];
    uint8_t lens[352];
};

__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct BitReader {   // lane 0 only; positions are relative to the segment's first compressed byte
    const uint8_t* in;   // the segment's first compressed byte
    uint32_t pos;        // offset of the next byte to load
    uint32_t end;        // the segment's compressed size
    uint32_t limit;      // bytes that may be loaded (to the end of the whole compressed buffer, capped)
    uint64_t acc;
    int nbits;
    // Call with nbits <= 32: one aligned 32-bit load brings the accumulator to >= 33 bits (7 instructions; the earlier
    // two-word 64-bit refill cost 23 and ran on 41 % of the symbols).  Bytes past `end` (the next segment's) may enter the
    // accumulator but are never consumed by a well-formed segment -- consumption is checked against `end` at every
    // block boundary.  Unaligned starts (segment start, after a stored block) and the last bytes of the buffer go bytewise.
    __device__ __forceinline__ void refill()
    {
        const uint8_t* p = in + pos;
        if ((reinterpret_cast<uintptr_t>(p) & 3u) == 0 && pos + 4 <= limit) {
            acc |= (uint64_t)__ldg(reinterpret_cast<const uint32_t*>(p)) << nbits;
            pos += 4;
            nbits += 32;
        } else {
            while (nbits <= 56 && pos < limit) {
                acc |= (uint64_t)in[pos++] << nbits;
                nbits += 8;
                if (nbits > 32 && (reinterpret_cast<uintptr_t>(in + pos) & 3u) == 0) break;
            }
        }
    }
    __device__ __forceinline__ bool need(int k)
    {
        if (nbits < k) refill();
        return nbits >= k;
    }
    __device__ __forceinline__ uint32_t peek(int k) const { return (uint32_t)acc & ((1u << k) - 1u); }   // k <= 16
    __device__ __forceinline__ void drop(int k) { acc >>= k; nbits -= k; }
    __device__ __forceinline__ uint32_t byte_pos() const { return pos - (uint32_t)(nbits >> 3); }
};

// Builds one canonical decoding table from sm.lens[base .. base+n).  Whole warp.
// Returns 0 ok, 1 all zero, -1 oversubscribed, -2 incomplete (InfTree.cs huft_build rules).
__device__ int build_table(WarpTables& t, int base, int n, bool dist, int lane, uint32_t gmask)
{
    uint16_t* fast = dist ? t.fast_d : t.fast_l;
    uint16_t* sorted = dist ? t.sorted_d : t.sorted_l;
    uint16_t* count = dist ? t.count_d : t.count_l;
    uint16_t* first = dist ? t.first_d : t.first_l;
    uint16_t* offset = dist ? t.offset_d : t.offset_l;
    const int fast_bits = dist ? FAST_D : FAST_L;
    for (int i = lane; i < (1 << fast_bits); i += GW) fast[i] = 0;
    int rc = 0;
    if (lane == 0) {
        for (int i = 0; i < 16; i++) count[i] = 0;
        for (int i = 0; i < n; i++) count[t.lens[base + i]]++;
        if (count[0] == n) rc = 1;
        else {
            int left = 1, maxlen = 0;
            for (int len = 1; len <= MAX_BITS; len++) {
                left <<= 1;
                left -= count[len];
                if (left < 0) { rc = -1; break; }
                if (count[len]) maxlen = len;
            }
            if (rc == 0) {
                count[0] = 0;
                uint32_t code = 0;
                offset[1] = 0;
                uint16_t next_off[16], next_code[16];
                for (int len = 1; len <= MAX_BITS; len++) {
                    code = (code + count[len - 1]) << 1;
                    first[len] = (uint16_t)code;
                    next_code[len] = (uint16_t)code;
                    if (len < MAX_BITS) offset[len + 1] = (uint16_t)(offset[len] + count[len]);
                }
                for (int len = 1; len <= MAX_BITS; len++) next_off[len] = offset[len];
                for (int i = 0; i < n; i++) {
                    int len = t.lens[base + i];
                    if (len) {
                        sorted[next_off[len]++] = (uint16_t)i;
                        t.code_tmp[i] = next_code[len]++;
                    }
                }
                if (left > 0 && maxlen != 1) rc = -2;
            }
        }
    }
    rc = __shfl_sync(gmask, rc, 0, GW);
    __syncwarp(gmask);
    if (rc == 0 || rc == -2) {
        for (int i = lane; i < n; i += GW) {
            const int len = t.lens[base + i];
            if (len && len <= fast_bits) {
                const uint32_t r = __brev((uint32_t)t.code_tmp[i]) >> (32 - len);
                for (uint32_t k = r; k < (1u << fast_bits); k += 1u << len) fast[k] = (uint16_t)((i << 4) | len);
            }
        }
    }
    __syncwarp(gmask);
    return rc;
}

// lane 0: decode one symbol. >= 0 symbol, -1 invalid code, -2 out of input
__device__ __forceinline__ int decode_sym(BitReader& b, const uint16_t* fast, int fast_bits, const uint16_t* count,
                                          const uint16_t* first, const uint16_t* offset, const uint16_t* sorted)
{
    if (b.nbits <= 32) b.refill();
    const uint32_t e = fast[b.acc & ((1u << fast_bits) - 1)];
    if (e) {
        const int len = (int)(e & 15);
        if (len > b.nbits) return -2;
        b.drop(len);
        return (int)(e >> 4);
    }
    uint32_t code = 0;
    uint64_t acc = b.acc;
    for (int len = 1; len <= MAX_BITS; len++) {
        if (len > b.nbits) return -2;
        code = (code << 1) | (uint32_t)(acc & 1);
        acc >>= 1;
        const uint32_t cnt = count[len];
        if (cnt && code >= first[len] && code - first[len] < cnt) {
            b.drop(len);
            return sorted[offset[len] + (code - first[len])];
        }
    }
    return -1;
}

enum { OP_LIT = 0, OP_MATCH = 1, OP_EOB = 2, OP_ERR = 3 };

}  // namespace

__global__ void __launch_bounds__(INF_THREADS) inflate_segments_kernel(InflateParams P)
{
    __shared__ WarpTables tabs[INF_GROUPS];
    const int lane = threadIdx.x % GW;   // lane within the group
    const uint32_t gmask = GW >= 32 ? 0xffffffffu : ((1u << (GW & 31)) - 1u) << ((threadIdx.x & 31) - lane);
    WarpTables& t = tabs[threadIdx.x / GW];

    for (;;) {
        uint32_t seg = 0;
        if (lane == 0) seg = atomicAdd(P.work_counter, 1u);
        seg = __shfl_sync(gmask, seg, 0, GW);
        if (seg >= P.nseg) break;

        const uint64_t c_begin = seg == 0 ? P.first_offset : P.seg_end[seg - 1];
        const uint64_t c_end = P.seg_end[seg];
        const uint64_t o_begin = (uint64_t)seg * P.seg_size;
        const uint64_t o_room = o_begin < P.out_cap ? P.out_cap - o_begin : 0;   // never past the caller's buffer
        const uint32_t o_cap = (uint32_t)(o_room < P.seg_size ? o_room : P.seg_size);
        uint8_t* out = P.out + o_begin;
        uint32_t o = 0;
        uint32_t carry = 0;   // Q14: leading bytes that repeat the previous segment's last byte (Rle streams)
        int err = IE_OK;
        uint32_t pend_at = 0;     // a short match copy whose loads are in flight (stores still to be issued)
        int pend_len = 0;
        uint8_t pend_val[PEND_PER_LANE] = {0};
        BitReader br;
        br.in = P.in + c_begin;
        br.pos = 0;
        br.acc = 0;
        br.nbits = 0;
        br.end = 0;
        br.limit = 0;
        if (c_end > P.n || c_begin > c_end || c_end - c_begin > 0xF0000000ull) err = IE_TRUNCATED;
        else {
            br.end = (uint32_t)(c_end - c_begin);
            const uint64_t room = P.n - c_begin;
            br.limit = room > 0xFFFFFF00ull ? 0xFFFFFF00u : (uint32_t)room;
        }

        // blocks until the segment's compressed bytes are used up (it ends with the full-flush marker) or, in a one-piece
        // stream, until its final block is done (Inflate.cs:286-300)
        bool final_seen = false;
        while (err == IE_OK && !final_seen) {
            int hdr = -1;
            if (lane == 0) {
                const uint32_t used = br.byte_pos();
                if (used > br.end) hdr = -3;            // ran past the segment's bytes
                else if (used == br.end) hdr = -2;      // clean end of segment (after the byte-aligned full-flush marker)
                else if (!br.need(3)) hdr = -3;
                else { hdr = (int)br.peek(3); br.drop(3); }
            }
            hdr = __shfl_sync(gmask, hdr, 0, GW);
            if (hdr == -2) break;
            if (hdr == -3) { err = IE_TRUNCATED; break; }
            const int type = hdr >> 1;
            final_seen = (hdr & 1) != 0;   // segments of a full-flush stream never hold one: the stream's `03 00` lies outside them
            if (type == 0) {
                // stored (InflateBlocks.cs:251-364)
                int len = -1;
                if (lane == 0) {
                    br.drop(br.nbits & 7);
                    if (br.need(32)) {
                        const uint32_t l = br.peek(16);
                        br.drop(16);
                        const uint32_t nl = br.peek(16);
                        br.drop(16);
                        len = (((~nl) & 0xFFFF) == l) ? (int)l : -2;
                    }
                }
                len = __shfl_sync(gmask, len, 0, GW);
                if (len == -1) { err = IE_TRUNCATED; break; }
                if (len == -2) { err = IE_STORED_LEN; break; }
                uint32_t src = 0, send = 0;
                if (lane == 0) { src = br.byte_pos(); send = br.end; }
                src = __shfl_sync(gmask, src, 0, GW);
                send = __shfl_sync(gmask, send, 0, GW);
                if (src + (uint32_t)len > send) { err = IE_TRUNCATED; break; }
                if (o + (uint32_t)len > o_cap) { err = IE_OUTPUT_OVERFLOW; break; }
                {
                    const uint8_t* sp = P.in + c_begin + src;
                    for (int i = lane; i < len; i += GW) out[o + i] = sp[i];
                }
                o += (uint32_t)len;
                if (lane == 0) { br.pos = src + (uint32_t)len; br.acc = 0; br.nbits = 0; }
                __syncwarp(gmask);
                continue;
            }
            if (type == 3) { err = IE_BLOCK_TYPE; break; }
            if (type == 1) {
                // fixed tables (InfTree.cs:16-96)
                for (int i = lane; i < 288; i += GW) t.lens[i] = (uint8_t)static_llen(i);
                for (int i = lane; i < 30; i += GW) t.lens[288 + i] = 5;
                __syncwarp(gmask);
                build_table(t, 0, 288, false, lane, gmask);
                build_table(t, 288, 30, true, lane, gmask);
            } else {
                // dynamic header (InflateBlocks.cs:366-626)
                int nl = 0, nd = 0, e = IE_OK;
                if (lane == 0) {
                    if (!br.need(14)) e = IE_TRUNCATED;
                    else {
                        const uint32_t v = br.peek(14);
                        br.drop(14);
                        if ((v & 0x1f) > 29 || ((v >> 5) & 0x1f) > 29) e = IE_TOO_MANY_SYMS;
                        else {
                            nl = 257 + (int)(v & 0x1f);
                            nd = 1 + (int)((v >> 5) & 0x1f);
                            const int nb = 4 + (int)(v >> 10);
                            for (int i = 0; i < 19; i++) t.lens[i] = 0;
                            for (int i = 0; i < nb && e == IE_OK; i++) {
                                if (!br.need(3)) e = IE_TRUNCATED;
                                else { t.lens[c_order[i]] = (uint8_t)br.peek(3); br.drop(3); }
                            }
                        }
                    }
                }
                e = __shfl_sync(gmask, e, 0, GW);
                if (e != IE_OK) { err = e; break; }
                nl = __shfl_sync(gmask, nl, 0, GW);
                nd = __shfl_sync(gmask, nd, 0, GW);
                __syncwarp(gmask);
                // the bit-length code is decoded through the distance-table slots
                int r = build_table(t, 0, 19, true, lane, gmask);
                if (r == -1) { err = IE_OVERSUB_BL; break; }
                if (r != 0) { err = IE_INCOMPLETE_BL; break; }
                if (lane == 0) {
                    int i = 0;
                    uint8_t prev = 0;
                    while (i < nl + nd && e == IE_OK) {
                        const int sym = decode_sym(br, t.fast_d, FAST_D, t.count_d, t.first_d, t.offset_d, t.sorted_d);
                        if (sym == -2) e = IE_TRUNCATED;
                        else if (sym < 0) e = IE_BAD_REPEAT;
                        else if (sym < 16) { t.lens[32 + i++] = prev = (uint8_t)sym; }
                        else {
                            const int eb = sym == 18 ? 7 : sym - 14, rb = sym == 18 ? 11 : 3;
                            if (!br.need(eb)) { e = IE_TRUNCATED; break; }
                            int rep = rb + (int)br.peek(eb);
                            br.drop(eb);
                            if (i + rep > nl + nd || (sym == 16 && i < 1)) { e = IE_BAD_REPEAT; break; }
                            const uint8_t v = sym == 16 ? prev : 0;
                            if (sym != 16) prev = 0;
                            while (rep--) t.lens[32 + i++] = v;
                        }
                    }
                }
                e = __shfl_sync(gmask, e, 0, GW);
                if (e != IE_OK) { err = e; break; }
                __syncwarp(gmask);
                r = build_table(t, 32, nl, false, lane, gmask);
                if (r == -1) { err = IE_OVERSUB_L; break; }
                if (r != 0) { err = IE_INCOMPLETE_L; break; }
                r = build_table(t, 32 + nl, nd, true, lane, gmask);
                if (r == -1) { err = IE_OVERSUB_D; break; }
                if (r == -2) { err = IE_INCOMPLETE_D; break; }
                if (r == 1 && nl > 257) { err = IE_EMPTY_D; break; }
            }
            // ---- symbols (InfCodes.cs:106-817) ----
            // Lane 0 decodes and stores literals on its own until it meets a match or the end of the block; the warp
            // then copies the match.  A copy of <= PEND_MAX bytes is split: its loads are issued at once, its stores wait
            // until the next match / end of block, so the L2 round trip of the back-reference hides behind the
            // literals decoded in between (one copy in flight: the next match may read what this one writes).
            for (;;) {
                uint32_t packed = 0;   // op | len << 2 | dist << 12   (or op | error << 2)
                uint32_t onow = 0;
                if (lane == 0) {
                    int op = OP_ERR, a = IE_BAD_LCODE, d = 0;
                    for (;;) {
                        if (br.nbits <= 32) br.refill();   // >= 33 bits: a literal/length code (<= 15) and its extra bits (<= 5)
                        uint32_t e = t.fast_l[br.acc & ((1u << FAST_L) - 1)];
                        int sym;
                        if (e) {
                            br.drop((int)(e & 15));
                            sym = (int)(e >> 4);
                        } else {
                            sym = decode_sym(br, t.fast_l, FAST_L, t.count_l, t.first_l, t.offset_l, t.sorted_l);
                            if (sym < 0) { a = sym == -2 ? IE_TRUNCATED : IE_BAD_LCODE; break; }
                        }
                        if (sym < 256) {
                            if (o >= o_cap) { a = IE_OUTPUT_OVERFLOW; break; }
                            out[o++] = (uint8_t)sym;
                            continue;
                        }
                        if (sym == 256) { op = OP_EOB; break; }
                        const int ls = sym - 257;
                        if (ls >= 29) { a = IE_BAD_LCODE; break; }
                        int eb = c_len_extra[ls];
                        const int len = c_len_base[ls] + (int)br.peek(eb);
                        br.drop(eb);
                        int ds;
                        if (br.nbits <= 32) br.refill();   // >= 33 bits: a distance code (<= 15) and its extra bits (<= 13)
                        e = t.fast_d[br.acc & ((1u << FAST_D) - 1)];
                        if (e) {
                            br.drop((int)(e & 15));
                            ds = (int)(e >> 4);
                        } else {
                            ds = decode_sym(br, t.fast_d, FAST_D, t.count_d, t.first_d, t.offset_d, t.sorted_d);
                            if (ds < 0) { a = ds == -2 ? IE_TRUNCATED : IE_BAD_DCODE; break; }
                        }
                        if (ds >= 30) { a = IE_BAD_DCODE; break; }
                        eb = c_dist_extra[ds];
                        d = c_dist_base[ds] + (int)br.peek(eb);
                        br.drop(eb);
                        if (br.nbits < 0) { a = IE_TRUNCATED; break; }
                        a = len;
                        op = OP_MATCH;
                        break;
                    }
                    if (op != OP_ERR && br.nbits < 0) { op = OP_ERR; a = IE_TRUNCATED; }
                    packed = (uint32_t)op | ((uint32_t)a << 2) | ((uint32_t)d << 12);
                    onow = o;
                }
                packed = __shfl_sync(gmask, packed, 0, GW);
                o = __shfl_sync(gmask, onow, 0, GW);
                const int op = (int)(packed & 3u);
                // finish the copy in flight before anything else touches the output
                if (pend_len) {
#pragma unroll
                    for (int j = 0; j < PEND_PER_LANE; j++) {
                        const int i = lane + j * GW;
                        if (i < pend_len) out[pend_at + i] = pend_val[j];
                    }
                    pend_len = 0;
                }
                __syncwarp(gmask);
                if (op == OP_EOB) break;
                if (op == OP_ERR) { err = (int)((packed >> 2) & 0x3ffu); break; }
                const int a = (int)((packed >> 2) & 0x3ffu), d = (int)(packed >> 12);
                if (o + (uint32_t)a > o_cap) { err = IE_OUTPUT_OVERFLOW; break; }
                if ((uint32_t)d > o || o - (uint32_t)d < carry) {
                    // Q14 (Deflate.Rle.cs:47-50): distance-1 runs of the previous segment's last byte open the segment;
                    // placeholders now, inflate_carry_fixup_kernel fills them in once every segment is decoded
                    if (!(d == 1 && o == carry && (seg > 0 || P.continuation))) { err = IE_DIST_TOO_FAR; break; }
                    for (int i = lane; i < a; i += GW) out[o + i] = 0;
                    carry += (uint32_t)a;
                    o += (uint32_t)a;
                    __syncwarp(gmask);
                    continue;
                }
                // out[o + i] = out[o - d + (i mod d)]: every source byte was written before this match
                if (a <= PEND_MAX) {
#pragma unroll
                    for (int j = 0; j < PEND_PER_LANE; j++) {
                        const int i = lane + j * GW;
                        if (i < a) {
                            int k = i;
                            if (k >= d) k %= d;   // overlapping copy (run): rare, so the division stays out of the common path
                            pend_val[j] = __ldcg(out + (o - (uint32_t)d) + k);
                        }
                    }
                    pend_at = o;
                    pend_len = a;
                } else {
                    for (int i = lane; i < a; i += GW) {
                        int k = i;
                        if (k >= d) k %= d;
                        out[o + i] = __ldcg(out + (o - (uint32_t)d) + k);
                    }
                    __syncwarp(gmask);
                }
                o += (uint32_t)a;
            }
            if (err != IE_OK) break;
        }
        // every segment but the last holds exactly seg_size bytes (that is what places the output)
        if (err == IE_OK && seg + 1 < P.nseg && o != P.seg_size) err = IE_SEG_SIZE;
        if (lane == 0) {
            P.seg_out_len[seg] = (uint32_t)o;
            P.seg_status[seg] = err;
            if (P.seg_carry) P.seg_carry[seg] = carry;
            if (carry && P.carry_flag) *P.carry_flag = 1;
        }
        __syncwarp(gmask);
    }
}

cudaError_t launch_inflate(const InflateParams& P, int num_sms, cudaStream_t stream)
{
    if (P.nseg == 0) return cudaSuccess;
    // queried per call: the answer belongs to the current device, and a process may hold contexts on several
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, inflate_segments_kernel, INF_THREADS, 0) != cudaSuccess || per_sm < 1)
        per_sm = 4;
    int ctas = num_sms * per_sm;   // persistent: every resident warp pulls segments from the work counter
    const int need = (int)((P.nseg + INF_GROUPS - 1) / INF_GROUPS);
    if (ctas > need) ctas = need;
    inflate_segments_kernel<<<ctas, INF_THREADS, 0, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace zsb
