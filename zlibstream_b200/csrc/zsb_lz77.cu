// zsb_lz77.cu -- LZ77 match finding for full-flush segments: one CTA per segment.
//
// Replaces Deflate.Fill_window / InsertString / Longest_match (Deflate.cs:866-1100), DeflateFast
// (Deflate.Fast.cs:20-128), DeflateSlow (Deflate.Slow.cs:18-159), DeflateRle (Deflate.Rle.cs:18-104),
// Compare256 / SlideHash / UpdateHash (Deflate.Intrinsics.cs) and Tr_tally_* (Deflate.cs:910-948).
//
// HBM layout: input is read once, 32 KiB at a time, straight into the shared-memory window with a bulk
// async copy (TMA engine, cp.async.bulk + mbarrier); the 64 KiB window, head[32768] and prev[32768]
// live in shared memory in the reference's own window coordinates (192 KiB per CTA, one CTA per SM),
// so the refill / slide schedule -- which the output depends on -- is reproduced
// literally.  Symbols (dist | lc << 16) are streamed to HBM; Adler-32 partial sums are taken from the
// staged window so the checksum costs no extra HBM pass.
#include "zsb_kernels.cuh"
#include "zsb_lz77_dev.cuh"

#include <cstdlib>
#include <cstring>

namespace zsb {

namespace {

constexpr int LZ_THREADS = 256;
constexpr int WIN_PAD = 64;
constexpr int SLOT_COUNT = 4096;
constexpr int TILE_MAX = 256;   // positions per DeflateSlow tile (= LZ_THREADS, < MIN_LOOKAHEAD)
constexpr uint32_t SLOT_MASK = SLOT_COUNT - 1;
constexpr int WIDE_TILE = 992;        // positions per tile of the pipelined DeflateSlow: one per thread of 31 producer warps
constexpr int WIDE_SLOTS = 1024;      // slot table of the 1024-thread kernel (only slow_tile_insert uses it there)
constexpr int SHADOW_N = 1024;        // >= WIDE_TILE, power of two

struct ParserState {
    int str_start, lookahead, block_start, match_length, prev_length, match_start, prev_match, match_available;
    int avail_in;      // bytes of the segment not yet read into the window
    int in_pos;        // segment-relative offset of the next byte to read
    int abs_base;      // segment-relative offset of window index 0
    int last_lit;      // symbols in the open block
    int sym_total;     // symbols emitted so far in this segment
    int nblocks;
    int done;
};

struct Smem {
    uint8_t* win;    // WINDOW_SIZE + WIN_PAD
    uint16_t* head;  // HASH_SIZE
    uint16_t* prev;  // W_SIZE
    uint16_t* crc;   // 4 * 256, low 16 bits of the slice-by-4 CRC-32C tables
    uint32_t* slot;  // SLOT_COUNT: first-lane-per-hash-slot arbitration of the batch parser
    uint32_t* batch_seq;
    uint16_t* tile_hh;   // TILE_MAX: what InsertString returned for each position of the current DeflateSlow tile
    uint2* tile_rec;     // TILE_MAX: Longest_match results of the tile (full / quarter chain budget)
    // pipelined DeflateSlow (1024-thread kernel): tiles lie on a fixed grid up to pipe_ins_limit and are inserted / searched
    // by producer warps while warp 0 replays; pipe_flags[0] = tiles produced, [1] = tiles consumed.  0 / nullptr otherwise.
    int pipe_ins_limit;
    int pipe_ntiles;
    volatile int* pipe_flags;
    uint32_t slot_mask;   // entries of `slot` - 1
    // prev[] slots overwritten by the inserts of the current wide tile keep their old value here (index: position & 1023),
    // so that a search may run after MORE than MIN_LOOKAHEAD later positions have been inserted.  nullptr otherwise.
    uint16_t* shadow;
    int shadow_lo, shadow_hi;   // the tile [lo, hi) whose inserts the shadow covers
    uint32_t* fast_masks; // 2 * 8 ballots of the DeflateFast batch (per warp)
    uint16_t* fast_endl;  // 64: next parse point of each batch position if it is one
    uint16_t* fast_mst;   // 64: match start of each batch position
};

__device__ __forceinline__ uint32_t hash_of(const Smem& sm, uint32_t v, int hash_variant)
{
    if (hash_variant == 0) {  // Sse42.Crc32(0, v): Deflate.Intrinsics.cs:298-301
        uint32_t h = sm.crc[(v & 0xFF)] ^ sm.crc[256 + ((v >> 8) & 0xFF)] ^ sm.crc[512 + ((v >> 16) & 0xFF)] ^
                     sm.crc[768 + (v >> 24)];
        return h & HASH_MASK;
    }
    return ((v * 2654435761u) >> 16) & HASH_MASK;  // Deflate.Intrinsics.cs:305
}

// Deflate.cs:866-877 (Q1: hashes the 4 bytes at str+2; Q2: no update when head already == str)
__device__ __forceinline__ int insert_string(const Smem& sm, int str, int hash_variant)
{
    uint32_t h = hash_of(sm, lds_u32_unaligned(sm.win, str + (MIN_MATCH - 1)), hash_variant);
    int cur = sm.head[h];
    if (cur != str) {
        sm.prev[str & W_MASK] = (uint16_t)cur;
        sm.head[h] = (uint16_t)str;
    }
    return cur;
}

// Compare256 (Deflate.Intrinsics.cs:19-162): common prefix length of two window strings, <= 256.
__device__ __forceinline__ int compare256(const uint8_t* win, int a, int b)
{
    int len = 0;
    while (len < 256) {
        uint32_t x = lds_u32_unaligned(win, a + len) ^ lds_u32_unaligned(win, b + len);
        if (x) return len + ((__ffs(x) - 1) >> 3);
        len += 4;
    }
    return 256;
}

__device__ __forceinline__ uint32_t lds_u16_unaligned(const uint8_t* win, int idx)
{
    return lds_u32_unaligned(win, idx) & 0xFFFF;
}

// Deflate.cs:1022-1100
__device__ int longest_match(const Smem& sm, ParserState& s, const LevelCfg& cfg, int cur_match)
{
    int chain_length = cfg.chain;
    const int scan = s.str_start;
    int best_len = s.prev_length;
    const int limit = s.str_start > MAX_DIST ? s.str_start - MAX_DIST : 0;
    int nice_match = cfg.nice;
    int match_start = s.match_start;
    if (best_len == 0) best_len = 1;
    const uint32_t scan_start = lds_u16_unaligned(sm.win, scan);
    uint32_t scan_end = lds_u16_unaligned(sm.win, scan + best_len - 1);
    if (s.prev_length >= cfg.good) chain_length >>= 2;
    if (nice_match > s.lookahead) nice_match = s.lookahead;
    do {
        if (cur_match >= s.str_start) break;
        if (lds_u16_unaligned(sm.win, cur_match + best_len - 1) != scan_end ||
            lds_u16_unaligned(sm.win, cur_match) != scan_start)
            continue;
        int len = compare256(sm.win, scan + 2, cur_match + 2) + 2;
        if (len > best_len) {
            match_start = cur_match;
            best_len = len;
            if (len >= nice_match) break;
            scan_end = lds_u16_unaligned(sm.win, scan + best_len - 1);
        }
    } while ((cur_match = sm.prev[cur_match & W_MASK]) > limit && --chain_length != 0);
    s.match_start = match_start;
    return best_len < s.lookahead ? best_len : s.lookahead;
}

struct SegOut {
    uint32_t* sym;
    BlockRec* blocks;
    int max_blocks;
};

// Flush_block_only (Deflate.cs:951-956) as far as the parser is concerned: close the block record.
__device__ __forceinline__ void close_block(ParserState& s, const SegOut& o)
{
    if (s.nblocks < o.max_blocks) {
        BlockRec& b = o.blocks[s.nblocks];
        b.sym_begin = (uint32_t)(s.sym_total - s.last_lit);
        b.sym_count = (uint32_t)s.last_lit;
        b.in_begin = (uint32_t)(s.block_start + s.abs_base);
        b.in_len = (uint32_t)(s.str_start - s.block_start);
        b.buf_valid = s.block_start >= 0;
    }
    s.nblocks++;
    s.last_lit = 0;
    s.block_start = s.str_start;
}

// Both return the reference's `bflush`: the caller closes the block where the reference calls
// Flush_block_only, i.e. after str_start has moved past the symbol.
__device__ __forceinline__ bool tally_dist(ParserState& s, const SegOut& o, int dist, int lc)  // Deflate.cs:910-926
{
    o.sym[s.sym_total++] = (uint32_t)dist | ((uint32_t)lc << 16);
    return ++s.last_lit == BLOCK_SYMS;
}
__device__ __forceinline__ bool tally_lit(ParserState& s, const SegOut& o, int c)  // Deflate.cs:935-948
{
    o.sym[s.sym_total++] = (uint32_t)c << 16;
    return ++s.last_lit == BLOCK_SYMS;
}

// True when Fill_window (Deflate.cs:967-1019) would do something that needs the whole CTA.
__device__ __forceinline__ bool fill_needed(const ParserState& s)
{
    return s.avail_in > 0 || s.str_start >= WINDOW_SIZE - MIN_LOOKAHEAD;
}

// One iteration of the DeflateFast loop body (Deflate.Fast.cs:43-112) at a position that does not need Fill_window.
__device__ void fast_step(const Smem& sm, ParserState& s, const SegOut& o, const LevelCfg& cfg, int strategy,
                          int hash_variant)
{
    int hash_head = 0;
    if (s.lookahead >= MIN_MATCH) hash_head = insert_string(sm, s.str_start, hash_variant);
    if (hash_head != 0 && (s.str_start - hash_head) <= MAX_DIST) {
        if (strategy != 2 /* HuffmanOnly */) s.match_length = longest_match(sm, s, cfg, hash_head);
    }
    bool bflush;
    if (s.match_length >= MIN_MATCH) {
        bflush = tally_dist(s, o, s.str_start - s.match_start, s.match_length - MIN_MATCH);
        s.lookahead -= s.match_length;
        if (s.match_length <= cfg.lazy && s.lookahead >= MIN_MATCH) {
            s.match_length--;
            do {
                s.str_start++;
                insert_string(sm, s.str_start, hash_variant);
            } while (--s.match_length != 0);
            s.str_start++;
        } else {
            s.str_start += s.match_length;
            s.match_length = 0;
        }
    } else {
        bflush = tally_lit(s, o, sm.win[s.str_start]);
        s.lookahead--;
        s.str_start++;
    }
    if (bflush) close_block(s, o);
}

// Common prefix length of the window strings at p and c, capped at MAX_MATCH (what Longest_match's pre-filters plus
// Compare256 amount to: a candidate replaces the best match iff this value exceeds best_len).
__device__ __forceinline__ int common_prefix258(const uint8_t* win, int p, int c)
{
    // eight bytes per step: three aligned words per side and two funnel shifts instead of two x (two words, one shift)
    int len = 0;
    while (len < MAX_MATCH + 2) {
        const uint32_t* qa = reinterpret_cast<const uint32_t*>(win + ((p + len) & ~3));
        const uint32_t* qb = reinterpret_cast<const uint32_t*>(win + ((c + len) & ~3));
        const int sa = ((p + len) & 3) * 8, sb = ((c + len) & 3) * 8;
        const uint32_t a0 = qa[0], a1 = qa[1], a2 = qa[2], b0 = qb[0], b1 = qb[1], b2 = qb[2];
        const uint32_t x0 = __funnelshift_r(a0, a1, sa) ^ __funnelshift_r(b0, b1, sb);
        if (x0) {
            len += (__ffs(x0) - 1) >> 3;
            break;
        }
        const uint32_t x1 = __funnelshift_r(a1, a2, sa) ^ __funnelshift_r(b1, b2, sb);
        if (x1) {
            len += 4 + ((__ffs(x1) - 1) >> 3);
            break;
        }
        len += 8;
    }
    return len < MAX_MATCH ? len : MAX_MATCH;
}

// 16 window bytes at any byte index (the array is padded), as four little-endian words.
__device__ __forceinline__ void lds16_unaligned(const uint8_t* base, int idx, uint32_t (&w)[4])
{
    const uint32_t* q = reinterpret_cast<const uint32_t*>(base + (idx & ~3));
    const uint32_t a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3], a4 = q[4];
    const int sh = (idx & 3) * 8;
    w[0] = __funnelshift_r(a0, a1, sh);
    w[1] = __funnelshift_r(a1, a2, sh);
    w[2] = __funnelshift_r(a2, a3, sh);
    w[3] = __funnelshift_r(a3, a4, sh);
}

// Common prefix of two 16-byte strings: 0..16.
__device__ __forceinline__ int prefix16(const uint32_t (&a)[4], const uint32_t (&b)[4])
{
    const uint32_t x0 = a[0] ^ b[0], x1 = a[1] ^ b[1], x2 = a[2] ^ b[2], x3 = a[3] ^ b[3];
    int L = 16;
    if (x3) L = 12 + ((__ffs(x3) - 1) >> 3);
    if (x2) L = 8 + ((__ffs(x2) - 1) >> 3);
    if (x1) L = 4 + ((__ffs(x1) - 1) >> 3);
    if (x0) L = (__ffs(x0) - 1) >> 3;
    return L;
}

// Whole warp: common prefix of the strings at p and c beyond their first `from` bytes (known equal), capped at MAX_MATCH.
__device__ __forceinline__ int warp_extend_match(const uint8_t* win, int p, int c, int from, int lane)
{
    int len = from;
    while (len < MAX_MATCH) {
        const int off = len + 4 * lane;
        const uint32_t x = lds_u32_unaligned(win, p + off) ^ lds_u32_unaligned(win, c + off);
        const uint32_t ne = __ballot_sync(0xffffffffu, x != 0);
        if (ne) {
            const int first = __ffs(ne) - 1;
            const uint32_t xf = __shfl_sync(0xffffffffu, x, first);
            len += 4 * first + ((__ffs(xf) - 1) >> 3);
            break;
        }
        len += 128;
    }
    return len < MAX_MATCH ? len : MAX_MATCH;
}

// DeflateFast (Deflate.Fast.cs:20-128) by FAST_WARPS warps: FAST_WARPS * 32 consecutive positions are evaluated
// speculatively as parse points (thread i = str_start + i) against the hash table as it stands before the batch; a
// mask walk then resolves the greedy parse, and the inserts of the positions the reference would have inserted are
// committed in position order.  A position whose hash also occurs earlier in the same batch cannot be speculated on
// (its chain would start inside the batch): when the parse reaches one, the batch ends there and the next one starts
// at it.  A position the Q3 pre-insert touched (head[h] >= position) runs the reference iteration literally.
// "Occurs earlier in the batch" is decided through a small slot table (atomicMax of a batch-sequence | index key per
// hash slot; __match_any_sync costs ~12 cycles per distinct value on sm_100, 385 cycles for 32 hashes): a thread is
// clean iff it is the first of the batch in its slot, which errs only towards "not clean".
// The warps exchange six ballots and one u16 per position through shared memory and each of them runs the (uniform)
// walk on its own, so three CTA-partial barriers per batch are all the synchronisation there is.
// Runs until the window needs refilling / sliding or the segment is done.
constexpr int FAST_WARPS = 2;
constexpr int FAST_NB = FAST_WARPS * 32;

__device__ __forceinline__ void fast_bar()
{
    if (FAST_WARPS == 1) __syncwarp();
    else asm volatile("bar.sync 1, %0;" ::"n"(FAST_NB) : "memory");
}

__device__ void parse_fast_multi(const Smem& sm, ParserState& s, ParserState* shared_state, const SegOut& o,
                                 const LevelCfg& cfg, int strategy, int hash_variant, unsigned long long* stats)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, idx = threadIdx.x;   // idx < FAST_NB
    const uint64_t below = (1ull << idx) - 1ull;
    const bool quick = cfg.chain <= 4 && cfg.nice <= 16;  // level 1: the whole chain is probed in one round of loads
    long long st_t0 = 0, st_t1 = 0, st_t2 = 0, st_t3 = 0;
    uint32_t seq = *sm.batch_seq;
    for (;;) {
        if (s.lookahead < MIN_LOOKAHEAD) {
            if (fill_needed(s)) break;
            if (s.lookahead == 0) { s.done = 1; break; }
        }
        // positions str_start .. str_start + nact - 1 reach the top of the loop without triggering Fill_window
        int nact;
        if (s.avail_in > 0) nact = s.lookahead - (MIN_LOOKAHEAD - 1);
        else {
            const int top = s.str_start + s.lookahead;
            int first_trigger = top - (MIN_LOOKAHEAD - 1);
            if (first_trigger < WINDOW_SIZE - MIN_LOOKAHEAD) first_trigger = WINDOW_SIZE - MIN_LOOKAHEAD;
            nact = first_trigger - s.str_start;
            if (nact > s.lookahead) nact = s.lookahead;
        }
        if (nact > FAST_NB) nact = FAST_NB;
        if (nact > BLOCK_SYMS - s.last_lit) nact = BLOCK_SYMS - s.last_lit;
        const uint64_t act = nact >= 64 ? ~0ull : (1ull << nact) - 1ull;

        const int S = s.str_start, LA = s.lookahead;
        const int p = S + idx, la = LA - idx;
        if (stats) st_t0 = clock64();
        // ---- speculative phase ----
        const bool hashed = la >= MIN_MATCH;
        uint32_t h = 0;
        int c0 = 0;
        uint32_t scan[4];
        lds16_unaligned(sm.win, p, scan);
        bool first_in_slot = true;
        seq++;
        const uint32_t key = (seq << 6) | (uint32_t)(63 - idx);
        if (hashed) {
            h = hash_of(sm, __funnelshift_r(scan[0], scan[1], 16), hash_variant);
            atomicMax(&sm.slot[h & sm.slot_mask], key);
            c0 = sm.head[h];
        }
        fast_bar();
        if (hashed) first_in_slot = sm.slot[h & sm.slot_mask] == key;
        const bool hard = hashed && c0 >= p;                 // Q2/Q3: the refill pre-inserted this position (or the next)
        const bool dirty = hashed && (!first_in_slot || hard);
        int best = MIN_MATCH - 1, mstart = 0;
        bool ext = false;  // best == 16 stands for "at least 16": the exact length is found if the position is a parse point
        if (hashed && !dirty && idx < nact && c0 != 0 && p - c0 <= MAX_DIST && strategy != 2) {
            const int nice = cfg.nice < la ? cfg.nice : la;
            const int limit = p > MAX_DIST ? p - MAX_DIST : 0;
            if (quick) {
                // Longest_match (Deflate.cs:1022-1100) with the <= 4 chain links fetched up front and their windows
                // compared side by side; the selection below visits them in the reference's order.
                const int c1 = sm.prev[c0 & W_MASK];
                const int c2 = sm.prev[c1 & W_MASK];
                const int c3 = sm.prev[c2 & W_MASK];
                uint32_t w0[4], w1[4], w2[4], w3[4];
                lds16_unaligned(sm.win, c0, w0);
                lds16_unaligned(sm.win, c1, w1);
                lds16_unaligned(sm.win, c2, w2);
                lds16_unaligned(sm.win, c3, w3);
                const int L0 = prefix16(scan, w0), L1 = prefix16(scan, w1), L2 = prefix16(scan, w2), L3 = prefix16(scan, w3);
                bool go = c0 < p;
                if (go && L0 > best) { best = L0; mstart = c0; if (L0 >= nice) go = false; }
                go = go && c1 > limit && cfg.chain > 1 && c1 < p;
                if (go && L1 > best) { best = L1; mstart = c1; if (L1 >= nice) go = false; }
                go = go && c2 > limit && cfg.chain > 2 && c2 < p;
                if (go && L2 > best) { best = L2; mstart = c2; if (L2 >= nice) go = false; }
                go = go && c3 > limit && cfg.chain > 3 && c3 < p;
                if (go && L3 > best) { best = L3; mstart = c3; }
                ext = best == 16;
            } else {
                int chain = cfg.chain;
                int cur = c0;
                do {
                    if (cur >= p) break;  // Deflate.cs:1067-1070 (reachable through the Q3 two-cycle)
                    uint32_t cw[4];
                    lds16_unaligned(sm.win, cur, cw);
                    const int nxt = sm.prev[cur & W_MASK];
                    int L = prefix16(scan, cw);
                    if (L == 16 && nice > 16) {  // nice_match exceeds the probe (level 3): finish the compare here
                        L = common_prefix258(sm.win, p, cur);
                        if (L > best) { best = L; mstart = cur; if (L >= nice) break; }
                    } else if (L > best) {
                        best = L;
                        mstart = cur;
                        if (L == 16) { ext = true; break; }
                        if (L >= nice) break;
                    }
                    cur = nxt;
                } while (cur > limit && --chain != 0);
            }
        }
        int len = best < la ? best : la;
        const bool is_match = len >= MIN_MATCH;
        {
            const uint32_t b0 = __ballot_sync(0xffffffffu, is_match), b1 = __ballot_sync(0xffffffffu, dirty);
            const uint32_t b2 = __ballot_sync(0xffffffffu, ext), b3 = __ballot_sync(0xffffffffu, hard);
            const uint32_t b4 = __ballot_sync(0xffffffffu, hashed), b5 = __ballot_sync(0xffffffffu, first_in_slot && !hard);
            if (lane == 0) {
                uint32_t* m = sm.fast_masks + warp * 8;
                m[0] = b0; m[1] = b1; m[2] = b2; m[3] = b3; m[4] = b4; m[5] = b5;
            }
            sm.fast_endl[idx] = (uint16_t)(idx + (is_match ? len : 1));   // index of the next parse point if this is one
            sm.fast_mst[idx] = (uint16_t)mstart;
        }
        fast_bar();
        uint64_t Mm, Dm, Xm, Hd, H3, Fm;
        {
            const uint32_t* m = sm.fast_masks;
            Mm = m[0]; Dm = m[1]; Xm = m[2]; Hd = m[3]; H3 = m[4]; Fm = m[5];
            if (FAST_WARPS > 1) {
                Mm |= (uint64_t)m[8] << 32; Dm |= (uint64_t)m[9] << 32; Xm |= (uint64_t)m[10] << 32;
                Hd |= (uint64_t)m[11] << 32; H3 |= (uint64_t)m[12] << 32; Fm |= (uint64_t)m[13] << 32;
            }
            Mm &= act;
            Dm &= act;
        }
        if (stats) st_t1 = clock64();
        // ---- resolve the greedy parse (every warp on its own, the walk is uniform): hop from match to match; the
        // literals in between are parse points; lanes under a taken match are not, but the interior of a short one
        // is still inserted (Deflate.Fast.cs:81-104) ----
        int cur = 0;
        uint64_t vis = 0, cover = 0, shortins = 0;
        int tail_last = 0;         // last interior position of a short match that runs past the batch
        bool hit_dirty = false;
        const uint64_t stops = Mm | Dm;
        for (;;) {
            const uint64_t stop = stops & (~0ull << cur);
            if (stop == 0) { cur = nact; break; }
            const int nxt = __ffsll((long long)stop) - 1;
            if ((Dm >> nxt) & 1ull) { cur = nxt; hit_dirty = true; break; }
            int e = sm.fast_endl[nxt];
            if ((Xm >> nxt) & 1ull) {  // long match at a parse point: the warp finishes the compare
                const int full = warp_extend_match(sm.win, S + nxt, sm.fast_mst[nxt], 16, lane);
                const int lx = full < LA - nxt ? full : LA - nxt;
                e = nxt + lx;
                if (idx == nxt) len = lx;
            }
            const int mlen = e - nxt;
            vis |= 1ull << nxt;
            const uint64_t rng = (e >= 64 ? ~0ull : (1ull << e) - 1ull) & ~((2ull << nxt) - 1ull);
            cover |= rng;
            if (mlen <= cfg.lazy && LA - nxt - mlen >= MIN_MATCH) {
                shortins |= rng;
                if (e > FAST_NB) tail_last = S + e - 1;
            }
            cur = e;
            if (cur >= nact) break;
        }
        const uint64_t done_mask = cur >= 64 ? ~0ull : (1ull << cur) - 1ull;
        const uint64_t lanes_mask = FAST_NB >= 64 ? ~0ull : (1ull << FAST_NB) - 1ull;
        const uint64_t pm = done_mask & ~cover & lanes_mask;             // parse points
        const uint64_t im = ((pm & H3) | shortins) & lanes_mask;         // inserted positions
        if (stats) st_t2 = clock64();
        // ---- symbols (Tr_tally_*): parse points in position order ----
        const int nsym = __popcll(pm);
        if ((pm >> idx) & 1ull) {
            const uint32_t v = ((vis >> idx) & 1ull) ? ((uint32_t)(p - mstart) | ((uint32_t)(len - MIN_MATCH) << 16))
                                                     : ((scan[0] & 0xFFu) << 16);
            o.sym[s.sym_total + __popcll(pm & below)] = v;
        }
        // ---- commit InsertString of every inserted position, in position order ----
        // Threads that are the first of their hash slot go side by side.  The others (a later position of a slot already
        // seen in the batch) follow one at a time, as the reference would execute them.
        const uint64_t serial_ins = im & ~Fm;
        if (((im & ~serial_ins) >> idx) & 1ull) {
            sm.prev[p & W_MASK] = (uint16_t)c0;
            sm.head[h] = (uint16_t)p;
        }
        if (serial_ins != 0) {
            for (uint64_t m = serial_ins; m != 0; m &= m - 1) {
                fast_bar();
                if (idx == __ffsll((long long)m) - 1) {
                    const int curh = sm.head[h];
                    if (curh != p) {   // Q2
                        sm.prev[p & W_MASK] = (uint16_t)curh;
                        sm.head[h] = (uint16_t)p;
                    }
                }
            }
        }
        s.sym_total += nsym;
        s.last_lit += nsym;
        s.str_start = S + cur;
        s.lookahead = LA - cur;
        s.match_length = 0;
        if (tail_last != 0) {   // the interior of a short match runs past the batch
            fast_bar();
            if (idx == 0)
                for (int q = S + FAST_NB; q <= tail_last; q++) insert_string(sm, q, hash_variant);
        }
        if (s.last_lit == BLOCK_SYMS) {
            ParserState t = s;
            if (idx == 0) close_block(t, o);
            else { t.nblocks++; t.last_lit = 0; t.block_start = t.str_start; }
            s = t;
        }
        if (stats) st_t3 = clock64();
        const bool serial = hit_dirty && ((Hd >> cur) & 1ull);
        fast_bar();   // the commits above are visible to every warp before the next batch reads head / prev
        if (serial) {
            if (idx == 0) {
                fast_step(sm, s, o, cfg, strategy, hash_variant);
                *shared_state = s;
            }
            fast_bar();
            s = *shared_state;
            fast_bar();
        }
        if (stats && idx == 0) {
            const long long t4 = clock64();
            atomicAdd(stats + 0, 1ull);
            atomicAdd(stats + 1, (unsigned long long)cur);
            atomicAdd(stats + 2, hit_dirty ? 1ull : 0ull);
            atomicAdd(stats + 3, (unsigned long long)nact);
            atomicAdd(stats + 4, (unsigned long long)(st_t1 - st_t0));
            atomicAdd(stats + 5, (unsigned long long)(st_t2 - st_t1));
            atomicAdd(stats + 6, (unsigned long long)(st_t3 - st_t2));
            atomicAdd(stats + 7, (unsigned long long)(t4 - st_t3));
            atomicAdd(stats + 8, (unsigned long long)__popcll(vis));
            atomicAdd(stats + 9, (unsigned long long)__popcll(Dm));
            atomicAdd(stats + 12, (unsigned long long)__popcll(Xm & vis));
            atomicAdd(stats + 13, serial ? 1ull : 0ull);
            atomicAdd(stats + 14, (unsigned long long)__popcll(serial_ins));
        }
    }
    if (idx == 0) *sm.batch_seq = seq;
    if (s.done) {
        if (idx == 0) close_block(s, o);  // the end-of-input Flush_block_only(false) (Deflate.Fast.cs:124)
        else { s.nblocks++; s.last_lit = 0; s.block_start = s.str_start; }
    }
}

// Deflate.Slow.cs:18-159
__device__ void parse_slow(const Smem& sm, ParserState& s, const SegOut& o, const LevelCfg& cfg, int strategy,
                           int hash_variant)
{
    int hash_head = 0;  // Q8: stale value only matters when lookahead < 3, where it cannot change the output
    for (;;) {
        if (s.lookahead < MIN_LOOKAHEAD) {
            if (fill_needed(s)) return;
            if (s.lookahead == 0) break;
        }
        if (s.lookahead >= MIN_MATCH) hash_head = insert_string(sm, s.str_start, hash_variant);
        s.prev_length = s.match_length;
        s.prev_match = s.match_start;
        s.match_length = MIN_MATCH - 1;
        if (hash_head != 0 && s.prev_length < cfg.lazy && s.str_start - hash_head <= MAX_DIST) {
            if (strategy != 2) s.match_length = longest_match(sm, s, cfg, hash_head);
            if (s.match_length <= 5 &&
                (strategy == 1 /* Filtered */ || (s.match_length == MIN_MATCH && s.str_start - s.match_start > 4096)))
                s.match_length = MIN_MATCH - 1;
        }
        if (s.prev_length >= MIN_MATCH && s.match_length <= s.prev_length) {
            const int max_insert = s.str_start + s.lookahead - MIN_MATCH;
            const bool bflush = tally_dist(s, o, s.str_start - 1 - s.prev_match, s.prev_length - MIN_MATCH);
            s.lookahead -= s.prev_length - 1;
            s.prev_length -= 2;
            do {
                if (++s.str_start <= max_insert) hash_head = insert_string(sm, s.str_start, hash_variant);
            } while (--s.prev_length != 0);
            s.match_available = 0;
            s.match_length = MIN_MATCH - 1;
            s.str_start++;
            if (bflush) close_block(s, o);
        } else if (s.match_available != 0) {
            if (tally_lit(s, o, sm.win[s.str_start - 1])) close_block(s, o);
            s.str_start++;
            s.lookahead--;
        } else {
            s.match_available = 1;
            s.str_start++;
            s.lookahead--;
        }
    }
    if (s.match_available != 0) {
        // `_ = Tr_tally_lit(...)`: the flush flag is ignored, the block is flushed right after anyway
        o.sym[s.sym_total++] = (uint32_t)sm.win[s.str_start - 1] << 16;
        s.last_lit++;
        s.match_available = 0;
    }
    close_block(s, o);
    s.done = 1;
}

// ---- DeflateSlow by tiles (levels 4-9) ---------------------------------------------------------------------
// The lazy parser inserts EVERY position exactly once and in position order (Deflate.Slow.cs:50-53,98-105), so the
// hash chains do not depend on the parse.  A tile of up to 256 positions (all with lookahead >= MIN_LOOKAHEAD, so no
// Fill_window can intervene; 256 < MIN_LOOKAHEAD keeps every prev[] slot a search may still read intact) is handled
// in three steps: (A) one warp inserts the tile's positions in order and keeps what InsertString returned for each,
// (B) all threads run Longest_match for one position each, (C) one thread replays DeflateSlow with the results.
// Longest_match depends on the parse only through prev_length: it starts from best_len = prev_length and visits
// chain >> 2 candidates instead of chain when prev_length >= good_match.  Its result is therefore the longest
// candidate among the first `budget` ones if that beats prev_length, and (B) keeps that pair for both budgets.

__device__ void slow_tile_insert(const Smem& sm, int S, int T, int hash_variant)   // (A), warp 0
{
    const int lane = threadIdx.x & 31;
    uint32_t seq = *sm.batch_seq;
    for (int sub = 0; sub < T; sub += 32) {
        const int i = sub + lane;
        const bool on = i < T;
        const int p = S + i;
        uint32_t h = 0;
        int c0 = 0;
        seq++;
        const uint32_t key = (seq << 5) | (uint32_t)(31 - lane);
        if (on) {
            h = hash_of(sm, lds_u32_unaligned(sm.win, p + (MIN_MATCH - 1)), hash_variant);
            atomicMax(&sm.slot[h & sm.slot_mask], key);
            c0 = sm.head[h];
        }
        __syncwarp();
        const bool first = on && sm.slot[h & sm.slot_mask] == key;
        if (first) {   // the only position of its hash so far in this group of 32: InsertString (Deflate.cs:866-877)
            if (c0 != p) {
                sm.prev[p & W_MASK] = (uint16_t)c0;
                sm.head[h] = (uint16_t)p;
            }
            sm.tile_hh[i] = (uint16_t)c0;
        }
        __syncwarp();
        for (uint32_t m = __ballot_sync(0xffffffffu, on && !first); m != 0; m &= m - 1) {
            if (lane == __ffs(m) - 1) {
                const int cur = sm.head[h];
                if (cur != p) {
                    sm.prev[p & W_MASK] = (uint16_t)cur;
                    sm.head[h] = (uint16_t)p;
                }
                sm.tile_hh[i] = (uint16_t)cur;
            }
            __syncwarp();
        }
    }
    if (lane == 0) *sm.batch_seq = seq;
}

// (B): Longest_match (Deflate.cs:1022-1100) at window position p from hash_head, for both chain budgets.
// Returns (length | match_start << 9) for the full budget in .x and for chain >> 2 in .y; length 2 = nothing found.
__device__ uint2 slow_tile_search(const Smem& sm, int p, int hash_head, const LevelCfg& cfg)
{
    int best = MIN_MATCH - 1, ms = 0;
    uint32_t early = 0;
    bool have_early = false;
    if (hash_head != 0 && p - hash_head <= MAX_DIST) {
        const int limit = p > MAX_DIST ? p - MAX_DIST : 0;
        const int early_budget = cfg.chain >> 2;
        int chain = cfg.chain, j = 0, cur = hash_head;
        do {
            if (cur >= p) break;
            j++;
            const int off = best >= 3 ? best - 3 : 0;
            const uint32_t mask = best >= 3 ? 0xffffffffu : 0x00ffffffu;
            const int nxt = sm.prev[cur & W_MASK];
            if (((lds_u32_unaligned(sm.win, cur + off) ^ lds_u32_unaligned(sm.win, p + off)) & mask) == 0) {
                const int L = common_prefix258(sm.win, p, cur);
                if (L > best) {
                    best = L;
                    ms = cur;
                    if (L >= cfg.nice) break;
                }
            }
            if (j == early_budget) {
                early = (uint32_t)best | ((uint32_t)ms << 9);
                have_early = true;
            }
            cur = nxt;
        } while (cur > limit && --chain != 0);
    }
    const uint32_t full = (uint32_t)best | ((uint32_t)ms << 9);
    return make_uint2(full, have_early ? early : full);
}

// (A) for the wide pipelined tile, by all 992 producer threads.  InsertString in position order means: position p gets
// (as return value and as prev link) the nearest earlier position of the tile with the same hash, or the head as it stood
// before the tile if there is none, and head[h] ends up at the tile's last position with hash h.  So: sort
// (hash << 10 | index) in shared memory (bitonic, 1024 keys; the tile_rec buffer of this tile is still free), then every
// position finds its predecessor next to it.  Q2 is kept (no prev write when the value found is the position itself, which
// happens for the position a refill pre-inserted, Q3), and the old content of every prev slot goes to the shadow first.
// Returns (my_pos, my_head) = the position of sorted rank ptid and its predecessor, my_pos < 0 if there is none: the caller
// searches THAT position, so that the lanes of a warp walk chains of the same or neighbouring hashes (similar lengths, the
// same prev[] entries and candidates) instead of 32 unrelated ones.
__device__ void slow_tile_insert_sorted(const Smem& sm, int S, int T, int hash_variant, int ptid, int& my_pos, int& my_head)
{
    uint32_t* keys = reinterpret_cast<uint32_t*>(sm.tile_rec);   // 1024 x u32 <= WIDE_TILE x 8 bytes
    // Bitonic sort of 1024 keys by the first 512 producer threads, two adjacent keys (2t, 2t + 1) per thread: partners at
    // distance 1 sit in the same thread, at distance 2..32 one shuffle away, only distance >= 64 goes through shared
    // memory (10 of the 55 stages).  The other producer threads wait at the barrier below.
    if (ptid < 512) {
        const int e0 = 2 * ptid;
        uint32_t k0 = 0xFFFFFFFFu, k1 = 0xFFFFFFFFu;
        if (e0 < T) k0 = (hash_of(sm, lds_u32_unaligned(sm.win, S + e0 + (MIN_MATCH - 1)), hash_variant) << 10) | (uint32_t)e0;
        if (e0 + 1 < T) k1 = (hash_of(sm, lds_u32_unaligned(sm.win, S + e0 + 1 + (MIN_MATCH - 1)), hash_variant) << 10) | (uint32_t)(e0 + 1);
        for (int k = 2; k <= 1024; k <<= 1) {
            const bool up = (e0 & k) == 0;                       // both keys of a thread lie in the same k-block (k >= 2)
            for (int j = k >> 1; j > 0; j >>= 1) {
                if (j == 1) {
                    if ((k0 > k1) == up) { const uint32_t x = k0; k0 = k1; k1 = x; }
                } else if (j <= 32) {
                    const uint32_t o0 = __shfl_xor_sync(0xffffffffu, k0, j >> 1), o1 = __shfl_xor_sync(0xffffffffu, k1, j >> 1);
                    const bool low = (e0 & j) == 0;              // this thread holds the lower element of each pair
                    // the lower element keeps the minimum in an ascending block, the maximum in a descending one
                    k0 = (low == up) ? min(k0, o0) : max(k0, o0);
                    k1 = (low == up) ? min(k1, o1) : max(k1, o1);
                } else {
                    keys[e0] = k0;
                    keys[e0 + 1] = k1;
                    asm volatile("bar.sync 4, 512;" ::: "memory");
                    const uint32_t o0 = keys[e0 ^ j], o1 = keys[(e0 + 1) ^ j];
                    asm volatile("bar.sync 4, 512;" ::: "memory");
                    const bool low = (e0 & j) == 0;
                    k0 = (low == up) ? min(k0, o0) : max(k0, o0);
                    k1 = (low == up) ? min(k1, o1) : max(k1, o1);
                }
            }
        }
        keys[e0] = k0;
        keys[e0 + 1] = k1;
    }
    asm volatile("bar.sync 3, 992;" ::: "memory");
    // predecessor of every position; all reads of the old head before any write
    int my_p[2] = {0, 0}, my_pred[2] = {0, 0}, my_h[2] = {0, 0};
    bool my_last[2] = {false, false}, my_on[2] = {false, false};
    int n = 0;
    for (int t = ptid; t < 1024; t += WIDE_TILE, n++) {
        const uint32_t key = keys[t];
        if (key == 0xFFFFFFFFu) continue;
        const int h = (int)(key >> 10), i = (int)(key & 1023u);
        const uint32_t before = t > 0 ? keys[t - 1] : 0xFFFFFFFFu, after = t + 1 < 1024 ? keys[t + 1] : 0xFFFFFFFFu;
        my_on[n] = true;
        my_h[n] = h;
        my_p[n] = S + i;
        my_pred[n] = (before != 0xFFFFFFFFu && (int)(before >> 10) == h) ? S + (int)(before & 1023u) : (int)sm.head[h];
        my_last[n] = after == 0xFFFFFFFFu || (int)(after >> 10) != h;
    }
    asm volatile("bar.sync 3, 992;" ::: "memory");
#pragma unroll
    for (int m = 0; m < 2; m++) {
        if (!my_on[m]) continue;
        const int p = my_p[m], pred = my_pred[m];
        sm.tile_hh[p - S] = (uint16_t)pred;
        if (sm.shadow) sm.shadow[p & (SHADOW_N - 1)] = sm.prev[p & W_MASK];
        if (pred != p) sm.prev[p & W_MASK] = (uint16_t)pred;
        if (my_last[m]) sm.head[my_h[m]] = (uint16_t)p;
    }
    my_pos = my_on[0] ? my_p[0] : -1;
    my_head = my_pred[0];
}

// (B) for the wide pipelined tile: slow_tile_search, except that a link whose prev[] slot has meanwhile been reused by a
// position of the same tile (32768 further on) is read from the shadow copy taken at that insert.
__device__ uint2 slow_tile_search_sh(const Smem& sm, int p, int hash_head, const LevelCfg& cfg)
{
    int best = MIN_MATCH - 1, ms = 0;
    uint32_t early = 0;
    bool have_early = false;
    if (hash_head != 0 && p - hash_head <= MAX_DIST) {
        const int limit = p > MAX_DIST ? p - MAX_DIST : 0;
        int cur = hash_head;
        // A candidate can only beat `best` if it agrees with the scan string at [best] and on the bytes up to it: one byte
        // load rejects most candidates, the scan side of both tests stays in registers until `best` changes.
        int off = 0;
        uint32_t mask = 0x00ffffffu;
        uint32_t scan_w = lds_u32_unaligned(sm.win, p);
        uint32_t scan_b = sm.win[p + best];
        const uint32_t sh_lo = (uint32_t)sm.shadow_lo, sh_n = (uint32_t)(sm.shadow_hi - sm.shadow_lo);
        // Two passes over the same loop body: the first chain >> 2 steps (after which the result for the reduced budget
        // is noted), then the rest -- the step counter of Longest_match's `chain_length` is the loop bound itself.
        bool stop = false;
#pragma unroll 1
        for (int pass = 0; pass < 2 && !stop; pass++) {
            int steps = pass == 0 ? (cfg.chain >> 2) : cfg.chain - (cfg.chain >> 2);
            for (;;) {
                if (cur >= p) { stop = true; break; }
                int nxt = sm.prev[cur & W_MASK];
                if ((uint32_t)(cur + W_SIZE) - sh_lo < sh_n) nxt = sm.shadow[(cur + W_SIZE) & (SHADOW_N - 1)];   // slot reused inside this tile
                if (sm.win[cur + best] == scan_b && ((lds_u32_unaligned(sm.win, cur + off) ^ scan_w) & mask) == 0) {
                    const int L = common_prefix258(sm.win, p, cur);
                    if (L > best) {
                        best = L;
                        ms = cur;
                        if (L >= cfg.nice) { stop = true; break; }
                        off = best - 3;
                        mask = 0xffffffffu;
                        scan_w = lds_u32_unaligned(sm.win, p + off);
                        scan_b = sm.win[p + best];
                    }
                }
                cur = nxt;
                --steps;
                if (!(cur > limit)) { stop = true; if (steps == 0 && pass == 0) { early = (uint32_t)best | ((uint32_t)ms << 9); have_early = true; } break; }
                if (steps == 0) break;
            }
            if (pass == 0 && !stop) {
                early = (uint32_t)best | ((uint32_t)ms << 9);
                have_early = true;
            }
        }
    }
    const uint32_t full = (uint32_t)best | ((uint32_t)ms << 9);
    return make_uint2(full, have_early ? early : full);
}

// Before the replay inserts a position itself (one past the pre-inserted region), every producer tile must be done:
// InsertString calls have to happen in position order.
__device__ __forceinline__ void pipe_wait_all(const Smem& sm)
{
    if (sm.pipe_flags) {
        while (sm.pipe_flags[0] < sm.pipe_ntiles) __nanosleep(100);
        __threadfence_block();
    }
}

// (C): DeflateSlow (Deflate.Slow.cs:18-159) over the tile [S, S + T), Longest_match / InsertString looked up.
__device__ void slow_tile_replay(const Smem& sm, ParserState& s, const SegOut& o, const LevelCfg& cfg, int strategy,
                                 int hash_variant, int S, int T, int until)
{
    const int tile_end = S + T;
    int hash_head = 0;
    while (s.str_start < until) {
        const int idx = s.str_start - S;
        hash_head = sm.tile_hh[idx];
        s.prev_length = s.match_length;
        s.prev_match = s.match_start;
        s.match_length = MIN_MATCH - 1;
        if (hash_head != 0 && s.prev_length < cfg.lazy && s.str_start - hash_head <= MAX_DIST) {
            if (strategy != 2) {
                const uint2 r = sm.tile_rec[idx];
                const uint32_t e = s.prev_length >= cfg.good ? r.y : r.x;
                int best = s.prev_length;
                if ((int)(e & 511u) > best) {
                    best = (int)(e & 511u);
                    s.match_start = (int)(e >> 9);
                }
                s.match_length = best < s.lookahead ? best : s.lookahead;
            }
            if (s.match_length <= 5 &&
                (strategy == 1 /* Filtered */ || (s.match_length == MIN_MATCH && s.str_start - s.match_start > 4096)))
                s.match_length = MIN_MATCH - 1;
        }
        if (s.prev_length >= MIN_MATCH && s.match_length <= s.prev_length) {
            const int max_insert = s.str_start + s.lookahead - MIN_MATCH;
            const bool bflush = tally_dist(s, o, s.str_start - 1 - s.prev_match, s.prev_length - MIN_MATCH);
            s.lookahead -= s.prev_length - 1;
            s.prev_length -= 2;
            do {
                if (++s.str_start <= max_insert)
                {
                    if (s.str_start < tile_end) hash_head = (int)sm.tile_hh[s.str_start - S];
                    else if (s.str_start < sm.pipe_ins_limit) hash_head = 0;   // a producer warp inserts it (value unused: the next visit reloads it)
                    else {
                        pipe_wait_all(sm);
                        hash_head = insert_string(sm, s.str_start, hash_variant);
                    }
                }
            } while (--s.prev_length != 0);
            s.match_available = 0;
            s.match_length = MIN_MATCH - 1;
            s.str_start++;
            if (bflush) close_block(s, o);
        } else if (s.match_available != 0) {
            if (tally_lit(s, o, sm.win[s.str_start - 1])) close_block(s, o);
            s.str_start++;
            s.lookahead--;
        } else {
            s.match_available = 1;
            s.str_start++;
            s.lookahead--;
        }
    }
}

__device__ __forceinline__ void bcast_state(ParserState& s)
{
    int* f = reinterpret_cast<int*>(&s);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(ParserState) / sizeof(int)); k++) f[k] = __shfl_sync(0xffffffffu, f[k], 0);
}

// (C) by one warp: the DeflateSlow state machine (Deflate.Slow.cs:40-146) over 32 positions at a time, lane = one visit
// (one loop iteration of the reference).  While no match is pending (match_length == 2) every position whose search
// finds nothing just emits the previous byte as a literal, so those runs are taken in one step with a ballot; around a
// match the lazy evaluation is stepped visit by visit, each lane looking up its own Longest_match result for the
// pending prev_length.  Every visit emits at most one symbol: they are stored in lane order.
__device__ void slow_tile_replay_warp(const Smem& sm, ParserState& s, const SegOut& o, const LevelCfg& cfg, int strategy,
                                      int hash_variant, int S, int T)
{
    const int lane = threadIdx.x & 31;
    const uint32_t below = (1u << lane) - 1u;
    const int tile_end = S + T;
    int ins_done_to = tile_end > sm.pipe_ins_limit ? tile_end : sm.pipe_ins_limit;   // positions >= tile_end below this one have been (or will be, by a producer) inserted
    while (s.str_start < tile_end) {
        const int g0 = s.str_start;
        const int gn = tile_end - g0 < 32 ? tile_end - g0 : 32;
        if (s.last_lit + 32 >= BLOCK_SYMS) {   // a block ends inside this stretch: the reference loop, literally
            if (lane == 0) slow_tile_replay(sm, s, o, cfg, strategy, hash_variant, S, T, g0 + gn);
            bcast_state(s);
            if (s.str_start > ins_done_to) ins_done_to = s.str_start;
            continue;
        }
        const int p = g0 + lane;
        const bool valid = lane < gn;
        int hh = 0;
        uint32_t full = MIN_MATCH - 1, early = MIN_MATCH - 1;
        if (valid) {
            hh = sm.tile_hh[p - S];
            const uint2 r = sm.tile_rec[p - S];
            full = r.x;
            early = r.y;
        }
        const uint32_t prev_byte = p > 0 ? sm.win[p - 1] : 0u;
        const bool searchable = valid && hh != 0 && p - hh <= MAX_DIST && strategy != 2;
        const int fullL = (int)(full & 511u), fullMs = (int)(full >> 9);
        // what the visit yields when nothing is pending (prev_length == 2)
        int D = MIN_MATCH - 1;
        if (searchable && fullL > D) {
            D = fullL;
            if (D <= 5 && (strategy == 1 || (D == MIN_MATCH && p - fullMs > 4096))) D = MIN_MATCH - 1;
        }
        const uint32_t NQ = __ballot_sync(0xffffffffu, D >= MIN_MATCH);
        uint32_t litmask = 0, matmask = 0;
        int my_len = 0, my_dist = 0;
        int cur = 0;
        int mlp = s.match_length, msp = s.match_start, avail = s.match_available;
        while (cur < gn) {
            if (mlp < MIN_MATCH) {
                const uint32_t rem = NQ & (0xffffffffu << cur);
                const int nxt = rem ? __ffs(rem) - 1 : gn;
                if (nxt > cur) {   // visits cur .. nxt-1 find nothing: literals of the bytes before them
                    uint32_t run = (nxt >= 32 ? 0xffffffffu : (1u << nxt) - 1u) & (0xffffffffu << cur);
                    if (!avail) run &= ~(1u << cur);
                    litmask |= run;
                    avail = 1;
                }
                if (nxt >= gn) { cur = gn; break; }
                if (avail) litmask |= 1u << nxt;
                avail = 1;
                mlp = __shfl_sync(0xffffffffu, D, nxt);
                msp = __shfl_sync(0xffffffffu, fullMs, nxt);
                cur = nxt + 1;
            } else {
                // visit `cur` with prev_length = mlp pending: Longest_match from best_len = prev_length, chain >> 2 if >= good
                int ml = MIN_MATCH - 1, msn = msp;
                if (searchable && mlp < cfg.lazy) {
                    const uint32_t e = mlp >= cfg.good ? early : full;
                    int best = mlp;
                    if ((int)(e & 511u) > best) {
                        best = (int)(e & 511u);
                        msn = (int)(e >> 9);
                    }
                    ml = best;
                    if (ml <= 5 && (strategy == 1 || (ml == MIN_MATCH && p - msn > 4096))) ml = MIN_MATCH - 1;
                }
                const int ml_c = __shfl_sync(0xffffffffu, ml, cur);
                const int ms_c = __shfl_sync(0xffffffffu, msn, cur);
                if (ml_c <= mlp) {   // the pending match stands: emitted at this visit, the parse skips to its end
                    matmask |= 1u << cur;
                    if (lane == cur) {
                        my_len = mlp;
                        my_dist = (g0 + cur - 1) - msp;
                    }
                    cur += mlp - 1;
                    avail = 0;
                    mlp = MIN_MATCH - 1;
                } else {
                    litmask |= 1u << cur;
                    mlp = ml_c;
                    cur++;
                }
                msp = ms_c;
            }
        }
        const uint32_t emit = litmask | matmask;
        if ((emit >> lane) & 1u) {
            const uint32_t v = ((matmask >> lane) & 1u) ? ((uint32_t)my_dist | ((uint32_t)(my_len - MIN_MATCH) << 16)) : (prev_byte << 16);
            o.sym[s.sym_total + __popc(emit & below)] = v;
        }
        const int nsym = __popc(emit);
        s.sym_total += nsym;
        s.last_lit += nsym;
        s.str_start = g0 + cur;
        s.lookahead -= cur;
        s.match_length = mlp;
        s.match_start = msp;
        s.match_available = avail;
    }
    if (s.str_start > ins_done_to) {   // positions past the tile that the last match skipped over (Deflate.Slow.cs:98-105)
        if (lane == 0) {
            pipe_wait_all(sm);
            for (int q = ins_done_to; q < s.str_start; q++) insert_string(sm, q, hash_variant);
        }
        __syncwarp();
    }
}

// Deflate.Rle.cs:18-104
__device__ void parse_rle(const Smem& sm, ParserState& s, const SegOut& o)
{
    for (;;) {
        if (s.lookahead <= MAX_MATCH) {
            if (fill_needed(s)) return;
        }
        if (s.lookahead == 0) break;
        s.match_length = 0;
        if (s.lookahead >= MIN_MATCH && s.str_start > 0) {  // Q14: window[strStart-1] may be the previous segment's last byte
            int scan = s.str_start - 1;
            const uint32_t prev = sm.win[scan];
            if (prev == sm.win[scan + 1] && prev == sm.win[scan + 2] && prev == sm.win[scan + 3]) {
                const int strend = s.str_start + MAX_MATCH;
                scan += 3;
                for (;;) {  // the reference's 8-way unrolled scan: stops at a mismatch or once scan >= strend
                    int k;
                    bool stop = false;
                    for (k = 0; k < 8; k++) {
                        if (prev != sm.win[++scan]) { stop = true; break; }
                    }
                    if (stop || !(scan < strend)) break;
                }
                s.match_length = MAX_MATCH - (strend - scan);
                if (s.match_length > s.lookahead) s.match_length = s.lookahead;
            }
        }
        bool bflush;
        if (s.match_length >= MIN_MATCH) {
            bflush = tally_dist(s, o, 1, s.match_length - MIN_MATCH);
            s.lookahead -= s.match_length;
            s.str_start += s.match_length;
            s.match_length = 0;
        } else {
            bflush = tally_lit(s, o, sm.win[s.str_start]);
            s.lookahead--;
            s.str_start++;
        }
        if (bflush) close_block(s, o);
    }
    close_block(s, o);
    s.done = 1;
}

}  // namespace

// NT = 256: every parser (DeflateFast needs the registers).  NT = 1024, SLOW_ONLY: levels 4-9 as a pipeline over wide
// tiles (insert / search / replay on different warps, one searched position per producer thread).
template <int NT, bool SLOW_ONLY>
__global__ void __launch_bounds__(NT, 1) lz77_segments_kernel(Lz77Params P)
{
    constexpr int LZ_THREADS = NT;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Smem sm;
    sm.win = smem_raw;
    sm.head = reinterpret_cast<uint16_t*>(smem_raw + WINDOW_SIZE + WIN_PAD);
    sm.prev = sm.head + HASH_SIZE;
    sm.crc = sm.prev + W_SIZE;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm.crc + 1024);
    uint32_t* red = reinterpret_cast<uint32_t*>(bar + 1);  // 2 * 8 * u64
    constexpr int REC_N = NT == 1024 ? 2 * WIDE_TILE : TILE_MAX;      // 1024-thread kernel: two wide tile buffers
    constexpr int SLOTS = NT == 1024 ? WIDE_SLOTS : SLOT_COUNT;
    sm.tile_rec = reinterpret_cast<uint2*>(red + 128);
    sm.slot = reinterpret_cast<uint32_t*>(sm.tile_rec + REC_N);
    sm.slot_mask = SLOTS - 1;
    sm.tile_hh = reinterpret_cast<uint16_t*>(sm.slot + SLOTS);
    uint16_t* shadow_buf = sm.tile_hh + REC_N;                         // SHADOW_N entries (1024-thread kernel only)
    sm.shadow = nullptr;
    sm.shadow_lo = sm.shadow_hi = 0;
    sm.fast_masks = reinterpret_cast<uint32_t*>(shadow_buf + (NT == 1024 ? SHADOW_N : 0));
    sm.fast_endl = reinterpret_cast<uint16_t*>(sm.fast_masks + 16);
    sm.fast_mst = sm.fast_endl + 64;
    ParserState* shared_state = reinterpret_cast<ParserState*>(sm.fast_mst + 64);
    int* seg_slot = reinterpret_cast<int*>(shared_state + 1);
    sm.batch_seq = reinterpret_cast<uint32_t*>(seg_slot + 1);
    volatile int* pipe_flags = reinterpret_cast<volatile int*>(sm.batch_seq + 1);   // 2 ints
    sm.pipe_ins_limit = 0;
    sm.pipe_ntiles = 0;
    sm.pipe_flags = nullptr;

    const int tid = threadIdx.x;
    for (int i = tid; i < 1024; i += LZ_THREADS) sm.crc[i] = P.crc_tab[i];
    for (int i = tid; i < SLOTS; i += LZ_THREADS) sm.slot[i] = 0;
    if (tid == 0) *sm.batch_seq = 0;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    uint32_t phase = 0;

    LevelCfg cfg;
    cfg.good = P.good; cfg.lazy = P.lazy; cfg.nice = P.nice; cfg.chain = P.chain; cfg.fn = P.fn;
    if (SLOW_ONLY) cfg.fn = FN_SLOW;

    for (;;) {
        if (tid == 0) *seg_slot = (int)atomicAdd(P.work_counter, 1u);
        __syncthreads();
        const int seg = *seg_slot;
        __syncthreads();
        if (seg >= (int)P.nseg) break;

        const size_t seg_off = (size_t)seg * P.seg_size;
        const int seg_len = (int)((P.n - seg_off) < (size_t)P.seg_size ? (P.n - seg_off) : (size_t)P.seg_size);
        const uint8_t* seg_in = P.in + seg_off;
        const bool stream_first = (P.first_seg_index + (uint64_t)seg) == 0;

        // ---- per-segment init: head zeroed by the previous full flush (Deflate.cs:596-605) ----
        // A SYNC flush (P.chained, Deflate.cs:583-595) resets nothing: this CTA has just parsed the previous segment (grid = 1,
        // segments in order), its hash table, prev[] and window are what the next Write finds, and the parse stands at
        // window index 32768 with an empty lookahead exactly as after a full flush.
        const bool carry_over = P.chained && seg > 0;
        if (!carry_over) {
            uint4 z = make_uint4(0, 0, 0, 0);
            uint4* h4 = reinterpret_cast<uint4*>(sm.head);
            for (int i = tid; i < HASH_SIZE * 2 / 16; i += LZ_THREADS) h4[i] = z;
            if (stream_first) {  // a freshly rented window is zero-filled (Q7)
                uint4* w4 = reinterpret_cast<uint4*>(sm.win);
                for (int i = tid; i < (WINDOW_SIZE + WIN_PAD) / 16; i += LZ_THREADS) w4[i] = z;
            } else {
                // Window as the previous segment left it (Q4): both halves hold its last 32 KiB.  Only two
                // things can observe that: Rle through window[32767] (Q14) and stale bytes past a short
                // final segment (Q7).
                const uint8_t* before = nullptr;  // the 32768 bytes preceding this segment in the stream
                if (seg > 0) before = seg_in - W_SIZE;
                else if (P.halo) before = P.halo;
                if (before && seg_len < W_SIZE + MIN_LOOKAHEAD) {
                    for (int i = tid; i < W_SIZE; i += LZ_THREADS) {
                        uint8_t v = before[i];
                        sm.win[i] = v;
                        sm.win[W_SIZE + i] = v;
                    }
                } else if (!before && seg_len < W_SIZE + MIN_LOOKAHEAD) {
                    // no halo supplied: deterministic (zero) history instead of the reference's stale bytes
                    for (int i = tid; i < WINDOW_SIZE; i += LZ_THREADS) sm.win[i] = 0;
                } else if (tid == 0) {
                    sm.win[W_SIZE - 1] = before ? before[W_SIZE - 1] : 0;
                }
                for (int i = tid; i < WIN_PAD; i += LZ_THREADS) sm.win[WINDOW_SIZE + i] = 0;
            }
        }
        uint32_t adler_a = 1, adler_b = 0;
        if (tid == 0) {
            ParserState s;
            s.str_start = stream_first ? 0 : W_SIZE;
            s.block_start = s.str_start;
            s.abs_base = stream_first ? 0 : -W_SIZE;
            s.lookahead = 0;
            s.match_length = s.prev_length = MIN_MATCH - 1;
            s.match_start = 0;
            s.prev_match = 0;
            s.match_available = 0;
            s.avail_in = seg_len;
            s.in_pos = 0;
            s.last_lit = 0;
            s.sym_total = 0;
            s.nblocks = 0;
            s.done = 0;
            *shared_state = s;
        }
        __syncthreads();

        SegOut o;
        o.sym = P.sym + (size_t)seg * P.sym_stride;
        o.blocks = P.blocks + (size_t)seg * P.max_blocks;
        o.max_blocks = (int)P.max_blocks;

        for (;;) {
            const long long st_f0 = P.stats ? clock64() : 0;
            // ---- Fill_window (Deflate.cs:967-1019), executed by the whole CTA ----
            {
                int str_start = shared_state->str_start, lookahead = shared_state->lookahead;
                int avail_in = shared_state->avail_in, in_pos = shared_state->in_pos;
                const int thresh = (cfg.fn == FN_RLE) ? MAX_MATCH + 1 : MIN_LOOKAHEAD;
                __syncthreads();
                if (lookahead < thresh) {
                    do {
                        int more = WINDOW_SIZE - lookahead - str_start;
                        if (str_start >= WINDOW_SIZE - MIN_LOOKAHEAD) {
                            // slide: window memcpy + SlideHash (Deflate.Intrinsics.cs:261-285)
                            {
                                const uint4* src = reinterpret_cast<const uint4*>(sm.win + W_SIZE);
                                uint4* dst = reinterpret_cast<uint4*>(sm.win);
                                for (int i = tid; i < W_SIZE / 16; i += LZ_THREADS) dst[i] = src[i];
                                uint32_t* hp = reinterpret_cast<uint32_t*>(sm.head);  // head and prev are contiguous
                                for (int i = tid; i < (HASH_SIZE + W_SIZE) / 2; i += LZ_THREADS) {
                                    uint32_t v = hp[i];
                                    uint32_t lo = v & 0xFFFF, hi = v >> 16;
                                    lo = lo >= W_SIZE ? lo - W_SIZE : 0;
                                    hi = hi >= W_SIZE ? hi - W_SIZE : 0;
                                    hp[i] = lo | (hi << 16);
                                }
                            }
                            str_start -= W_SIZE;
                            if (tid == 0) {
                                shared_state->match_start -= W_SIZE;
                                shared_state->block_start -= W_SIZE;
                                shared_state->abs_base += W_SIZE;
                            }
                            more += W_SIZE;
                            __syncthreads();
                        }
                        if (avail_in == 0) break;
                        int n = avail_in < more ? avail_in : more;
                        stage_input<NT>(sm.win + str_start + lookahead, seg_in + in_pos, n, bar, phase);
                        __syncthreads();
                        adler_chunk<NT>(sm.win + str_start + lookahead, n, red, adler_a, adler_b);
                        avail_in -= n;
                        in_pos += n;
                        lookahead += n;
                        // Q3: every refill inserts string str_start + 1
                        if (tid == 0 && lookahead >= MIN_MATCH) insert_string(sm, str_start + 1, P.hash_variant);
                        __syncthreads();
                    } while (lookahead < MIN_LOOKAHEAD && avail_in != 0);
                    if (tid == 0) {
                        shared_state->str_start = str_start;
                        shared_state->lookahead = lookahead;
                        shared_state->avail_in = avail_in;
                        shared_state->in_pos = in_pos;
                    }
                }
                __syncthreads();
            }
            const long long st_f1 = P.stats ? clock64() : 0;
            // ---- parse until the next refill / slide or the end of the segment ----
            if (!SLOW_ONLY && cfg.fn == FN_FAST) {
                if (tid < FAST_NB) {
                    ParserState s = *shared_state;
                    fast_bar();
                    parse_fast_multi(sm, s, shared_state, o, cfg, P.strategy, P.hash_variant, P.stats);
                    fast_bar();
                    if (tid == 0) *shared_state = s;
                }
            } else if (cfg.fn == FN_SLOW && NT == 1024 && P.strategy != 2) {
                // Every level 4-9.  Variants that were measured and dropped (profiles/r01j_slow_pipeline_stats.txt): four threads
                // per position with a quad merge (more instructions than it saves on short compares), a dedicated insert warp
                // running one tile ahead of the search (the insert is a chain of dependent shared-memory round trips and
                // crawls under the load of 30 search warps: L6 text 91 -> 200 ms per 256 MiB).
                // ---- pipelined tiles: warp 0 replays tile k while warps 1..31 insert and search tile k + 1 -----------------
                constexpr int TP = WIDE_TILE;   // positions per tile: one per thread of the 31 producer warps
                const int S0 = shared_state->str_start;
                const int avail = shared_state->lookahead - (MIN_LOOKAHEAD - 1);
                if (tid == 0) { pipe_flags[0] = 0; pipe_flags[1] = 0; }
                __syncthreads();
                if (avail < 16) {   // too close to a refill / the end of the segment: the reference loop, literally
                    if (tid == 0) {
                        ParserState s = *shared_state;
                        parse_slow(sm, s, o, cfg, P.strategy, P.hash_variant);
                        *shared_state = s;
                    }
                } else {
                    const int L = S0 + avail, ntiles = (avail + TP - 1) / TP;
                    Smem smb = sm;
                    smb.pipe_ins_limit = L;
                    smb.pipe_ntiles = ntiles;
                    smb.pipe_flags = pipe_flags;
                    if (tid < 32) {
                        ParserState s = *shared_state;
                        for (int k = 0; k < ntiles; k++) {
                            const int tS = S0 + k * TP, tT = L - tS < TP ? L - tS : TP;
                            const long long tw0 = P.stats ? clock64() : 0;
                            if (tid == 0) while (pipe_flags[0] < k + 1) __nanosleep(100);   // back off: a spinning warp takes issue slots from the producers
                            __syncwarp();
                            const long long tw1 = P.stats ? clock64() : 0;
                            __threadfence_block();
                            smb.tile_rec = sm.tile_rec + (k & 1) * WIDE_TILE;
                            smb.tile_hh = sm.tile_hh + (k & 1) * WIDE_TILE;
                            if (s.str_start < tS + tT) slow_tile_replay_warp(smb, s, o, cfg, P.strategy, P.hash_variant, tS, tT);
                            __syncwarp();
                            if (tid == 0) {
                                __threadfence_block();
                                pipe_flags[1] = k + 1;
                                if (P.stats) {
                                    atomicAdd(P.stats + 0, 1ull);
                                    atomicAdd(P.stats + 3, (unsigned long long)tT);
                                    atomicAdd(P.stats + 8, (unsigned long long)(tw1 - tw0));          // replay waits for a tile
                                    atomicAdd(P.stats + 6, (unsigned long long)(clock64() - tw1));    // replay
                                }
                            }
                        }
                        if (tid == 0) *shared_state = s;
                    } else {
                        const int ptid = tid - 32;
                        for (int k = 0; k < ntiles; k++) {
                            const int tS = S0 + k * TP, tT = L - tS < TP ? L - tS : TP;
                            // Two buffers: tile k reuses the buffer of tile k - 2.  No deadlock with pipe_wait_all: a match can run
                            // past L only from the last two tiles (a tile is 992 positions, a match at most 258).
                            const long long tp0 = P.stats ? clock64() : 0;
                            if (ptid == 0 && k >= 2) while (pipe_flags[1] < k - 1) __nanosleep(100);
                            asm volatile("bar.sync 3, 992;" ::: "memory");
                            const long long tp1 = P.stats ? clock64() : 0;
                            smb.tile_rec = sm.tile_rec + (k & 1) * WIDE_TILE;
                            smb.tile_hh = sm.tile_hh + (k & 1) * WIDE_TILE;
                            smb.shadow = shadow_buf;
                            smb.shadow_lo = tS;
                            smb.shadow_hi = tS + tT;
                            int my_pos, my_head;
                            slow_tile_insert_sorted(smb, tS, tT, P.hash_variant, ptid, my_pos, my_head);
                            asm volatile("bar.sync 3, 992;" ::: "memory");
                            const long long tp2 = P.stats ? clock64() : 0;
                            if (my_pos >= 0) smb.tile_rec[my_pos - tS] = slow_tile_search_sh(smb, my_pos, my_head, cfg);
                            asm volatile("bar.sync 3, 992;" ::: "memory");
                            if (ptid == 0) {
                                __threadfence_block();
                                pipe_flags[0] = k + 1;
                                if (P.stats) {
                                    atomicAdd(P.stats + 7, (unsigned long long)(tp1 - tp0));          // producers wait for a buffer
                                    atomicAdd(P.stats + 4, (unsigned long long)(tp2 - tp1));          // insert
                                    atomicAdd(P.stats + 5, (unsigned long long)(clock64() - tp2));    // search
                                }
                            }
                        }
                    }
                }
            } else if (cfg.fn == FN_SLOW) {
                for (;;) {
                    const int S = shared_state->str_start;
                    int T = shared_state->lookahead - (MIN_LOOKAHEAD - 1);
                    if (T > TILE_MAX) T = TILE_MAX;
                    __syncthreads();
                    if (T < 16) {   // too close to a refill / the end of the segment: the reference loop, literally
                        if (tid == 0) {
                            ParserState s = *shared_state;
                            parse_slow(sm, s, o, cfg, P.strategy, P.hash_variant);
                            *shared_state = s;
                        }
                        break;
                    }
                    const long long tA = P.stats ? clock64() : 0;
                    // HuffmanOnly never searches (Deflate.Slow.cs:66) and the hash table cannot reach the output: replay only
                    if (P.strategy != 2) {
                        if (tid < 32) slow_tile_insert(sm, S, T, P.hash_variant);
                        __syncthreads();
                    }
                    const long long tB = P.stats ? clock64() : 0;
                    if (P.strategy != 2) {
                        if (tid < T) sm.tile_rec[tid] = slow_tile_search(sm, S + tid, sm.tile_hh[tid], cfg);
                        __syncthreads();
                    }
                    const long long tC = P.stats ? clock64() : 0;
                    if (tid < 32) {
                        ParserState s = *shared_state;
                        __syncwarp();
                        slow_tile_replay_warp(sm, s, o, cfg, P.strategy, P.hash_variant, S, T);
                        __syncwarp();
                        if (tid == 0) *shared_state = s;
                    }
                    __syncthreads();
                    if (P.stats && tid == 0) {
                        atomicAdd(P.stats + 0, 1ull);
                        atomicAdd(P.stats + 1, (unsigned long long)(shared_state->str_start - S));
                        atomicAdd(P.stats + 3, (unsigned long long)T);
                        atomicAdd(P.stats + 4, (unsigned long long)(tB - tA));
                        atomicAdd(P.stats + 5, (unsigned long long)(tC - tB));
                        atomicAdd(P.stats + 6, (unsigned long long)(clock64() - tC));
                    }
                }
            } else if (!SLOW_ONLY && tid == 0) {
                ParserState s = *shared_state;
                parse_rle(sm, s, o);
                *shared_state = s;
            }
            __syncthreads();
            if (P.stats && tid == 0) {
                atomicAdd(P.stats + 10, (unsigned long long)(st_f1 - st_f0));
                atomicAdd(P.stats + 11, (unsigned long long)(clock64() - st_f1));
            }
            if (shared_state->done) break;
        }
        if (tid == 0) {
            P.nblocks[seg] = (uint32_t)shared_state->nblocks;
            P.seg_adler[seg] = make_uint2(adler_a, adler_b);
        }
        __syncthreads();
    }
}

size_t lz77_smem_bytes(int threads)
{
    const size_t rec_n = threads == 1024 ? 2 * WIDE_TILE : TILE_MAX, slots = threads == 1024 ? WIDE_SLOTS : SLOT_COUNT;
    return WINDOW_SIZE + WIN_PAD + HASH_SIZE * 2 + W_SIZE * 2 + 1024 * 2 + 8 + 128 * 4 + sizeof(ParserState) + 16 + slots * 4 + 16 +
           rec_n * 8 + rec_n * 2 + (threads == 1024 ? SHADOW_N * 2 : 0) + 64 + 128 + 128;
}
size_t lz77_smem_bytes() { return lz77_smem_bytes(256); }

// HuffmanOnly / Rle -> the parallel symbolizer; level 1 -> the two-kernel path (lists + resolve) for the full segments and
// this file's batch parser for a short last segment; everything else -> this file's kernels.
static cudaError_t launch_lz77_classic(const Lz77Params& P, int num_sms, cudaStream_t stream);

cudaError_t launch_lz77(const Lz77Params& P, int num_sms, cudaStream_t stream)
{
    // ZSB_NO_SYMBOLIZE=1 keeps HuffmanOnly / Rle on the serial path below (A/B measurements, parity cross-checks)
    if (P.fn == FN_STORED) return launch_stored_blocks(P, stream);
    static const bool no_symbolize = getenv("ZSB_NO_SYMBOLIZE") != nullptr;
    if (!no_symbolize && !P.stats && symbolize_applies(P)) return launch_symbolize(P, num_sms, stream);
    // ZSB_L1_KERNEL=batch keeps level 1 on the batch parser (A/B measurements, parity cross-checks); read per call
    const char* l1 = getenv("ZSB_L1_KERNEL");
    if (lz77_lists_applies(P) && !P.chained && !(l1 && strcmp(l1, "batch") == 0)) {
        const uint32_t nfull = (uint32_t)(P.n / P.seg_size < P.nseg ? P.n / P.seg_size : P.nseg);
        if (nfull > 0) {
            Lz77Params F = P;
            F.nseg = nfull;
            cudaError_t e = launch_lz77_lists(F, num_sms, stream);
            if (e != cudaSuccess) return e;
            if (P.ev_mid) cudaEventRecord(P.ev_mid, stream);
            e = launch_lz77_resolve(F, num_sms, stream);
            if (e != cudaSuccess) return e;
        }
        if (nfull == P.nseg) return cudaSuccess;
        Lz77Params T = P;   // the short last segment
        const size_t off = (size_t)nfull * P.seg_size;
        T.in = P.in + off;
        T.n = P.n - off;
        T.nseg = P.nseg - nfull;
        T.first_seg_index = P.first_seg_index + nfull;
        if (nfull > 0) T.halo = T.in - W_SIZE;
        T.sym = P.sym + (size_t)nfull * P.sym_stride;
        T.blocks = P.blocks + (size_t)nfull * P.max_blocks;
        T.nblocks = P.nblocks + nfull;
        T.seg_adler = P.seg_adler + nfull;
        T.work_counter = P.work_counter + 2;
        return launch_lz77_classic(T, num_sms, stream);
    }
    return launch_lz77_classic(P, num_sms, stream);
}

static cudaError_t launch_lz77_classic(const Lz77Params& P, int num_sms, cudaStream_t stream)
{
    const size_t smem = lz77_smem_bytes(256), smem_wide = lz77_smem_bytes(1024);
    {   // per-device attributes: set on every launch (a process may drive several GPUs, one zsb_ctx each)
        cudaError_t e = cudaFuncSetAttribute(lz77_segments_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(lz77_segments_kernel<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_wide);
        if (e != cudaSuccess) return e;
    }
    int grid = (int)(P.nseg < (uint32_t)num_sms ? P.nseg : (uint32_t)num_sms);
    if (grid < 1 || P.chained) grid = 1;   // chained segments (sync flush) are one CTA's, in order
    // ZSB_SLOW_THREADS=256 keeps levels 4-9 on the 256-thread kernel (A/B measurements)
    static const bool slow_wide = [] {
        const char* e = getenv("ZSB_SLOW_THREADS");
        return !(e && atoi(e) == 256);
    }();
    if (P.fn == FN_SLOW && slow_wide) lz77_segments_kernel<1024, true><<<grid, 1024, smem_wide, stream>>>(P);
    else lz77_segments_kernel<256, false><<<grid, 256, smem, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace zsb
