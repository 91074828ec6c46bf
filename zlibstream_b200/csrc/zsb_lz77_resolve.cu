// zsb_lz77_resolve.cu -- level 1 (DeflateFast), kernel B of two: the greedy parse, one WARP per segment.
//
// Input: the 32-byte candidate records of zsb_lz77_lists.cu (for every position its 10 nearest predecessors in the
// all-position hash chain, each with the common prefix length) and the input bytes (literals).  State per segment: an "inserted" bitmap over the last
// 32768 positions (4 KiB of shared memory) -- the reference's head[] / prev[] chain (Deflate.cs:866-877) is the
// all-position chain filtered by that bitmap: parse points and the interiors of matches <= max_lazy are inserted,
// the interiors of longer matches are not (Deflate.Fast.cs:81-104).  Longest_match (Deflate.cs:1022-1100) visits the
// first max_chain inserted nodes (first one at distance <= MAX_DIST, later ones < MAX_DIST), keeps the first strictly
// longer one and stops at nice_match.  The Fill_window pre-insert of strStart + 1 is three rules:
//   - the refill at parse point r sets the inserted bit of r + 1;
//   - H(r) != H(r + 1): r + 1 searches nothing if it becomes a parse point (InsertString finds itself, Q2);
//   - H(r) == H(r + 1): r searches nothing, and a chain walk that has visited r ends there (prev[] two-cycle).
// Window index 0 is NIL: in the stream's first segment position 0 is never a candidate (Q5).
// tools/ab_model/ab_model.c is the CPU model of exactly these rules (byte-identical to the oracle).
//
// The warp takes 32 aligned positions at a time: every lane evaluates its own position as if it were a parse point
// (bitmap lookups for candidates in earlier groups, the live mask for candidates inside the group), then the parse is
// resolved left to right: literal runs in one step, a lane whose candidates lie inside the group is re-evaluated when
// the parse reaches it (all earlier bits are final then), a lane whose record does not settle the outcome (fewer than
// max_chain inserted nodes among 10, chain goes on) takes the warp-cooperative slow path over the records of the chain.
// Records are prefetched one group ahead; groups a match jumps over are never read.
//
// Most groups are "plain": no Fill_window call, no block end, no pre-inserted position, no two-cycle node and no NIL index
// can matter inside them.  They take the short path: ten bitmap lookups per lane, one evaluation over the bits of "the
// first max_chain inserted candidates that agree on three bytes" and a walk that takes literal runs in one step and
// one shuffle per match.  Everything else runs the general loop, which spells every rule out per parse point.
#include "zsb_kernels.cuh"

#ifndef ZSB_RB_PREFETCH
#define ZSB_RB_PREFETCH "prefetch.global.L2"
#endif

namespace zsb {

namespace {

constexpr int RB_WARPS = 8;
#ifndef ZSB_RB_BLOCKS
#define ZSB_RB_BLOCKS 4   // CTAs per SM: 32 warps at 64 registers (5 = 40 warps at 48 registers with spills measured slower)
#endif
constexpr int RING_WORDS = W_SIZE / 32;   // inserted bits of the last 32768 positions

__device__ __forceinline__ uint32_t ldg_u32_unaligned(const uint8_t* q)
{
    const uintptr_t u = reinterpret_cast<uintptr_t>(q);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(u & ~(uintptr_t)3);
    const uint32_t lo = __ldg(w), hi = __ldg(w + 1);
    return __funnelshift_r(lo, hi, (int)(u & 3) * 8);
}

// Common prefix (capped at MAX_MATCH) of the strings at segment positions p and c < p.  Bytes past the end of the segment
// are what the reference's window holds there: the byte 32768 positions earlier.
__device__ int prefix_global(const uint8_t* s, int N, int p, int c)
{
    int n = 0;
    while (n < MAX_MATCH && p + n + 8 <= N) {
        const uint32_t x = ldg_u32_unaligned(s + p + n) ^ ldg_u32_unaligned(s + c + n);
        if (x) {
            n += (__ffs(x) - 1) >> 3;
            return n < MAX_MATCH ? n : MAX_MATCH;
        }
        n += 4;
    }
    while (n < MAX_MATCH) {
        const int i = p + n, j = c + n;
        const uint8_t a = s[i < N ? i : i - W_SIZE], b = s[j < N ? j : j - W_SIZE];
        if (a != b) break;
        n++;
    }
    return n < MAX_MATCH ? n : MAX_MATCH;
}

// candidate j of a record: distance (15 bits) and prefix length - 3
__device__ __forceinline__ uint32_t dist_at(const uint4& r0, uint32_t r1x, int j)
{
    const uint32_t w = j < 4 ? (j < 2 ? r0.x : r0.y) : (j < 8 ? (j < 6 ? r0.z : r0.w) : r1x);
    return (w >> ((j & 1) * 16)) & 0xFFFFu;
}
__device__ __forceinline__ uint32_t len8_at(const uint4& r1, int j)
{
    const uint32_t w = j < 4 ? r1.y : (j < 8 ? r1.z : r1.w);
    return (w >> ((j & 3) * 8)) & 0xFFu;
}

// inserted bit of position a - d, shifted into I from the right (candidate 0 ends up in the highest of the ten bits)
__device__ __forceinline__ uint32_t look(const uint32_t* ring, int a, uint32_t d, uint32_t I)
{
    const int c = a - (int)d;
    const uint32_t w = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(ring) + ((c >> 3) & ((RING_WORDS - 1) * 4)));
    return __funnelshift_l(w << (~c & 31), I, 1);
}

}  // namespace

// Slow path of one parse point ac (warp-cooperative): the record of the parse point did not settle Longest_match; follow the
// all-position chain through the records of its nodes, up to 10 nodes per round, lane j = node j.
__device__ __forceinline__ void slow_search(const uint4* rec, const uint8_t* seg_in, const uint32_t* ring, int N, int G, uint32_t insm,
                                            int ac, int node, int nice, int chain, bool first, int cut0, int cut1, int lane, int& cnt,
                                            int& best, int& seld)
{
    const uint32_t lt = (1u << lane) - 1u;
    bool done = false;
    while (!done) {
        const uint8_t* rb = reinterpret_cast<const uint8_t*>(rec + (size_t)node * 2);
        const uint32_t fl = __ldg(reinterpret_cast<const uint32_t*>(rb) + 7);
        const int nkc = (int)(fl >> 27);
        const bool morec = ((fl >> 26) & 1u) != 0;
        const bool valid = lane < nkc;
        int d = 0, cj = 0, Lj = 0;
        bool in = false;
        if (valid) {
            cj = node - (int)__ldg(reinterpret_cast<const uint16_t*>(rb) + lane);
            d = ac - cj;
            if (d <= MAX_DIST) {
                in = cj >= G ? ((insm >> (cj - G)) & 1u) != 0 : ((ring[(cj >> 5) & (RING_WORDS - 1)] >> (cj & 31)) & 1u) != 0;
                if (in) Lj = prefix_global(seg_in, N, ac, cj);
            }
        }
        const uint32_t far = __ballot_sync(0xffffffffu, valid && d > MAX_DIST);   // the chain has left the window
        const uint32_t V = __ballot_sync(0xffffffffu, in);
        const int rank = cnt + __popc(V & lt);   // inserted nodes visited before this one
        const bool stop_before = in && ((rank == 0 ? d > MAX_DIST : d >= MAX_DIST) || (first && cj == 0));
        const bool stop_after = in && (Lj >= nice || cj == cut0 || cj == cut1 || rank + 1 == chain);
        const uint32_t SB = __ballot_sync(0xffffffffu, stop_before) | far;
        const uint32_t SA = __ballot_sync(0xffffffffu, stop_after);
        const int fb = SB ? __ffs(SB) - 1 : 32;     // lanes below fb may be visited
        const int fa = SA ? __ffs(SA) - 1 : 32;     // lanes up to fa may be visited
        const int lim = fb - 1 < fa ? fb - 1 : fa;  // last lane that is visited if inserted
        const uint32_t inc = lim >= 31 ? V : (lim < 0 ? 0u : V & ((2u << lim) - 1u));
        const uint32_t key = ((inc >> lane) & 1u) ? ((uint32_t)Lj << 5) | (uint32_t)(31 - lane) : 0u;
        const uint32_t top = __reduce_max_sync(0xffffffffu, key);
        if ((int)(top >> 5) > best) {
            best = (int)(top >> 5);
            seld = __shfl_sync(0xffffffffu, d, 31 - (int)(top & 31u));
        }
        cnt += __popc(inc);
        if (SB || SA || !morec || nkc == 0) done = true;
        else node = __shfl_sync(0xffffffffu, cj, nkc - 1);
    }
}

template <bool STATS>
__global__ void __launch_bounds__(RB_WARPS * 32, ZSB_RB_BLOCKS) lz77_resolve_kernel(Lz77Params P)
{
    __shared__ uint32_t ring_all[RB_WARPS][RING_WORDS];
    const int lane = threadIdx.x & 31;
    uint32_t* ring = ring_all[threadIdx.x >> 5];
    const uint32_t lt = (1u << lane) - 1u;
    const int N = (int)P.seg_size;
    const int chain = P.chain, nice_cfg = P.nice, lazy = P.lazy;
    unsigned long long st_groups = 0, st_evals = 0, st_steps = 0, st_slow = 0, st_general = 0;

    for (;;) {
        uint32_t seg = 0;
        if (lane == 0) seg = atomicAdd(P.work_counter + 1, 1u);
        seg = __shfl_sync(0xffffffffu, seg, 0);
        if (seg >= P.nseg) break;
        const uint4* rec = P.lists + (size_t)seg * P.seg_size * 2;
        const uint8_t* seg_in = P.in + (size_t)seg * P.seg_size;
        uint32_t* sym = P.sym + (size_t)seg * P.sym_stride;
        BlockRec* blocks = P.blocks + (size_t)seg * P.max_blocks;
        const bool first = (P.first_seg_index + (uint64_t)seg) == 0;

        // parser state (uniform across the warp)
        int pos = 0;                       // next parse point
        int loaded_end = 0;                // bytes Fill_window has read so far
        int base = first ? 0 : -W_SIZE;    // segment position of window index 0
        int forced = -1, cut0 = -1, cut1 = -1;
        int last_lit = 0, sym_total = 0, nblocks = 0, blk_start = 0;
        uint32_t carry_ins = 0;            // inserted bits already known for the group after the current one
        int pf_G = -1;                     // group whose records are in (n0, n1, nlit)
        uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
        uint32_t nlit = 0;

        while (pos < N) {
            const int G = pos & ~31;
            const int a = G + lane;
            const int la = N - a;
            uint4 r0, r1;
            uint32_t lit;
            if (pf_G == G) { r0 = n0; r1 = n1; lit = nlit; }
            else { r0 = __ldg(rec + (size_t)a * 2); r1 = __ldg(rec + (size_t)a * 2 + 1); lit = __ldg(seg_in + a); }
            if (G + 32 < N) { n0 = __ldg(rec + (size_t)(a + 32) * 2); n1 = __ldg(rec + (size_t)(a + 32) * 2 + 1); nlit = __ldg(seg_in + a + 32); pf_G = G + 32; }
            else pf_G = -1;
            if (STATS) st_groups++;

            // ---- the record: 10 x u16 distance | 10 x u8 (prefix - 3) | 10 bits "prefix >= 3" | chain goes on | count ----
            const int nk = (int)(r1.w >> 27);
            const bool more = ((r1.w >> 26) & 1u) != 0;
            const uint32_t usem = (r1.w >> 16) & 0x3FFu;
            const uint32_t d0 = r0.x & 0xFFFFu;

            // ---- inserted bits of the candidates (those inside this group are read from the live mask instead) ----
            uint32_t I = 0;
            I = look(ring, a, r0.x & 0xFFFFu, I); I = look(ring, a, r0.x >> 16, I);
            I = look(ring, a, r0.y & 0xFFFFu, I); I = look(ring, a, r0.y >> 16, I);
            I = look(ring, a, r0.z & 0xFFFFu, I); I = look(ring, a, r0.z >> 16, I);
            I = look(ring, a, r0.w & 0xFFFFu, I); I = look(ring, a, r0.w >> 16, I);
            I = look(ring, a, r1.x & 0xFFFFu, I); I = look(ring, a, r1.x >> 16, I);
            const uint32_t V = (__brev(I) >> 22) & ((1u << nk) - 1u);
            uint32_t ING = 0;   // candidates inside this group (a prefix of the list: distances ascend)
            if (__ballot_sync(0xffffffffu, nk > 0 && d0 <= (uint32_t)lane)) {
#pragma unroll
                for (int j = 0; j < LZ_LIST_K; j++)
                    if (j < nk && dist_at(r0, r1.x, j) <= (uint32_t)lane) ING |= 1u << j;
            }

            uint32_t insm = carry_ins;   // inserted bits of this group (final below the parse point)
            carry_ins = 0;
            uint32_t ppm = 0;            // parse points of this group
            const int sym_group = sym_total;
            int cur = pos - G;
            int r_len = MIN_MATCH - 1, r_dist = 0, r_cnt = 0, r_best = MIN_MATCH - 1;
            bool r_over = false;

            const bool plain = G + 32 + MIN_LOOKAHEAD <= loaded_end && forced < G && (cut0 < 0 || cut0 < G - W_SIZE) &&
                               !(first && G < W_SIZE) && last_lit + 32 < BLOCK_SYMS;
            if (plain) {
                // ================= short path =================
                int eval_at = cur;
                for (;;) {
                    if (STATS) st_evals++;
                    if (lane >= eval_at) {
                        uint32_t inm = V;
                        if (ING) {
                            inm &= ~ING;
                            uint32_t m = ING;
                            while (m) {
                                const int j = __ffs(m) - 1;
                                m &= m - 1;
                                inm |= ((insm >> (lane - (int)dist_at(r0, r1.x, j))) & 1u) << j;
                            }
                        }
                        // Longest_match visits the first max_chain (= 4) inserted nodes; only those with a prefix >= 3 can win
                        uint32_t t = inm;
                        t &= t - 1; t &= t - 1; t &= t - 1; t &= t - 1;
                        const uint32_t V4 = inm ^ t;
                        uint32_t m = V4 & usem;
                        int best = MIN_MATCH - 1, seld = 0;
                        bool nice_hit = false;
                        while (m) {
                            const int j = __ffs(m) - 1;
                            m &= m - 1;
                            const int L = (int)len8_at(r1, j) + MIN_MATCH;
                            if (L > best) { best = L; seld = (int)dist_at(r0, r1.x, j); }
                            if (L >= nice_cfg) { nice_hit = true; break; }
                        }
                        r_best = best;
                        r_len = best;
                        r_dist = seld;
                        r_cnt = __popc(V4);
                        r_over = more && !nice_hit && r_cnt < chain;
                        // the walk will need the record of the last listed node if the parse lands here: ask for it now
                        if (r_over && nk > 0) {
                            const uint4* nr = rec + (size_t)(a - (int)dist_at(r0, r1.x, nk - 1)) * 2;
                            asm volatile(ZSB_RB_PREFETCH " [%0];" ::"l"(nr));
                        }
                    }
                    uint32_t OVm = __ballot_sync(0xffffffffu, r_over);
                    uint32_t Mm = __ballot_sync(0xffffffffu, r_len >= MIN_MATCH);
                    // a lane whose nearest candidate lies at or beyond the point of this evaluation is evaluated again when reached
                    const uint32_t RE = __ballot_sync(0xffffffffu, ING != 0 && lane - (int)d0 >= eval_at);
                    bool again = false;
                    while (cur < 32) {
                        if (STATS) st_steps++;
                        const uint32_t fin_lit = ~(Mm | OVm | RE);
                        const uint32_t r = fin_lit >> cur;
                        const int run = r == 0xffffffffu ? 32 : __ffs(~r) - 1;
                        if (run > 0) {   // a run of final literals: all parse points, all inserted
                            const uint32_t m = (run >= 32 ? 0xffffffffu : (1u << run) - 1u) << cur;
                            ppm |= m;
                            insm |= m;
                            cur += run;
                            continue;
                        }
                        if ((RE >> cur) & 1u) { eval_at = cur; again = true; break; }
                        if ((OVm >> cur) & 1u) {
                            if (STATS) st_slow++;
                            int cnt = __shfl_sync(0xffffffffu, r_cnt, cur);
                            int best = __shfl_sync(0xffffffffu, r_best, cur);
                            int seld = __shfl_sync(0xffffffffu, r_dist, cur);
                            const int node = __shfl_sync(0xffffffffu, nk > 0 ? a - (int)dist_at(r0, r1.x, nk - 1) : a, cur);
                            slow_search(rec, seg_in, ring, N, G, insm, G + cur, node, nice_cfg, chain, first, cut0, cut1, lane, cnt, best, seld);
                            if (lane == cur) { r_len = best; r_dist = seld; }
                            OVm &= ~(1u << cur);
                            if (best >= MIN_MATCH) Mm |= 1u << cur;
                            else Mm &= ~(1u << cur);
                            continue;
                        }
                        const int L = __shfl_sync(0xffffffffu, r_len, cur);
                        ppm |= 1u << cur;
                        insm |= 1u << cur;
                        if (L <= lazy) {   // the interior of a short match is inserted (Deflate.Fast.cs:81-96)
                            const unsigned long long m = ((1ull << (L - 1)) - 1ull) << (cur + 1);
                            insm |= (uint32_t)m;
                            carry_ins |= (uint32_t)(m >> 32);
                        }
                        cur += L;
                    }
                    if (!again) break;
                }
                last_lit += __popc(ppm);
            } else {
                // ================= general path: every rule per parse point =================
                if (STATS) st_general++;
                uint32_t dist[LZ_LIST_K], len8[LZ_LIST_K];
                dist[0] = r0.x & 0xFFFF; dist[1] = r0.x >> 16; dist[2] = r0.y & 0xFFFF; dist[3] = r0.y >> 16;
                dist[4] = r0.z & 0xFFFF; dist[5] = r0.z >> 16; dist[6] = r0.w & 0xFFFF; dist[7] = r0.w >> 16;
                dist[8] = r1.x & 0xFFFF; dist[9] = r1.x >> 16;
                len8[0] = r1.y & 0xFF; len8[1] = (r1.y >> 8) & 0xFF; len8[2] = (r1.y >> 16) & 0xFF; len8[3] = r1.y >> 24;
                len8[4] = r1.z & 0xFF; len8[5] = (r1.z >> 8) & 0xFF; len8[6] = (r1.z >> 16) & 0xFF; len8[7] = r1.z >> 24;
                len8[8] = r1.w & 0xFF; len8[9] = (r1.w >> 8) & 0xFF;
                const uint32_t I_out = V & ~ING;
                const uint32_t DM = __ballot_sync(0xffffffffu, ING != 0);   // lanes that depend on this group's own bits
                const int lastc = nk > 0 ? a - (int)dist_at(r0, r1.x, nk - 1) : a;   // last node of the record (only used when `more`)
                bool need_eval = true;
                uint32_t CLM = 0;            // lanes whose outcome is a literal and final (clean, no slow path, not forced, insertable)

                while (cur < 32 && G + cur < N) {
                    const int ac = G + cur;
                    // ---- Fill_window (Deflate.cs:967-1019) as far as the parse can see it ----
                    if (loaded_end - ac < MIN_LOOKAHEAD && (loaded_end < N || ac - base >= WINDOW_SIZE - MIN_LOOKAHEAD)) {
                        if (ac - base >= WINDOW_SIZE - MIN_LOOKAHEAD) base += W_SIZE;   // slide
                        if (loaded_end < N) {
                            int room = WINDOW_SIZE - (loaded_end - ac) - (ac - base);
                            int n = N - loaded_end < room ? N - loaded_end : room;
                            loaded_end += n;
                            if (loaded_end - ac >= MIN_MATCH) {   // Q3: the refill inserts string ac + 1
                                // same hash as ac  <=>  ac is the nearest predecessor of ac + 1 in the all-position chain
                                uint32_t d0n;
                                if (cur < 31) d0n = __shfl_sync(0xffffffffu, nk > 0 ? d0 : 0u, cur + 1);
                                else d0n = __shfl_sync(0xffffffffu, (n1.w >> 27) ? (n0.x & 0xFFFFu) : 0u, 0);
                                if (cur < 31) insm |= 1u << (cur + 1);
                                else carry_ins |= 1u;
                                if (d0n == 1u) { forced = ac; cut1 = cut0; cut0 = ac; }
                                else forced = ac + 1;
                                need_eval = true;
                            }
                        }
                    }
                    // ---- (re-)evaluate the lanes from the parse point on ----
                    if (need_eval || ((DM >> cur) & 1u)) {
                        if (STATS) st_evals++;
                        if (lane >= cur && la > 0) {
                            int cnt = 0, best = MIN_MATCH - 1, seld = 0;
                            bool done = false;
                            const int nice = la < nice_cfg ? la : nice_cfg;
#pragma unroll 1   // the rare path: a rolled loop keeps its ten candidates out of the registers of the short path
                            for (int j = 0; j < LZ_LIST_K; j++) {
                                if (!done && j < nk) {
                                    const int d = (int)dist[j];
                                    const bool in = ((ING >> j) & 1u) ? ((insm >> ((lane - d) & 31)) & 1u) != 0 : ((I_out >> j) & 1u) != 0;
                                    if (in) {
                                        const int c = a - d;
                                        if (cnt == 0 ? d > MAX_DIST : d >= MAX_DIST) done = true;
                                        else if (first && c == 0) done = true;
                                        else {
                                            cnt++;
                                            const int L = ((usem >> j) & 1u) ? (int)len8[j] + MIN_MATCH : 0;
                                            if (L > best) {
                                                best = L;
                                                seld = d;
                                                if (L >= nice) done = true;
                                            }
                                            if (c == cut0 || c == cut1 || cnt == chain) done = true;
                                        }
                                    }
                                }
                            }
                            r_over = !done && more && la >= MIN_MATCH;
                            r_cnt = cnt;
                            r_best = best;
                            r_dist = seld;
                            r_len = la >= MIN_MATCH ? (best < la ? best : la) : MIN_MATCH - 1;
                            if (a == forced) { r_len = MIN_MATCH - 1; r_over = false; }
                        }
                        need_eval = false;
                        CLM = __ballot_sync(0xffffffffu, ING == 0 && !r_over && r_len < MIN_MATCH && la >= MIN_MATCH);
                    }
                    // ---- a run of final literals: all parse points, all inserted ----
                    {
                        // lanes at or beyond the next Fill_window trigger are taken one by one
                        uint32_t ok = CLM;
                        const bool fill_pending = loaded_end < N || N - base > WINDOW_SIZE - MIN_LOOKAHEAD;   // a refill, or the last slide
                        if (fill_pending) {
                            const int thr = loaded_end - (MIN_LOOKAHEAD - 1) - G;   // first lane with lookahead < MIN_LOOKAHEAD
                            if (thr < 32) ok &= thr <= 0 ? 0u : (1u << thr) - 1u;
                        }
                        const uint32_t r = ok >> cur;
                        int run = r == 0xffffffffu ? 32 : __ffs(~r) - 1;   // consecutive ones from lane cur
                        const int roomb = BLOCK_SYMS - last_lit;
                        if (run > roomb) run = roomb;
                        if (run > 0) {
                            const uint32_t m = (run >= 32 ? 0xffffffffu : (1u << run) - 1u) << cur;
                            ppm |= m;
                            insm |= m;
                            cur += run;
                            last_lit += run;
                            if (last_lit == BLOCK_SYMS) {   // Flush_block_only (Deflate.cs:951-956)
                                if (lane == 0 && nblocks < (int)P.max_blocks) {
                                    BlockRec& b = blocks[nblocks];
                                    b.sym_begin = (uint32_t)(sym_group + __popc(ppm) - BLOCK_SYMS);
                                    b.sym_count = BLOCK_SYMS;
                                    b.in_begin = (uint32_t)blk_start;
                                    b.in_len = (uint32_t)(G + cur - blk_start);
                                    b.buf_valid = blk_start - base >= 0;
                                }
                                nblocks++;
                                last_lit = 0;
                                blk_start = G + cur;
                            }
                            continue;
                        }
                    }
                    // ---- one parse point ----
                    const int la_c = N - ac;
                    int L = __shfl_sync(0xffffffffu, r_len, cur);
                    const bool ov = __shfl_sync(0xffffffffu, (int)r_over, cur) != 0;
                    if (ov) {
                        if (STATS) st_slow++;
                        int cnt = __shfl_sync(0xffffffffu, r_cnt, cur);
                        int best = __shfl_sync(0xffffffffu, r_best, cur);
                        int seld = __shfl_sync(0xffffffffu, r_dist, cur);
                        const int node = __shfl_sync(0xffffffffu, lastc, cur);
                        const int nice = la_c < nice_cfg ? la_c : nice_cfg;
                        slow_search(rec, seg_in, ring, N, G, insm, ac, node, nice, chain, first, cut0, cut1, lane, cnt, best, seld);
                        L = best < la_c ? best : la_c;
                        if (lane == cur) { r_len = L; r_dist = seld; }
                    }
                    ppm |= 1u << cur;
                    if (la_c >= MIN_MATCH) insm |= 1u << cur;
                    if (L >= MIN_MATCH) {
                        if (L <= lazy && la_c - L >= MIN_MATCH) {   // the interior of a short match is inserted (Deflate.Fast.cs:81-96)
                            const unsigned long long m = ((1ull << (L - 1)) - 1ull) << (cur + 1);
                            insm |= (uint32_t)m;
                            carry_ins |= (uint32_t)(m >> 32);
                        }
                        cur += L;
                    } else cur += 1;
                    if (++last_lit == BLOCK_SYMS) {
                        if (lane == 0 && nblocks < (int)P.max_blocks) {
                            BlockRec& b = blocks[nblocks];
                            b.sym_begin = (uint32_t)(sym_group + __popc(ppm) - BLOCK_SYMS);
                            b.sym_count = BLOCK_SYMS;
                            b.in_begin = (uint32_t)blk_start;
                            b.in_len = (uint32_t)(G + cur - blk_start);
                            b.buf_valid = blk_start - base >= 0;
                        }
                        nblocks++;
                        last_lit = 0;
                        blk_start = G + cur;
                    }
                }
            }
            // ---- symbols of the group (Tr_tally_*, Deflate.cs:910-948), bitmap word, next parse point ----
            if ((ppm >> lane) & 1u)
                sym[sym_group + __popc(ppm & lt)] = r_len >= MIN_MATCH ? ((uint32_t)r_dist | ((uint32_t)(r_len - MIN_MATCH) << 16)) : (lit << 16);
            sym_total = sym_group + __popc(ppm);
            pos = G + cur;
            if (pos > N) pos = N;
            {
                // groups the last match jumped over hold no inserted positions (beyond the carry of a short match)
                const int g0 = G >> 5, g1 = (pos < N ? pos : N - 1) >> 5;
                if (lane == 0) ring[g0 & (RING_WORDS - 1)] = insm;
                if (lane >= 1 && g0 + lane < g1) ring[(g0 + lane) & (RING_WORDS - 1)] = lane == 1 ? carry_ins : 0u;
                if (g1 > g0 + 1) carry_ins = 0;
                else if (g1 == g0) carry_ins = 0;   // end of the segment
            }
            __syncwarp();
        }
        // ---- end of input: Flush_block_only of whatever is open (Deflate.Fast.cs:124) ----
        // the Fill_window call that finds no more input still slides when the parse stands in the upper window half (Q4)
        if (N - base >= WINDOW_SIZE - MIN_LOOKAHEAD) base += W_SIZE;
        if (lane == 0) {
            if (nblocks < (int)P.max_blocks) {
                BlockRec& b = blocks[nblocks];
                b.sym_begin = (uint32_t)(sym_total - last_lit);
                b.sym_count = (uint32_t)last_lit;
                b.in_begin = (uint32_t)blk_start;
                b.in_len = (uint32_t)(N - blk_start);
                b.buf_valid = blk_start - base >= 0;
            }
            P.nblocks[seg] = (uint32_t)(nblocks + 1);
        }
        __syncwarp();
    }
    if (STATS && lane == 0) {
        atomicAdd(P.stats + 24, st_groups);
        atomicAdd(P.stats + 25, st_evals);
        atomicAdd(P.stats + 26, st_steps);
        atomicAdd(P.stats + 27, st_slow);
        atomicAdd(P.stats + 28, st_general);
    }
}

bool lz77_lists_applies(const Lz77Params& P)
{
    // level 1 only (max_chain 4, nice 8: the 10-node records settle most searches), full segments that refill like the
    // reference's window does for multiples of 32 KiB
    return P.lists != nullptr && P.fn == FN_FAST && P.strategy != 2 && P.chain == 4 && P.nice == 8 && P.lazy == 4 &&
           P.seg_size % W_SIZE == 0 && P.seg_size >= (uint32_t)WINDOW_SIZE && P.seg_size <= (1u << 24);
}

size_t lz77_lists_bytes_per_segment(uint32_t seg_size) { return (size_t)seg_size * 32; }

cudaError_t launch_lz77_resolve(const Lz77Params& P, int num_sms, cudaStream_t stream)
{
    int per_sm = 4;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lz77_resolve_kernel<false>, RB_WARPS * 32, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    int grid = num_sms * per_sm;
    const int need = (int)((P.nseg + RB_WARPS - 1) / RB_WARPS);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (P.stats) lz77_resolve_kernel<true><<<grid, RB_WARPS * 32, 0, stream>>>(P);
    else lz77_resolve_kernel<false><<<grid, RB_WARPS * 32, 0, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace zsb
