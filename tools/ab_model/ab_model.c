/*
 * ab_model.c -- CPU model of the two-kernel level-1 LZ77 (zsb_lz77_lists.cu / zsb_lz77_resolve.cu).
 *
 * DEVELOPMENT TOOL, not product code and not part of the oracle: it includes the oracle's source, swaps its DeflateFast
 * for the restructured algorithm through the oracle's test hook, and checks on real files that both produce the same
 * bytes.
 *
 * Kernel A ("lists"): for EVERY position a of a segment, the K nearest earlier positions with the same 15-bit hash
 *   (a chain over all positions -- independent of the parse), each with the common prefix length of the two strings,
 *   plus a flag "the chain goes on".  Fully parallel.
 * Kernel B ("resolve"): the greedy parse of DeflateFast (Deflate.Fast.cs:20-128).  The reference's hash chain holds
 *   only the INSERTED positions (parse points and the interiors of matches <= max_lazy): it is the all-position chain
 *   filtered by an "inserted" bitmap that B maintains.  Longest_match (Deflate.cs:1022-1100) visits the first
 *   max_chain inserted nodes: first node at distance <= MAX_DIST, later ones < MAX_DIST, stops at the first node whose
 *   length reaches nice_match.  The Fill_window pre-insert is modelled by three rules:
 *     - the refill at parse point r sets the inserted bit of r + 1;
 *     - H(r) != H(r+1): r + 1 searches nothing if it is a parse point (InsertString finds itself, Q2);
 *     - H(r) == H(r+1): r searches nothing, and a walk that has visited r ends there (the prev[] two-cycle).
 *   In the stream's first segment position 0 is NIL (window index 0).
 *
 * gcc -O2 -o /tmp/ab_model tools/ab_model/ab_model.c oracle/zso_adler32.c && /tmp/ab_model file seg_size [K]
 */
#include "../../oracle/zso_deflate.c"
#include <stdio.h>

#define MAXSEG (1 << 20)
#define KMAX 32
static int K = 10;
#define MAXD (32768 - MIN_LOOKAHEAD)   /* 32506 */

static struct {
    const uint8_t* seg;
    int N, first, fresh;
    /* kernel A */
    int prev_all[MAXSEG];
    uint16_t hash[MAXSEG];
    uint16_t dist[MAXSEG][KMAX];
    uint16_t len[MAXSEG][KMAX];   /* common prefix, 0..258 */
    uint8_t nk[MAXSEG], more[MAXSEG];
    /* kernel B */
    uint8_t ins[MAXSEG + 8];
    int base, loaded_end, cut0, cut1, forced;
    int blk_start;
} M;
static uint64_t st_search, st_overflow, st_over_nodes, st_pp, st_ov_ins, st_ov_notuseful, st_ov_exact, st_ov_unknown;

static inline int BY(int a) { return a < M.N ? M.seg[a] : M.seg[a - 32768]; }   /* stale bytes past the end */
static int cp258(int a, int c)
{
    int n = 0;
    while (n < MAX_MATCH && BY(a + n) == BY(c + n)) n++;
    return n;
}

static void kernel_a(const zso_deflater* s)
{
    static int head_all[32768];
    for (int i = 0; i < 32768; i++) head_all[i] = -1;
    for (int a = 0; a < M.N; a++) {
        uint32_t v = (uint32_t)BY(a + 2) | ((uint32_t)BY(a + 3) << 8) | ((uint32_t)BY(a + 4) << 16) | ((uint32_t)BY(a + 5) << 24);
        uint32_t h = update_hash(s, v) & s->hash_mask;
        M.hash[a] = (uint16_t)h;
        M.prev_all[a] = head_all[h];
        head_all[h] = a;
        int c = M.prev_all[a], k = 0;
        while (c >= 0 && a - c <= (k == 0 ? MAXD : MAXD - 1) && k < K) {
            M.dist[a][k] = (uint16_t)(a - c);
            M.len[a][k] = (uint16_t)cp258(a, c);
            k++;
            c = M.prev_all[c];
        }
        M.nk[a] = (uint8_t)k;
        M.more[a] = (uint8_t)(c >= 0 && a - c <= MAXD);
    }
}

/* kernel B's decision at parse point p with la bytes of lookahead: returns the match length (< 3: literal) */
static int resolve_at(const zso_deflater* s, int p, int la, int* mstart)
{
    if (p == M.forced) return 2;
    int cnt = 0, best = 2, sel = -1;
    const int nice = s->nice_match < la ? s->nice_match : la;
    int c = -1, L = 0, j = 0, slow = 0, ov_anchor = -1, ov_left = 0;
    st_search++;
    for (;;) {
        if (j < M.nk[p]) { c = p - M.dist[p][j]; L = M.len[p][j]; j++; }
        else if (j == M.nk[p] && !M.more[p]) break;
        else {   /* overflow: follow the all-position chain beyond the K-th node */
            if (!slow) { st_overflow++; slow = 1; }
            st_over_nodes++;
            if (j != KMAX + 1) { ov_anchor = c; ov_left = K; }           /* first step: the anchor is the record's last node */
            else if (ov_left == 0) { ov_anchor = c; ov_left = K; }       /* next round: the anchor is the previous round's last node */
            c = M.prev_all[c];
            if (c < 0 || p - c > MAXD) break;
            ov_left--;
            L = cp258(p, c);
            if (M.ins[c]) {   /* what kernel B could tell from the records alone: lcp(p,c) vs min(lcp(p,anchor), lcp(anchor,c)) */
                int lpa = cp258(p, ov_anchor), lac = cp258(ov_anchor, c);
                st_ov_ins++;
                if ((lpa >= 3 && lac < 3) || (lpa < 3 && lac >= 3)) st_ov_notuseful++;
                else if (lpa >= 3 && lac >= 3 && lpa != lac) st_ov_exact++;
                else st_ov_unknown++;
            }
            j = KMAX + 1;
        }
        if (!M.ins[c]) continue;
        const int d = p - c;
        if (cnt == 0 ? d > MAXD : d >= MAXD) break;
        if (M.first && c == 0) break;   /* window index 0 is NIL */
        cnt++;
        if (L > best) { best = L; sel = c; if (L >= nice) break; }
        if (c == M.cut0 || c == M.cut1) break;
        if (cnt == s->max_chain) break;
    }
    if (sel >= 0) *mstart = sel;
    return best < la ? best : la;
}

static int deflate_fast_ab(zso_deflater* s, int flush)
{
    if (s->strategy == ZSO_HUFFMAN_ONLY) return deflate_fast(s, flush);
    if (M.fresh) {
        M.fresh = 0;
        M.seg = s->next_in;
        M.N = s->avail_in;
        kernel_a(s);
        memset(M.ins, 0, sizeof M.ins);
        M.base = M.first ? 0 : -32768;
        M.loaded_end = 0;
        M.cut0 = M.cut1 = M.forced = -1;
        M.blk_start = 0;
    }
    for (;;) {
        int bflush;
        const int p = s->str_start + M.base;   /* absolute position (before a possible slide: base moves with it) */
        if (s->lookahead < MIN_LOOKAHEAD) {
            int av = s->avail_in, ss = s->str_start;
            /* B's own prediction of what Fill_window does, checked against the oracle's Fill_window */
            const int will_slide = (p - M.base) >= 65536 - MIN_LOOKAHEAD;
            const int will_read = M.loaded_end < M.N;
            fill_window(s);
            if ((s->str_start < ss) != will_slide) { fprintf(stderr, "model: slide prediction wrong at %d\n", p); abort(); }
            if ((av != s->avail_in) != will_read) { fprintf(stderr, "model: refill prediction wrong at %d\n", p); abort(); }
            if (will_slide) M.base += 32768;
            if (will_read) {
                M.loaded_end += av - s->avail_in;
                if (s->lookahead >= MIN_MATCH) {   /* Q3 */
                    M.ins[p + 1] = 1;
                    if (M.hash[p] == M.hash[p + 1]) { M.forced = p; M.cut1 = M.cut0; M.cut0 = p; }
                    else M.forced = p + 1;
                }
            }
            if (s->str_start + M.base != p) { fprintf(stderr, "model: base tracking wrong\n"); abort(); }
            if (s->lookahead < MIN_LOOKAHEAD && flush == ZSO_NO_FLUSH) return BS_NEED_MORE;
            if (s->lookahead == 0) break;
        }
        const int la = M.N - p;
        if (la != s->lookahead + (M.N - M.loaded_end)) { fprintf(stderr, "model: lookahead tracking wrong\n"); abort(); }
        int mlen = 2, mstart = 0;
        st_pp++;
        if (la >= MIN_MATCH) {
            M.ins[p] = 1;
            mlen = resolve_at(s, p, la, &mstart);
        }
        if (mlen >= MIN_MATCH) {
            bflush = tally_dist(s, p - mstart, mlen - MIN_MATCH);
            s->lookahead -= mlen;
            if (mlen <= s->max_lazy && la - mlen >= MIN_MATCH)
                for (int q = p + 1; q < p + mlen; q++) M.ins[q] = 1;
            s->str_start += mlen;
            s->match_length = 0;
        } else {
            bflush = tally_lit(s, s->window[s->str_start]);
            s->lookahead--;
            s->str_start++;
        }
        if (bflush) {
            if (s->block_start + M.base != M.blk_start) { fprintf(stderr, "model: block start tracking wrong\n"); abort(); }
            flush_block_only(s, 0);
            M.blk_start = s->str_start + M.base;
            if (s->avail_out == 0) return BS_NEED_MORE;
        }
    }
    flush_block_only(s, flush == ZSO_FINISH);
    return finish_state(s, flush);
}

static uint8_t* run(const uint8_t* in, size_t n, size_t seg, int level, int use_model, size_t* out_len)
{
    zso_deflater* d = zso_deflater_new(level, 0, ZSO_FULL_FLUSH, 0);
    g_fast_hook = use_model ? deflate_fast_ab : NULL;
    size_t k = 0;
    for (size_t off = 0; off < n; off += seg, k++) {
        size_t len = n - off < seg ? n - off : seg;
        M.fresh = 1;
        M.first = k == 0;
        /* the model covers full segments whose size is a multiple of 32 KiB and >= 64 KiB; others take the literal loop */
        if (use_model && (len != seg || seg % 32768 != 0 || seg < 65536)) g_fast_hook = NULL;
        if (zso_deflater_write(d, in + off, len) != 0) { fprintf(stderr, "write failed\n"); exit(2); }
    }
    if (zso_deflater_finish(d) != 0) exit(2);
    uint8_t* out = malloc(d->out_len);
    memcpy(out, d->out, d->out_len);
    *out_len = d->out_len;
    zso_deflater_free(d);
    return out;
}

/* The 32-byte records of kernel A for one segment (tools/l1_diag.py diffs the device's records against these).
 * Build as a library:  gcc -O2 -shared -fPIC -DABM_LIB -o tools/ab_model/libabmodel.so tools/ab_model/ab_model.c oracle/zso_adler32.c */
int abm_records(const uint8_t* seg, int n, int hash_variant, uint8_t* out)
{
    zso_deflater* d = zso_deflater_new(1, 0, ZSO_FULL_FLUSH, hash_variant);
    if (!d || n > MAXSEG) return -1;
    K = 10;
    M.seg = seg;
    M.N = n;
    kernel_a(d);
    for (int a = 0; a < n; a++) {
        uint8_t* r = out + (size_t)a * 32;
        memset(r, 0, 32);
        unsigned usem = 0;
        for (int k = 0; k < M.nk[a]; k++) {
            r[2 * k] = (uint8_t)M.dist[a][k];
            r[2 * k + 1] = (uint8_t)(M.dist[a][k] >> 8);
            r[20 + k] = (uint8_t)(M.len[a][k] >= 3 ? M.len[a][k] - 3 : 0);
            if (M.len[a][k] >= 3) usem |= 1u << k;
        }
        /* bytes 30..31: ten flags "prefix >= 3", "the chain goes on", the number of candidates */
        unsigned hi = usem | ((unsigned)M.more[a] << 10) | ((unsigned)M.nk[a] << 11);
        r[30] = (uint8_t)hi;
        r[31] = (uint8_t)(hi >> 8);
    }
    zso_deflater_free(d);
    return 0;
}

#ifndef ABM_LIB
int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: ab_model file seg_size [K] [level]\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    fseek(f, 0, SEEK_END);
    size_t n = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t* in = malloc(n + 1);
    if (fread(in, 1, n, f) != n) return 2;
    fclose(f);
    size_t seg = (size_t)atol(argv[2]);
    if (argc > 3) K = atoi(argv[3]);
    int level = argc > 4 ? atoi(argv[4]) : 1;
    size_t la, lb;
    uint8_t* a = run(in, n, seg, level, 0, &la);
    uint8_t* b = run(in, n, seg, level, 1, &lb);
    int same = la == lb && memcmp(a, b, la) == 0;
    printf("%s seg=%zu L%d K=%d: ref %zu B, model %zu B: %s;  searches/pp %.3f overflow/search %.4f nodes/overflow %.1f\n", argv[1], seg,
           level, K, la, lb, same ? "IDENTICAL" : "DIFFERENT", (double)st_search / (double)(st_pp ? st_pp : 1),
           (double)st_overflow / (double)(st_search ? st_search : 1), (double)st_over_nodes / (double)(st_overflow ? st_overflow : 1));
    printf("  overflow walks: inserted nodes %llu: not useful by the min rule %.3f, exact by the min rule %.3f, need bytes %.3f\n",
           (unsigned long long)st_ov_ins, (double)st_ov_notuseful / (double)(st_ov_ins ? st_ov_ins : 1),
           (double)st_ov_exact / (double)(st_ov_ins ? st_ov_ins : 1), (double)st_ov_unknown / (double)(st_ov_ins ? st_ov_ins : 1));
    return same ? 0 : 1;
}
#endif
