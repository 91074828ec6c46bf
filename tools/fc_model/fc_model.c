/*
 * fc_model.c -- CPU model of the batch DeflateFast used by the CUDA kernel (zsb_lz77.cu, level 1).
 *
 * DEVELOPMENT TOOL, not product code and not part of the oracle: it includes the oracle's source, swaps its DeflateFast
 * for the restructured parse through the oracle's test hook, and checks on real files that both produce the same bytes.
 *
 * The restructured parse (what the kernel does, phase by phase, here with plain loops over "lanes"):
 *   - phase A links EVERY position of the batch into head/prev in position order ("temporary links"), by rounds of slot
 *     arbitration (the first lane of a slot wins a round); after R_MAX rounds the batch is truncated at the first
 *     unresolved lane and the later links are undone.  Each lane now knows its predecessor c0 in the full chain.
 *   - phase B: every lane collects up to A_MAX in-batch nodes (whether the reference inserts them is not known yet),
 *     then the first 4 nodes of the real chain as it stood before the batch; 8-byte compares (nice_match = 8 ends the
 *     reference's scan).  A lane whose outcome does not depend on the unknown statuses is "clean"; otherwise it is
 *     "ambiguous" (resolved when the parse reaches it) or "hard" (more than A_MAX in-batch nodes: exact search then).
 *   - walk: the greedy parse over the batch.
 *   - commit: the temporary links are replaced by the reference's: every inserted lane points to its nearest inserted
 *     in-batch predecessor (or to the pre-batch head), head[h] to the last inserted lane of h.
 *
 * Build / run:  gcc -O2 -o /tmp/fc_model tools/fc_model/fc_model.c oracle/zso_adler32.c && /tmp/fc_model <file> <seg_size>
 */
#include "../../oracle/zso_deflate.c"

#include <stdio.h>

#ifndef NB
#define NB 128
#endif
#ifndef A_MAX
#define A_MAX 3
#endif
#ifndef R_MAX
#define R_MAX 4
#endif
#define SLOTS 2048
#define MAX_DIST_(s) ((s)->w_size - MIN_LOOKAHEAD)

static int g_lit_until;
static uint64_t c_batches, c_lanes, c_consumed, c_amb, c_amb_hit, c_hard, c_hard_hit, c_trunc, c_lit_steps, c_rounds, c_unamb;

static int prefix_cap(const uint8_t* a, const uint8_t* b, int cap)
{
    int n = 0;
    while (n < cap && a[n] == b[n]) n++;
    return n;
}

static inline uint32_t fc_hash(const zso_deflater* s, int p)
{
    return update_hash(s, load_le32(s->window + p + (MIN_MATCH - 1))) & s->hash_mask;
}

/* One iteration of Deflate.Fast.cs:43-112.  Returns bflush. */
static int fc_literal_step(zso_deflater* s)
{
    int bflush, hash_head = 0;
    c_lit_steps++;
    if (s->lookahead >= MIN_MATCH) hash_head = insert_string(s, s->str_start);
    if (hash_head != 0 && (s->str_start - hash_head) <= MAX_DIST_(s)) s->match_length = longest_match(s, hash_head);
    if (s->match_length >= MIN_MATCH) {
        bflush = tally_dist(s, s->str_start - s->match_start, s->match_length - MIN_MATCH);
        s->lookahead -= s->match_length;
        if (s->match_length <= s->max_lazy && s->lookahead >= MIN_MATCH) {
            s->match_length--;
            do {
                s->str_start++;
                insert_string(s, s->str_start);
            } while (--s->match_length != 0);
            s->str_start++;
        } else {
            s->str_start += s->match_length;
            s->match_length = 0;
        }
    } else {
        bflush = tally_lit(s, s->window[s->str_start]);
        s->lookahead--;
        s->str_start++;
    }
    return bflush;
}

typedef struct {
    int nU, nK;
    int Uidx[A_MAX], Ulen[A_MAX];
    int Kpos[4], Klen[4], Kedge[4];
    int base;         /* first chain node before the batch (the real head[h] when the batch began) */
    int hard;
    int clean;        /* outcome independent of the unknown statuses */
} lane_rec;

/* selection of Longest_match given the statuses of the lane's in-batch nodes; returns candidate position or 0 */
static int lane_select(const lane_rec* r, int S, unsigned umask, int* len8)
{
    int cnt = 0, best = MIN_MATCH - 1, sel = 0;
    for (int j = 0; j < r->nU; j++) {
        if (!((umask >> j) & 1)) continue;
        cnt++;
        if (r->Ulen[j] > best) { best = r->Ulen[j]; sel = S + r->Uidx[j]; if (best >= 8) goto done; }
    }
    for (int k = 0; k < r->nK; k++) {
        if (cnt == 4) break;
        if (r->Kedge[k] && cnt > 0) break;
        cnt++;
        if (r->Klen[k] > best) { best = r->Klen[k]; sel = r->Kpos[k]; if (best >= 8) goto done; }
    }
done:
    *len8 = best;
    return sel;
}

/* Longest_match for a hard lane: in-batch nodes (temporary links) count only if inserted, then the real chain. */
static int hard_search(zso_deflater* s, int S, int p, int node, const uint8_t* status, int* mstart)
{
    const int max_dist = MAX_DIST_(s);
    int best = MIN_MATCH - 1, chain = s->max_chain, first = 1;
    const int nice = s->nice_match;   /* lookahead >= MIN_LOOKAHEAD for every lane of a batch */
    const int limit = p > max_dist ? p - max_dist : 0;
    for (;;) {
        while (node >= S && node != 0 && !status[node - S]) node = s->prev[node & s->w_mask];
        if (node == 0) break;
        if (first) { if (p - node > max_dist) break; }
        else if (!(node > limit)) break;
        if (node >= p) break;
        int len = prefix_cap(s->window + p, s->window + node, MAX_MATCH);
        if (len > best) { best = len; *mstart = node; if (len >= nice) break; }
        first = 0;
        if (--chain == 0) break;
        node = s->prev[node & s->w_mask];
    }
    return best;
}

static int fc_batch(zso_deflater* s, int nact)
{
    const int S = s->str_start, max_dist = MAX_DIST_(s);
    static uint32_t slot[SLOTS];
    uint32_t h[NB];
    int c0[NB], linked[NB];
    lane_rec rec[NB];
    c_batches++;
    /* ---- phase A ---- */
    for (int i = 0; i < nact; i++) { h[i] = fc_hash(s, S + i); linked[i] = 0; }
    int unresolved = nact;
    for (int r = 0; r < R_MAX && unresolved; r++) {
        c_rounds++;
        for (int i = 0; i < SLOTS; i++) slot[i] = 0xFFFFFFFFu;
        for (int i = nact - 1; i >= 0; i--) if (!linked[i]) slot[h[i] & (SLOTS - 1)] = (uint32_t)i;   /* min index wins */
        for (int i = 0; i < nact; i++) {
            if (linked[i] || slot[h[i] & (SLOTS - 1)] != (uint32_t)i) continue;
            int p = S + i;
            c0[i] = s->head[h[i]];
            if (c0[i] >= p) { fprintf(stderr, "model: c0 >= p in a batch\n"); abort(); }
            s->prev[p & s->w_mask] = (uint16_t)c0[i];
            s->head[h[i]] = (uint16_t)p;
            linked[i] = 1;
            unresolved--;
        }
    }
    if (unresolved) {
        int m = 0;
        while (linked[m]) m++;
        for (int i = m; i < nact; i++)
            if (linked[i] && c0[i] < S + m) s->head[h[i]] = (uint16_t)c0[i];
        nact = m;
        c_trunc++;
    }
    c_lanes += (uint64_t)nact;
    /* ---- phase B ---- */
    for (int i = 0; i < nact; i++) {
        lane_rec* r = &rec[i];
        const int p = S + i;
        const int limit_lo = p - max_dist;
        memset(r, 0, sizeof *r);
        int cur = c0[i];
        while (cur >= S && cur != 0) {
            if (r->nU == A_MAX) { r->hard = 1; break; }
            r->Uidx[r->nU] = cur - S;
            r->Ulen[r->nU] = prefix_cap(s->window + p, s->window + cur, 8);
            r->nU++;
            cur = s->prev[cur & s->w_mask];
        }
        if (r->hard) {   /* the base is still needed by the commit */
            while (cur >= S && cur != 0) cur = s->prev[cur & s->w_mask];
            r->base = cur;
            c_hard++;
            continue;
        }
        r->base = cur;
        while (cur != 0 && r->nK < 4) {
            if (cur < limit_lo) break;
            int L = prefix_cap(s->window + p, s->window + cur, 8);
            r->Kpos[r->nK] = cur;
            r->Klen[r->nK] = L;
            r->Kedge[r->nK] = cur == limit_lo;
            r->nK++;
            if (L >= 8 || cur == limit_lo) break;
            if (cur >= p) break;   /* unreachable here: every pre-batch node is below S */
            cur = s->prev[cur & s->w_mask];
        }
        int l0;
        int s0 = lane_select(r, S, 0, &l0);
        r->clean = 1;
        for (unsigned um = 1; um < (1u << r->nU); um++) {
            int l1;
            if (lane_select(r, S, um, &l1) != s0) r->clean = 0;
        }
        if (r->nU && r->clean) c_unamb++;
        if (!r->clean) c_amb++;
    }
    /* ---- walk ---- */
    int cur = 0, bflush = 0, last_short_end = 0;
    uint8_t status[NB];
    memset(status, 0, sizeof status);
    while (cur < nact) {
        lane_rec* r = &rec[cur];
        const int p = S + cur;
        int mpos = 0, len = MIN_MATCH - 1;
        if (r->hard) {
            c_hard_hit++;
            len = hard_search(s, S, p, c0[cur], status, &mpos);
        } else {
            unsigned um = 0;
            if (!r->clean) c_amb_hit++;
            for (int j = 0; j < r->nU; j++) um |= (unsigned)status[r->Uidx[j]] << j;
            int l8;
            mpos = lane_select(r, S, um, &l8);
            if (mpos) len = prefix_cap(s->window + p, s->window + mpos, MAX_MATCH);
        }
        if (len >= MIN_MATCH) {
            bflush = tally_dist(s, p - mpos, len - MIN_MATCH);
            status[cur] = 1;
            const int is_short = len <= s->max_lazy;
            for (int j = 1; j < len; j++)
                if (cur + j < nact) status[cur + j] = (uint8_t)is_short;
            if (is_short && cur + len > nact) last_short_end = S + cur + len;
            cur += len;
        } else {
            bflush = tally_lit(s, s->window[p]);
            status[cur] = 1;
            cur++;
        }
    }
    /* ---- commit: the reference's links instead of the temporary ones ---- */
    for (int i = 0; i < nact; i++) {
        const int p = S + i;
        /* nearest inserted in-batch predecessor, else the base */
        int node = c0[i];
        while (node >= S && node != 0 && !status[node - S]) node = c0[node - S];
        if (status[i]) s->prev[p & s->w_mask] = (uint16_t)node;
        if (s->head[h[i]] == (uint16_t)p && !status[i]) s->head[h[i]] = (uint16_t)node;   /* last lane of its hash */
    }
    /* interior of a short match past the linked region: InsertString */
    for (int q = S + nact; q < last_short_end; q++) insert_string(s, q);
    s->str_start = S + cur;
    s->lookahead -= cur;
    s->match_length = 0;
    c_consumed += (uint64_t)cur;
    return bflush;
}

static int deflate_fast_fc(zso_deflater* s, int flush)
{
    if (s->strategy == ZSO_HUFFMAN_ONLY || s->level != 1) return deflate_fast(s, flush);
    for (;;) {
        int bflush;
        if (s->lookahead < MIN_LOOKAHEAD) {
            int before = s->avail_in;
            fill_window(s);
            if (before != s->avail_in) g_lit_until = s->str_start + 2;   /* a refill pre-inserted str_start + 1 (Q3) */
            if (s->lookahead < MIN_LOOKAHEAD && flush == ZSO_NO_FLUSH) return BS_NEED_MORE;
            if (s->lookahead == 0) break;
        }
        int nact = s->lookahead - (MIN_LOOKAHEAD - 1);
        if (nact > NB) nact = NB;
        if (nact > s->lit_bufsize - 1 - s->last_lit) nact = s->lit_bufsize - 1 - s->last_lit;
        if (nact >= 1 && s->str_start >= g_lit_until) bflush = fc_batch(s, nact);
        else bflush = fc_literal_step(s);
        if (bflush) {
            flush_block_only(s, 0);
            if (s->avail_out == 0) return BS_NEED_MORE;
        }
    }
    flush_block_only(s, flush == ZSO_FINISH);
    return finish_state(s, flush);
}

static uint8_t* run(const uint8_t* in, size_t n, size_t seg, int level, int use_fc, size_t* out_len)
{
    size_t cap = zso_deflate_bound(n, seg);
    uint8_t* out = malloc(cap);
    uint32_t ad;
    g_fast_hook = use_fc ? deflate_fast_fc : NULL;
    g_lit_until = 0;
    int rc = seg ? zso_deflate_segments(in, n, seg, level, 0, 0, out, cap, out_len, NULL, &ad)
                 : zso_deflate_oneshot(in, n, level, 0, 0, out, cap, out_len, &ad);
    if (rc != 0) { fprintf(stderr, "rc %d\n", rc); exit(2); }
    return out;
}

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: fc_model file seg_size [level]\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    fseek(f, 0, SEEK_END);
    size_t n = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t* in = malloc(n + 1);
    if (fread(in, 1, n, f) != n) return 2;
    fclose(f);
    size_t seg = (size_t)atol(argv[2]);
    int level = argc > 3 ? atoi(argv[3]) : 1;
    size_t la, lb;
    uint8_t* a = run(in, n, seg, level, 0, &la);
    uint8_t* b = run(in, n, seg, level, 1, &lb);
    int same = la == lb && memcmp(a, b, la) == 0;
    printf("%s seg=%zu L%d: ref %zu B, fc %zu B: %s\n", argv[1], seg, level, la, lb, same ? "IDENTICAL" : "DIFFERENT");
    printf("  batches %llu lanes/batch %.1f consumed/batch %.1f rounds/batch %.2f truncated %.3f\n",
           (unsigned long long)c_batches, (double)c_lanes / (double)(c_batches ? c_batches : 1),
           (double)c_consumed / (double)(c_batches ? c_batches : 1), (double)c_rounds / (double)(c_batches ? c_batches : 1),
           (double)c_trunc / (double)(c_batches ? c_batches : 1));
    printf("  per batch: ambiguous lanes %.2f (reached as parse point %.2f), in-batch-but-clean %.2f, hard %.2f (reached %.2f); literal steps %.4f per byte\n",
           (double)c_amb / (double)c_batches, (double)c_amb_hit / (double)c_batches, (double)c_unamb / (double)c_batches,
           (double)c_hard / (double)c_batches, (double)c_hard_hit / (double)c_batches, (double)c_lit_steps / (double)n);
    return same ? 0 : 1;
}
