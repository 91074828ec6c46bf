/*
 * ab_stats.c -- DEVELOPMENT TOOL (not product, not oracle): statistics for the "all-positions chain + inserted
 * bitmap" formulation of DeflateFast.  Includes the oracle's source and replaces its DeflateFast through the test hook
 * with a copy that also keeps a chain over EVERY position, and counts how many all-position nodes a filtered walk has
 * to look at before the reference's Longest_match outcome is certain.
 *
 * gcc -O2 -o /tmp/ab_stats tools/ab_model/ab_stats.c oracle/zso_adler32.c && /tmp/ab_stats file seg_size [level]
 */
#include "../../oracle/zso_deflate.c"
#include <stdio.h>

#define MAXSEG (1 << 20)
static int ap_head[32768];
static int ap_prev[MAXSEG + 70000];   /* by absolute segment position + 32768 */
static uint8_t ins_bit[MAXSEG + 70000];
static int abs_base;                  /* absolute position (+32768) of window index 0 */
static int ap_done;                   /* all positions < ap_done (absolute) are linked */
static uint64_t hist[66], n_search, n_end_nice, n_end_budget, n_end_chain, sum_need, n_pp, n_inserted, n_pos;
static uint64_t hist_budget[66], hist_chain[66], hist_nice[66];

static void ap_link_upto(zso_deflater* s, int abs_to)   /* link absolute positions < abs_to */
{
    for (; ap_done < abs_to; ap_done++) {
        int idx = ap_done - abs_base;
        uint32_t h = update_hash(s, load_le32(s->window + idx + 2)) & s->hash_mask;
        ap_prev[ap_done] = ap_head[h];
        ap_head[h] = ap_done;
    }
}

static int deflate_fast_stats(zso_deflater* s, int flush)
{
    /* new segment? (head zeroed by the full flush: detect by block_start == str_start and lookahead == 0 at entry) */
    for (;;) {
        int bflush;
        if (s->lookahead < MIN_LOOKAHEAD) {
            int ss = s->str_start, av = s->avail_in;
            fill_window(s);
            if (s->str_start < ss) abs_base += s->w_size;
            if (av != s->avail_in && s->lookahead >= MIN_MATCH) ins_bit[s->str_start + 1 + abs_base] = 1;   /* Q3 */
            if (s->lookahead < MIN_LOOKAHEAD && flush == ZSO_NO_FLUSH) return BS_NEED_MORE;
            if (s->lookahead == 0) break;
        }
        int hash_head = 0;
        const int p = s->str_start, pabs = p + abs_base;
        ap_link_upto(s, pabs);
        n_pp++;
        if (s->lookahead >= MIN_MATCH) { hash_head = insert_string(s, s->str_start); ins_bit[pabs] = 1; }
        if (hash_head != 0 && (s->str_start - hash_head) <= s->w_size - MIN_LOOKAHEAD) {
            /* emulate: walk the all-position chain */
            if (s->strategy != ZSO_HUFFMAN_ONLY) {
                const int max_dist = s->w_size - MIN_LOOKAHEAD;
                int need = 0, cnt = 0, best = 2, why = 2;
                uint32_t h = update_hash(s, load_le32(s->window + p + 2)) & s->hash_mask;
                int node = ap_head[h];
                while (node != 0) {
                    int nidx = node - abs_base;
                    if (nidx <= 0) break;
                    if (p - nidx > max_dist || (cnt > 0 && p - nidx >= max_dist)) break;
                    need++;
                    if (ins_bit[node]) {
                        cnt++;
                        int len = 0;
                        while (len < 258 && s->window[p + len] == s->window[nidx + len]) len++;
                        if (len > best) { best = len; if (len >= s->nice_match) { why = 0; break; } }
                        if (cnt == s->max_chain) { why = 1; break; }
                    }
                    node = ap_prev[node];
                }
                n_search++;
                sum_need += need;
                hist[need > 64 ? 65 : need]++;
                if (why == 0) { n_end_nice++; hist_nice[need > 64 ? 65 : need]++; }
                else if (why == 1) { n_end_budget++; hist_budget[need > 64 ? 65 : need]++; }
                else { n_end_chain++; hist_chain[need > 64 ? 65 : need]++; }
                s->match_length = longest_match(s, hash_head);
            }
        }
        if (s->match_length >= MIN_MATCH) {
            bflush = tally_dist(s, s->str_start - s->match_start, s->match_length - MIN_MATCH);
            s->lookahead -= s->match_length;
            if (s->match_length <= s->max_lazy && s->lookahead >= MIN_MATCH) {
                s->match_length--;
                do {
                    s->str_start++;
                    ins_bit[s->str_start + abs_base] = 1;
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
        if (bflush) {
            flush_block_only(s, 0);
            if (s->avail_out == 0) return BS_NEED_MORE;
        }
    }
    flush_block_only(s, flush == ZSO_FINISH);
    return finish_state(s, flush);
}

int main(int argc, char** argv)
{
    FILE* f = fopen(argv[1], "rb");
    fseek(f, 0, SEEK_END);
    size_t n = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t* in = malloc(n + 1);
    if (fread(in, 1, n, f) != n) return 2;
    size_t seg = (size_t)atol(argv[2]);
    int level = argc > 3 ? atoi(argv[3]) : 1;
    g_fast_hook = deflate_fast_stats;
    size_t cap = zso_deflate_bound(seg, seg);
    uint8_t* out = malloc(cap);
    for (size_t off = 0; off < n; off += seg) {
        size_t len = n - off < seg ? n - off : seg, ol;
        uint32_t ad;
        memset(ap_head, 0, sizeof ap_head);
        memset(ins_bit, 0, sizeof ins_bit);
        abs_base = 32768;   /* stream-first segment: window index 0 = absolute 32768 (we run every segment as a fresh stream) */
        ap_done = 32768;
        zso_deflate_segments(in + off, len, seg, level, 0, 0, out, cap, &ol, NULL, &ad);
        for (size_t i = 0; i < len; i++) n_inserted += ins_bit[32768 + i];
        n_pos += len;
    }
    printf("%s seg=%zu L%d: parse points/B %.3f inserted/B %.3f searches/B %.3f avg need %.2f\n", argv[1], seg, level,
           (double)n_pp / n_pos, (double)n_inserted / n_pos, (double)n_search / n_pos, (double)sum_need / n_search);
    printf("  end by nice %.3f budget %.3f chain-end %.3f\n", (double)n_end_nice / n_search, (double)n_end_budget / n_search, (double)n_end_chain / n_search);
    uint64_t c = 0;
    printf("  P(need > K): ");
    for (int k = 0; k <= 64; k++) {
        c += hist[k];
        if (k == 4 || k == 6 || k == 8 || k == 10 || k == 12 || k == 16 || k == 24 || k == 32 || k == 64) printf("K=%d: %.4f  ", k, 1.0 - (double)c / n_search);
    }
    printf("\n");
    return 0;
}
