/*
 * zso_deflate.c -- CPU ORACLE (test infrastructure only, see zso.h): restatement of the
 * reference's managed deflate as one explicit state machine.
 *
 * Every function names the reference code it restates (paths relative to
 * /root/reference/src/ZlibStream/).  The reference's deflate is its own dialect (zlib 1.1.3
 * skeleton + zlib-ng style pieces); it is NOT bit-compatible with stock zlib, so every quirk
 * below is deliberate.
 */
#include "zso.h"

#include <stdlib.h>
#include <string.h>

/* ---- constants (Deflate.cs:17-78, Trees.cs:15-70) ---------------------------------- */
#define MIN_MATCH 3
#define MAX_MATCH 258
#define MIN_LOOKAHEAD (MAX_MATCH + MIN_MATCH + 1)
#define MAX_BITS 15
#define MAX_BL_BITS 7
#define LITERALS 256
#define END_BLOCK 256
#define LENGTH_CODES 29
#define L_CODES (LITERALS + 1 + LENGTH_CODES)
#define D_CODES 30
#define BL_CODES 19
#define HEAP_SIZE (2 * L_CODES + 1)
#define REP_3_6 16
#define REPZ_3_10 17
#define REPZ_11_138 18

#define ST_INIT 42
#define ST_BUSY 113
#define ST_FINISH 666

enum { BS_NEED_MORE = 0, BS_BLOCK_DONE = 1, BS_FINISH_STARTED = 2, BS_FINISH_DONE = 3 };
enum { FN_STORED = 0, FN_FAST = 1, FN_SLOW = 2 };

#define CHUNK 512 /* ZlibOutputStream.cs:53 */

typedef struct { uint16_t good, lazy, nice, chain; int fn; } level_cfg;

/* Deflate.cs:80-98 (USE_QUICK is not defined: level 1 is FAST) */
static const level_cfg k_cfg[10] = {
    {0, 0, 0, 0, FN_STORED},      {4, 4, 8, 4, FN_FAST},        {4, 5, 16, 8, FN_FAST},
    {4, 6, 32, 32, FN_FAST},      {4, 4, 16, 16, FN_SLOW},      {8, 16, 32, 32, FN_SLOW},
    {8, 16, 128, 128, FN_SLOW},   {8, 32, 128, 256, FN_SLOW},   {32, 128, 258, 1024, FN_SLOW},
    {32, 258, 258, 4096, FN_SLOW}};

/* Trees.cs:85-124 */
static const uint8_t k_extra_lbits[LENGTH_CODES] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2,
                                                    2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint8_t k_extra_dbits[D_CODES] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4,  4,  5,  5,  6,
                                               6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t k_extra_blbits[BL_CODES] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 3, 7};
static const uint8_t k_bl_order[BL_CODES] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

/* Derived tables (Trees.cs:75-183 hold them as literals; rebuilt here from the RFC 1951 code layout). */
static uint8_t g_length_code[256];
static uint8_t g_base_length[LENGTH_CODES];
static uint8_t g_dist_code[512];
static uint16_t g_base_dist[D_CODES];

typedef struct { uint16_t fc; /* Freq | Code */ uint16_t dl; /* Dad | Len */ } code_data; /* Trees.Static.cs:98-124 */

static code_data g_static_ltree[L_CODES + 2];
static code_data g_static_dtree[D_CODES];
static uint16_t g_crc_tab[4][256]; /* low 16 bits are all the 15-bit mask can see, but keep 32 for the API */
static uint32_t g_crc32c[256];
static int g_hw_crc; /* set once in the table initialiser: SSE4.2 present (x86 only) */
static int g_tables_ready = 0;

static unsigned bit_reverse(unsigned code, int len) /* Trees.cs:269-281 */
{
    unsigned r = 0;
    while (len-- > 0) {
        r = (r << 1) | (code & 1u);
        code >>= 1;
    }
    return r;
}

static void gen_codes(code_data* tree, int max_code, const uint16_t* bl_count) /* Trees.cs:1123-1151 */
{
    uint16_t next_code[MAX_BITS + 1];
    uint16_t code = 0;
    next_code[0] = 0;
    for (int bits = 1; bits <= MAX_BITS; bits++) {
        code = (uint16_t)((code + bl_count[bits - 1]) << 1);
        next_code[bits] = code;
    }
    for (int n = 0; n <= max_code; n++) {
        int len = tree[n].dl;
        if (len == 0) continue;
        tree[n].fc = (uint16_t)bit_reverse(next_code[len]++, len);
    }
}

static void build_static_tables(void)
{
    if (g_tables_ready) return;
    /* length codes: RFC 1951 3.2.5 */
    int length = 0;
    for (int code = 0; code < LENGTH_CODES - 1; code++) {
        g_base_length[code] = (uint8_t)length;
        for (int n = 0; n < (1 << k_extra_lbits[code]); n++) g_length_code[length++] = (uint8_t)code;
    }
    g_length_code[length - 1] = (uint8_t)(LENGTH_CODES - 1); /* length 258 has its own code 285 */
    g_base_length[LENGTH_CODES - 1] = 0;
    /* distance codes */
    int dist = 0, code = 0;
    for (; code < 16; code++) {
        g_base_dist[code] = (uint16_t)dist;
        for (int n = 0; n < (1 << k_extra_dbits[code]); n++) g_dist_code[dist++] = (uint8_t)code;
    }
    dist >>= 7;
    for (; code < D_CODES; code++) {
        g_base_dist[code] = (uint16_t)(dist << 7);
        for (int n = 0; n < (1 << (k_extra_dbits[code] - 7)); n++) g_dist_code[256 + dist++] = (uint8_t)code;
    }
    /* Trees.Static.cs:44-92 */
    uint16_t bl_count[MAX_BITS + 1];
    memset(bl_count, 0, sizeof bl_count);
    int n = 0;
    while (n <= 143) { g_static_ltree[n++].dl = 8; bl_count[8]++; }
    while (n <= 255) { g_static_ltree[n++].dl = 9; bl_count[9]++; }
    while (n <= 279) { g_static_ltree[n++].dl = 7; bl_count[7]++; }
    while (n <= 287) { g_static_ltree[n++].dl = 8; bl_count[8]++; }
    gen_codes(g_static_ltree, L_CODES + 1, bl_count);
    for (n = 0; n < D_CODES; n++) {
        g_static_dtree[n].dl = 5;
        g_static_dtree[n].fc = (uint16_t)bit_reverse((unsigned)n, 5);
    }
    /* CRC-32C (Castagnoli), reflected polynomial 0x82F63B78: what Sse42.Crc32 computes */
    for (unsigned i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
        g_crc32c[i] = c;
    }
    (void)g_crc_tab;
#ifdef __x86_64__
    g_hw_crc = __builtin_cpu_supports("sse4.2") ? 1 : 0;
#endif
    g_tables_ready = 1;
}

uint32_t zso_crc32c_u32(uint32_t crc, uint32_t v)
{
    build_static_tables();
    for (int i = 0; i < 4; i++) {
        crc = g_crc32c[(crc ^ v) & 0xFF] ^ (crc >> 8);
        v >>= 8;
    }
    return crc;
}

static inline int length_code(int lc) { return g_length_code[lc]; }              /* Trees.cs:218 */
static inline int dist_code(int d) { return g_dist_code[d < 256 ? d : 256 + (d >> 7)]; } /* Trees.cs:193 */

/* ---- state --------------------------------------------------------------------------- */

typedef struct {
    code_data* tree;
    int max_code;
} dyn_desc;

typedef struct {
    const code_data* stree; /* NULL for the bit-length tree */
    const uint8_t* extra;
    int extra_base, elems, max_length;
} static_desc;

static const static_desc k_l_desc = {g_static_ltree, k_extra_lbits, LITERALS + 1, L_CODES, MAX_BITS};
static const static_desc k_d_desc = {g_static_dtree, k_extra_dbits, 0, D_CODES, MAX_BITS};
static const static_desc k_bl_desc = {NULL, k_extra_blbits, 0, BL_CODES, MAX_BL_BITS};

struct zso_deflater {
    /* ZlibOutputStream */
    int flush_mode;
    int finished;
    uint8_t chunk[CHUNK];
    uint8_t* out;
    size_t out_len, out_cap;
    int oom;
    /* ZLibStream (ZlibStream.cs:34-99) */
    const uint8_t* next_in;
    int next_in_idx, avail_in;
    int64_t total_in;
    uint8_t* next_out;
    int next_out_idx, avail_out;
    int64_t total_out;
    const char* msg;
    uint32_t adler;
    /* Deflate (Deflate.cs:115-214) */
    int status, last_flush, no_header, level, strategy, hash_variant;
    int w_size, w_bits, w_mask, window_size;
    int hash_bits, hash_size;
    uint32_t hash_mask;
    int block_start, match_length, prev_match, match_available, str_start, match_start, lookahead, prev_length;
    int max_chain, max_lazy, good_match, nice_match;
    int last_eob_len, bi_valid;
    uint64_t bi_buf;
    int lit_bufsize, last_lit, d_buf, l_buf, matches;
    int pending, pending_out, pending_size;
    int opt_len, static_len, heap_len, heap_max;
    int slow_hash_head; /* Deflate.Slow.cs:20: hash_head lives outside the loop but inside the call */
    uint8_t* window;
    uint16_t* prev;
    uint16_t* head;
    uint8_t* pending_buf;
    code_data ltree[1024], dtree[128], bltree[64]; /* ArrayPool rent sizes for 573 / 61 / 39 */
    dyn_desc l_desc, d_desc, bl_desc;
    uint16_t bl_count[MAX_BITS + 1];
    int heap[HEAP_SIZE];
    uint8_t depth[HEAP_SIZE];
    /* instrumentation */
    zso_block_cb cb;
    void* cb_user;
    zso_counters cnt;
    uint16_t* tr_dist;
    uint8_t* tr_lc;
};

/* ---- pending buffer / bit output (Deflate.cs:640-705, 754-821) ----------------------- */

static inline void put_byte(zso_deflater* s, unsigned c) { s->pending_buf[s->pending++] = (uint8_t)c; }
static inline void put_le(zso_deflater* s, uint64_t w, int nbytes)
{
    for (int i = 0; i < nbytes; i++) put_byte(s, (unsigned)(w >> (8 * i)) & 0xFF);
}
static inline void put_short_msb(zso_deflater* s, unsigned b)
{
    put_byte(s, (b >> 8) & 0xFF);
    put_byte(s, b & 0xFF);
}

static void send_bits(zso_deflater* s, unsigned value, int length) /* Deflate.cs:799-821 */
{
    uint64_t val = value;
    int total = s->bi_valid + length;
    if (total < 64) {
        s->bi_buf |= val << s->bi_valid;
        s->bi_valid = total;
    } else if (s->bi_valid == 64) {
        put_le(s, s->bi_buf, 8);
        s->bi_buf = val;
        s->bi_valid = length;
    } else {
        s->bi_buf |= val << s->bi_valid;
        put_le(s, s->bi_buf, 8);
        s->bi_buf = val >> (64 - s->bi_valid);
        s->bi_valid = total - 64;
    }
}
static inline void send_code(zso_deflater* s, int c, const code_data* tree) { send_bits(s, tree[c].fc, tree[c].dl); }

static void bi_flush(zso_deflater* s) /* Deflate.cs:640-671 */
{
    if (s->bi_valid == 64) {
        put_le(s, s->bi_buf, 8);
        s->bi_buf = 0;
        s->bi_valid = 0;
        return;
    }
    if (s->bi_valid >= 32) { put_le(s, s->bi_buf, 4); s->bi_buf >>= 32; s->bi_valid -= 32; }
    if (s->bi_valid >= 16) { put_le(s, s->bi_buf, 2); s->bi_buf >>= 16; s->bi_valid -= 16; }
    if (s->bi_valid >= 8)  { put_le(s, s->bi_buf, 1); s->bi_buf >>= 8;  s->bi_valid -= 8; }
}

static void bi_windup(zso_deflater* s) /* Deflate.cs:675-705 */
{
    if (s->bi_valid > 56) {
        put_le(s, s->bi_buf, 8);
    } else {
        if (s->bi_valid > 24) { put_le(s, s->bi_buf, 4); s->bi_buf >>= 32; s->bi_valid -= 32; }
        if (s->bi_valid > 8)  { put_le(s, s->bi_buf, 2); s->bi_buf >>= 16; s->bi_valid -= 16; }
        if (s->bi_valid > 0)  { put_le(s, s->bi_buf, 1); }
    }
    s->bi_buf = 0;
    s->bi_valid = 0;
}

static void flush_pending(zso_deflater* s) /* Deflate.cs:828-854 */
{
    bi_flush(s);
    int len = s->pending;
    if (len > s->avail_out) len = s->avail_out;
    if (len == 0) return;
    memcpy(s->next_out + s->next_out_idx, s->pending_buf + s->pending_out, (size_t)len);
    s->next_out_idx += len;
    s->pending_out += len;
    s->total_out += len;
    s->avail_out -= len;
    s->pending -= len;
    if (s->pending == 0) s->pending_out = 0;
}

/* ---- trees (Trees.cs) ------------------------------------------------------------------ */

static void init_block(zso_deflater* s) /* Trees.cs:777-802 */
{
    for (int i = 0; i < L_CODES; i++) s->ltree[i].fc = 0;
    for (int i = 0; i < D_CODES; i++) s->dtree[i].fc = 0;
    for (int i = 0; i < BL_CODES; i++) s->bltree[i].fc = 0;
    s->ltree[END_BLOCK].fc = 1;
    s->opt_len = s->static_len = 0;
    s->last_lit = s->matches = 0;
}

static inline int node_smaller(const code_data* tree, int n, int m, const uint8_t* depth) /* Trees.cs:556 */
{
    return tree[n].fc < tree[m].fc || (tree[n].fc == tree[m].fc && depth[n] <= depth[m]);
}

static void pq_down_heap(zso_deflater* s, const code_data* tree, int k) /* Trees.cs:513-544 */
{
    int v = s->heap[k];
    int j = k << 1;
    while (j <= s->heap_len) {
        if (j < s->heap_len && node_smaller(tree, s->heap[j + 1], s->heap[j], s->depth)) j++;
        if (node_smaller(tree, v, s->heap[j], s->depth)) break;
        s->heap[k] = s->heap[j];
        k = j;
        j <<= 1;
    }
    s->heap[k] = v;
}

static void gen_bitlen(zso_deflater* s, dyn_desc* desc, const static_desc* st) /* Trees.cs:999-1109 */
{
    code_data* tree = desc->tree;
    int max_code = desc->max_code;
    int h, n, m, bits, xbits, overflow = 0;
    for (bits = 0; bits <= MAX_BITS; bits++) s->bl_count[bits] = 0;
    tree[s->heap[s->heap_max]].dl = 0; /* root */
    for (h = s->heap_max + 1; h < HEAP_SIZE; h++) {
        n = s->heap[h];
        bits = tree[tree[n].dl].dl + 1;
        if (bits > st->max_length) { bits = st->max_length; overflow++; }
        tree[n].dl = (uint16_t)bits;
        if (n > max_code) continue; /* internal node */
        s->bl_count[bits]++;
        xbits = (n >= st->extra_base) ? st->extra[n - st->extra_base] : 0;
        unsigned f = tree[n].fc;
        s->opt_len += (int)(f * (unsigned)(bits + xbits));
        if (st->stree) s->static_len += (int)(f * (unsigned)(st->stree[n].dl + xbits));
    }
    if (overflow == 0) return;
    do {
        bits = st->max_length - 1;
        while (s->bl_count[bits] == 0) bits--;
        s->bl_count[bits]--;
        s->bl_count[bits + 1] += 2;
        s->bl_count[st->max_length]--;
        overflow -= 2;
    } while (overflow > 0);
    for (bits = st->max_length; bits != 0; bits--) {
        n = s->bl_count[bits];
        while (n != 0) {
            m = s->heap[--h];
            if (m > max_code) continue;
            if (tree[m].dl != (unsigned)bits) {
                s->opt_len += bits * (int)tree[m].fc;
                s->opt_len -= (int)tree[m].dl * (int)tree[m].fc;
                tree[m].dl = (uint16_t)bits;
            }
            n--;
        }
    }
}

static void build_tree(zso_deflater* s, dyn_desc* desc, const static_desc* st) /* Trees.cs:404-501 */
{
    code_data* tree = desc->tree;
    int elems = st->elems;
    int n, m, max_code = -1, node;
    s->heap_len = 0;
    s->heap_max = HEAP_SIZE;
    for (n = 0; n < elems; n++) {
        if (tree[n].fc != 0) {
            s->heap[++s->heap_len] = max_code = n;
            s->depth[n] = 0;
        } else {
            tree[n].dl = 0;
        }
    }
    while (s->heap_len < 2) { /* force at least two codes */
        node = s->heap[++s->heap_len] = (max_code < 2 ? ++max_code : 0);
        tree[node].fc = 1;
        s->depth[node] = 0;
        s->opt_len--;
        if (st->stree) s->static_len -= st->stree[node].dl;
    }
    desc->max_code = max_code;
    for (n = s->heap_len / 2; n >= 1; n--) pq_down_heap(s, tree, n);
    node = elems;
    do {
        n = s->heap[1];
        s->heap[1] = s->heap[s->heap_len--];
        pq_down_heap(s, tree, 1);
        m = s->heap[1];
        s->heap[--s->heap_max] = n;
        s->heap[--s->heap_max] = m;
        tree[node].fc = (uint16_t)(tree[n].fc + tree[m].fc);
        s->depth[node] = (uint8_t)((s->depth[n] >= s->depth[m] ? s->depth[n] : s->depth[m]) + 1);
        tree[n].dl = tree[m].dl = (uint16_t)node;
        s->heap[1] = node++;
        pq_down_heap(s, tree, 1);
    } while (s->heap_len >= 2);
    s->heap[--s->heap_max] = s->heap[1];
    gen_bitlen(s, desc, st);
    gen_codes(tree, max_code, s->bl_count);
}

static void scan_tree(zso_deflater* s, code_data* tree, int max_code) /* Trees.cs:290-357 */
{
    int prevlen = -1, curlen, nextlen = tree[0].dl;
    unsigned count = 0;
    int max_count = 7, min_count = 4;
    if (nextlen == 0) { max_count = 138; min_count = 3; }
    tree[max_code + 1].dl = 0xFFFF; /* guard */
    for (int n = 0; n <= max_code; n++) {
        curlen = nextlen;
        nextlen = tree[n + 1].dl;
        if (++count < (unsigned)max_count && curlen == nextlen) continue;
        else if (count < (unsigned)min_count) s->bltree[curlen].fc = (uint16_t)(s->bltree[curlen].fc + count);
        else if (curlen != 0) {
            if (curlen != prevlen) s->bltree[curlen].fc++;
            s->bltree[REP_3_6].fc++;
        } else if (count <= 10) s->bltree[REPZ_3_10].fc++;
        else s->bltree[REPZ_11_138].fc++;
        count = 0;
        prevlen = curlen;
        if (nextlen == 0) { max_count = 138; min_count = 3; }
        else if (curlen == nextlen) { max_count = 6; min_count = 3; }
        else { max_count = 7; min_count = 4; }
    }
}

static void send_tree(zso_deflater* s, const code_data* tree, int max_code) /* Trees.cs:879-952 */
{
    int prevlen = -1, curlen, nextlen = tree[0].dl;
    int count = 0, max_count = 7, min_count = 4;
    if (nextlen == 0) { max_count = 138; min_count = 3; }
    for (int n = 0; n <= max_code; n++) {
        curlen = nextlen;
        nextlen = tree[n + 1].dl;
        if (++count < max_count && curlen == nextlen) continue;
        else if (count < min_count) {
            do send_code(s, curlen, s->bltree); while (--count != 0);
        } else if (curlen != 0) {
            if (curlen != prevlen) { send_code(s, curlen, s->bltree); count--; }
            send_code(s, REP_3_6, s->bltree);
            send_bits(s, (unsigned)(count - 3), 2);
        } else if (count <= 10) {
            send_code(s, REPZ_3_10, s->bltree);
            send_bits(s, (unsigned)(count - 3), 3);
        } else {
            send_code(s, REPZ_11_138, s->bltree);
            send_bits(s, (unsigned)(count - 11), 7);
        }
        count = 0;
        prevlen = curlen;
        if (nextlen == 0) { max_count = 138; min_count = 3; }
        else if (curlen == nextlen) { max_count = 6; min_count = 3; }
        else { max_count = 7; min_count = 4; }
    }
}

static int build_bl_tree(zso_deflater* s) /* Trees.cs:361-391 */
{
    int max_blindex;
    scan_tree(s, s->ltree, s->l_desc.max_code);
    scan_tree(s, s->dtree, s->d_desc.max_code);
    build_tree(s, &s->bl_desc, &k_bl_desc);
    for (max_blindex = BL_CODES - 1; max_blindex >= 3; max_blindex--)
        if (s->bltree[k_bl_order[max_blindex]].dl != 0) break;
    s->opt_len += 3 * (max_blindex + 1) + 5 + 5 + 4;
    return max_blindex;
}

static void send_all_trees(zso_deflater* s, int lcodes, int dcodes, int blcodes) /* Trees.cs:856-870 */
{
    send_bits(s, (unsigned)(lcodes - 257), 5);
    send_bits(s, (unsigned)(dcodes - 1), 5);
    send_bits(s, (unsigned)(blcodes - 4), 4);
    for (int rank = 0; rank < blcodes; rank++) send_bits(s, s->bltree[k_bl_order[rank]].dl, 3);
    send_tree(s, s->ltree, lcodes - 1);
    send_tree(s, s->dtree, dcodes - 1);
}

static void compress_block(zso_deflater* s, const code_data* ltree, const code_data* dtree) /* Trees.cs:956-989, 707-732 */
{
    int lx = 0;
    if (s->last_lit != 0) {
        const uint8_t* pb = s->pending_buf; /* symbols live in an overlay of the pending buffer */
        do {
            int dist = (pb[s->d_buf + lx * 2] << 8) | pb[s->d_buf + lx * 2 + 1];
            int lc = pb[s->l_buf + lx];
            lx++;
            if (dist == 0) {
                send_code(s, lc, ltree);
            } else {
                int code = length_code(lc);
                send_code(s, code + LITERALS + 1, ltree);
                int extra = k_extra_lbits[code];
                if (extra != 0) send_bits(s, (unsigned)(lc - g_base_length[code]), extra);
                dist--;
                code = dist_code(dist);
                send_code(s, code, dtree);
                extra = k_extra_dbits[code];
                if (extra != 0) send_bits(s, (unsigned)(dist - g_base_dist[code]), extra);
            }
        } while (lx < s->last_lit);
    }
    send_code(s, END_BLOCK, ltree);
    s->last_eob_len = ltree[END_BLOCK].dl;
}

static void copy_block(zso_deflater* s, int buf, int len, int header) /* Deflate.cs:710-722 */
{
    bi_windup(s);
    s->last_eob_len = 8;
    if (header) {
        put_le(s, (uint16_t)len, 2);
        put_le(s, (uint16_t)~len, 2);
    }
    memcpy(s->pending_buf + s->pending, s->window + buf, (size_t)len);
    s->pending += len;
}

static void tr_stored_block(zso_deflater* s, int buf, int stored_len, int eof) /* Trees.cs:742-746 */
{
    send_bits(s, (unsigned)((0 << 1) + (eof ? 1 : 0)), 3);
    copy_block(s, buf, stored_len, 1);
}

static void tr_align(zso_deflater* s) /* Trees.cs:658-680 */
{
    send_bits(s, (1 << 1) + 0, 3);
    send_code(s, END_BLOCK, g_static_ltree);
    bi_flush(s);
    if (1 + s->last_eob_len + 10 - s->bi_valid < 9) {
        send_bits(s, (1 << 1) + 0, 3);
        send_code(s, END_BLOCK, g_static_ltree);
        bi_flush(s);
    }
    s->last_eob_len = 7;
}

static void trace_block(zso_deflater* s, int type, int eof, int stored_len)
{
    s->cnt.blocks++;
    if (!s->cb) return;
    zso_block_trace t;
    t.type = type;
    t.last = eof;
    t.n_symbols = s->last_lit;
    t.stored_len = stored_len;
    /* bits already handed out: total_out bytes left the engine, pending bytes wait, bi_valid bits in the accumulator */
    t.bit_start = ((uint64_t)s->total_out + (uint64_t)s->pending) * 8u + (uint64_t)s->bi_valid;
    for (int i = 0; i < s->last_lit; i++) {
        s->tr_dist[i] = (uint16_t)((s->pending_buf[s->d_buf + 2 * i] << 8) | s->pending_buf[s->d_buf + 2 * i + 1]);
        s->tr_lc[i] = s->pending_buf[s->l_buf + i];
    }
    s->cb(s->cb_user, &t, s->tr_dist, s->tr_lc);
}

static void tr_flush_block(zso_deflater* s, int buf, int stored_len, int eof) /* Trees.cs:568-643 */
{
    int opt_lenb, static_lenb, max_blindex = 0;
    if (s->level > 0) {
        build_tree(s, &s->l_desc, &k_l_desc);
        build_tree(s, &s->d_desc, &k_d_desc);
        max_blindex = build_bl_tree(s);
        opt_lenb = (s->opt_len + 3 + 7) >> 3;
        static_lenb = (s->static_len + 3 + 7) >> 3;
        if (static_lenb <= opt_lenb) opt_lenb = static_lenb;
    } else {
        opt_lenb = static_lenb = stored_len + 5;
    }
    if (stored_len + 4 <= opt_lenb && buf != -1) {
        trace_block(s, 0, eof, stored_len);
        tr_stored_block(s, buf, stored_len, eof);
    } else if (s->strategy == ZSO_FIXED || static_lenb == opt_lenb) {
        trace_block(s, 1, eof, stored_len);
        send_bits(s, (unsigned)((1 << 1) + (eof ? 1 : 0)), 3);
        compress_block(s, g_static_ltree, g_static_dtree);
    } else {
        trace_block(s, 2, eof, stored_len);
        send_bits(s, (unsigned)((2 << 1) + (eof ? 1 : 0)), 3);
        send_all_trees(s, s->l_desc.max_code + 1, s->d_desc.max_code + 1, max_blindex + 1);
        compress_block(s, s->ltree, s->dtree);
    }
    init_block(s);
    if (eof) bi_windup(s);
}

/* ---- LZ77 (Deflate.cs:866-1100) -------------------------------------------------------- */

static inline uint32_t load_le32(const uint8_t* p)
{
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

#if defined(__x86_64__) || defined(__i386__)
#include <nmmintrin.h>
__attribute__((target("sse4.2"))) static inline uint32_t crc32c_hw(uint32_t val) { return _mm_crc32_u32(0, val); }
#define ZSO_HAVE_HW_CRC 1
#endif

static inline uint32_t update_hash(const zso_deflater* s, uint32_t val) /* Deflate.Intrinsics.cs:295-307 */
{
    if (s->hash_variant == ZSO_HASH_CRC32C) {
#ifdef ZSO_HAVE_HW_CRC
        /* the reference itself executes this as one SSE4.2 crc32 instruction (Sse42.Crc32); same value as the table walk */
        if (g_hw_crc) return crc32c_hw(val);
#endif
        uint32_t crc = 0;
        for (int i = 0; i < 4; i++) {
            crc = g_crc32c[(crc ^ val) & 0xFF] ^ (crc >> 8);
            val >>= 8;
        }
        return crc;
    }
    return (val * 2654435761u) >> 16;
}

/* Deflate.cs:866-877.  Q1: hashes the four bytes at str+2.  Q2: no update when head already == str. */
static inline int insert_string(zso_deflater* s, int str)
{
    uint32_t h = update_hash(s, load_le32(s->window + str + (MIN_MATCH - 1))) & s->hash_mask;
    uint16_t cur = s->head[h];
    s->cnt.inserts++;
    if (cur != (uint16_t)str) {
        s->prev[str & s->w_mask] = cur;
        s->head[h] = (uint16_t)str;
    }
    return cur;
}

static void slide_hash(zso_deflater* s) /* Deflate.Intrinsics.cs:261-285 */
{
    int w = s->w_size;
    for (int i = 0; i < s->hash_size; i++) s->head[i] = (uint16_t)(s->head[i] >= w ? s->head[i] - w : 0);
    for (int i = 0; i < w; i++) s->prev[i] = (uint16_t)(s->prev[i] >= w ? s->prev[i] - w : 0);
}

static int read_buffer(zso_deflater* s, uint8_t* dst, int size) /* ZlibStream.cs:197-222 */
{
    int len = s->avail_in;
    if (len > size) len = size;
    if (len == 0) return 0;
    s->avail_in -= len;
    if (s->no_header == 0) s->adler = zso_adler32(s->adler, s->next_in + s->next_in_idx, (size_t)len);
    memcpy(dst, s->next_in + s->next_in_idx, (size_t)len);
    s->next_in_idx += len;
    s->total_in += len;
    return len;
}

static void fill_window(zso_deflater* s) /* Deflate.cs:967-1019 */
{
    do {
        int more = s->window_size - s->lookahead - s->str_start;
        if (s->str_start >= s->w_size + s->w_size - MIN_LOOKAHEAD) {
            memcpy(s->window, s->window + s->w_size, (size_t)s->w_size);
            s->match_start -= s->w_size;
            s->str_start -= s->w_size;
            s->block_start -= s->w_size;
            slide_hash(s);
            more += s->w_size;
        }
        if (s->avail_in == 0) return;
        int n = read_buffer(s, s->window + s->str_start + s->lookahead, more);
        s->lookahead += n;
        /* Q3: the reference inserts string str_start+1 here (stock zlib only primes ins_h) */
        if (s->lookahead >= MIN_MATCH) insert_string(s, s->str_start + 1);
    } while (s->lookahead < MIN_LOOKAHEAD && s->avail_in != 0);
}

static inline int compare256(const uint8_t* a, const uint8_t* b) /* Deflate.Intrinsics.cs:19-162 */
{
    /* 8 bytes per step (the reference uses 32-byte AVX2 / 16-byte SSE2 steps); same result as the byte loop */
    int len = 0;
    while (len < 256) {
        uint64_t x, y;
        memcpy(&x, a + len, 8);
        memcpy(&y, b + len, 8);
        x ^= y;
        if (x) {
#if defined(__BYTE_ORDER__) && __BYTE_ORDER__ == __ORDER_LITTLE_ENDIAN__
            return len + (__builtin_ctzll(x) >> 3);
#else
            while (a[len] == b[len]) len++;
            return len;
#endif
        }
        len += 8;
    }
    return 256;
}

static inline unsigned load_le16(const uint8_t* p) { return (unsigned)p[0] | ((unsigned)p[1] << 8); }

static int longest_match(zso_deflater* s, int cur_match) /* Deflate.cs:1022-1100 */
{
    const uint8_t* window = s->window;
    int chain_length = s->max_chain;
    const uint8_t* scan = window + s->str_start;
    int best_len = s->prev_length;
    int limit = s->str_start > (s->w_size - MIN_LOOKAHEAD) ? s->str_start - (s->w_size - MIN_LOOKAHEAD) : 0;
    int nice_match = s->nice_match;
    int match_start = s->match_start;
    s->cnt.searches++;
    if (best_len == 0) best_len = 1;
    unsigned scan_start = load_le16(scan);
    unsigned scan_end = load_le16(scan + best_len - 1);
    if (s->prev_length >= s->good_match) chain_length >>= 2;
    if (nice_match > s->lookahead) nice_match = s->lookahead;
    do {
        s->cnt.chain_steps++;
        if (cur_match >= s->str_start) break;
        const uint8_t* match = window + cur_match;
        if (load_le16(match + best_len - 1) != scan_end || load_le16(match) != scan_start) continue;
        int len = compare256(scan + 2, match + 2) + 2;
        s->cnt.compare_bytes += (uint64_t)len;
        if (len > best_len) {
            match_start = cur_match;
            best_len = len;
            if (len >= nice_match) break;
            scan_end = load_le16(scan + best_len - 1);
        }
    } while ((cur_match = s->prev[cur_match & s->w_mask]) > limit && --chain_length != 0);
    s->match_start = match_start;
    return best_len < s->lookahead ? best_len : s->lookahead;
}

static inline int tally_dist(zso_deflater* s, int dist, int len) /* Deflate.cs:910-926 */
{
    int di = s->d_buf + s->last_lit * 2;
    s->pending_buf[di] = (uint8_t)(dist >> 8);
    s->pending_buf[di + 1] = (uint8_t)dist;
    s->pending_buf[s->l_buf + s->last_lit++] = (uint8_t)len;
    s->matches++;
    dist--;
    s->ltree[length_code(len) + LITERALS + 1].fc++;
    s->dtree[dist_code(dist)].fc++;
    s->cnt.matches++;
    s->cnt.match_bytes += (uint64_t)(len + MIN_MATCH);
    return s->last_lit == s->lit_bufsize - 1;
}

static inline int tally_lit(zso_deflater* s, unsigned c) /* Deflate.cs:935-948 */
{
    int di = s->d_buf + s->last_lit * 2;
    s->pending_buf[di] = 0;
    s->pending_buf[di + 1] = 0;
    s->pending_buf[s->l_buf + s->last_lit++] = (uint8_t)c;
    s->ltree[c].fc++;
    s->cnt.literals++;
    return s->last_lit == s->lit_bufsize - 1;
}

static void flush_block_only(zso_deflater* s, int eof) /* Deflate.cs:951-956 */
{
    tr_flush_block(s, s->block_start >= 0 ? s->block_start : -1, s->str_start - s->block_start, eof);
    s->block_start = s->str_start;
    flush_pending(s);
}

static int finish_state(const zso_deflater* s, int flush)
{
    if (s->avail_out == 0) return flush == ZSO_FINISH ? BS_FINISH_STARTED : BS_NEED_MORE;
    return flush == ZSO_FINISH ? BS_FINISH_DONE : BS_BLOCK_DONE;
}

static int deflate_stored(zso_deflater* s, int flush) /* Deflate.Stored.cs:24-84 */
{
    int max_block_size = s->pending_size - 5 < s->w_size ? s->pending_size - 5 : s->w_size;
    for (;;) {
        if (s->lookahead <= 1) {
            fill_window(s);
            if (s->lookahead == 0 && flush == ZSO_NO_FLUSH) return BS_NEED_MORE;
            if (s->lookahead == 0) break;
        }
        s->str_start += s->lookahead;
        s->lookahead = 0;
        int max_start = s->block_start + max_block_size;
        if (s->str_start == 0 || s->str_start >= max_start) {
            s->lookahead = s->str_start - max_start;
            s->str_start = max_start;
            flush_block_only(s, 0);
            if (s->avail_out == 0) return BS_NEED_MORE;
        }
        if (s->str_start - s->block_start >= s->w_size - MIN_LOOKAHEAD) {
            flush_block_only(s, 0);
            if (s->avail_out == 0) return BS_NEED_MORE;
        }
    }
    flush_block_only(s, flush == ZSO_FINISH);
    return finish_state(s, flush);
}

static int deflate_fast(zso_deflater* s, int flush) /* Deflate.Fast.cs:20-128 */
{
    for (;;) {
        int bflush;
        if (s->lookahead < MIN_LOOKAHEAD) {
            fill_window(s);
            if (s->lookahead < MIN_LOOKAHEAD && flush == ZSO_NO_FLUSH) return BS_NEED_MORE;
            if (s->lookahead == 0) break;
        }
        int hash_head = 0;
        if (s->lookahead >= MIN_MATCH) hash_head = insert_string(s, s->str_start);
        if (hash_head != 0 && (s->str_start - hash_head) <= s->w_size - MIN_LOOKAHEAD) {
            if (s->strategy != ZSO_HUFFMAN_ONLY) s->match_length = longest_match(s, hash_head);
        }
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
                /* Deflate.Fast.cs:100 computes a hash and discards it (Q8): no effect */
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

static int deflate_slow(zso_deflater* s, int flush) /* Deflate.Slow.cs:18-159 */
{
    int hash_head = 0; /* Q8: declared outside the loop, not reset per iteration */
    for (;;) {
        int bflush;
        if (s->lookahead < MIN_LOOKAHEAD) {
            fill_window(s);
            if (s->lookahead < MIN_LOOKAHEAD && flush == ZSO_NO_FLUSH) return BS_NEED_MORE;
            if (s->lookahead == 0) break;
        }
        if (s->lookahead >= MIN_MATCH) hash_head = insert_string(s, s->str_start);
        s->prev_length = s->match_length;
        s->prev_match = s->match_start;
        s->match_length = MIN_MATCH - 1;
        if (hash_head != 0 && s->prev_length < s->max_lazy && s->str_start - hash_head <= s->w_size - MIN_LOOKAHEAD) {
            if (s->strategy != ZSO_HUFFMAN_ONLY) s->match_length = longest_match(s, hash_head);
            if (s->match_length <= 5 &&
                (s->strategy == ZSO_FILTERED || (s->match_length == MIN_MATCH && s->str_start - s->match_start > 4096))) {
                s->match_length = MIN_MATCH - 1;
            }
        }
        if (s->prev_length >= MIN_MATCH && s->match_length <= s->prev_length) {
            int max_insert = s->str_start + s->lookahead - MIN_MATCH;
            bflush = tally_dist(s, s->str_start - 1 - s->prev_match, s->prev_length - MIN_MATCH);
            s->lookahead -= s->prev_length - 1;
            s->prev_length -= 2;
            do {
                if (++s->str_start <= max_insert) hash_head = insert_string(s, s->str_start);
            } while (--s->prev_length != 0);
            s->match_available = 0;
            s->match_length = MIN_MATCH - 1;
            s->str_start++;
            if (bflush) {
                flush_block_only(s, 0);
                if (s->avail_out == 0) return BS_NEED_MORE;
            }
        } else if (s->match_available != 0) {
            bflush = tally_lit(s, s->window[s->str_start - 1]);
            if (bflush) flush_block_only(s, 0);
            s->str_start++;
            s->lookahead--;
            if (s->avail_out == 0) return BS_NEED_MORE;
        } else {
            s->match_available = 1;
            s->str_start++;
            s->lookahead--;
        }
    }
    if (s->match_available != 0) {
        (void)tally_lit(s, s->window[s->str_start - 1]);
        s->match_available = 0;
    }
    flush_block_only(s, flush == ZSO_FINISH);
    return finish_state(s, flush);
}

static int deflate_rle(zso_deflater* s, int flush) /* Deflate.Rle.cs:18-104 */
{
    for (;;) {
        int bflush;
        if (s->lookahead <= MAX_MATCH) {
            fill_window(s);
            if (s->lookahead <= MAX_MATCH && flush == ZSO_NO_FLUSH) return BS_NEED_MORE;
        }
        if (s->lookahead == 0) break;
        s->match_length = 0;
        if (s->lookahead >= MIN_MATCH && s->str_start > 0) { /* Q14: may reach into the previous segment */
            const uint8_t* scan = s->window + s->str_start - 1;
            unsigned prev = *scan;
            if (prev == scan[1] && prev == scan[2] && prev == scan[3]) {
                /* the reference's 8-way unrolled scan stops at the first mismatch or once scan >= strend */
                const uint8_t* strend = s->window + s->str_start + MAX_MATCH;
                scan += 3;
                for (;;) {
                    int k, stop = 0;
                    for (k = 0; k < 8; k++) {
                        if (prev != *++scan) { stop = 1; break; }
                    }
                    if (stop || !(scan < strend)) break;
                }
                s->match_length = MAX_MATCH - (int)(strend - scan);
                if (s->match_length > s->lookahead) s->match_length = s->lookahead;
            }
        }
        if (s->match_length >= MIN_MATCH) {
            bflush = tally_dist(s, 1, s->match_length - MIN_MATCH);
            s->lookahead -= s->match_length;
            s->str_start += s->match_length;
            s->match_length = 0;
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

/* ---- deflate() (Deflate.cs:436-636) ---------------------------------------------------- */

static const char* const k_errmsg[10] = {"need dictionary", "stream end", "", "file error", "stream error",
                                         "data error", "insufficient memory", "buffer error", "incompatible version", ""};
#define ERRMSG(code) k_errmsg[ZSO_NEED_DICT - (code)]

/* Test-only hook: tools/fc_model/fc_model.c includes this file and swaps in an experimental DeflateFast to check
 * that a restructured parse (the one the CUDA kernel uses) is equivalent to the literal one below.  Always NULL in
 * libzso.so. */
static int (*g_fast_hook)(zso_deflater* s, int flush);

static int deflate_call(zso_deflater* s, int flush)
{
    if (flush > ZSO_FINISH || flush < 0) return ZSO_STREAM_ERROR;
    if (s->next_out == NULL || (s->next_in == NULL && s->avail_in != 0) || (s->status == ST_FINISH && flush != ZSO_FINISH)) {
        s->msg = ERRMSG(ZSO_STREAM_ERROR);
        return ZSO_STREAM_ERROR;
    }
    if (s->avail_out == 0) {
        s->msg = ERRMSG(ZSO_BUF_ERROR);
        return ZSO_BUF_ERROR;
    }
    int old_flush = s->last_flush;
    s->last_flush = flush;
    if (s->status == ST_INIT) { /* Q11: old-zlib FLEVEL, no strategy adjustment */
        int header = (8 + ((s->w_bits - 8) << 4)) << 8;
        int level_flags = ((s->level - 1) & 0xff) >> 1;
        if (level_flags > 3) level_flags = 3;
        header |= level_flags << 6;
        if (s->str_start != 0) header |= 0x20;
        header += 31 - (header % 31);
        s->status = ST_BUSY;
        put_short_msb(s, (unsigned)header);
        if (s->str_start != 0) {
            put_short_msb(s, s->adler >> 16);
            put_short_msb(s, s->adler & 0xFFFF);
        }
        s->adler = 1;
    }
    if (s->pending != 0) {
        flush_pending(s);
        if (s->avail_out == 0) {
            s->last_flush = -1;
            return ZSO_OK;
        }
    } else if (s->avail_in == 0 && flush <= old_flush && flush != ZSO_FINISH) {
        s->msg = ERRMSG(ZSO_BUF_ERROR);
        return ZSO_BUF_ERROR;
    }
    if (s->status == ST_FINISH && s->avail_in != 0) {
        s->msg = ERRMSG(ZSO_BUF_ERROR);
        return ZSO_BUF_ERROR;
    }
    if (s->avail_in != 0 || s->lookahead != 0 || (flush != ZSO_NO_FLUSH && s->status != ST_FINISH)) {
        int bstate;
        if (s->strategy == ZSO_RLE) bstate = deflate_rle(s, flush);
        else if (k_cfg[s->level].fn == FN_STORED) bstate = deflate_stored(s, flush);
        else if (k_cfg[s->level].fn == FN_FAST) bstate = g_fast_hook ? g_fast_hook(s, flush) : deflate_fast(s, flush);
        else bstate = deflate_slow(s, flush);
        if (bstate == BS_FINISH_STARTED || bstate == BS_FINISH_DONE) s->status = ST_FINISH;
        if (bstate == BS_NEED_MORE || bstate == BS_FINISH_STARTED) {
            if (s->avail_out == 0) s->last_flush = -1;
            return ZSO_OK;
        }
        if (bstate == BS_BLOCK_DONE) {
            if (flush == ZSO_PARTIAL_FLUSH) {
                tr_align(s);
            } else { /* Q12 */
                tr_stored_block(s, 0, 0, 0);
                if (flush == ZSO_FULL_FLUSH) memset(s->head, 0, sizeof(uint16_t) * (size_t)s->hash_size);
            }
            flush_pending(s);
            if (s->avail_out == 0) {
                s->last_flush = -1;
                return ZSO_OK;
            }
        }
    }
    if (flush != ZSO_FINISH) return ZSO_OK;
    if (s->no_header != 0) return ZSO_STREAM_END;
    put_short_msb(s, s->adler >> 16);
    put_short_msb(s, s->adler & 0xFFFF);
    flush_pending(s);
    s->no_header = -1;
    return s->pending != 0 ? ZSO_OK : ZSO_STREAM_END;
}

/* ---- ZlibOutputStream (ZlibOutputStream.cs) -------------------------------------------- */

static void sink_write(zso_deflater* d, const uint8_t* p, size_t n)
{
    if (n == 0) return;
    if (d->out_len + n > d->out_cap) {
        size_t cap = d->out_cap ? d->out_cap * 2 : 1 << 16;
        while (cap < d->out_len + n) cap *= 2;
        uint8_t* q = (uint8_t*)realloc(d->out, cap);
        if (!q) { d->oom = 1; return; }
        d->out = q;
        d->out_cap = cap;
    }
    memcpy(d->out + d->out_len, p, n);
    d->out_len += n;
}

zso_deflater* zso_deflater_new(int level, int strategy, int flush_mode, int hash_variant)
{
    build_static_tables();
    if (level == -1) level = 6; /* Deflate.cs:241-244 */
    if (level < 0 || level > 9 || strategy < 0 || strategy > ZSO_FIXED) return NULL; /* Deflate.cs:273-281 */
    if (flush_mode < 0 || flush_mode > ZSO_FINISH) return NULL;
    zso_deflater* s = (zso_deflater*)calloc(1, sizeof *s);
    if (!s) return NULL;
    int mem_level = level == 0 ? 7 : 8; /* Deflate.cs:19-20,246-249 */
    s->flush_mode = flush_mode;
    s->level = level;
    s->strategy = strategy;
    s->hash_variant = hash_variant;
    s->w_bits = 15;
    s->w_size = 1 << 15;
    s->w_mask = s->w_size - 1;
    s->window_size = 2 * s->w_size;
    s->hash_bits = mem_level + 7;
    s->hash_size = 1 << s->hash_bits;
    s->hash_mask = (uint32_t)s->hash_size - 1;
    s->lit_bufsize = 1 << (mem_level + 6);
    s->d_buf = s->lit_bufsize;
    s->l_buf = 3 * s->lit_bufsize;
    s->pending_size = 4 * s->lit_bufsize;
    /* zero-filled like a first ArrayPool rent (Q7); slack so stale/out-of-range reads stay defined */
    s->window = (uint8_t*)calloc((size_t)s->window_size + 1024, 1);
    s->prev = (uint16_t*)calloc((size_t)s->w_size, sizeof(uint16_t));
    s->head = (uint16_t*)calloc((size_t)s->hash_size, sizeof(uint16_t));
    s->pending_buf = (uint8_t*)calloc((size_t)s->pending_size + 1024, 1);
    s->tr_dist = (uint16_t*)calloc((size_t)s->lit_bufsize, sizeof(uint16_t));
    s->tr_lc = (uint8_t*)calloc((size_t)s->lit_bufsize, 1);
    if (!s->window || !s->prev || !s->head || !s->pending_buf || !s->tr_dist || !s->tr_lc) {
        zso_deflater_free(s);
        return NULL;
    }
    s->l_desc.tree = s->ltree;
    s->d_desc.tree = s->dtree;
    s->bl_desc.tree = s->bltree;
    /* DeflateReset (Deflate.cs:879-900) + Tr_init + Lm_init */
    s->status = ST_INIT;
    s->adler = 1;
    s->last_flush = ZSO_NO_FLUSH;
    s->last_eob_len = 8;
    init_block(s);
    s->max_lazy = k_cfg[level].lazy;
    s->good_match = k_cfg[level].good;
    s->nice_match = k_cfg[level].nice;
    s->max_chain = k_cfg[level].chain;
    s->match_length = s->prev_length = MIN_MATCH - 1;
    return s;
}

void zso_deflater_free(zso_deflater* d)
{
    if (!d) return;
    free(d->window);
    free(d->prev);
    free(d->head);
    free(d->pending_buf);
    free(d->tr_dist);
    free(d->tr_lc);
    free(d->out);
    free(d);
}

void zso_deflater_set_trace(zso_deflater* d, zso_block_cb cb, void* user)
{
    d->cb = cb;
    d->cb_user = user;
}

int zso_deflater_write(zso_deflater* d, const uint8_t* buf, size_t n) /* ZlibOutputStream.cs:125-168 */
{
    if (n == 0) return ZSO_OK;
    if (n > 0x7FFFFFFFu) return ZSO_STREAM_ERROR;
    d->next_in = buf;
    d->next_in_idx = 0;
    d->avail_in = (int)n;
    int state;
    do {
        d->next_out = d->chunk;
        d->next_out_idx = 0;
        d->avail_out = CHUNK;
        state = deflate_call(d, d->flush_mode);
        if (state != ZSO_OK && state != ZSO_STREAM_END) return state;
        sink_write(d, d->chunk, (size_t)(CHUNK - d->avail_out));
        if (state == ZSO_STREAM_END) break;
    } while (d->avail_in > 0 || d->avail_out == 0);
    return d->oom ? ZSO_MEM_ERROR : ZSO_OK;
}

int zso_deflater_finish(zso_deflater* d) /* ZlibOutputStream.cs:213-256 */
{
    if (d->finished) return ZSO_OK;
    int state;
    do {
        d->next_out = d->chunk;
        d->next_out_idx = 0;
        d->avail_out = CHUNK;
        state = deflate_call(d, ZSO_FINISH);
        if (state != ZSO_STREAM_END && state != ZSO_OK) return state;
        sink_write(d, d->chunk, (size_t)(CHUNK - d->avail_out));
        if (state == ZSO_STREAM_END) break;
    } while (d->avail_in > 0 || d->avail_out == 0);
    d->finished = 1;
    return d->oom ? ZSO_MEM_ERROR : ZSO_OK;
}

const uint8_t* zso_deflater_output(const zso_deflater* d, size_t* len)
{
    if (len) *len = d->out_len;
    return d->out;
}
uint64_t zso_deflater_total_in(const zso_deflater* d) { return (uint64_t)d->total_in; }
uint64_t zso_deflater_total_out(const zso_deflater* d) { return (uint64_t)d->total_out; }
uint32_t zso_deflater_adler(const zso_deflater* d) { return d->adler; }
const char* zso_deflater_message(const zso_deflater* d) { return d->msg; }
void zso_deflater_counters(const zso_deflater* d, zso_counters* out) { *out = d->cnt; }

size_t zso_deflate_bound(size_t n, size_t seg_size)
{
    size_t segs = seg_size ? (n + seg_size - 1) / seg_size : 1;
    /* stored fallback: 5 B per <=16383-symbol block, plus per segment: empty blocks + markers */
    return n + (n / 16383 + 1) * 5 + segs * 32 + 64;
}

int zso_deflate_segments(const uint8_t* in, size_t n, size_t seg_size, int level, int strategy, int hash_variant,
                         uint8_t* out, size_t out_cap, size_t* out_len, uint64_t* seg_end, uint32_t* adler)
{
    if (seg_size == 0) return ZSO_STREAM_ERROR;
    zso_deflater* d = zso_deflater_new(level, strategy, ZSO_FULL_FLUSH, hash_variant);
    if (!d) return ZSO_STREAM_ERROR;
    int rc = ZSO_OK;
    size_t k = 0;
    for (size_t off = 0; off < n && rc == ZSO_OK; off += seg_size, k++) {
        size_t len = n - off < seg_size ? n - off : seg_size;
        rc = zso_deflater_write(d, in + off, len);
        if (seg_end) seg_end[k] = zso_deflater_total_out(d);
    }
    if (rc == ZSO_OK) rc = zso_deflater_finish(d);
    if (rc == ZSO_OK) {
        if (out_len) *out_len = d->out_len;
        if (adler) *adler = d->adler;
        if (out) {
            if (d->out_len > out_cap) rc = ZSO_BUF_ERROR;
            else memcpy(out, d->out, d->out_len);
        }
    }
    zso_deflater_free(d);
    return rc;
}

int zso_deflate_oneshot(const uint8_t* in, size_t n, int level, int strategy, int hash_variant, uint8_t* out,
                        size_t out_cap, size_t* out_len, uint32_t* adler)
{
    zso_deflater* d = zso_deflater_new(level, strategy, ZSO_NO_FLUSH, hash_variant);
    if (!d) return ZSO_STREAM_ERROR;
    int rc = zso_deflater_write(d, in, n);
    if (rc == ZSO_OK) rc = zso_deflater_finish(d);
    if (rc == ZSO_OK) {
        if (out_len) *out_len = d->out_len;
        if (adler) *adler = d->adler;
        if (out) {
            if (d->out_len > out_cap) rc = ZSO_BUF_ERROR;
            else memcpy(out, d->out, d->out_len);
        }
    }
    zso_deflater_free(d);
    return rc;
}
