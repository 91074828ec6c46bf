/*
 * zso_inflate.c -- CPU ORACLE (test infrastructure only, see zso.h): RFC 1950 / RFC 1951 decoder.
 *
 * The reference's own inflater (Inflate.cs, InflateBlocks.cs, InfCodes.cs, InfTree.cs) is a zlib 1.1.3
 * port whose dynamic-table construction is broken as written (InfTree.cs:312,341 copy 3 *bytes* into an
 * int[]); its contract is therefore "RFC 1951 decode + Adler-32 check".  This file restates that contract
 * and keeps the reference's error vocabulary (Inflate.cs:134-339, InflateBlocks.cs:237-569,
 * InfTree.cs:377-427, InfCodes.cs:294,349) and its acceptance rules (InfTree.cs huft_build: an
 * incomplete code is only legal when it is a single 1-bit code).
 * Pinned against system zlib in tests/test_oracle_inflate.py.
 */
#include "zso.h"

#include <string.h>

#define FAST_BITS 10
#define MAX_BITS 15

typedef struct {
    const uint8_t* in;
    size_t n, pos;
    uint64_t acc;
    int nbits;
} bitreader;

typedef struct {
    uint16_t fast[1 << FAST_BITS]; /* (sym << 4) | len, 0 = not in the fast table */
    uint16_t count[MAX_BITS + 1];
    uint16_t first[MAX_BITS + 1];  /* first canonical code of each length */
    uint16_t offset[MAX_BITS + 1]; /* index into sorted[] of the first symbol of each length */
    uint16_t sorted[288];
} hufftab;

static const uint16_t k_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t k_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t k_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t k_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t k_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

static inline void refill(bitreader* b)
{
    while (b->nbits <= 56 && b->pos < b->n) {
        b->acc |= (uint64_t)b->in[b->pos++] << b->nbits;
        b->nbits += 8;
    }
}
static inline int need(bitreader* b, int k)
{
    if (b->nbits < k) refill(b);
    return b->nbits >= k;
}
static inline unsigned peek(const bitreader* b, int k) { return (unsigned)(b->acc & ((1ull << k) - 1)); }
static inline void drop(bitreader* b, int k) { b->acc >>= k; b->nbits -= k; }

static unsigned rev_bits(unsigned v, int len)
{
    unsigned r = 0;
    for (int i = 0; i < len; i++) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}

/* Returns 0 ok, 1 all lengths zero, -1 oversubscribed, -2 incomplete (and not the single 1-bit code case). */
static int build_table(hufftab* t, const uint8_t* lens, int n)
{
    memset(t->count, 0, sizeof t->count);
    memset(t->fast, 0, sizeof t->fast);
    for (int i = 0; i < n; i++) t->count[lens[i]]++;
    if (t->count[0] == n) return 1;
    int left = 1, maxlen = 0;
    for (int len = 1; len <= MAX_BITS; len++) {
        left <<= 1;
        left -= t->count[len];
        if (left < 0) return -1;
        if (t->count[len]) maxlen = len;
    }
    unsigned code = 0;
    t->offset[1] = 0;
    t->count[0] = 0;
    for (int len = 1; len <= MAX_BITS; len++) {
        code = (code + t->count[len - 1]) << 1;
        t->first[len] = (uint16_t)code;
        if (len < MAX_BITS) t->offset[len + 1] = (uint16_t)(t->offset[len] + t->count[len]);
    }
    uint16_t next[MAX_BITS + 1];
    memcpy(next, t->offset, sizeof next);
    for (int i = 0; i < n; i++)
        if (lens[i]) t->sorted[next[lens[i]]++] = (uint16_t)i;
    /* fast table */
    uint16_t nc[MAX_BITS + 1];
    memcpy(nc, t->first, sizeof nc);
    for (int i = 0; i < n; i++) {
        int len = lens[i];
        if (!len) continue;
        unsigned c = nc[len]++;
        if (len <= FAST_BITS) {
            unsigned r = rev_bits(c, len);
            for (unsigned k = r; k < (1u << FAST_BITS); k += 1u << len) t->fast[k] = (uint16_t)((i << 4) | len);
        }
    }
    if (left > 0 && maxlen != 1) return -2; /* huft_build: `y != 0 && g != 1 ? Z_BUF_ERROR : Z_OK` */
    return 0;
}

/* Decodes one symbol; -1 = invalid code, -2 = out of input. */
static inline int decode_sym(bitreader* b, const hufftab* t)
{
    refill(b);
    unsigned e = t->fast[b->acc & ((1u << FAST_BITS) - 1)];
    if (e) {
        int len = (int)(e & 15);
        if (len > b->nbits) return -2;
        drop(b, len);
        return (int)(e >> 4);
    }
    unsigned code = 0;
    uint64_t acc = b->acc;
    for (int len = 1; len <= MAX_BITS; len++) {
        if (len > b->nbits) return -2;
        code = (code << 1) | (unsigned)(acc & 1);
        acc >>= 1;
        unsigned cnt = t->count[len];
        if (cnt && code >= t->first[len] && code - t->first[len] < cnt) {
            drop(b, len);
            return t->sorted[t->offset[len] + (code - t->first[len])];
        }
    }
    return -1;
}

static hufftab g_fixed_l, g_fixed_d;
static int g_fixed_ready = 0;

static void build_fixed(void)
{
    if (g_fixed_ready) return;
    uint8_t lens[288];
    int i = 0;
    for (; i < 144; i++) lens[i] = 8;
    for (; i < 256; i++) lens[i] = 9;
    for (; i < 280; i++) lens[i] = 7;
    for (; i < 288; i++) lens[i] = 8;
    build_table(&g_fixed_l, lens, 288);
    for (i = 0; i < 30; i++) lens[i] = 5;
    build_table(&g_fixed_d, lens, 30);
    g_fixed_ready = 1;
}

#define FAIL(m) do { if (msg) *msg = (m); return ZSO_DATA_ERROR; } while (0)
#define TRUNC() do { if (msg) *msg = "buffer error"; return ZSO_BUF_ERROR; } while (0)

static int inflate_codes(bitreader* b, const hufftab* tl, const hufftab* td, uint8_t* out, size_t cap, size_t* opos,
                         const char** msg)
{
    size_t o = *opos;
    for (;;) {
        int sym = decode_sym(b, tl);
        if (sym == -2) { *opos = o; TRUNC(); }
        if (sym < 0) { *opos = o; FAIL("invalid literal/length code"); }
        if (sym < 256) {
            if (o >= cap) { *opos = o; TRUNC(); }
            out[o++] = (uint8_t)sym;
            continue;
        }
        if (sym == 256) break;
        sym -= 257;
        if (sym >= 29) { *opos = o; FAIL("invalid literal/length code"); }
        int eb = k_len_extra[sym];
        if (!need(b, eb)) { *opos = o; TRUNC(); }
        unsigned len = k_len_base[sym] + peek(b, eb);
        drop(b, eb);
        int ds = decode_sym(b, td);
        if (ds == -2) { *opos = o; TRUNC(); }
        if (ds < 0 || ds >= 30) { *opos = o; FAIL("invalid distance code"); }
        eb = k_dist_extra[ds];
        if (!need(b, eb)) { *opos = o; TRUNC(); }
        size_t dist = k_dist_base[ds] + peek(b, eb);
        drop(b, eb);
        if (dist > o) { *opos = o; FAIL("invalid distance too far back"); } /* not checked by the reference (InfCodes.cs:236-241) */
        if (o + len > cap) { *opos = o; TRUNC(); }
        for (unsigned i = 0; i < len; i++, o++) out[o] = out[o - dist];
    }
    *opos = o;
    return ZSO_OK;
}

/* Decodes blocks until BFINAL.  On success the reader is positioned right after the last block (not byte aligned). */
static int inflate_blocks(bitreader* b, uint8_t* out, size_t cap, size_t* opos, const char** msg)
{
    build_fixed();
    hufftab tl, td, tb;
    int last;
    do {
        if (!need(b, 3)) TRUNC();
        last = (int)peek(b, 1);
        int type = (int)(peek(b, 3) >> 1);
        drop(b, 3);
        if (type == 0) {
            drop(b, b->nbits & 7);
            if (!need(b, 32)) TRUNC();
            unsigned len = peek(b, 16);
            drop(b, 16);
            unsigned nlen = peek(b, 16);
            drop(b, 16);
            if (((~nlen) & 0xFFFF) != len) FAIL("invalid stored block lengths");
            if (*opos + len > cap) TRUNC();
            for (unsigned i = 0; i < len; i++) {
                if (!need(b, 8)) TRUNC();
                out[(*opos)++] = (uint8_t)peek(b, 8);
                drop(b, 8);
            }
        } else if (type == 1) {
            int rc = inflate_codes(b, &g_fixed_l, &g_fixed_d, out, cap, opos, msg);
            if (rc != ZSO_OK) return rc;
        } else if (type == 2) {
            if (!need(b, 14)) TRUNC();
            unsigned t = peek(b, 14);
            drop(b, 14);
            if ((t & 0x1f) > 29 || ((t >> 5) & 0x1f) > 29) FAIL("too many length or distance symbols");
            int nl = 257 + (int)(t & 0x1f), nd = 1 + (int)((t >> 5) & 0x1f), nb = 4 + (int)(t >> 10);
            uint8_t lens[320];
            memset(lens, 0, sizeof lens);
            for (int i = 0; i < nb; i++) {
                if (!need(b, 3)) TRUNC();
                lens[k_order[i]] = (uint8_t)peek(b, 3);
                drop(b, 3);
            }
            int r = build_table(&tb, lens, 19);
            if (r == -1) FAIL("oversubscribed dynamic bit lengths tree");
            if (r != 0) FAIL("incomplete dynamic bit lengths tree");
            int i = 0;
            memset(lens, 0, sizeof lens);
            while (i < nl + nd) {
                int sym = decode_sym(b, &tb);
                if (sym == -2) TRUNC();
                if (sym < 0) FAIL("invalid bit length repeat");
                if (sym < 16) {
                    lens[i++] = (uint8_t)sym;
                } else {
                    int eb = sym == 18 ? 7 : sym - 14, base = sym == 18 ? 11 : 3;
                    if (!need(b, eb)) TRUNC();
                    int rep = base + (int)peek(b, eb);
                    drop(b, eb);
                    if (i + rep > nl + nd || (sym == 16 && i < 1)) FAIL("invalid bit length repeat");
                    uint8_t v = sym == 16 ? lens[i - 1] : 0;
                    while (rep--) lens[i++] = v;
                }
            }
            r = build_table(&tl, lens, nl);
            if (r == -1) FAIL("oversubscribed literal/length tree");
            if (r != 0) FAIL("incomplete literal/length tree");
            r = build_table(&td, lens + nl, nd);
            if (r == -1) FAIL("oversubscribed distance tree");
            if (r == -2) FAIL("incomplete distance tree");
            if (r == 1 && nl > 257) FAIL("empty distance tree with lengths");
            int rc = inflate_codes(b, &tl, &td, out, cap, opos, msg);
            if (rc != ZSO_OK) return rc;
        } else {
            FAIL("invalid block type");
        }
    } while (!last);
    return ZSO_OK;
}

int zso_inflate_raw(const uint8_t* in, size_t n, uint8_t* out, size_t out_cap, size_t* out_len, size_t* in_used,
                    const char** msg)
{
    bitreader b = {in, n, 0, 0, 0};
    size_t o = 0;
    if (msg) *msg = NULL;
    int rc = inflate_blocks(&b, out, out_cap, &o, msg);
    if (out_len) *out_len = o;
    if (in_used) *in_used = b.pos - (size_t)(b.nbits >> 3);
    return rc == ZSO_OK ? ZSO_STREAM_END : rc;
}

int zso_inflate(const uint8_t* in, size_t n, uint8_t* out, size_t out_cap, size_t* out_len, size_t* in_used,
                uint32_t* adler, const char** msg)
{
    if (msg) *msg = NULL;
    if (out_len) *out_len = 0;
    if (n < 2) TRUNC();
    unsigned cmf = in[0], flg = in[1];
    if ((cmf & 0xf) != 8) FAIL("unknown compression method");   /* Inflate.cs:130-137 */
    if ((cmf >> 4) + 8 > 15) FAIL("invalid window size");         /* Inflate.cs:139-145 */
    if (((cmf << 8) + flg) % 31 != 0) FAIL("incorrect header check"); /* Inflate.cs:162-169 */
    if (flg & 0x20) {                                                /* Inflate.cs:171-243 */
        if (msg) *msg = "need dictionary";
        return ZSO_NEED_DICT;
    }
    bitreader b = {in, n, 2, 0, 0};
    size_t o = 0;
    int rc = inflate_blocks(&b, out, out_cap, &o, msg);
    if (out_len) *out_len = o;
    if (rc != ZSO_OK) return rc;
    drop(&b, b.nbits & 7);
    if (!need(&b, 32)) TRUNC();
    uint32_t want = 0;
    for (int i = 0; i < 4; i++) {
        want = (want << 8) | peek(&b, 8);
        drop(&b, 8);
    }
    uint32_t got = zso_adler32(1, out, o);
    if (adler) *adler = got;
    if (in_used) *in_used = b.pos - (size_t)(b.nbits >> 3);
    if (got != want) FAIL("incorrect data check");                 /* Inflate.cs:335-342 */
    return ZSO_STREAM_END;
}
