// zsb_huffman.cu -- Huffman stage of the segmented deflate: per-block trees, output layout, bit packing.
//
// Replaces Trees.Tr_flush_block / Build_tree / Pqdownheap / Gen_bitlen / Gen_codes / Scan_tree /
// Build_bl_tree / Send_all_trees / Send_tree / Compress_block (Trees.cs:290-1151), Send_bits / Bi_flush /
// Bi_windup / Copy_block (Deflate.cs:640-821), the full-flush marker and head reset of Deflate.Compress
// (Deflate.cs:583-614) and -- because it decides the bit stream -- the 512-byte drain loop of
// ZlibOutputStream.WriteCore (ZlibOutputStream.cs:140-165).
//
// Three kernels:
//   block_trees_kernel  one CTA per deflate block: shared-memory symbol histogram, the reference's heap
//                       construction run literally (its tie-breaks decide the code lengths), block type.
//   layout_kernel       per segment: bit offset of every block, the extra empty static blocks / repeated
//                       markers the 512-byte output chunk causes, compressed size; exclusive scan of sizes.
//   pack_kernel         one CTA per segment: per-symbol code+extra bits, CTA prefix sum of bit lengths,
//                       shared-memory staging, word stores straight into the final stream position.
#include "zsb_kernels.cuh"

namespace zsb {

namespace {

struct CodeData { uint16_t fc; uint16_t dl; };  // Trees.Static.cs:98-124: Freq|Code, Dad|Len

__constant__ uint8_t c_extra_lbits[LENGTH_CODES] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint8_t c_extra_dbits[D_CODES] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_extra_blbits[BL_CODES] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 3, 7};
__constant__ uint8_t c_bl_order[BL_CODES] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct TreeWork {            // scratch of one Build_tree run (Deflate.Buffers.cs:17-85)
    uint16_t heap[HEAP_SIZE + 3];   // the nodes in the order they leave the heap (Gen_bitlen walks it), filled from the top
    // The live heap, one word per entry: frequency << 18 | depth << 10 | node.  Smaller (Trees.cs:556: lower frequency, or the
    // same frequency and depth not greater) is then `(a >> 10) <= (b >> 10)` on words that are already loaded -- one shared-
    // memory round trip per heap level instead of three dependent ones (index -> frequency -> depth).  Frequencies that are
    // compared stay below 2^14 (a block holds 16383 symbols + END_BLOCK; only the root reaches 16384 and it is never compared).
    uint32_t hk[HEAP_SIZE + 3];
    uint16_t bl_count[MAX_BITS + 1];
    int heap_len, heap_max;
    int opt_len, static_len;
};

enum { TK_L = 0, TK_D = 1, TK_BL = 2 };

__device__ __forceinline__ int tree_extra(int kind, int n)
{
    if (kind == TK_L) return n >= LITERALS + 1 ? c_extra_lbits[n - (LITERALS + 1)] : 0;
    if (kind == TK_D) return c_extra_dbits[n];
    return c_extra_blbits[n];
}
__device__ __forceinline__ int tree_static_len(int kind, int n) { return kind == TK_L ? static_llen(n) : 5; }

__device__ __forceinline__ uint32_t heap_entry(uint32_t freq, uint32_t depth, int node) { return (freq << 18) | (depth << 10) | (uint32_t)node; }

__device__ void pq_down_heap(TreeWork& w, int k)  // Trees.cs:513-544
{
    const uint32_t v = w.hk[k];
    int j = k << 1;
    while (j <= w.heap_len) {
        uint32_t e = w.hk[j];
        if (j < w.heap_len) {
            const uint32_t e2 = w.hk[j + 1];
            if ((e2 >> 10) <= (e >> 10)) { j++; e = e2; }
        }
        if ((v >> 10) <= (e >> 10)) break;
        w.hk[k] = e;
        k = j;
        j <<= 1;
    }
    w.hk[k] = v;
}

// Trees.cs:404-501 (Build_tree) + 999-1109 (Gen_bitlen) + 1123-1151 (Gen_codes), run by one thread.
// On return tree[n].dl = code length, tree[n].fc = bit-reversed code for n <= max_code.
__device__ int build_tree(TreeWork& w, CodeData* tree, int kind)
{
    const int elems = kind == TK_L ? L_CODES : kind == TK_D ? D_CODES : BL_CODES;
    const int max_length = kind == TK_BL ? MAX_BL_BITS : MAX_BITS;
    const bool has_stree = kind != TK_BL;
    int n, m, max_code = -1, node;
    w.heap_len = 0;
    w.heap_max = HEAP_SIZE;
    for (n = 0; n < elems; n++) {
        if (tree[n].fc != 0) {
            w.hk[++w.heap_len] = heap_entry(tree[n].fc, 0, max_code = n);
        } else {
            tree[n].dl = 0;
        }
    }
    while (w.heap_len < 2) {  // Trees.cs:439-455: force two codes
        node = max_code < 2 ? ++max_code : 0;
        w.hk[++w.heap_len] = heap_entry(1, 0, node);
        tree[node].fc = 1;
        w.opt_len--;
        if (has_stree) w.static_len -= tree_static_len(kind, node);
    }
    for (n = w.heap_len / 2; n >= 1; n--) pq_down_heap(w, n);
    node = elems;
    do {
        const uint32_t en = w.hk[1];
        w.hk[1] = w.hk[w.heap_len--];
        pq_down_heap(w, 1);
        const uint32_t em = w.hk[1];
        n = (int)(en & 1023u);
        m = (int)(em & 1023u);
        w.heap[--w.heap_max] = (uint16_t)n;
        w.heap[--w.heap_max] = (uint16_t)m;
        const uint32_t f = (en >> 18) + (em >> 18);
        const uint32_t dn = (en >> 10) & 255u, dm = (em >> 10) & 255u;
        tree[node].fc = (uint16_t)f;
        tree[n].dl = tree[m].dl = (uint16_t)node;
        w.hk[1] = heap_entry(f, (dn >= dm ? dn : dm) + 1, node);
        node++;
        pq_down_heap(w, 1);
    } while (w.heap_len >= 2);
    w.heap[--w.heap_max] = (uint16_t)(w.hk[1] & 1023u);

    // Gen_bitlen
    int h, bits, overflow = 0;
    for (bits = 0; bits <= MAX_BITS; bits++) w.bl_count[bits] = 0;
    tree[w.heap[w.heap_max]].dl = 0;
    for (h = w.heap_max + 1; h < HEAP_SIZE; h++) {
        n = w.heap[h];
        bits = tree[tree[n].dl].dl + 1;
        if (bits > max_length) { bits = max_length; overflow++; }
        tree[n].dl = (uint16_t)bits;
        if (n > max_code) continue;
        w.bl_count[bits]++;
        int xbits = tree_extra(kind, n);
        int f = tree[n].fc;
        w.opt_len += f * (bits + xbits);
        if (has_stree) w.static_len += f * (tree_static_len(kind, n) + xbits);
    }
    if (overflow != 0) {
        do {
            bits = max_length - 1;
            while (w.bl_count[bits] == 0) bits--;
            w.bl_count[bits]--;
            w.bl_count[bits + 1] += 2;
            w.bl_count[max_length]--;
            overflow -= 2;
        } while (overflow > 0);
        for (bits = max_length; bits != 0; bits--) {
            n = w.bl_count[bits];
            while (n != 0) {
                m = w.heap[--h];
                if (m > max_code) continue;
                if (tree[m].dl != (unsigned)bits) {
                    w.opt_len += bits * (int)tree[m].fc;
                    w.opt_len -= (int)tree[m].dl * (int)tree[m].fc;
                    tree[m].dl = (uint16_t)bits;
                }
                n--;
            }
        }
    }
    // Gen_codes
    uint16_t next_code[MAX_BITS + 1];
    uint16_t code = 0;
    next_code[0] = 0;
    for (bits = 1; bits <= MAX_BITS; bits++) {
        code = (uint16_t)((code + w.bl_count[bits - 1]) << 1);
        next_code[bits] = code;
    }
    for (n = 0; n <= max_code; n++) {
        int len = tree[n].dl;
        if (len == 0) continue;
        tree[n].fc = (uint16_t)bit_reverse(next_code[len]++, len);
    }
    return max_code;
}

// Trees.cs:290-357: adds the run-length symbols of one tree to the bit-length tree's frequencies.
__device__ void scan_tree(CodeData* tree, int max_code, CodeData* bltree)
{
    int prevlen = -1, curlen, nextlen = tree[0].dl;
    int count = 0, max_count = 7, min_count = 4;
    if (nextlen == 0) { max_count = 138; min_count = 3; }
    tree[max_code + 1].dl = 0xFFFF;  // guard
    for (int n = 0; n <= max_code; n++) {
        curlen = nextlen;
        nextlen = tree[n + 1].dl;
        if (++count < max_count && curlen == nextlen) continue;
        else if (count < min_count) bltree[curlen].fc = (uint16_t)(bltree[curlen].fc + count);
        else if (curlen != 0) {
            if (curlen != prevlen) bltree[curlen].fc++;
            bltree[16].fc++;
        } else if (count <= 10) bltree[17].fc++;
        else bltree[18].fc++;
        count = 0;
        prevlen = curlen;
        if (nextlen == 0) { max_count = 138; min_count = 3; }
        else if (curlen == nextlen) { max_count = 6; min_count = 3; }
        else { max_count = 7; min_count = 4; }
    }
}

struct BitSink {   // LSB-first bit string in shared memory (one writer thread)
    uint32_t* w;
    uint32_t nbits;
    __device__ void put(uint32_t value, int len)
    {
        uint32_t pos = nbits & 31, idx = nbits >> 5;
        w[idx] |= value << pos;
        if (pos + len > 32) w[idx + 1] |= value >> (32 - pos);
        nbits += len;
    }
};

// Trees.cs:879-952
__device__ void send_tree(BitSink& o, const CodeData* tree, int max_code, const CodeData* bltree)
{
    int prevlen = -1, curlen, nextlen = tree[0].dl;
    int count = 0, max_count = 7, min_count = 4;
    if (nextlen == 0) { max_count = 138; min_count = 3; }
    for (int n = 0; n <= max_code; n++) {
        curlen = nextlen;
        nextlen = tree[n + 1].dl;
        if (++count < max_count && curlen == nextlen) continue;
        else if (count < min_count) {
            do o.put(bltree[curlen].fc, bltree[curlen].dl); while (--count != 0);
        } else if (curlen != 0) {
            if (curlen != prevlen) { o.put(bltree[curlen].fc, bltree[curlen].dl); count--; }
            o.put(bltree[16].fc, bltree[16].dl);
            o.put((uint32_t)(count - 3), 2);
        } else if (count <= 10) {
            o.put(bltree[17].fc, bltree[17].dl);
            o.put((uint32_t)(count - 3), 3);
        } else {
            o.put(bltree[18].fc, bltree[18].dl);
            o.put((uint32_t)(count - 11), 7);
        }
        count = 0;
        prevlen = curlen;
        if (nextlen == 0) { max_count = 138; min_count = 3; }
        else if (curlen == nextlen) { max_count = 6; min_count = 3; }
        else { max_count = 7; min_count = 4; }
    }
}

#ifndef ZSB_BT_THREADS
#define ZSB_BT_THREADS 128
#endif
constexpr int BT_THREADS = ZSB_BT_THREADS;

struct TreesSmem {
    uint32_t lfreq[L_CODES + 2];
    uint32_t dfreq[D_CODES + 2];
    CodeData ltree[HEAP_SIZE + 3];
    CodeData dtree[2 * D_CODES + 3];
    CodeData bltree[2 * BL_CODES + 3];
    TreeWork wl, wd;
    uint32_t hdr[kHdrWords];
    int l_max, d_max;
    uint32_t type, hdr_bits, body_bits;
};

}  // namespace

// One CTA per (segment, block).  Tr_flush_block (Trees.cs:568-643) up to, but not including, the emission.
__global__ void __launch_bounds__(BT_THREADS) block_trees_kernel(HuffParams P)
{
    __shared__ TreesSmem sm;
    const int tid = threadIdx.x;
    const uint32_t seg = blockIdx.x / P.max_blocks, blk = blockIdx.x % P.max_blocks;
    if (seg >= P.nseg || blk >= P.nblocks[seg]) return;
    BlockRec& rec = P.blocks[(size_t)seg * P.max_blocks + blk];
    if (P.level0) {   // Tr_flush_block with level == 0 (Trees.cs:586-594): opt_lenb = static_lenb = stored_len + 5, so stored wins
        if (tid == 0) {
            rec.type = BT_STORED;
            rec.hdr_bits = 0;
            rec.body_bits = 0;
        }
        return;
    }
    const uint32_t* sym = P.sym + (size_t)seg * P.sym_stride + rec.sym_begin;
    const int nsym = (int)rec.sym_count;

    for (int i = tid; i < L_CODES + 2; i += BT_THREADS) sm.lfreq[i] = 0;
    for (int i = tid; i < D_CODES + 2; i += BT_THREADS) sm.dfreq[i] = 0;
    for (int i = tid; i < kHdrWords; i += BT_THREADS) sm.hdr[i] = 0;
    __syncthreads();
    // Tr_tally_* frequency counts (Deflate.cs:922-923,945) as a shared-memory histogram
    // eight loads in flight per thread: one load per trip left every warp waiting a full HBM round trip 128 times
    constexpr int HU = 8;
    for (int base = 0; base < nsym; base += BT_THREADS * HU) {
        uint32_t v[HU];
#pragma unroll
        for (int k = 0; k < HU; k++) {
            const int i = base + k * BT_THREADS + tid;
            v[k] = i < nsym ? __ldg(sym + i) : 0xFFFFFFFFu;   // no symbol has lc > 255
        }
#pragma unroll
        for (int k = 0; k < HU; k++) {
            const uint32_t s = v[k];
            if (s == 0xFFFFFFFFu) continue;
            const uint32_t dist = s & 0xFFFF, lc = s >> 16;
            if (dist == 0) {
                atomicAdd(&sm.lfreq[lc], 1u);
            } else {
                atomicAdd(&sm.lfreq[length_code_of((int)lc) + LITERALS + 1], 1u);
                atomicAdd(&sm.dfreq[dist_code_of((int)dist - 1)], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < L_CODES; i += BT_THREADS) { sm.ltree[i].fc = (uint16_t)sm.lfreq[i]; sm.ltree[i].dl = 0; }
    for (int i = tid; i < D_CODES; i += BT_THREADS) { sm.dtree[i].fc = (uint16_t)sm.dfreq[i]; sm.dtree[i].dl = 0; }
    for (int i = tid; i < BL_CODES; i += BT_THREADS) { sm.bltree[i].fc = 0; sm.bltree[i].dl = 0; }
    __syncthreads();
    if (tid == 0) sm.ltree[END_BLOCK].fc = 1;  // Trees.cs:799
    __syncthreads();

    // The literal and distance trees are independent: one thread of two different warps each.
    if (tid == 0) {
        sm.wl.opt_len = 0;
        sm.wl.static_len = 0;
        sm.l_max = build_tree(sm.wl, sm.ltree, TK_L);
    } else if (tid == 32) {
        sm.wd.opt_len = 0;
        sm.wd.static_len = 0;
        sm.d_max = build_tree(sm.wd, sm.dtree, TK_D);
    }
    __syncthreads();
    if (tid == 0) {
        int opt_len = sm.wl.opt_len + sm.wd.opt_len;
        int static_len = sm.wl.static_len + sm.wd.static_len;
        const int body_dyn = opt_len;
        // Build_bl_tree (Trees.cs:361-391)
        scan_tree(sm.ltree, sm.l_max, sm.bltree);
        scan_tree(sm.dtree, sm.d_max, sm.bltree);
        sm.wl.opt_len = 0;
        sm.wl.static_len = 0;
        build_tree(sm.wl, sm.bltree, TK_BL);
        opt_len += sm.wl.opt_len;
        int max_blindex;
        for (max_blindex = BL_CODES - 1; max_blindex >= 3; max_blindex--)
            if (sm.bltree[c_bl_order[max_blindex]].dl != 0) break;
        opt_len += 3 * (max_blindex + 1) + 5 + 5 + 4;
        int opt_lenb = (opt_len + 3 + 7) >> 3;
        const int static_lenb = (static_len + 3 + 7) >> 3;
        if (static_lenb <= opt_lenb) opt_lenb = static_lenb;
        const int stored_len = (int)rec.in_len;
        uint32_t type;
        if (stored_len + 4 <= opt_lenb && rec.buf_valid) type = BT_STORED;
        else if (P.strategy == 4 /* Fixed */ || static_lenb == opt_lenb) type = BT_STATIC;
        else type = BT_DYNAMIC;
        sm.type = type;
        sm.hdr_bits = 0;
        if (type == BT_DYNAMIC) {  // Send_all_trees (Trees.cs:856-870)
            BitSink o{sm.hdr, 0};
            const int lcodes = sm.l_max + 1, dcodes = sm.d_max + 1, blcodes = max_blindex + 1;
            o.put((uint32_t)(lcodes - 257), 5);
            o.put((uint32_t)(dcodes - 1), 5);
            o.put((uint32_t)(blcodes - 4), 4);
            for (int rank = 0; rank < blcodes; rank++) o.put(sm.bltree[c_bl_order[rank]].dl, 3);
            send_tree(o, sm.ltree, lcodes - 1, sm.bltree);
            send_tree(o, sm.dtree, dcodes - 1, sm.bltree);
            sm.hdr_bits = o.nbits;
            sm.body_bits = (uint32_t)body_dyn;
        } else if (type == BT_STATIC) {
            sm.body_bits = (uint32_t)static_len;
        } else {
            sm.body_bits = 0;
        }
        rec.type = type;
        rec.hdr_bits = sm.hdr_bits;
        rec.body_bits = sm.body_bits;
    }
    __syncthreads();
    BlockCodesDev& out = P.codes[(size_t)seg * P.max_blocks + blk];
    if (sm.type == BT_DYNAMIC) {
        for (int i = tid; i < L_CODES; i += BT_THREADS)
            out.lcode[i] = (i <= sm.l_max && sm.ltree[i].dl != 0 && sm.ltree[i].dl != 0xFFFF)
                               ? (uint32_t)sm.ltree[i].fc | ((uint32_t)sm.ltree[i].dl << 16) : 0u;
        for (int i = tid; i < D_CODES; i += BT_THREADS)
            out.dcode[i] = (i <= sm.d_max && sm.dtree[i].dl != 0 && sm.dtree[i].dl != 0xFFFF)
                               ? (uint32_t)sm.dtree[i].fc | ((uint32_t)sm.dtree[i].dl << 16) : 0u;
        for (int i = tid; i < kHdrWords; i += BT_THREADS) out.hdr[i] = sm.hdr[i];
    } else if (sm.type == BT_STATIC) {
        for (int i = tid; i < L_CODES; i += BT_THREADS) out.lcode[i] = static_lcode(i) | ((uint32_t)static_llen(i) << 16);
        for (int i = tid; i < D_CODES; i += BT_THREADS) out.dcode[i] = bit_reverse((uint32_t)i, 5) | (5u << 16);
    }
}

// ---- layout -----------------------------------------------------------------------------------------

namespace {

constexpr int LAYOUT_THREADS = 1024;

struct TailSink {
    uint32_t w[kTailWords];
    uint32_t nbits;
    __device__ void put(uint32_t value, int len)
    {
        uint32_t pos = nbits & 31, idx = nbits >> 5;
        if (idx < kTailWords) w[idx] |= value << pos;
        if (pos + len > 32 && idx + 1 < kTailWords) w[idx + 1] |= value >> (32 - pos);
        nbits += len;
    }
};

// The 512-byte drain arithmetic of ZlibOutputStream.WriteCore + Deflate.Flush_pending.
struct Drain {
    uint32_t a;           // free bytes in the current 512-byte output chunk
    uint64_t accounted;   // whole bytes already handed to Flush_pending
    __device__ bool step(uint64_t bitpos)
    {
        uint64_t whole = bitpos >> 3;
        uint64_t p = whole - accounted;
        accounted = whole;
        if (p < a) { a -= (uint32_t)p; return false; }
        uint32_t r = (uint32_t)((p - a) % CHUNK_OUT);
        a = r ? CHUNK_OUT - r : CHUNK_OUT;
        return true;
    }
};

}  // namespace

__global__ void __launch_bounds__(LAYOUT_THREADS) layout_kernel(HuffParams P)
{
    __shared__ uint64_t warp_sums[LAYOUT_THREADS / 32];
    __shared__ uint64_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = *P.total_out;  // bytes already in `out`: header, earlier batches
    __syncthreads();
    if (P.partial_flush) {
        // FlushMode.PartialFlush (Deflate.cs:566-581, Trees.cs:658-680): every Write ends with Tr_align -- an empty static
        // block, and a second one if the inflater could not yet see the end of the last real block -- instead of the stored
        // marker.  Nothing aligns the stream to a byte, so a segment starts where the previous one stopped, bit for bit: one
        // thread walks the segments in order.  TotalOut after a Write counts the whole bytes that have left (Bi_flush).
        if (tid == 0) {
            uint64_t stream_bits = carry_s * 8;   // header bytes of this call (2) are already in `out`
            for (uint32_t seg = 0; seg < P.nseg; seg++) {
                BlockRec* blocks = P.blocks + (size_t)seg * P.max_blocks;
                const uint32_t nb = P.nblocks[seg];
                SegLayout lay;
                lay.out_offset = stream_bits >> 3;
                lay.bit_shift = (uint32_t)(stream_bits & 7);
                lay.pad_ = 0;
                Drain dr;
                dr.a = (P.first_seg_index + seg == 0) ? CHUNK_OUT - 2 : CHUNK_OUT;
                dr.accounted = 0;
                uint64_t bitpos = lay.bit_shift;   // relative to the first bit of byte out_offset
                bool drained = false;
                int last_eob_len = 8;
                for (uint32_t b = 0; b < nb; b++) {
                    BlockRec& r = blocks[b];
                    r.bit_offset = bitpos - lay.bit_shift;
                    if (r.type == BT_STORED) {
                        bitpos = ((bitpos + 3 + 7) & ~7ull) + 32 + 8ull * r.in_len;
                        last_eob_len = 8;
                    } else {
                        bitpos += 3 + r.hdr_bits + r.body_bits;
                        last_eob_len = r.type == BT_STATIC ? 7 : (int)(P.codes[(size_t)seg * P.max_blocks + b].lcode[END_BLOCK] >> 16);
                    }
                    drained = dr.step(bitpos);
                }
                TailSink t;
                for (int i = 0; i < kTailWords; i++) t.w[i] = 0;
                t.nbits = 0;
                const uint64_t tail_start = bitpos;
                auto empty_static = [&]() {
                    t.put(0x2u, 3);
                    t.put(0u, 7);
                    bitpos += 10;
                };
                while (drained) {      // the flush re-run with nothing to compress: an empty static block
                    empty_static();
                    last_eob_len = 7;
                    drained = dr.step(bitpos);
                }
                for (;;) {
                    empty_static();    // Tr_align; bi_valid after Bi_flush = the bits of the unfinished byte
                    if (1 + last_eob_len + 10 - (int)(bitpos & 7) < 9) empty_static();
                    last_eob_len = 7;
                    drained = dr.step(bitpos);
                    if (!drained) break;
                    do {
                        empty_static();
                        drained = dr.step(bitpos);
                    } while (drained);
                }
                lay.out_bytes = (uint32_t)(((bitpos + 7) >> 3));   // bytes this segment touches, the shared last one included
                lay.tail_nbits = t.nbits;
                lay.tail_bit_offset = tail_start - lay.bit_shift;
                for (int i = 0; i < kTailWords; i++) lay.tail[i] = t.w[i];
                P.layout[seg] = lay;
                stream_bits = lay.out_offset * 8 + bitpos;
                if (P.seg_end) P.seg_end[seg] = stream_bits >> 3;
            }
            *P.total_out = (stream_bits + 7) >> 3;
            if (P.total_bits) *P.total_bits = stream_bits;
            if (((stream_bits + 7) >> 3) > P.out_cap) *P.status = -5;  // ZSB_BUF_ERROR
        }
        return;
    }
    for (uint32_t base = 0; base < P.nseg; base += LAYOUT_THREADS) {
        const uint32_t seg = base + tid;
        uint64_t bytes = 0;
        SegLayout lay;
        lay.bit_shift = 0;
        lay.pad_ = 0;
        if (seg < P.nseg) {
            BlockRec* blocks = P.blocks + (size_t)seg * P.max_blocks;
            const uint32_t nb = P.nblocks[seg];
            Drain dr;
            dr.a = (P.first_seg_index + seg == 0) ? CHUNK_OUT - 2 : CHUNK_OUT;  // 2-byte header drained first
            dr.accounted = 0;
            uint64_t bitpos = 0;
            bool drained = false;
            for (uint32_t b = 0; b < nb; b++) {
                BlockRec& r = blocks[b];
                r.bit_offset = bitpos;
                if (r.type == BT_STORED) {
                    bitpos = ((bitpos + 3 + 7) & ~7ull) + 32 + 8ull * r.in_len;  // Trees.cs:742-746, Deflate.cs:710-722
                } else {
                    bitpos += 3 + r.hdr_bits + r.body_bits;
                }
                drained = dr.step(bitpos);
            }
            TailSink t;
            for (int i = 0; i < kTailWords; i++) t.w[i] = 0;
            t.nbits = 0;
            const uint64_t tail_start = bitpos;
            if (P.finish_mode) {
                // NoFlush + Finish (Deflate.cs:616-635): the last block was sent with BFINAL and Bi_windup pads to a byte; nothing
                // is flushed in between, so the 512-byte chunk arithmetic never adds a block
                bitpos = (bitpos + 7) & ~7ull;
                drained = false;
            }
            // the block a repeated flush finds nothing to put into: empty static (3 + 7 bits), or at level 0 empty stored
            auto empty_block = [&]() {
                if (P.level0) {
                    t.put(0u, 3);
                    bitpos += 3;
                    const uint32_t pad0 = (uint32_t)((8 - (bitpos & 7)) & 7);
                    if (pad0) t.put(0u, (int)pad0);
                    bitpos += pad0;
                    t.put(0xFFFF0000u, 32);
                    bitpos += 32;
                } else {
                    t.put(0x2u, 3);  // BFINAL=0, static
                    t.put(0u, 7);    // END_BLOCK of the fixed tree
                    bitpos += 10;
                }
            };
            // `do { block with the remaining symbols; step } while (drained)`: the first one is block nb-1
            while (drained) {
                empty_block();
                drained = dr.step(bitpos);
            }
            for (; !P.finish_mode;) {
                // Tr_stored_block(0, 0, false): 3 type bits, Bi_windup, LEN=0000 NLEN=FFFF
                t.put(0u, 3);
                bitpos += 3;
                uint32_t pad = (uint32_t)((8 - (bitpos & 7)) & 7);
                if (pad) t.put(0u, (int)pad);
                bitpos += pad;
                t.put(0xFFFF0000u, 32);
                bitpos += 32;
                drained = dr.step(bitpos);
                if (!drained) break;
                do {
                    empty_block();
                    drained = dr.step(bitpos);
                } while (drained);
            }
            bytes = bitpos >> 3;
            lay.out_bytes = (uint32_t)bytes;
            lay.tail_nbits = t.nbits;
            lay.tail_bit_offset = tail_start;
            for (int i = 0; i < kTailWords; i++) lay.tail[i] = t.w[i];
        }
        // CTA-wide exclusive scan of `bytes`
        uint64_t incl = bytes;
        for (int off = 1; off < 32; off <<= 1) {
            uint64_t v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint64_t ws = warp_sums[lane];
            uint64_t wi = ws;
            for (int off = 1; off < 32; off <<= 1) {
                uint64_t v = __shfl_up_sync(0xffffffffu, wi, off);
                if (lane >= off) wi += v;
            }
            warp_sums[lane] = wi - ws;  // exclusive
        }
        __syncthreads();
        const uint64_t carry = carry_s;
        const uint64_t excl = carry + warp_sums[warp] + incl - bytes;
        if (seg < P.nseg) {
            lay.out_offset = excl;
            P.layout[seg] = lay;
            if (P.seg_end) P.seg_end[seg] = excl + bytes;
        }
        __syncthreads();
        if (tid == LAYOUT_THREADS - 1) carry_s = excl + bytes;
        __syncthreads();
    }
    if (tid == 0) {
        *P.total_out = carry_s;
        if (carry_s > P.out_cap) *P.status = -5;  // ZSB_BUF_ERROR
    }
}

// ---- bit packer ---------------------------------------------------------------------------------------

namespace {

constexpr int PK_THREADS = 256;
constexpr int PK_SYMS_PER_THREAD = 4;
constexpr int PK_CHUNK = PK_THREADS * PK_SYMS_PER_THREAD;  // 1024 symbols, <= 48 bits each
constexpr int PK_STAGE_WORDS = 2048;                        // 65536 bits

struct PackSmem {
    uint32_t stage[PK_STAGE_WORDS + 4];
    uint32_t lcode[L_CODES];
    uint32_t dcode[D_CODES];
    uint32_t warp_sums[PK_THREADS / 32];
};

struct Packer {
    PackSmem* sm;
    uint8_t* out;            // global output
    uint64_t seg_byte0;      // first output byte of the segment
    uint64_t seg_byte_end;   // one past the last
    uint64_t stage_word0;    // global 32-bit word index that stage[0] maps to
    uint64_t bitpos;         // next free bit, absolute (global byte offset * 8 + bit)
    bool shared_edges;       // FlushMode.PartialFlush: the first and the last byte may hold a neighbour's bits too (OR them in)

    __device__ __forceinline__ void or_bits(uint64_t at, uint64_t value, int len)
    {
        // at = absolute bit position; value has `len` (<= 57) valid bits
        uint64_t rel = at - stage_word0 * 32;
        uint32_t idx = (uint32_t)(rel >> 5), sh = (uint32_t)(rel & 31);
        uint64_t lo = value << sh;
        atomicOr(&sm->stage[idx], (uint32_t)lo);
        if (sh + len > 32) atomicOr(&sm->stage[idx + 1], (uint32_t)(lo >> 32));
        if (sh + len > 64) atomicOr(&sm->stage[idx + 2], (uint32_t)(value >> (64 - sh)));
    }

    // Writes every staged word that is complete (or all of them when `final`) and re-bases the stage.
    __device__ void flush(bool final)
    {
        __syncthreads();
        const int tid = threadIdx.x;
        const uint64_t end_word = final ? (bitpos + 31) >> 5 : bitpos >> 5;  // words [stage_word0, end_word) are written
        const uint32_t nwords = (uint32_t)(end_word - stage_word0);
        for (uint32_t i = tid; i < nwords; i += PK_THREADS) {
            const uint64_t gw = stage_word0 + i;
            const uint64_t b0 = gw * 4;
            const uint32_t v = sm->stage[i];
            if (shared_edges) {
                // the output was zeroed; a word that holds the segment's first or last byte is merged, the others are stored
                if (b0 > seg_byte0 && b0 + 4 < seg_byte_end) {
                    *reinterpret_cast<uint32_t*>(out + b0) = v;
                } else {
                    uint32_t m = 0;
                    for (int k = 0; k < 4; k++) {
                        const uint64_t b = b0 + k;
                        if (b >= seg_byte0 && b < seg_byte_end) m |= 0xFFu << (8 * k);
                    }
                    if (v & m) atomicOr(reinterpret_cast<uint32_t*>(out + b0), v & m);
                }
            } else if (b0 >= seg_byte0 && b0 + 4 <= seg_byte_end) {
                *reinterpret_cast<uint32_t*>(out + b0) = v;
            } else {
                for (int k = 0; k < 4; k++) {
                    uint64_t b = b0 + k;
                    if (b >= seg_byte0 && b < seg_byte_end) out[b] = (uint8_t)(v >> (8 * k));
                }
            }
        }
        __syncthreads();
        const uint32_t keep = final ? 0u : sm->stage[nwords];
        __syncthreads();
        for (uint32_t i = tid; i < PK_STAGE_WORDS + 4; i += PK_THREADS) sm->stage[i] = 0;
        __syncthreads();
        if (tid == 0) sm->stage[0] = keep;
        stage_word0 = end_word;
        __syncthreads();
    }
    __device__ __forceinline__ uint64_t stage_bits_free() const { return (stage_word0 + PK_STAGE_WORDS) * 32 - bitpos; }
};

}  // namespace

__global__ void __launch_bounds__(PK_THREADS) pack_kernel(HuffParams P)
{
    __shared__ PackSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t seg = blockIdx.x;
    if (seg >= P.nseg) return;
    if (*P.status != 0) return;
    const SegLayout& lay = P.layout[seg];
    const BlockRec* blocks = P.blocks + (size_t)seg * P.max_blocks;
    const uint32_t nb = P.nblocks[seg];
    const uint32_t* seg_sym = P.sym + (size_t)seg * P.sym_stride;
    const uint8_t* seg_in = P.in + (size_t)seg * P.seg_size;

    Packer pk;
    pk.sm = &sm;
    pk.out = P.out;
    pk.seg_byte0 = lay.out_offset;
    pk.seg_byte_end = lay.out_offset + lay.out_bytes;
    pk.bitpos = lay.out_offset * 8 + lay.bit_shift;
    pk.stage_word0 = pk.bitpos >> 5;
    pk.shared_edges = P.partial_flush != 0;
    for (uint32_t i = tid; i < PK_STAGE_WORDS + 4; i += PK_THREADS) sm.stage[i] = 0;
    __syncthreads();
    const uint64_t seg_bit0 = lay.out_offset * 8 + lay.bit_shift;

    for (uint32_t b = 0; b < nb; b++) {
        const BlockRec& r = blocks[b];
        const BlockCodesDev& codes = P.codes[(size_t)seg * P.max_blocks + b];
        // block type bits: BFINAL = 0 (Trees.cs:755-756)
        if (pk.stage_bits_free() < 4096) pk.flush(false);
        if (tid == 0) pk.or_bits(seg_bit0 + r.bit_offset, (uint64_t)(r.type << 1) | (uint64_t)((P.finish_mode && b + 1 == nb) ? 1u : 0u), 3);
        pk.bitpos = seg_bit0 + r.bit_offset + 3;
        if (r.type == BT_STORED) {
            // Copy_block (Deflate.cs:710-722): byte align, LEN, NLEN, raw bytes
            pk.bitpos = (pk.bitpos + 7) & ~7ull;
            const uint32_t len = r.in_len;
            if (tid == 0) pk.or_bits(pk.bitpos, (uint64_t)(len & 0xFFFF) | ((uint64_t)((~len) & 0xFFFF) << 16), 32);
            pk.bitpos += 32;
            pk.flush(true);  // everything so far is whole bytes
            const uint64_t dst = pk.bitpos >> 3;
            const uint8_t* src = seg_in + r.in_begin;
            for (uint32_t i = tid; i < len; i += PK_THREADS) P.out[dst + i] = src[i];
            pk.bitpos += 8ull * len;
            pk.stage_word0 = pk.bitpos >> 5;
            // the bytes of the word that precede bitpos belong to the copy above: they are masked out by the
            // byte-range check in flush() only if we move seg_byte0 forward
            pk.seg_byte0 = pk.bitpos >> 3;
            __syncthreads();
            continue;
        }
        for (int i = tid; i < L_CODES; i += PK_THREADS) sm.lcode[i] = codes.lcode[i];
        for (int i = tid; i < D_CODES; i += PK_THREADS) sm.dcode[i] = codes.dcode[i];
        if (r.type == BT_DYNAMIC) {
            // tree description, 32 bits per thread
            const uint32_t hb = r.hdr_bits;
            for (uint32_t i = tid; i * 32 < hb; i += PK_THREADS) {
                uint32_t w = codes.hdr[i];
                int len = (int)(hb - i * 32 < 32 ? hb - i * 32 : 32);
                if (len < 32) w &= (1u << len) - 1;
                pk.or_bits(pk.bitpos + i * 32ull, w, len);
            }
            pk.bitpos += hb;
        }
        __syncthreads();
        // symbols, PK_CHUNK at a time: Compress_block (Trees.cs:956-989) with the bit offsets from a prefix sum
        const uint32_t* sym = seg_sym + r.sym_begin;
        const uint32_t nsym = r.sym_count;
        // the symbols of the NEXT chunk are loaded while this one is packed (one HBM round trip per chunk otherwise)
        uint32_t nxt[PK_SYMS_PER_THREAD];
#pragma unroll
        for (int k = 0; k < PK_SYMS_PER_THREAD; k++) {
            const uint32_t i = tid * PK_SYMS_PER_THREAD + k;
            nxt[k] = i < nsym ? __ldg(sym + i) : 0u;
        }
        for (uint32_t c0 = 0; c0 < nsym; c0 += PK_CHUNK) {
            if (pk.stage_bits_free() < (uint64_t)PK_CHUNK * 48 + 64) pk.flush(false);
            uint64_t bits[PK_SYMS_PER_THREAD];
            int lens[PK_SYMS_PER_THREAD];
            uint32_t total = 0;
            uint32_t cur_s[PK_SYMS_PER_THREAD];
#pragma unroll
            for (int k = 0; k < PK_SYMS_PER_THREAD; k++) {
                cur_s[k] = nxt[k];
                const uint32_t i2 = c0 + PK_CHUNK + tid * PK_SYMS_PER_THREAD + k;
                nxt[k] = i2 < nsym ? __ldg(sym + i2) : 0u;
            }
#pragma unroll
            for (int k = 0; k < PK_SYMS_PER_THREAD; k++) {
                const uint32_t i = c0 + tid * PK_SYMS_PER_THREAD + k;
                bits[k] = 0;
                lens[k] = 0;
                if (i < nsym) {
                    const uint32_t s = cur_s[k];
                    const uint32_t dist = s & 0xFFFF, lc = s >> 16;
                    if (dist == 0) {
                        const uint32_t e = sm.lcode[lc];
                        bits[k] = e & 0xFFFF;
                        lens[k] = (int)(e >> 16);
                    } else {
                        const int code = length_code_of((int)lc);
                        uint32_t e = sm.lcode[code + LITERALS + 1];
                        uint64_t v = e & 0xFFFF;
                        int n = (int)(e >> 16);
                        int extra = length_extra_bits(code);
                        if (extra) { v |= (uint64_t)(lc - (uint32_t)length_base(code)) << n; n += extra; }
                        const int d = (int)dist - 1;
                        const int dc = dist_code_of(d);
                        e = sm.dcode[dc];
                        v |= (uint64_t)(e & 0xFFFF) << n;
                        n += (int)(e >> 16);
                        extra = dist_extra_bits(dc);
                        if (extra) { v |= (uint64_t)(d - dist_base(dc)) << n; n += extra; }
                        bits[k] = v;
                        lens[k] = n;
                    }
                    total += (uint32_t)lens[k];
                }
            }
            // exclusive prefix sum of `total` over the CTA
            uint32_t incl = total;
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += v;
            }
            if (lane == 31) sm.warp_sums[warp] = incl;
            __syncthreads();
            uint32_t warp_base = 0, chunk_total = 0;
            for (int w = 0; w < PK_THREADS / 32; w++) {
                uint32_t v = sm.warp_sums[w];
                if (w < warp) warp_base += v;
                chunk_total += v;
            }
            uint64_t at = pk.bitpos + warp_base + incl - total;
#pragma unroll
            for (int k = 0; k < PK_SYMS_PER_THREAD; k++) {
                if (lens[k]) pk.or_bits(at, bits[k], lens[k]);
                at += (uint32_t)lens[k];
            }
            pk.bitpos += chunk_total;
            __syncthreads();
        }
        // END_BLOCK
        if (pk.stage_bits_free() < 64) pk.flush(false);
        if (tid == 0) pk.or_bits(pk.bitpos, sm.lcode[END_BLOCK] & 0xFFFF, (int)(sm.lcode[END_BLOCK] >> 16));
        pk.bitpos += sm.lcode[END_BLOCK] >> 16;
        __syncthreads();
    }
    // tail: extra empty static blocks and the full-flush marker(s)
    if (pk.stage_bits_free() < 512) pk.flush(false);
    {
        const uint32_t tb = lay.tail_nbits;
        const uint64_t t0 = seg_bit0 + lay.tail_bit_offset;
        for (uint32_t i = tid; i * 32 < tb; i += PK_THREADS) {
            uint32_t w = lay.tail[i];
            int len = (int)(tb - i * 32 < 32 ? tb - i * 32 : 32);
            if (len < 32) w &= (1u << len) - 1;
            pk.or_bits(t0 + i * 32ull, w, len);
        }
        pk.bitpos = t0 + tb;
    }
    pk.flush(true);
}

cudaError_t launch_block_trees(const HuffParams& P, cudaStream_t stream)
{
    const uint64_t grid = (uint64_t)P.nseg * P.max_blocks;
    if (grid == 0) return cudaSuccess;
    block_trees_kernel<<<(unsigned)grid, BT_THREADS, 0, stream>>>(P);
    return cudaGetLastError();
}
cudaError_t launch_layout(const HuffParams& P, cudaStream_t stream)
{
    layout_kernel<<<1, LAYOUT_THREADS, 0, stream>>>(P);
    return cudaGetLastError();
}
cudaError_t launch_pack(const HuffParams& P, cudaStream_t stream)
{
    if (P.nseg == 0) return cudaSuccess;
    pack_kernel<<<P.nseg, PK_THREADS, 0, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace zsb
