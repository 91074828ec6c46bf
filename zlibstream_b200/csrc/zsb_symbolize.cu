// zsb_symbolize.cu -- the parse-free strategies, fully parallel: HuffmanOnly and Rle.
//
// Replaces, for these two strategies, what zsb_lz77.cu does with one thread: DeflateRle (Deflate.Rle.cs:18-104) and the
// match-less DeflateFast / DeflateSlow of HuffmanOnly (Deflate.Fast.cs:57-63, Deflate.Slow.cs:66), plus Tr_tally_*
// (Deflate.cs:910-948), the block cut at 16383 symbols (Q9) and the Adler-32 of the input (ZlibStream.cs:213-216).
//
// Neither strategy looks at the hash table, so the symbol stream is a pure function of the segment's bytes (and, for
// Rle, of the byte before the segment, Q14):
//   HuffmanOnly  every byte is a literal;
//   Rle          let eq[p] = (byte[p] == byte[p-1]).  A position with eq = 0 is a literal.  A maximal stretch of M
//                positions with eq = 1 is parsed greedily from its first position: matches of distance 1 and length 258
//                while at least 258 remain, then one match of the remainder r = M mod 258 if r >= 3, else r literals
//                (Deflate.Rle.cs:47-76: the scan needs three equal bytes, is capped at MAX_MATCH and at lookahead).
// So every position can decide on its own whether a symbol starts there and which one: it needs the previous and the
// next position with eq = 0 (two scans over the segment's eq bit mask), and its symbol index is a prefix sum of the
// start flags.  Blocks are cut every 16383 symbols; whether the block's first byte is still inside the window when it
// is flushed (`buf != -1`, Trees.cs:608) follows from the slide schedule of Fill_window (Deflate.cs:967-1019), which
// one thread replays over the parse points.
//
// One CTA per segment, the segment's masks in shared memory (20 bytes per 32 input bytes); segments above 256 KiB keep
// using the serial path in zsb_lz77.cu.
#include "zsb_kernels.cuh"

namespace zsb {

namespace {

constexpr int SY_THREADS = 256;
constexpr uint32_t SY_MAX_SEG = 262144;
constexpr int SY_MAX_BLOCKS = SY_MAX_SEG / BLOCK_SYMS + 3;

struct SySmem {
    uint32_t* eq;       // [nwords] bit i of word w: byte[32w+i] == byte[32w+i-1]   (0 beyond the segment)
    uint32_t* start;    // [nwords] a symbol starts here
    uint32_t* base;     // [nwords] symbols that start before this word
    int* zprev;         // [nwords] last position with eq = 0 in words <= w, -1 if none
    int* znext;         // [nwords] first position with eq = 0 in words >= w, seg_len if none
};

// Block-wide exclusive scan helpers over one value per thread (SY_THREADS threads).
__device__ __forceinline__ uint32_t block_excl_sum(uint32_t v, uint32_t* red, uint32_t& total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) red[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
    for (int w = 0; w < SY_THREADS / 32; w++) {
        const uint32_t t = red[w];
        if (w < warp) woff += t;
        tot += t;
    }
    __syncthreads();
    total = tot;
    return woff + inc - v;
}
__device__ __forceinline__ int block_excl_max(int v, int* red)   // max of the values of earlier threads, -1 if none
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = max(inc, t);
    }
    if (lane == 31) red[warp] = inc;
    __syncthreads();
    int woff = -1;
    for (int w = 0; w < warp; w++) woff = max(woff, red[w]);
    __syncthreads();
    int prev = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) prev = -1;
    return max(woff, prev);
}
__device__ __forceinline__ int block_excl_min_rev(int v, int* red, int none)   // min of the values of later threads
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_down_sync(0xffffffffu, inc, o);
        if (lane + o < 32) inc = min(inc, t);
    }
    if (lane == 0) red[warp] = inc;
    __syncthreads();
    int woff = none;
    for (int w = warp + 1; w < SY_THREADS / 32; w++) woff = min(woff, red[w]);
    __syncthreads();
    int next = __shfl_down_sync(0xffffffffu, inc, 1);
    if (lane == 31) next = none;
    return min(woff, next);
}

// What starts at position p of a stretch of eq = 1 positions that began at a + 1 and ends before e:
// 0 nothing (inside a match), 1 a literal, >= 3 a match of that length.
__device__ __forceinline__ int rle_symbol_at(int p, int a, int e)
{
    const int j = p - a - 1, M = e - a - 1, rem = e - p;
    const int jm = j % MAX_MATCH;
    if (jm == 0 && rem >= MIN_MATCH) return rem < MAX_MATCH ? rem : MAX_MATCH;
    const int r = M % MAX_MATCH;
    if (r < MIN_MATCH && j >= M - r) return 1;
    return 0;
}

}  // namespace

__global__ void __launch_bounds__(SY_THREADS) symbolize_segments_kernel(Lz77Params P)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ uint32_t red[16];
    __shared__ unsigned long long red64[2 * SY_THREADS / 32];
    __shared__ int blk_first[SY_MAX_BLOCKS], blk_last[SY_MAX_BLOCKS];
    const int tid = threadIdx.x;
    const int nwords = (int)(P.seg_size / 32);
    const int K = nwords / SY_THREADS;   // words per thread (seg_size is a multiple of 32768)
    SySmem sm;
    sm.eq = reinterpret_cast<uint32_t*>(smem_raw);
    sm.start = sm.eq + nwords;
    sm.base = sm.start + nwords;
    sm.zprev = reinterpret_cast<int*>(sm.base + nwords);
    sm.znext = sm.zprev + nwords;
    const bool rle = P.fn == FN_RLE;

    for (uint32_t seg = blockIdx.x; seg < P.nseg; seg += gridDim.x) {
        const size_t seg_off = (size_t)seg * P.seg_size;
        const int n = (int)((P.n - seg_off) < (size_t)P.seg_size ? (P.n - seg_off) : (size_t)P.seg_size);
        const uint8_t* in = P.in + seg_off;
        const bool stream_first = (P.first_seg_index + (uint64_t)seg) == 0;
        // Q14: the byte before the segment (window[32767] of the serial path)
        uint32_t before = 0x100;   // matches nothing
        if (rle && !stream_first) before = seg > 0 ? in[-1] : (P.halo ? P.halo[W_SIZE - 1] : 0);
        for (int i = tid; i < SY_MAX_BLOCKS; i += SY_THREADS) blk_first[i] = blk_last[i] = 0;

        // ---- A: eq masks + Adler partial sums --------------------------------------------------------------------
        unsigned long long s1 = 0, s2 = 0;
        const bool aligned = (reinterpret_cast<uintptr_t>(in) & 15) == 0;
        for (int k = 0; k < K; k++) {
            const int w = tid * K + k, p0 = w * 32;
            uint32_t E = 0;
            if (p0 < n) {
                uint32_t x[8];
                if (aligned && p0 + 32 <= n) {
                    const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(in + p0)), v1 = __ldg(reinterpret_cast<const uint4*>(in + p0 + 16));
                    x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w;
                    x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
                } else {
                    for (int q = 0; q < 8; q++) {
                        uint32_t v = 0;
                        for (int b = 0; b < 4; b++) {
                            const int p = p0 + q * 4 + b;
                            if (p < n) v |= (uint32_t)in[p] << (8 * b);
                        }
                        x[q] = v;
                    }
                }
                uint32_t carry = p0 == 0 ? before : (uint32_t)in[p0 - 1];
                const int valid = n - p0 < 32 ? n - p0 : 32;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const uint32_t v = x[q];
                    if (rle) {
                        const uint32_t sh = (v << 8) | (carry & 0xFFu);
                        uint32_t m = __vcmpeq4(v, sh);                  // 0xFF per byte equal to its predecessor
                        if (carry > 0xFFu) m &= 0xFFFFFF00u;             // no predecessor
                        m &= 0x01010101u;
                        E |= ((m * 0x01020408u) >> 24) << (4 * q);
                        carry = v >> 24;
                    }
                    const uint32_t b0 = v & 0xFF, b1 = (v >> 8) & 0xFF, b2 = (v >> 16) & 0xFF, b3 = v >> 24;
                    const uint32_t t1 = b0 + b1 + b2 + b3;
                    s1 += t1;   // bytes beyond n are zero in x[]
                    s2 += (unsigned long long)(uint32_t)(n - (p0 + q * 4)) * t1 - (b1 + 2 * b2 + 3 * b3);
                }
                if (valid < 32) E &= (1u << valid) - 1u;
            }
            sm.eq[w] = E;
        }
        // Adler-32 halves of the segment, seed 1: a = 1 + s1, b = n + s2   (mod 65521)
        for (int off = 16; off > 0; off >>= 1) {
            s1 += __shfl_down_sync(0xffffffffu, s1, off);
            s2 += __shfl_down_sync(0xffffffffu, s2, off);
        }
        if ((tid & 31) == 0) {
            red64[(tid >> 5) * 2] = s1;
            red64[(tid >> 5) * 2 + 1] = s2;
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long t1 = 0, t2 = 0;
            for (int w = 0; w < SY_THREADS / 32; w++) {
                t1 += red64[w * 2];
                t2 += red64[w * 2 + 1];
            }
            P.seg_adler[seg] = make_uint2((uint32_t)((1 + t1) % ADLER_BASE), (uint32_t)(((unsigned long long)n + t2) % ADLER_BASE));
        }

        // ---- B: previous / next position with eq = 0, per word ------------------------------------------------------
        {
            int run = -1;
            for (int k = 0; k < K; k++) {
                const int w = tid * K + k;
                const uint32_t Z = ~sm.eq[w];
                if (Z) run = w * 32 + 31 - __clz(Z);
                sm.zprev[w] = run;
            }
            const int carry = block_excl_max(run, reinterpret_cast<int*>(red));
            for (int k = 0; k < K; k++) {
                const int w = tid * K + k;
                if (sm.zprev[w] < 0) sm.zprev[w] = carry;
                else break;
            }
            run = 0x7fffffff;
            for (int k = K - 1; k >= 0; k--) {
                const int w = tid * K + k;
                const uint32_t Z = ~sm.eq[w];
                if (Z) run = w * 32 + __ffs(Z) - 1;
                sm.znext[w] = run;
            }
            const int carry2 = block_excl_min_rev(run, reinterpret_cast<int*>(red), n);
            for (int k = K - 1; k >= 0; k--) {
                const int w = tid * K + k;
                if (sm.znext[w] == 0x7fffffff) sm.znext[w] = carry2;
                else break;
            }
        }
        __syncthreads();

        // ---- C: where symbols start ---------------------------------------------------------------------------------
        uint32_t my_count = 0;
        for (int k = 0; k < K; k++) {
            const int w = tid * K + k, p0 = w * 32;
            uint32_t S = 0;
            if (p0 < n) {
                const uint32_t E = sm.eq[w], Z = ~E;
                const int valid = n - p0 < 32 ? n - p0 : 32;
                const uint32_t vmask = valid < 32 ? (1u << valid) - 1u : 0xffffffffu;
                S = Z & vmask;   // literals at eq = 0
                uint32_t todo = E & vmask;
                int a = w > 0 ? sm.zprev[w - 1] : -1;
                const int e_out = w + 1 < nwords ? min(sm.znext[w + 1], n) : n;
                while (todo) {
                    const int i = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const uint32_t zb = Z & ((1u << i) - 1u);
                    const int aa = zb ? p0 + 31 - __clz(zb) : a;
                    const uint32_t za = i < 31 ? Z & (0xffffffffu << (i + 1)) : 0;
                    const int ee = za ? min(p0 + __ffs(za) - 1, n) : e_out;
                    if (rle_symbol_at(p0 + i, aa, ee)) S |= 1u << i;
                }
            }
            sm.start[w] = S;
            my_count += __popc(S);
        }
        uint32_t total = 0;
        {
            uint32_t run = block_excl_sum(my_count, red, total);
            for (int k = 0; k < K; k++) {
                const int w = tid * K + k;
                sm.base[w] = run;
                run += __popc(sm.start[w]);
            }
        }
        __syncthreads();

        // ---- D: write the symbols (Tr_tally_lit / Tr_tally_dist), note where blocks begin and end ----------------------
        uint32_t* sym = P.sym + (size_t)seg * P.sym_stride;
        const bool sym_aligned = (reinterpret_cast<uintptr_t>(sym) & 15) == 0;
        for (int k = 0; k < K; k++) {
            const int w = tid * K + k, p0 = w * 32;
            uint32_t S = sm.start[w];
            if (!S) continue;
            const uint32_t E = sm.eq[w], Z = ~E;
            uint32_t idx = sm.base[w];
            // the word's 32 bytes, once (the per-literal byte loads of the first version went to L2 one at a time)
            uint32_t x[8];
            const bool whole = aligned && p0 + 32 <= n;
            if (whole) {
                const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(in + p0)), v1 = __ldg(reinterpret_cast<const uint4*>(in + p0 + 16));
                x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w;
                x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
            }
            if (whole && S == 0xffffffffu && E == 0) {
                // 32 literals in a row (HuffmanOnly always, Rle on run-free data): records straight from the registers
                const uint32_t blk0 = idx / BLOCK_SYMS, off0 = idx - blk0 * BLOCK_SYMS;
                if (off0 == 0 && blk0 < SY_MAX_BLOCKS) blk_first[blk0] = p0;
                const uint32_t j = BLOCK_SYMS - 1 - off0;   // the record that completes block blk0, if it is one of these
                if (j < 32) {
                    if (blk0 < SY_MAX_BLOCKS) blk_last[blk0] = p0 + (int)j;
                    if (j + 1 < 32 && blk0 + 1 < SY_MAX_BLOCKS) blk_first[blk0 + 1] = p0 + (int)j + 1;
                }
                if (sym_aligned && (idx & 3u) == 0) {
                    uint4* d = reinterpret_cast<uint4*>(sym + idx);
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const uint32_t v = x[q];
                        d[q] = make_uint4((v & 0xFFu) << 16, (v & 0xFF00u) << 8, v & 0xFF0000u, (v >> 8) & 0xFF0000u);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; i++) sym[idx + i] = ((x[i >> 2] >> (8 * (i & 3))) & 0xFFu) << 16;
                }
                continue;
            }
            const int a = w > 0 ? sm.zprev[w - 1] : -1;
            const int e_out = w + 1 < nwords ? min(sm.znext[w + 1], n) : n;
            while (S) {
                const int i = __ffs(S) - 1;
                S &= S - 1;
                const int p = p0 + i;
                int s = 1;
                if (!((Z >> i) & 1u)) {
                    const uint32_t zb = Z & ((1u << i) - 1u);
                    const int aa = zb ? p0 + 31 - __clz(zb) : a;
                    const uint32_t za = i < 31 ? Z & (0xffffffffu << (i + 1)) : 0;
                    const int ee = za ? min(p0 + __ffs(za) - 1, n) : e_out;
                    s = rle_symbol_at(p, aa, ee);
                }
                uint32_t rec;
                if (s >= MIN_MATCH) rec = 1u | ((uint32_t)(s - MIN_MATCH) << 16);
                else if (whole) {
                    const int q = i >> 2;   // byte i of the registers, without indexing them dynamically
                    const uint32_t lo = q & 1 ? ((q & 2) ? ((q & 4) ? x[7] : x[3]) : ((q & 4) ? x[5] : x[1]))
                                              : ((q & 2) ? ((q & 4) ? x[6] : x[2]) : ((q & 4) ? x[4] : x[0]));
                    rec = ((lo >> (8 * (i & 3))) & 0xFFu) << 16;
                } else rec = (uint32_t)__ldg(in + p) << 16;
                sym[idx] = rec;
                const uint32_t blk = idx / BLOCK_SYMS, off = idx - blk * BLOCK_SYMS;
                if (off == 0 && blk < SY_MAX_BLOCKS) blk_first[blk] = p;
                if (off == BLOCK_SYMS - 1 && blk < SY_MAX_BLOCKS) blk_last[blk] = p;
                idx++;
            }
        }
        __syncthreads();

        // ---- E: block records (Flush_block_only, Deflate.cs:951-956) ---------------------------------------------------
        if (tid == 0) {
            const uint32_t full = total / BLOCK_SYMS, rem = total - full * BLOCK_SYMS;
            // DeflateSlow ignores the flush request of its very last literal (Deflate.Slow.cs:148-156): no empty block then
            uint32_t nb = full + 1;
            if (!rle && P.fn == FN_SLOW && rem == 0 && total > 0) nb = full;
            // Rle: replay the slides of Fill_window over the parse points (see the header of this file)
            int ev[SY_MAX_SEG / W_SIZE + 4];
            int nev = 0;
            const int base0 = stream_first ? 0 : W_SIZE;
            if (rle) {
                int base = base0;
                for (;;) {
                    const int F = min(n, WINDOW_SIZE - base);
                    int X = F < n ? (WINDOW_SIZE - MAX_MATCH) - base : max(n - MAX_MATCH, (WINDOW_SIZE - MIN_LOOKAHEAD) - base);
                    if (X < 0) X = 0;
                    if (X > n) break;
                    // first parse point >= X (the loop's last visit, at n, counts)
                    int Pp = n;
                    for (int w = X / 32; w < nwords && w * 32 < n; w++) {
                        uint32_t S = sm.start[w];
                        if (w == X / 32) S &= 0xffffffffu << (X & 31);
                        if (S) {
                            Pp = w * 32 + __ffs(S) - 1;
                            break;
                        }
                    }
                    if (nev < (int)(sizeof ev / sizeof ev[0])) ev[nev++] = Pp;
                    base -= W_SIZE;
                }
            }
            BlockRec* blocks = P.blocks + (size_t)seg * P.max_blocks;
            for (uint32_t k = 0; k < nb && k < P.max_blocks; k++) {
                BlockRec& b = blocks[k];
                const uint32_t cnt = k < full ? (uint32_t)BLOCK_SYMS : rem;
                const int in_begin = k * BLOCK_SYMS < total ? (k == 0 ? 0 : blk_first[k]) : n;
                const bool last = k + 1 == nb;
                const int in_end = (!last && (k + 1) * BLOCK_SYMS < total) ? blk_first[k + 1] : n;
                b.sym_begin = k * BLOCK_SYMS;
                b.sym_count = cnt;
                b.in_begin = (uint32_t)in_begin;
                b.in_len = (uint32_t)(in_end - in_begin);
                int valid = 1;
                if (rle) {
                    const int flush_at = last ? n : blk_last[k];
                    int base = base0;
                    for (int i = 0; i < nev; i++)
                        if (ev[i] <= flush_at) base -= W_SIZE;
                    valid = in_begin + base >= 0;
                }
                b.buf_valid = (uint32_t)valid;
            }
            P.nblocks[seg] = nb;
        }
        __syncthreads();
    }
}

bool symbolize_applies(const Lz77Params& P)
{
    if (P.seg_size > SY_MAX_SEG || P.seg_size % W_SIZE != 0) return false;
    return P.fn == FN_RLE || (P.strategy == 2 && (P.fn == FN_FAST || P.fn == FN_SLOW));
}

cudaError_t launch_symbolize(const Lz77Params& P, int num_sms, cudaStream_t stream)
{
    const size_t smem = (size_t)(P.seg_size / 32) * 20;
    {   // a per-device attribute: set on every launch (a process may drive several GPUs, one zsb_ctx each)
        cudaError_t e = cudaFuncSetAttribute(symbolize_segments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, symbolize_segments_kernel, SY_THREADS, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    int grid = num_sms * per_sm;
    if ((uint32_t)grid > P.nseg) grid = (int)P.nseg;
    if (grid < 1) grid = 1;
    symbolize_segments_kernel<<<grid, SY_THREADS, smem, stream>>>(P);
    return cudaGetLastError();
}

// ---- level 0 (DeflateStored, Deflate.Stored.cs:24-84) -----------------------------------------------------------------------
// One thread per segment replays the loop's arithmetic: Fill_window's refill / slide schedule (Deflate.cs:967-1019; level 0
// calls it when lookahead <= 1) and the two flush rules -- a block is cut at max_block_size = pending buffer - 5 = 32763
// bytes (memLevel is 7 at level 0, Deflate.cs:19-20,246-249, so the pending buffer is 32 KiB) and whenever w_size -
// MIN_LOOKAHEAD bytes are waiting (:69-76); the end-of-input flush (:80) closes whatever is left, possibly nothing.
__global__ void __launch_bounds__(256) stored_blocks_kernel(Lz77Params P)
{
    const uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= P.nseg) return;
    const size_t off = (size_t)seg * P.seg_size;
    const int len = (int)((P.n - off) < (size_t)P.seg_size ? (P.n - off) : (size_t)P.seg_size);
    const bool first = (P.first_seg_index + (uint64_t)seg) == 0;
    BlockRec* b = P.blocks + (size_t)seg * P.max_blocks;
    constexpr int MAX_BLOCK = 32768 - 5;
    uint32_t nb = 0;
    int str_start = first ? 0 : W_SIZE, block_start = str_start, lookahead = 0, avail = len;
    int base = first ? 0 : -W_SIZE;   // segment offset of window index 0
    auto flush = [&]() {
        if (nb < P.max_blocks) {
            BlockRec r{};
            r.in_begin = (uint32_t)(block_start + base);
            r.in_len = (uint32_t)(str_start - block_start);
            r.buf_valid = block_start >= 0;
            b[nb] = r;
        }
        nb++;
        block_start = str_start;
    };
    for (;;) {
        if (lookahead <= 1) {
            do {   // Fill_window
                int more = WINDOW_SIZE - lookahead - str_start;
                if (str_start >= WINDOW_SIZE - MIN_LOOKAHEAD) {
                    str_start -= W_SIZE;
                    block_start -= W_SIZE;
                    base += W_SIZE;
                    more += W_SIZE;
                }
                if (avail == 0) break;
                const int n = avail < more ? avail : more;
                avail -= n;
                lookahead += n;
            } while (lookahead < MIN_LOOKAHEAD && avail != 0);
            if (lookahead == 0) break;
        }
        str_start += lookahead;
        lookahead = 0;
        const int max_start = block_start + MAX_BLOCK;
        if (str_start == 0 || str_start >= max_start) {
            lookahead = str_start - max_start;
            str_start = max_start;
            flush();
        }
        if (str_start - block_start >= W_SIZE - MIN_LOOKAHEAD) flush();
    }
    flush();
    P.nblocks[seg] = nb;
}

cudaError_t launch_stored_blocks(const Lz77Params& P, cudaStream_t stream)
{
    if (P.nseg == 0) return cudaSuccess;
    stored_blocks_kernel<<<(P.nseg + 255) / 256, 256, 0, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace zsb
