// zsb_lz77_lists.cu -- level 1 (DeflateFast), kernel A of two: candidate lists for EVERY position of a segment.
//
// The reference's hash chain (InsertString, Deflate.cs:866-877) links only the positions the greedy parse inserts
// (parse points and the interiors of matches <= max_lazy, Deflate.Fast.cs:81-104), so it depends on the parse.  The chain
// over ALL positions with the same 15-bit hash does not, and the reference's chain is that chain filtered by an
// "inserted" bit per position.  This kernel builds the all-position chain of a segment and stores, for every position,
// its LK nearest predecessors with the common prefix length of the two strings (what Longest_match + Compare256 would
// compute, Deflate.cs:1022-1100) -- fully parallel work.  Kernel B (zsb_lz77_resolve.cu) then runs the greedy parse
// with nothing but these records and a bitmap: no window, no head[], no prev[] -- 4 KiB of state per segment instead
// of 192 KiB, so that dozens of segments are parsed per SM at once.  tools/ab_model/ab_model.c is the CPU model of
// both kernels (byte-identical to the oracle on the corpus and the synthetic inputs).
//
// One CTA of 1024 threads per segment (persistent, dynamic work counter).  Shared memory: a 64 KiB + 1 KiB window that
// slides by 32 KiB like the reference's (bulk async copy refills), head[32768] (latest position of every hash, window
// coordinates, rebased at a slide like SlideHash), and a ring of 34816 distances "position -> its predecessor" -- 2048
// more slots than the 32506 a chain may reach back, which is what lets 32 warps walk chains while later positions are
// being linked.  Linking is the one ordered step: groups of 32 consecutive positions take turns, and the turn passes from
// warp to warp through named hardware barriers (a polled shared-memory flag cost half of the SM's issue slots); inside a
// group, equal hashes are found with one warp match.  Adler-32 partial sums are taken from the staged bytes (no extra
// pass over the input).
//
// Record (32 bytes per position, HBM): 10 x u16 distance, 10 x u8 (prefix - 3), then 16 bits: ten flags "prefix >= 3",
// "the chain goes on", the number of candidates.
#include "zsb_kernels.cuh"
#include "zsb_lz77_dev.cuh"

namespace zsb {

namespace {

constexpr int LA_THREADS = 1024;
constexpr int WIN_AHEAD = 1024;                       // bytes staged beyond the chunk being linked (>= 258 + 16 + 5)
constexpr int WIN_BYTES = WINDOW_SIZE + WIN_AHEAD;    // 66560
constexpr int WIN_PADDED = WIN_BYTES + 64;
constexpr int NO_LINK = 0xFFFF;                       // ring entry of a position without a predecessor in reach (> MAX_DIST)
#ifndef ZSB_LA_NEAR
#define ZSB_LA_NEAR 12
#endif
constexpr int LA_NEAR = ZSB_LA_NEAR;                           // turns away at which a warp enters its barrier (< 15 - 1)
constexpr int RING = W_SIZE + 2048;                   // slots of the distance ring: MAX_DIST + twice the run-ahead of the linking

__device__ __forceinline__ uint32_t hash4(const uint16_t* crc, uint32_t v, int hash_variant)
{
    if (hash_variant == 0)   // Sse42.Crc32(0, v): Deflate.Intrinsics.cs:298-301
        return (crc[v & 0xFF] ^ crc[256 + ((v >> 8) & 0xFF)] ^ crc[512 + ((v >> 16) & 0xFF)] ^ crc[768 + (v >> 24)]) & HASH_MASK;
    return ((v * 2654435761u) >> 16) & HASH_MASK;   // Deflate.Intrinsics.cs:305
}

__device__ __forceinline__ void lds16(const uint8_t* base, int idx, uint32_t (&w)[4])
{
    const uint32_t* q = reinterpret_cast<const uint32_t*>(base + (idx & ~3));
    const uint32_t a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3], a4 = q[4];
    const int sh = (idx & 3) * 8;
    w[0] = __funnelshift_r(a0, a1, sh);
    w[1] = __funnelshift_r(a1, a2, sh);
    w[2] = __funnelshift_r(a2, a3, sh);
    w[3] = __funnelshift_r(a3, a4, sh);
}

// common prefix of the window strings at p and c, from byte `from` on (known equal before), capped at MAX_MATCH
__device__ __forceinline__ int prefix_from(const uint8_t* win, int p, int c, int from)
{
    int len = from;
    while (len < MAX_MATCH) {
        const uint32_t x = lds_u32_unaligned(win, p + len) ^ lds_u32_unaligned(win, c + len);
        if (x) {
            len += (__ffs(x) - 1) >> 3;
            break;
        }
        len += 4;
    }
    return len < MAX_MATCH ? len : MAX_MATCH;
}

}  // namespace

__global__ void __launch_bounds__(LA_THREADS, 1) lz77_lists_kernel(Lz77Params P)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* win = smem_raw;
    uint16_t* head = reinterpret_cast<uint16_t*>(smem_raw + WIN_PADDED);
    uint16_t* ring = head + HASH_SIZE;
    uint16_t* crc = ring + RING;
    uint64_t* bar = reinterpret_cast<uint64_t*>(crc + 1024);
    uint32_t* red = reinterpret_cast<uint32_t*>(bar + 1);   // 64 x u64 for adler_chunk
    volatile int* turn = reinterpret_cast<volatile int*>(red + 128);
    int* seg_slot = const_cast<int*>(turn) + 1;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    for (int i = tid; i < 1024; i += LA_THREADS) crc[i] = P.crc_tab[i];
    for (int i = tid; i < 64; i += LA_THREADS) win[WIN_BYTES + i] = 0;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    uint32_t phase = 0;
    const int N = (int)P.seg_size;
    const int nchunks = N / W_SIZE;

    for (;;) {
        if (tid == 0) *seg_slot = (int)atomicAdd(P.work_counter, 1u);
        __syncthreads();
        const int seg = *seg_slot;
        __syncthreads();
        if (seg >= (int)P.nseg) break;
        const uint8_t* seg_in = P.in + (size_t)seg * P.seg_size;
        uint4* seg_rec = P.lists + (size_t)seg * P.seg_size * 2;

        {   // head zeroed by the previous full flush (Deflate.cs:596-605); the ring needs no reset (only linked slots are read)
            const uint4 z = make_uint4(0, 0, 0, 0);
            uint4* h4 = reinterpret_cast<uint4*>(head);
            for (int i = tid; i < HASH_SIZE * 2 / 16; i += LA_THREADS) h4[i] = z;
        }
        uint32_t adler_a = 1, adler_b = 0;
        int loaded = 0;   // bytes of the segment staged so far
        __syncthreads();

        for (int c = 0; c < nchunks; c++) {
            // ---- stage: window index of segment position a is a - wbase, wbase = 32768 * (c - 1) ----
            const int wbase = W_SIZE * (c - 1);
            if (c > 0) {
                const uint4* src = reinterpret_cast<const uint4*>(win + W_SIZE);
                uint4* dst = reinterpret_cast<uint4*>(win);
                // (W_SIZE + WIN_AHEAD) / 16 = 2112 vectors by 1024 threads; the regions overlap by 1 KiB: read everything before any write
                constexpr int NV = (W_SIZE + WIN_AHEAD) / 16;
                const uint4 k0 = src[tid], k1 = src[tid + LA_THREADS];
                uint4 k2 = make_uint4(0, 0, 0, 0);
                if (tid + 2 * LA_THREADS < NV) k2 = src[tid + 2 * LA_THREADS];
                __syncthreads();
                dst[tid] = k0;
                dst[tid + LA_THREADS] = k1;
                if (tid + 2 * LA_THREADS < NV) dst[tid + 2 * LA_THREADS] = k2;
                uint32_t* hp = reinterpret_cast<uint32_t*>(head);   // SlideHash (Deflate.Intrinsics.cs:261-285) for the head table
                for (int i = tid; i < HASH_SIZE / 2; i += LA_THREADS) {
                    const uint32_t v = hp[i];
                    uint32_t lo = v & 0xFFFF, hi = v >> 16;
                    lo = lo >= W_SIZE ? lo - W_SIZE : 0;
                    hi = hi >= W_SIZE ? hi - W_SIZE : 0;
                    hp[i] = lo | (hi << 16);
                }
                __syncthreads();
            }
            {
                int want = W_SIZE * (c + 1) + WIN_AHEAD;
                if (want > N) want = N;
                const int n = want - loaded;
                if (n > 0) {
                    uint8_t* dst = win + (loaded - wbase);
                    stage_input<LA_THREADS>(dst, seg_in + loaded, n, bar, phase);
                    __syncthreads();
                    adler_chunk<LA_THREADS>(dst, n, red, adler_a, adler_b);
                    loaded += n;
                }
                if (loaded == N) {
                    // bytes past the end of the segment as the reference's window holds them when it gets there: the copy of
                    // the last 32 KiB that the final slide leaves in the upper half
                    const int e = N - wbase;
                    for (int i = e + tid; i < WIN_BYTES; i += LA_THREADS) win[i] = win[i - W_SIZE];
                }
                if (tid == 0) *turn = 0;
                __syncthreads();
            }

            // ---- link and walk: groups of 32 positions, warp w takes groups w, w + 32, ... of this 32 KiB chunk ----
            for (int t = 0; t < W_SIZE / LA_THREADS; t++) {
                const int g = t * 32 + warp;
                const int idx = W_SIZE + g * 32 + lane;          // window index
                const int a = W_SIZE * c + g * 32 + lane;        // position in the segment
                uint32_t d;
                int rs;   // ring slot of the node the walk stands on
                {
                    uint32_t scan[4];
                    lds16(win, idx, scan);
                    const uint32_t h = hash4(crc, __funnelshift_r(scan[0], scan[1], 16), P.hash_variant);   // Q1: bytes +2..+5
                    // lanes of this group with my hash (one MATCH instruction; 15 ballots cost ~100)
                    const uint32_t same = __match_any_sync(0xffffffffu, h);
                    const uint32_t below = same & lt;
                    const bool leader = below == 0, tail = (same >> lane) == 1u;
                    const int src = leader ? lane : 31 - __clz(below);
                    uint32_t slot = (uint32_t)a % RING;
                    int bar_in = 1 + g % 15, bar_out = 1 + (g + 1) % 15;
                    uint32_t pidx = __shfl_sync(0xffffffffu, (uint32_t)idx, src);   // predecessor inside the group (non-leaders)
                    asm volatile("" : "+r"(slot), "+r"(bar_in), "+r"(bar_out), "+r"(pidx));   // computed here, not inside the ordered step
                    // -- ordered: InsertString of every position, in position order --
                    // The turn passes through a hardware barrier (no polling, ~50 % of the issue slots went there): the holder
                    // of turn g - 1 arrives at barrier 1 + g % 15, the holder of turn g waits on it.  Fifteen barriers serve the
                    // turns that are at most NEAR away; a warp further ahead than that sleeps on the counter first, so a barrier
                    // is never entered for turn g + 15 before turn g is through.  The barrier orders the shared-memory accesses
                    // of the two warps (PTX: barrier.sync / barrier.arrive), turn after turn.
                    if (g > 0) {
                        if (lane == 0) {
                            int tnow;
                            while (g - (tnow = *turn) > LA_NEAR) __nanosleep(g - tnow > 2 * LA_NEAR ? 800 : 200);
                        }
                        __syncwarp();
                        asm volatile("barrier.sync %0, 64;" ::"r"(bar_in) : "memory");
                    }
                    const uint32_t pred = leader ? (uint32_t)head[h] : pidx;
                    if (tail) head[h] = (uint16_t)idx;
                    d = pred ? (uint32_t)idx - pred : (uint32_t)NO_LINK;
                    if (d > (uint32_t)MAX_DIST) d = NO_LINK;   // no predecessor, or one out of reach: the chain ends here
                    ring[slot] = (uint16_t)d;
                    rs = (int)slot;
                    if (g + 1 < W_SIZE / 32) asm volatile("barrier.arrive %0, 64;" ::"r"(bar_out) : "memory");
                    if (lane == 0) *turn = g + 1;
                }
                uint32_t scan[4];
                lds16(win, idx, scan);   // again: cheaper than keeping four registers across the ordered step (measured)
                // -- the LK nearest predecessors and their common prefix lengths --
                uint32_t dd[LZ_LIST_K], ll[LZ_LIST_K];
                uint32_t usem = 0, edge = 0;
                int total = 0, nk = 0;
#pragma unroll
                for (int k = 0; k < LZ_LIST_K; k++) {
                    dd[k] = 0;
                    ll[k] = 0;
                    // a node at exactly MAX_DIST can only be visited first (Deflate.cs:1031,1096): listed as candidate 0 only,
                    // otherwise left to kernel B's chain walk ("the chain goes on")
                    const int nt = total + (int)d;   // NO_LINK takes it past every limit, and stays there
                    if (k > 0 && nt == MAX_DIST) edge = 1u;
                    if (nt < MAX_DIST + (k == 0 ? 1 : 0)) {
                        total += (int)d;
                        const int cidx = idx - total;
                        // the next link first: the walk is a chain of dependent shared-memory reads, the compare is not on it
                        rs -= (int)d;
                        rs += rs < 0 ? RING : 0;
                        const uint32_t d_next = ring[rs];
                        uint32_t useful = 0, L = 0;
                        if (((lds_u32_unaligned(win, cidx) ^ scan[0]) & 0x00ffffffu) == 0) {
                            uint32_t cw[4];
                            lds16(win, cidx, cw);
                            const uint32_t x0 = cw[0] ^ scan[0], x1 = cw[1] ^ scan[1], x2 = cw[2] ^ scan[2], x3 = cw[3] ^ scan[3];
                            // first differing word without branches: the lanes of a warp stop in different words
                            uint32_t m = x2 ? x2 : x3;
                            int len = x2 ? 8 : 12;
                            m = x1 ? x1 : m;
                            len = x1 ? 4 : len;
                            m = x0 ? x0 : m;
                            len = x0 ? 0 : len;
                            if (m) len += (__ffs(m) - 1) >> 3;
                            else len = prefix_from(win, idx, cidx, 16);
                            useful = 1;
                            L = (uint32_t)(len - MIN_MATCH);
                        }
                        dd[k] = (uint32_t)total;
                        usem |= useful << k;
                        ll[k] = L;
                        nk = k + 1;
                        d = d_next;
                    } else d = NO_LINK;
                }
                const uint32_t more = total + (int)d <= MAX_DIST ? 1u : edge;
                uint4 r0, r1;
                r0.x = dd[0] | (dd[1] << 16);
                r0.y = dd[2] | (dd[3] << 16);
                r0.z = dd[4] | (dd[5] << 16);
                r0.w = dd[6] | (dd[7] << 16);
                r1.x = dd[8] | (dd[9] << 16);
                r1.y = ll[0] | (ll[1] << 8) | (ll[2] << 16) | (ll[3] << 24);
                r1.z = ll[4] | (ll[5] << 8) | (ll[6] << 16) | (ll[7] << 24);
                r1.w = ll[8] | (ll[9] << 8) | (usem << 16) | (more << 26) | ((uint32_t)nk << 27);
                seg_rec[(size_t)a * 2] = r0;
                seg_rec[(size_t)a * 2 + 1] = r1;
            }
            __syncthreads();
        }
        if (tid == 0) P.seg_adler[seg] = make_uint2(adler_a, adler_b);
        __syncthreads();
    }
}

size_t lz77_lists_smem_bytes()
{
    return WIN_PADDED + HASH_SIZE * 2 + RING * 2 + 1024 * 2 + 8 + 128 * 4 + 16;
}

cudaError_t launch_lz77_lists(const Lz77Params& P, int num_sms, cudaStream_t stream)
{
    const size_t smem = lz77_lists_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(lz77_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int grid = (int)(P.nseg < (uint32_t)num_sms ? P.nseg : (uint32_t)num_sms);
    if (grid < 1) grid = 1;
    lz77_lists_kernel<<<grid, LA_THREADS, smem, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace zsb
