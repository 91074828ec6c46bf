// zsb_inflate_lanes.cu -- the lane-per-segment inflate kernel (state machine and table build: zsb_inflate_lanes.cuh).
//
// One CTA per SM with as many warps as shared memory holds tables for, one segment per lane; each lane pulls its next
// segment from the work counter as soon as it is done, so a warp never waits for its slowest lane beyond the iteration in
// flight, and the warps of a CTA never wait for each other except for the build scratch.  Shared memory per CTA: 32
// LaneTables per warp (29 KiB; 15.5 KiB in the small instance) + ONE build scratch (2 KiB): a build takes ~2 k cycles
// and a warp asks for one every ~1 M cycles, so the warps take turns on it behind a lock.  (One warp per CTA with a
// scratch each, the round-1 layout, pays 1 KiB of per-CTA reservation + 2 KiB of scratch per warp: 12 warps per SM for
// the small instance instead of 14.)
#include "zsb_inflate_lanes.cuh"

#include <cstdlib>

namespace zsb {

using namespace ZSB_LN_NS;

// names of this translation unit's kernel and launcher (zsb_inflate_lanes_small.cu renames them)
#ifndef ZSB_LN_KERNEL
#define ZSB_LN_KERNEL inflate_lanes_kernel
#define ZSB_LN_LAUNCH launch_inflate_lanes
#endif

namespace {

static_assert(sizeof(LaneTables) % 4 == 0 && sizeof(WarpScratch) % 8 == 0, "shared-memory layout: scratch first, then the tables");
constexpr int MAX_WARPS = 16;   // per CTA (= per SM): 118 registers per thread allow no more
// Warps per SM at 32 segments per warp.  The tables of the small instance would fit 14 times, but past 196 KiB of shared
// memory the SM's L1 shrinks from 60 to 28 KiB and every lane slows down by a third (the compressed input is read 4 bytes at
// a time and the match sources 8: both live on L1 hits).  16 GiB of 64 KiB segments on a B200: 8 warps 111 GB/s, 10 126,
// 11 131, 12 143, 13 131, 14 134.
#ifndef ZSB_LN_WARPS_FULL
#define ZSB_LN_WARPS_FULL MAX_WARPS
#endif
size_t lane_smem(int warps, int segs_per_warp, int nscratch) { return (size_t)nscratch * sizeof(WarpScratch) + (size_t)warps * segs_per_warp * sizeof(LaneTables); }

__device__ __forceinline__ void next_segment(Lane& L, const InflateParams& P)
{
    const uint32_t seg = atomicAdd(P.work_counter, 1u);
    if (seg >= P.nseg) L.st = S_EXIT;
    else begin_segment(L, P, seg);
}

}  // namespace

// segs_per_warp (S): lanes 0..S-1 decode a segment each, the others only help with the table builds.  Fewer segments
// per warp = fewer divergent paths serialised per iteration and more warps for the same number of segments; the
// launcher picks S from the number of segments so that the SMs have enough warps to hide latency.
// nscratch: build scratches of the CTA; warp w uses scratch w % nscratch (behind its lock when it shares it).
__global__ void __launch_bounds__(32 * MAX_WARPS) ZSB_LN_KERNEL(InflateParams P, uint8_t* scratch, int segs_per_warp, int nscratch)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratch& sc = reinterpret_cast<WarpScratch*>(smem)[warp % nscratch];
    const size_t gwarp = (size_t)blockIdx.x * (blockDim.x >> 5) + warp;
    LaneTables* tabs = reinterpret_cast<LaneTables*>(smem + (size_t)nscratch * sizeof(WarpScratch)) + warp * segs_per_warp;
    if (threadIdx.x < nscratch) reinterpret_cast<WarpScratch*>(smem)[threadIdx.x].lock = 0;
    __syncthreads();
    uint8_t* my_scratch = scratch + (gwarp * 32 + lane) * SCRATCH_STRIDE;
    uint8_t* glens = my_scratch;
    uint16_t* glong = reinterpret_cast<uint16_t*>(my_scratch + LENS_STRIDE);
    LaneTables& T = tabs[lane < segs_per_warp ? lane : 0];

    Lane L;
    L.st = S_EXIT;
    if (lane < segs_per_warp) next_segment(L, P);

    for (;;) {
        lane_step(L, T, glens, glong);
        if (L.st == S_FINISH) {
            finish_segment(L, P);
            next_segment(L, P);
        }
        // serve the lanes that wait for a table, one after the other, with the whole warp
        unsigned need = __ballot_sync(0xffffffffu, L.st == S_BUILD_BL || L.st == S_BUILD_LD);
        const bool locked = need != 0 && (int)(blockDim.x >> 5) > nscratch;
        if (locked) {   // the scratch is the CTA's: every word of it is written by the build before the build reads it
            if (lane == 0)
                while (atomicCAS(&sc.lock, 0, 1) != 0) __nanosleep(200);
            __syncwarp();
        }
        while (need) {
            const int t = __ffs(need) - 1;
            need &= need - 1;
            const bool ld = __shfl_sync(0xffffffffu, L.st, t) == S_BUILD_LD;
            const uint64_t bl = __shfl_sync(0xffffffffu, (unsigned long long)L.bl_lens, t);
            const int nl = __shfl_sync(0xffffffffu, L.nl, t), nd = __shfl_sync(0xffffffffu, L.nd, t);
            const int fixed = __shfl_sync(0xffffffffu, L.fixed, t);
            uint8_t* t_scratch = scratch + (gwarp * 32 + t) * SCRATCH_STRIDE;
            LaneTables& TT = tabs[t];
            HelperRegs R;
            __syncwarp();   // the target's code lengths (global) and everybody's table reads are done
            build_p0(lane, sc, TT, R, ld, bl, t_scratch, nl, nd, fixed);
            __syncwarp();
            build_p1(lane, sc, ld);
            __syncwarp();
            build_p2(lane, sc, TT, ld);
            __syncwarp();
            build_p3(lane, sc, TT, R, ld, reinterpret_cast<uint16_t*>(t_scratch + LENS_STRIDE), fixed);
            __syncwarp();
            const int rc_l = sc.rc_l, rc_d = sc.rc_d;
            if (lane == t) after_build(L, ld, rc_l, rc_d);
            __syncwarp();
        }
        if (locked) {
            __threadfence_block();
            __syncwarp();
            if (lane == 0) atomicExch(&sc.lock, 0);
        }
        if (__all_sync(0xffffffffu, L.st == S_EXIT)) break;
    }
}

#ifndef ZSB_LN_SECONDARY
// Q14: one warp per segment that opens with a carried run.
__global__ void __launch_bounds__(256) inflate_carry_fixup_kernel(InflateParams P)
{
    const uint32_t warps = gridDim.x * (blockDim.x / 32), lane = threadIdx.x & 31;
    for (uint32_t k = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32; k < P.nseg; k += warps) {
        const uint32_t c = P.seg_carry[k];
        if (!c || P.seg_status[k] != IE_OK) continue;
        uint8_t val = 0;
        if (!carry_fill_value(P, k, val)) continue;
        uint8_t* o = P.out + (uint64_t)k * P.seg_size;
        for (uint32_t i = lane; i < c; i += 32) o[i] = val;
    }
}

cudaError_t launch_inflate_carry_fixup(const InflateParams& P, cudaStream_t stream)
{
    if (P.nseg == 0 || !P.seg_carry) return cudaSuccess;
    int ctas = (int)((P.nseg + 7) / 8);
    if (ctas > 148 * 8) ctas = 148 * 8;
    inflate_carry_fixup_kernel<<<ctas, 256, 0, stream>>>(P);
    return cudaGetLastError();
}

// ---- segment index for streams without a side-car (mirrors InflateSync, Inflate.cs:386-448) --------------
// Every full flush ends with an empty stored block whose LEN/NLEN bytes are 00 00 FF FF (Deflate.cs:710-712).  This kernel
// lists the END offset of every occurrence; the host sorts them, has the pieces in between decoded in count-only mode and
// merges pieces into segments (zsb_api.cu: build_segment_index).
__global__ void __launch_bounds__(256) index_markers_kernel(const uint8_t* in, size_t begin, size_t end, uint64_t* cand,
                                                            uint32_t cand_cap, uint32_t* ncand)
{
    // thread = one 4-byte aligned word and the 4 marker positions that start inside it
    const uintptr_t base = reinterpret_cast<uintptr_t>(in) & ~(uintptr_t)3;
    const size_t delta = reinterpret_cast<uintptr_t>(in) - base;       // in = base + delta
    const uint32_t* words = reinterpret_cast<const uint32_t*>(base);
    const size_t w_first = (begin + delta) / 4, w_last = (end + delta + 3) / 4;   // words [w_first, w_last)
    for (size_t w = w_first + (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < w_last; w += (size_t)gridDim.x * blockDim.x) {
        const uint64_t lo = __ldg(words + w);
        const uint64_t hi = w + 1 < w_last ? __ldg(words + w + 1) : 0;
        const uint64_t v = lo | (hi << 32);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if ((uint32_t)(v >> (8 * k)) != 0xFFFF0000u) continue;
            const size_t p = w * 4 + k;               // offset from base
            if (p < begin + delta || p + 4 > end + delta) continue;
            const uint32_t slot = atomicAdd(ncand, 1u);
            if (slot < cand_cap) cand[slot] = (uint64_t)(p - delta + 4);
        }
    }
}

cudaError_t launch_index_markers(const uint8_t* in, size_t begin, size_t end, uint64_t* cand, uint32_t cand_cap, uint32_t* ncand,
                                 cudaStream_t stream)
{
    if (end < begin + 4) return cudaSuccess;
    const size_t words = (end - begin) / 4 + 2;
    size_t ctas = (words + 255) / 256;
    if (ctas > 148 * 16) ctas = 148 * 16;
    index_markers_kernel<<<(int)ctas, 256, 0, stream>>>(in, begin, end, cand, cand_cap, ncand);
    return cudaGetLastError();
}

size_t inflate_lanes_scratch_bytes(int num_sms)
{
    return (size_t)num_sms * 32 * 32 * SCRATCH_STRIDE;   // up to 32 resident CTAs per SM
}

#endif  // ZSB_LN_SECONDARY

cudaError_t ZSB_LN_LAUNCH(const InflateParams& P, int num_sms, uint8_t* scratch, size_t scratch_bytes, cudaStream_t stream)
{
    if (P.nseg == 0) return cudaSuccess;
    int smem_max = 0, dev = 0;
    {   // a per-device attribute: set on every launch (a process may drive several GPUs, one zsb_ctx each)
        cudaError_t e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ZSB_LN_KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
        if (e != cudaSuccess) return e;
    }
    // warps one SM holds with S segments per warp: shared memory for the tables, MAX_WARPS for the registers
    auto warps_per_sm = [&](int S) {
        int w = (int)(((size_t)smem_max - sizeof(WarpScratch)) / ((size_t)S * sizeof(LaneTables)));   // with one scratch; more if room is left
        const char* e = getenv("ZSB_INFLATE_WARPS");   // experiments: cap the warps per SM
        const int cap = e ? atoi(e) : 0;
        if (cap >= 1 && w > cap) w = cap;
        else if (!cap && S == 32 && w > ZSB_LN_WARPS_FULL) w = ZSB_LN_WARPS_FULL;
        return w > MAX_WARPS ? MAX_WARPS : w < 1 ? 1 : w;
    };
    // Segments per warp: as many as still leave enough warps per SM (ZSB_INFLATE_SEGS_PER_WARP overrides).
    const int forced = [] {
        const char* e = getenv("ZSB_INFLATE_SEGS_PER_WARP");
        const int v = e ? atoi(e) : 0;
        return v >= 1 && v <= 32 ? v : 0;
    }();
    int S = forced;
    if (!S) {
        // measured on B200 (profiles/r01i_inflate_sweep.txt, r01n_inflate_sweep.txt): with enough segments to keep every SM's resident
        // warps full, 32 per warp is fastest (issue slots are shared 32 ways); below that, 16 per warp trades issue
        // efficiency for more warps in flight (16384 segments: 58.5 GB/s at 16, 55.9 at 32, 33.9 at 6)
        S = (size_t)P.nseg * 4 >= (size_t)num_sms * warps_per_sm(32) * 32 * 3 ? 32 : 16;
    }
    int W = warps_per_sm(S);
    const size_t need = ((size_t)P.nseg + S - 1) / S;            // warps that have a first segment to take
    int ctas = num_sms;
    if (need < (size_t)num_sms * W) {
        W = (int)((need + num_sms - 1) / num_sms);
        ctas = (int)((need + W - 1) / W);
    }
    if ((size_t)ctas * W * 32 * SCRATCH_STRIDE > scratch_bytes) return cudaErrorInvalidValue;
    // Build scratches: one per warp if shared memory has the room (the lanes of equal segments reach their block headers
    // together, and a single scratch then serialises 7 x 32 builds: 32 768 segments 98 instead of 105 GB/s); but never across
    // the 196 KiB carve-out step, which would halve the L1 (see ZSB_LN_WARPS_FULL).
    const size_t tables = (size_t)W * S * sizeof(LaneTables), step = (size_t)195 * 1024;   // 196 KiB minus the 1 KiB the system reserves
    const size_t limit = tables + sizeof(WarpScratch) <= step ? step : (size_t)smem_max;
    int NS = (int)((limit - tables) / sizeof(WarpScratch));
    NS = NS >= W ? W : 1;   // shared anyway: one scratch (four for twelve warps measured 141.5 against 143.1 GB/s at 16 GiB)
    ZSB_LN_KERNEL<<<ctas, 32 * W, lane_smem(W, S, NS), stream>>>(P, scratch, S, NS);
    return cudaGetLastError();
}

}  // namespace zsb
