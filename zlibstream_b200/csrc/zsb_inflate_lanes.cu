// zsb_inflate_lanes.cu -- the lane-per-segment inflate kernel (state machine and table build: zsb_inflate_lanes.cuh).
//
// One warp per CTA, one segment per lane; each lane pulls its next segment from the work counter as soon as it is done,
// so a warp never waits for its slowest lane beyond the iteration in flight.  Shared memory per CTA: 32 LaneTables
// (45 KiB) + the build scratch (2 KiB) => 4 CTAs per SM.
#include "zsb_inflate_lanes.cuh"

namespace zsb {

using namespace lanes;

namespace {

constexpr size_t kLaneSmem = 32 * sizeof(LaneTables) + sizeof(WarpScratch);
static_assert((32 * sizeof(LaneTables)) % 8 == 0, "WarpScratch follows the tables and holds 64-bit words");

__device__ __forceinline__ void next_segment(Lane& L, const InflateParams& P)
{
    const uint32_t seg = atomicAdd(P.work_counter, 1u);
    if (seg >= P.nseg) L.st = S_EXIT;
    else begin_segment(L, P, seg);
}

}  // namespace

__global__ void __launch_bounds__(32) inflate_lanes_kernel(InflateParams P, uint8_t* scratch)
{
    extern __shared__ __align__(16) uint8_t smem[];
    LaneTables* tabs = reinterpret_cast<LaneTables*>(smem);
    WarpScratch& sc = *reinterpret_cast<WarpScratch*>(smem + 32 * sizeof(LaneTables));
    const int lane = threadIdx.x;
    uint8_t* my_scratch = scratch + ((size_t)blockIdx.x * 32 + lane) * SCRATCH_STRIDE;
    uint8_t* glens = my_scratch;
    uint16_t* glong = reinterpret_cast<uint16_t*>(my_scratch + LENS_STRIDE);
    LaneTables& T = tabs[lane];

    Lane L;
    L.st = S_EXIT;
    next_segment(L, P);

    for (;;) {
        lane_step(L, T, glens, glong);
        if (L.st == S_FINISH) {
            finish_segment(L, P);
            next_segment(L, P);
        }
        // serve the lanes that wait for a table, one after the other, with the whole warp
        unsigned need = __ballot_sync(0xffffffffu, L.st == S_BUILD_BL || L.st == S_BUILD_LD);
        while (need) {
            const int t = __ffs(need) - 1;
            need &= need - 1;
            const bool ld = __shfl_sync(0xffffffffu, L.st, t) == S_BUILD_LD;
            const uint64_t bl = __shfl_sync(0xffffffffu, (unsigned long long)L.bl_lens, t);
            const int nl = __shfl_sync(0xffffffffu, L.nl, t), nd = __shfl_sync(0xffffffffu, L.nd, t);
            const int fixed = __shfl_sync(0xffffffffu, L.fixed, t);
            uint8_t* t_scratch = scratch + ((size_t)blockIdx.x * 32 + t) * SCRATCH_STRIDE;
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
        if (__all_sync(0xffffffffu, L.st == S_EXIT)) break;
    }
}

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

size_t inflate_lanes_scratch_bytes(int num_sms)
{
    return (size_t)num_sms * 8 * 32 * SCRATCH_STRIDE;   // up to 8 resident CTAs per SM
}

cudaError_t launch_inflate_lanes(const InflateParams& P, int num_sms, uint8_t* scratch, size_t scratch_bytes, cudaStream_t stream)
{
    if (P.nseg == 0) return cudaSuccess;
    static int per_sm = 0;
    if (per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(inflate_lanes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLaneSmem);
        if (e != cudaSuccess) return e;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, inflate_lanes_kernel, 32, kLaneSmem) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        if (per_sm > 8) per_sm = 8;
    }
    int ctas = num_sms * per_sm;
    const int need = (int)((P.nseg + 31) / 32);
    if (ctas > need) ctas = need;
    if ((size_t)ctas * 32 * SCRATCH_STRIDE > scratch_bytes) return cudaErrorInvalidValue;
    inflate_lanes_kernel<<<ctas, 32, kLaneSmem, stream>>>(P, scratch);
    return cudaGetLastError();
}

}  // namespace zsb
