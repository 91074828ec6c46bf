// Host emulation of the lane-per-segment inflate kernel (zlibstream_b200/csrc/zsb_inflate_lanes.cuh), compiled by nvcc
// for the HOST only.  TEST INFRASTRUCTURE: it lets `-m "not gpu"` tests run the kernel's state machine and table build
// (the very same __host__ __device__ code) against the oracle's streams.  Nothing in the product links or calls this.
#include <vector>

#include "../../zlibstream_b200/csrc/zsb_inflate_lanes.cuh"

using namespace zsb;
using namespace zsb::lanes;

extern "C" __attribute__((visibility("default"))) int zsb_lanes_host_inflate(const uint8_t* in, size_t n, const uint64_t* seg_end,
                                                                             uint32_t nseg, uint64_t first_offset,
                                                                             uint32_t seg_size, uint8_t* out, size_t out_cap,
                                                                             uint32_t* seg_out_len, int* seg_status)
{
    unsigned int counter = 0;
    InflateParams P;
    P.in = in;
    P.n = n;
    P.seg_end = seg_end;
    P.first_offset = first_offset;
    P.nseg = nseg;
    P.seg_size = seg_size;
    P.out = out;
    P.out_cap = out_cap;
    P.seg_out_len = seg_out_len;
    P.seg_adler = nullptr;
    P.seg_status = seg_status;
    P.work_counter = &counter;
    std::vector<uint32_t> carry(nseg ? nseg : 1, 0);
    unsigned int carry_flag = 0;
    P.seg_carry = carry.data();
    P.carry_flag = &carry_flag;
    std::vector<LaneTables> tabs(32);
    std::vector<uint8_t> scratch(32 * SCRATCH_STRIDE);
    WarpScratch sc;
    Lane L[32];
    auto next = [&](Lane& l) {
        const uint32_t seg = counter++;
        if (seg >= nseg) l.st = S_EXIT;
        else begin_segment(l, P, seg);
    };
    for (int i = 0; i < 32; i++) {
        L[i].st = S_EXIT;
        next(L[i]);
    }
    for (;;) {
        bool all_exit = true;
        for (int i = 0; i < 32; i++) {
            uint8_t* s = scratch.data() + i * SCRATCH_STRIDE;
            lane_step(L[i], tabs[i], s, reinterpret_cast<uint16_t*>(s + LENS_STRIDE));
            if (L[i].st == S_FINISH) {
                finish_segment(L[i], P);
                next(L[i]);
            }
        }
        for (int t = 0; t < 32; t++) {
            if (L[t].st != S_BUILD_BL && L[t].st != S_BUILD_LD) continue;
            const bool ld = L[t].st == S_BUILD_LD;
            uint8_t* s = scratch.data() + t * SCRATCH_STRIDE;
            HelperRegs R[32];
            for (int h = 0; h < 32; h++) build_p0(h, sc, tabs[t], R[h], ld, L[t].bl_lens, s, L[t].nl, L[t].nd, L[t].fixed);
            for (int h = 0; h < 32; h++) build_p1(h, sc, ld);
            for (int h = 0; h < 32; h++) build_p2(h, sc, tabs[t], ld);
            for (int h = 0; h < 32; h++) build_p3(h, sc, tabs[t], R[h], ld, reinterpret_cast<uint16_t*>(s + LENS_STRIDE), L[t].fixed);
            after_build(L[t], ld, sc.rc_l, sc.rc_d);
        }
        for (int i = 0; i < 32; i++) all_exit = all_exit && L[i].st == S_EXIT;
        if (all_exit) break;
    }
    if (carry_flag)
        for (uint32_t k = 0; k < nseg; k++) {   // what inflate_carry_fixup_kernel does
            uint8_t val = 0;
            if (carry[k] && seg_status[k] == IE_OK && carry_fill_value(P, k, val)) memset(out + (uint64_t)k * seg_size, val, carry[k]);
        }
    return 0;
}
